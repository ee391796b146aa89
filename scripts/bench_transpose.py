"""Sweep of the transpose kernel's tile -> CTA maps / sub-tile shapes (MARLIN_B200_TRANSPOSE_VARIANT), one subprocess per
variant.  Prints GB/s at 8192^2 and 16384^2 (algorithmic bytes = 2*8*M*N) and checks the result bit for bit."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
import marlin_b200 as mb
from marlin_b200 import _native as nat
rt = mb.Runtime.get()
res = {}
for n, m in ((8192, 8192), (16384, 16384), (16384, 4096)):
    A = mb.MTUtils.randomBlockMatrix(None, n, m, 1, 1, seed=1).blocks[0][1]
    T = mb.SubMatrix.empty(m, n)
    rt.sync_stream()
    f = lambda: nat.check(rt.lib.mb_block_transpose(rt.ctx, A.handle(), T.handle()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ok = bool(torch.equal(T.buf.view(n, m), A.buf.view(m, n).t()))
    res[f"{n}x{m}"] = {"ms": ms, "GB/s": n * m * 16 / ms / 1e6, "exact": ok}
    del A, T
print("RES " + json.dumps(res))
''' % str(ROOT)

out = {}
for v in [int(a) for a in sys.argv[1:]] or list(range(10)):
    env = dict(os.environ, MARLIN_B200_TRANSPOSE_VARIANT=str(v))
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    line = [l for l in p.stdout.splitlines() if l.startswith("RES ")]
    out[v] = json.loads(line[0][4:]) if line else {"error": p.stdout[-500:]}
    print(v, json.dumps(out[v]), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "transpose_sweep.json").write_text(json.dumps(out, indent=1))
