"""fp64-on-int8-tensor-cores (Ozaki split) vs the native DMMA kernel: time and error, 1 GPU."""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import marlin_b200 as mb
from marlin_b200 import _native as nat

rt = mb.Runtime.get()
lib, ctx = rt.lib, rt.ctx
out = {}


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


sizes = [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192, 16384]
for n in sizes:
    A = mb.MTUtils.randomBlockMatrix(None, n, n, 1, 1, seed=1).blocks[0][1]
    B = mb.MTUtils.randomBlockMatrix(None, n, n, 1, 1, seed=2).blocks[0][1]
    Cn = mb.SubMatrix.empty(n, n)
    Co = mb.SubMatrix.empty(n, n)
    rt.sync_stream()
    fl = 2.0 * n ** 3
    iters = 5 if n <= 8192 else 2
    nat.check(lib.mb_set_fp64_mode(ctx, 0, 7))
    t_nat = timeit(lambda: A.multiply(B, out=Cn), iters)
    res = {"native_ms": t_nat, "native_tflops": fl / t_nat / 1e9}
    ref = Cn.buf[: n * n]
    for mode, s in ((1, 6), (1, 7), (2, 5), (2, 6), (2, 4)):
        nat.check(lib.mb_set_fp64_mode(ctx, mode, s))
        t = timeit(lambda: A.multiply(B, out=Co), iters)
        got = Co.buf[: n * n]
        rel = ((got - ref).abs() / ref.abs()).max().item()      # U[0,1) inputs: ref = (|A||B|)_ij
        tag = f"b{7 if mode == 1 else 8}s{s}"
        res[f"{tag}_ms"] = round(t, 3)
        res[f"{tag}_tflops_equiv"] = round(fl / t / 1e9, 2)
        res[f"{tag}_int8_tops"] = round(fl * (s * (s + 1) // 2) / t / 1e9, 1)
        res[f"{tag}_max_scaled_err_vs_native"] = rel
    nat.check(lib.mb_set_fp64_mode(ctx, 0, 7))
    out[str(n)] = res
    print(n, json.dumps(res), flush=True)
    del A, B, Cn, Co
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/ozaki.json").write_text(json.dumps(out, indent=1))
