"""First-contact GPU probe: DMMA peak, cuBLAS dgemm yardstick, and a quick check/timing of our fp64 GEMM."""
import ctypes as C
import json
import subprocess
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from marlin_b200 import _native as nat

out = {}
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total", "--format=csv"],
                     capture_output=True, text=True).stdout)
print(subprocess.run([str(Path(__file__).parent / "bin" / "dmma_bench")], capture_output=True, text=True).stdout)

lib = nat.load(build_if_missing=False)
ctx = nat.c_ctx()
nat.check(lib.mb_init(0, C.byref(ctx)))
torch.cuda.set_device(0)
lib.mb_set_stream(ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream))


def ours(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha=1.0, beta=0.0, generic=False):
    fn = lib.mb_dgemm_device_generic if generic else lib.mb_dgemm_device
    nat.check(fn(ctx, ta.encode(), tb.encode(), m, n, k, alpha, C.c_void_p(A.data_ptr()), lda,
                 C.c_void_p(B.data_ptr()), ldb, beta, C.c_void_p(Cm.data_ptr()), ldc))


def check(m, n, k, ta="N", tb="N", generic=False):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    # column-major storage: tensor of shape (cols, ld) viewed transposed
    Ast = torch.rand((m if ta == "T" else k), (k if ta == "T" else m), device="cuda", dtype=torch.float64, generator=g) - 0.3
    Bst = torch.rand((k if tb == "T" else n), (n if tb == "T" else k), device="cuda", dtype=torch.float64, generator=g) - 0.3
    A = Ast.t() if ta == "N" else Ast          # logical M x K
    Bm = Bst.t() if tb == "N" else Bst         # logical K x N
    Cst = torch.full((n, m), float("nan"), device="cuda", dtype=torch.float64)
    ours(ta, tb, m, n, k, Ast, Ast.shape[1], Bst, Bst.shape[1], Cst, m, generic=generic)
    torch.cuda.synchronize()
    ref = A @ Bm
    got = Cst.t()
    denom = (A.abs() @ Bm.abs()).clamp_min(1e-300)
    err = ((got - ref).abs() / denom).max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    return err, rel


fails = 0
for (m, n, k) in [(4, 4, 4), (8, 8, 4), (100, 100, 100), (128, 128, 16), (130, 126, 50), (257, 511, 1000), (1024, 1024, 1024), (2, 2, 2), (64, 32, 8)]:
    for ta in "NT":
        for tb in "NT":
            if (ta == "T" and m % 2) or (tb == "T" and False):
                pass
            try:
                err, rel = check(m, n, k, ta, tb)
                ok = err < 1e-13 and rel < 1e-13
            except Exception as e:  # noqa
                err, rel, ok = str(e), None, False
            fails += (not ok)
            print(f"check {m}x{n}x{k} {ta}{tb}: scaled_err={err} rel={rel} {'OK' if ok else 'FAIL'}")
err, rel = check(300, 200, 100, "N", "N", generic=True)
print("generic kernel:", err, rel)
out["fails"] = fails


def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n in (4096, 8192, 16384):
    A = torch.rand(n, n, device="cuda", dtype=torch.float64)
    B = torch.rand(n, n, device="cuda", dtype=torch.float64)
    Cm = torch.empty(n, n, device="cuda", dtype=torch.float64)
    iters = 5 if n <= 8192 else 2
    t_cublas = timeit(lambda: torch.matmul(A, B, out=Cm), iters)
    t_ours = timeit(lambda: ours("N", "N", n, n, n, A, n, B, n, Cm, n), iters)
    fl = 2.0 * n ** 3
    out[f"n{n}"] = {"cublas_ms": t_cublas, "cublas_tflops": fl / t_cublas / 1e9, "ours_ms": t_ours, "ours_tflops": fl / t_ours / 1e9}
    print(n, out[f"n{n}"])
    # full-size sanity vs cuBLAS (both col-major interpretations: torch row-major => C^T = B^T A^T; compare accordingly)
    ours("N", "N", n, n, n, A, n, B, n, Cm, n)
    ref = torch.matmul(B, A)   # col-major view: A^T_rm... (A_cm = A_rm^T): C_cm = A_cm B_cm = (B_rm A_rm)^T -> stored as B_rm @ A_rm
    rel = ((Cm - ref).norm() / ref.norm()).item()
    print(f"  full-size rel err vs cuBLAS: {rel:.3e}")
    out[f"n{n}"]["rel_vs_cublas"] = rel
    del A, B, Cm, ref

Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/probe.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
