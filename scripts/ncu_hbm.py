import ctypes as C, sys
sys.path.insert(0, '.')
import marlin_b200 as mb
from marlin_b200 import _native as nat
import torch
rt = mb.Runtime.get()
n = 8192
A = mb.MTUtils.randomBlockMatrix(None, n, n, 1, 1, seed=1).blocks[0][1]
B = mb.MTUtils.randomBlockMatrix(None, n, n, 1, 1, seed=2).blocks[0][1]
for _ in range(2):
    A.add(B); A.multiply(2.0); A.transpose(); A.copy(); A.sum()
torch.cuda.synchronize()
# vector side: gemv (plain and row-major view), rank-1 product, dot
x, y = mb.SubMatrix.empty(n, 1), mb.SubMatrix.empty(n, 1)
nat.check(rt.lib.mb_fill_uniform(rt.ctx, x.handle(), 3, 0, 0.0, 1.0, 0))
for _ in range(2):
    A.multiply(x, out=y); A.t.multiply(x, out=y); x.outer(x); x.dot(x)
    mb.SubMatrix(buf=A.buf, rows=n * n, cols=1, ld=n * n).dot(mb.SubMatrix(buf=B.buf, rows=n * n, cols=1, ld=n * n))
torch.cuda.synchronize()
