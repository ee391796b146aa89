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
