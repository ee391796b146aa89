"""examples/BLAS3.scala:14-60 — `BLAS3 <A rows> <A cols> <B cols> <mode> <m> <k> <n>`:
mode 1 local multiply, mode 2 broadcast multiply (DenseVecMatrix x local matrix), mode 3 shuffle multiply with (m,k,n)."""
import sys

import numpy as np

from ._common import millis, start, stop


def main(args):
    if len(args) < 4:
        print("usage: blas3 <rows of A> <cols of A> <cols of B> <mode> [<m> [<k> <n>]]")
        print("  mode 1: both operands as one block each on one GPU;  mode 2: row-sharded A times a replicated local B (m shards);")
        print("  mode 3: both operands distributed, split (m, k, n).   e.g. blas3 10000 10000 10000 3 2 2 2")
        sys.exit(1)
    mb, rank = start()
    rowA, colA, colB, mode = int(args[0]), int(args[1]), int(args[2]), int(args[3])
    if rank == 0:
        print(f"matrixA: {rowA} by {colA} ; matrixB: {colA} by {colB}, mode: {mode}")
    if mode == 1:                                           # :29-35 — one block on one GPU
        t0 = millis()
        a = mb.MTUtils.randomBlockMatrix(None, rowA, colA, 1, 1)
        b = mb.MTUtils.randomBlockMatrix(None, colA, colB, 1, 1)
        a.multiply(b)
        used = millis() - t0
    elif mode == 2:                                         # :36-45
        m = int(args[4])
        matrixA = mb.MTUtils.randomDenVecMatrix(None, rowA, colA, m)
        matrixB = np.random.default_rng(0).random((colA, colB))
        t0 = millis()
        matrixA.multiply(matrixB)
        used = millis() - t0
    else:                                                   # :46-56
        m, k, n = int(args[4]), int(args[5]), int(args[6])
        matrixA = mb.MTUtils.randomDenVecMatrix(None, rowA, colA)
        matrixB = mb.MTUtils.randomDenVecMatrix(None, colA, colB)
        t0 = millis()
        matrixA.multiply(matrixB, (m, k, n))
        used = millis() - t0
    if rank == 0:
        print(f"multiplication in mode {mode} used time {used:.0f} millis")
    stop()


if __name__ == "__main__":
    main(sys.argv[1:])
