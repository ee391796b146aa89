"""examples/RMMcompare.scala:14-55 — `RMMcompare <A rows> <A cols> <B cols> <mode> <m> <k> <n>`; only mode 2 (RMMv2 =
BlockMatrix.multiply(other: BlockMatrix)) is live in the reference."""
import sys

from ._common import millis, start, stop


def main(args):
    if len(args) < 7:
        print("usage: rmm_compare <rows of A> <cols of A> <cols of B> <mode> <m> <k> <n>   (mode 2 = BlockMatrix x BlockMatrix,")
        print("  the only mode the reference's RMMcompare still runs), e.g. rmm_compare 30000 30000 30000 2 6 6 6")
        sys.exit(1)
    mb, rank = start()
    rowA, colA, colB, mode = int(args[0]), int(args[1]), int(args[2]), int(args[3])
    m, k, n = int(args[4]), int(args[5]), int(args[6])
    matrixA = mb.MTUtils.randomBlockMatrix(None, rowA, colA, m, k)
    matrixB = mb.MTUtils.randomBlockMatrix(None, colA, colB, k, n)
    if rank == 0:
        print("=========================================")
        print(f"RMMcompare matrixA: {rowA} by {colA} ; matrixB: {colA} by {colB}, mode: {mode} m, k, n: {m}, {k}, {n}")
    if mode == 2:
        t0 = millis()
        result = matrixA.multiply(matrixB)
        mb.MTUtils.evaluate(result)
        if rank == 0:
            print(f"RMMv2 in mode {mode} used time {millis() - t0:.0f} millis")
    elif rank == 0:
        print("only mode 2 is implemented by the reference as shipped")
    stop()


if __name__ == "__main__":
    main(sys.argv[1:])
