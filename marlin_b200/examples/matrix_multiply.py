"""examples/MatrixMultiply.scala:16-49 — `MatrixMultiply <A rows> <A cols / B rows> <B cols> <cores> [<broadcast threshold>]`:
two random DenseVecMatrix, the auto-strategy multiply, then a count of the result."""
import sys

from ._common import start, stop


def main(args):
    if len(args) < 4:
        sys.stderr.write("usage: matrix_multiply <rows of A> <cols of A = rows of B> <cols of B> <cores> [<broadcast threshold, MB>]\n")
        sys.exit(-1)
    mb, rank = start()
    rowA, colA, colB = int(args[0]), int(args[1]), int(args[2])
    ma = mb.MTUtils.randomDenVecMatrix(None, rowA, colA)
    mbm = mb.MTUtils.randomDenVecMatrix(None, colA, colB)
    threshold = int(args[4]) if len(args) >= 5 else 300       # (the reference reads args(5) here, an off-by-one at :45)
    result = ma.multiply(mbm, int(args[3]), threshold)
    count = result.elementsCount() if hasattr(result, "elementsCount") else result.numRows()
    if rank == 0:
        print(f"Result RDD counts: {count}")
    stop()


if __name__ == "__main__":
    main(sys.argv[1:])
