"""Golden vectors of the reference's own suite for the multiply / transpose / add path.

Every literal below is taken from /root/reference/src/test/scala/edu/nju/pasalab/marlin/matrix/
DistributedMatrixSuite.scala (cited per item).  All expected values are small integers (or .5), so the
assertions are exact regardless of summation order — this is what pins the oracle and the CUDA path.
"""
import numpy as np

# DistributedMatrixSuite.scala:15-19 — indexed rows, deliberately out of order
DATA_ROWS = [(0, [0.0, 1.0, 2.0, 3.0]), (2, [3.0, 2.0, 1.0, 0.0]), (3, [1.0, 1.0, 1.0, 1.0]), (1, [2.0, 3.0, 4.0, 5.0])]
# :20-24 — the same matrix as a 2x2 grid of 2x2 blocks
BLKS = [((0, 0), [[0.0, 1.0], [2.0, 3.0]]), ((0, 1), [[2.0, 3.0], [4.0, 5.0]]),
        ((1, 0), [[3.0, 2.0], [1.0, 1.0]]), ((1, 1), [[1.0, 0.0], [1.0, 1.0]])]
# :75-79
EXPECTED_DENSE = np.array([[0.0, 1.0, 2.0, 3.0], [2.0, 3.0, 4.0, 5.0], [3.0, 2.0, 1.0, 0.0], [1.0, 1.0, 1.0, 1.0]])
# :228-232 (and :241-245, :258-262, :273-277, :293-297, :426-430, :441-445)
EXPECTED_PRODUCT = np.array([[11.0, 10.0, 9.0, 8.0], [23.0, 24.0, 25.0, 26.0], [7.0, 11.0, 15.0, 19.0], [6.0, 7.0, 8.0, 9.0]])
# :283-286 — result blocks of BlockMatrix x BlockMatrix
EXPECTED_PRODUCT_BLOCKS = {(0, 0): [[11.0, 10.0], [23.0, 24.0]], (0, 1): [[9.0, 8.0], [25.0, 26.0]],
                           (1, 0): [[7.0, 11.0], [6.0, 7.0]], (1, 1): [[15.0, 19.0], [8.0, 9.0]]}
# :306-307 — DenseVecMatrix.transpose() -> 1x2 grid of 4x2 blocks
EXPECTED_T_DVM_BLOCKS = {(0, 0): [[0.0, 2.0], [1.0, 3.0], [2.0, 4.0], [3.0, 5.0]],
                         (0, 1): [[3.0, 1.0], [2.0, 1.0], [1.0, 1.0], [0.0, 1.0]]}
# :312-315 — BlockMatrix.transpose()
EXPECTED_T_BLK_BLOCKS = {(0, 0): [[0.0, 2.0], [1.0, 3.0]], (0, 1): [[3.0, 1.0], [2.0, 1.0]],
                         (1, 0): [[2.0, 4.0], [3.0, 5.0]], (1, 1): [[1.0, 1.0], [0.0, 1.0]]}
# :166-188
ELE_ADD1 = EXPECTED_DENSE + 1.0
ADD_SELF = np.array([[0.0, 2.0, 4.0, 6.0], [4.0, 6.0, 8.0, 10.0], [6.0, 4.0, 2.0, 0.0], [2.0, 2.0, 2.0, 2.0]])
ELE_SUB1 = np.array([[-1.0, 0.0, 1.0, 2.0], [1.0, 2.0, 3.0, 4.0], [2.0, 1.0, 0.0, -1.0], [0.0, 0.0, 0.0, 0.0]])
DIVIDE2 = np.array([[0.0, 0.5, 1.0, 1.5], [1.0, 1.5, 2.0, 2.5], [1.5, 1.0, 0.5, 0.0], [0.5, 0.5, 0.5, 0.5]])
# :329-333
DOT_PRODUCT = np.array([[0.0, 1.0, 4.0, 9.0], [4.0, 9.0, 16.0, 25.0], [9.0, 4.0, 1.0, 0.0], [1.0, 1.0, 1.0, 1.0]])
SUM = 30.0   # :322-323

# known answers for data/a.100.100 x data/b.100.100 (NumPy fp64 during the survey; not a
# reference-suite vector, tolerance 1e-12 relative)
CFG1 = {"sumA": -8.317421000000003, "sumB": 47.658801000000025, "sumAplusB": 39.34138,
        "C00": 1.3037986919580005, "C01": -1.868456453296, "C9999": 2.7134692855449996,
        "sumC": 449.0644284029081, "traceC": -54.83429525636101, "frobC": 333.20829594580283}
SHA256 = {"a.100.100": "4cf2e6c125eb237810d5729c8154ed5ff97b08737363c581302bfa1bb7c76eb9",
          "b.100.100": "447f2e645563e0e08ca61bc748b0633da9f4865b8a11f83e4b0758ce8cb0e6b1"}

# DistributedMatrixSuite.scala:121-143 "disVec to disVec": three pieces re-split into four
DISVEC_PIECES = [(0, [0.0, 1.0, 2.0, 3.0]), (1, [4.0, 5.0, 6.0, 7.0]), (2, [8.0, 9.0, 10.0, 11.0])]
DISVEC_SPLIT_STATUS = [[(0, (0, 2), (0, 2)), (1, (3, 3), (0, 0))],
                       [(1, (0, 1), (1, 2)), (2, (2, 3), (0, 1))],
                       [(2, (0, 0), (2, 2)), (3, (1, 3), (0, 2))]]
# :390-409 "BLAS1 distributed vector multiplication"
BLAS1_PIECES = [(0, [1.0, 2.0]), (1, [3.0, 4.0])]
BLAS1_OUTER = np.array([[1.0, 2.0, 3.0, 4.0], [2.0, 4.0, 6.0, 8.0], [3.0, 6.0, 9.0, 12.0], [4.0, 8.0, 12.0, 16.0]])
BLAS1_INNER = 30.0
# not in the reference suite (it has no matrix x vector test): A . [1,2,3,4] for the 4x4 matrix above, exact
MATVEC_X = [1.0, 2.0, 3.0, 4.0]
MATVEC_Y = EXPECTED_DENSE @ np.array(MATVEC_X)
