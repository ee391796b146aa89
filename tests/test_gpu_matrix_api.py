"""GPU tests of the host mirror (marlin_b200.BlockMatrix / DenseVecMatrix / MTUtils): the reference's own
DistributedMatrixSuite cases (exact), BASELINE config[0] from the data files, and randomized parity with the
oracle executing the same algorithm."""
import numpy as np
import pytest

from tests import marlin_cases as mc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import marlin_b200 as mb
    mb.Runtime.get()
    return mb


def dvm(M):
    return M.DenseVecMatrix([(i, np.array(v)) for i, v in mc.DATA_ROWS])


def blk(M):
    return M.BlockMatrix([(M.BlockID(*k), M.SubMatrix(np.array(v))) for k, v in mc.BLKS])


def blocks_of(bm):
    return {(b.row, b.column): s.toBreeze() for b, s in bm.blocks}


def assert_blocks(bm, expected):
    got = blocks_of(bm)
    assert set(got) == set(expected)
    for k, v in expected.items():
        assert np.array_equal(got[k], np.array(v)), (k, got[k])


def test_matrix_size(M):                         # DMS.scala:42-51
    mat, ma = dvm(M), blk(M)
    assert (mat.numRows(), mat.numCols()) == (4, 4)
    assert (ma.numRows(), ma.numCols(), ma.numBlksByRow(), ma.numBlksByCol()) == (4, 4, 2, 2)


def test_empty_rows(M):                          # :53-71
    for obj in (M.DenseVecMatrix([]), M.BlockMatrix([])):
        with pytest.raises(RuntimeError):
            obj.numRows()
        with pytest.raises(RuntimeError):
            obj.numCols()


def test_to_breeze(M):                           # :73-84
    assert np.array_equal(dvm(M).toBreeze(), mc.EXPECTED_DENSE)
    assert np.array_equal(blk(M).toBreeze(), mc.EXPECTED_DENSE)


def test_to_block_matrix(M):                     # :86-105
    mat = dvm(M)
    b22 = mat.toBlockMatrix(2, 2)
    assert (b22.numRows(), b22.numCols()) == (4, 4)
    assert_blocks(b22, dict(mc.BLKS))
    assert np.array_equal(mat.toBlockMatrix(1, 4).toBreeze(), mc.EXPECTED_DENSE)
    assert np.array_equal(b22.toBreeze(), mc.EXPECTED_DENSE)


def test_to_dense_vec_matrix(M):                 # :108-119
    d = blk(M).toDenseVecMatrix()
    assert (d.numRows(), d.numCols()) == (4, 4)
    assert np.array_equal(d.toBreeze(), mc.EXPECTED_DENSE)
    arr = d.data.toBreeze()
    for pos, i in enumerate(d.ids):
        assert np.array_equal(arr[pos], mc.EXPECTED_DENSE[int(i)])


def test_elementwise(M):                         # :164-205
    mat, ma = dvm(M), blk(M)
    for x in (mat, ma):
        assert np.array_equal(x.add(1).toBreeze(), mc.ELE_ADD1)
        assert np.array_equal(x.add(x).toBreeze(), mc.ADD_SELF)
        assert np.array_equal(x.subtract(1).toBreeze(), mc.ELE_SUB1)
        assert np.array_equal(x.subtract(x).toBreeze(), np.zeros((4, 4)))
        assert np.array_equal(x.multiply(2).toBreeze(), mc.ADD_SELF)
        assert np.array_equal(x.divide(2).toBreeze(), mc.DIVIDE2)
    assert np.array_equal(ma.add(mat).toBreeze(), mc.ADD_SELF)
    assert np.array_equal(ma.subtract(mat).toBreeze(), np.zeros((4, 4)))


def test_elementwise_join_drops_unmatched_rows(M):
    """DenseVecMatrix.add/subtract/dotProduct are `rows.join(that.rows)` (DenseVecMatrix.scala:777-780): an INNER join,
    so a row id that exists on one side only does not appear in the result (it stays zero in toBreeze())."""
    rng = np.random.default_rng(4)
    A, B = rng.random((6, 5)), rng.random((6, 5))
    a = M.DenseVecMatrix([(i, A[i]) for i in (0, 1, 2, 4, 5)], 6, 5)          # row 3 missing
    b = M.DenseVecMatrix([(i, B[i]) for i in (5, 3, 2, 1, 0)], 6, 5)          # row 4 missing, other order
    want = np.zeros((6, 5))
    for i in (0, 1, 2, 5):
        want[i] = A[i] + B[i]
    got = a.add(b)
    assert sorted(got.ids.tolist()) == [0, 1, 2, 5]
    assert np.array_equal(got.toBreeze(), want)
    for i in (0, 1, 2, 5):
        want[i] = A[i] * B[i]
    assert np.array_equal(a.dotProduct(b).toBreeze(), want)


def test_multiply_selects_broadcast(M):          # :225-234
    mat = dvm(M)
    res = mat.multiply(mat, 2)
    assert isinstance(res, M.DenseVecMatrix)
    assert np.array_equal(res.toBreeze(), mc.EXPECTED_PRODUCT)


@pytest.mark.parametrize("split", [(2, 2, 1), (2, 1, 2), (2, 2, 2)])
def test_new_matrix_multiplication(M, split):    # :236-249
    mat = dvm(M)
    res = mat.multiply(mat, split)
    assert isinstance(res, M.BlockMatrix)
    assert np.array_equal(res.toBreeze(), mc.EXPECTED_PRODUCT)


def test_multiply_local_matrix(M):               # :251-267
    assert np.array_equal(dvm(M).multiply(mc.EXPECTED_DENSE).toBreeze(), mc.EXPECTED_PRODUCT)


def test_multiply_block_matrix(M):               # :269-287
    mat, ma = dvm(M), blk(M)
    assert np.array_equal(mat.multiply(ma, 2).toBreeze(), mc.EXPECTED_PRODUCT)
    assert_blocks(ma.multiply(ma), mc.EXPECTED_PRODUCT_BLOCKS)


def test_block_times_densevec_broadcast(M):      # :289-299
    assert np.array_equal(blk(M).multiply(dvm(M), 2).toBreeze(), mc.EXPECTED_PRODUCT)


def test_transpose(M):                           # :302-316
    assert_blocks(dvm(M).transpose(), mc.EXPECTED_T_DVM_BLOCKS)
    t = blk(M).transpose()
    assert_blocks(t, mc.EXPECTED_T_BLK_BLOCKS)
    assert (t.numRows(), t.numCols(), t.numBlksByRow(), t.numBlksByCol()) == (4, 4, 2, 2)


def test_sum_and_dot_product(M):                 # :319-338
    mat, ma = dvm(M), blk(M)
    assert mat.sum() == mc.SUM and ma.sum() == mc.SUM
    for a, b in ((mat, mat), (mat, ma), (ma, mat), (ma, ma)):
        assert np.array_equal(a.dotProduct(b).toBreeze(), mc.DOT_PRODUCT)


def test_block_to_block_and_regrid_multiply(M):  # :411-432
    mat = dvm(M)
    b1 = mat.toBlockMatrix(2, 2)
    assert np.array_equal(b1.toBlockMatrix(1, 4).toBreeze(), mc.EXPECTED_DENSE)
    assert np.array_equal(b1.toBlockMatrix(4, 1).toBreeze(), mc.EXPECTED_DENSE)
    m = b1.toBlockMatrix(2, 1)
    assert np.array_equal(m.multiply(mat.toBlockMatrix(1, 4)).toBreeze(), mc.EXPECTED_PRODUCT)


def test_block_multiply_broadcast(M):            # :434-448
    assert np.array_equal(blk(M).multiply(mc.EXPECTED_DENSE).toBreeze(), mc.EXPECTED_PRODUCT)


def test_dimension_mismatch_is_illegal_argument(M):
    a = M.DenseVecMatrix([(0, [1.0, 2.0, 3.0]), (1, [4.0, 5.0, 6.0])])
    with pytest.raises(ValueError, match="Dimension mismatch during matrix-matrix multiplication: 3 vs 2"):
        a.multiply(a, 2)
    with pytest.raises(ValueError):
        a.toBlockMatrix(1, 1).multiply(a.toBlockMatrix(1, 1))
    with pytest.raises(ValueError, match="currently not supported"):
        big = M.MTUtils.randomBlockMatrix(None, 12, 12, 3, 3, seed=1)
        big.multiply(M.MTUtils.randomBlockMatrix(None, 12, 12, 2, 2, seed=2))


def test_config1_data_files(M, golden_dir):
    """BASELINE config[0]: loadMatrixFile -> multiply(b, 2) (broadcast branch) and multiply(b, (2,2,2))."""
    a = M.MTUtils.loadMatrixFile(None, str(golden_dir / "a.100.100"))
    b = M.MTUtils.loadMatrixFile(None, str(golden_dir / "b.100.100"))
    assert (a.numRows(), a.numCols()) == (100, 100)
    for res in (a.multiply(b, 2), a.multiply(b, (2, 2, 2))):
        c = res.toBreeze()
        assert c[0, 0] == pytest.approx(mc.CFG1["C00"], rel=1e-12)
        assert c[0, 1] == pytest.approx(mc.CFG1["C01"], rel=1e-12)
        assert c[99, 99] == pytest.approx(mc.CFG1["C9999"], rel=1e-12)
        assert c.sum() == pytest.approx(mc.CFG1["sumC"], rel=1e-11)
        assert np.trace(c) == pytest.approx(mc.CFG1["traceC"], rel=1e-11)
        assert np.linalg.norm(c) == pytest.approx(mc.CFG1["frobC"], rel=1e-12)
    assert a.add(b).sum() == pytest.approx(mc.CFG1["sumAplusB"], rel=1e-12)
    with pytest.raises(ValueError):
        M.MTUtils.loadMatrixFile(None, "relative/path")              # MTUtils.scala:287-290


@pytest.mark.parametrize("dims", [(37, 53, 29, (3, 2, 2)), (64, 64, 64, (2, 2, 2)), (101, 17, 90, (4, 1, 3)), (50, 200, 10, (1, 5, 1))])
def test_random_parity_with_oracle_algorithm(M, oracle, dims):
    """Same inputs through the oracle's restatement of DenseVecMatrix.multiply(other, (m,k,n)) (ragged blocks)."""
    Mr, K, N, split = dims
    rng = np.random.default_rng(Mr + K + N)
    A, B = rng.random((Mr, K)) * 2 - 1, rng.random((K, N)) * 2 - 1
    perm = rng.permutation(Mr)
    ga = M.DenseVecMatrix([(int(i), A[i]) for i in perm])
    gb = M.DenseVecMatrix(list(enumerate(B)))
    oa = oracle.DenseVecMatrix([(int(i), A[i]) for i in perm])
    ob = oracle.DenseVecMatrix(list(enumerate(B)))
    ref = oa.multiply_split(ob, split, gemm="f2j")
    got = ga.multiply(gb, split)
    assert (got.numBlksByRow(), got.numBlksByCol()) == (ref.num_blks_by_row(), ref.num_blks_by_col())
    gblocks, rblocks = blocks_of(got), dict(ref.blocks)
    assert set(gblocks) == set(rblocks)
    denom = np.abs(A) @ np.abs(B)
    for key in rblocks:
        assert gblocks[key].shape == rblocks[key].shape
    err = np.abs(got.toBreeze() - ref.to_breeze()) / denom
    assert err.max() <= 1e-10
    # broadcast path (DenseVecMatrix.scala:1660-1680) and the transposes/adds around it
    ref2 = oa.multiply_local(B, gemm="f2j").to_breeze()
    assert (np.abs(ga.multiply(B).toBreeze() - ref2) / denom).max() <= 1e-10
    assert np.array_equal(ga.transpose().toBreeze(), A.T)
    assert np.array_equal(ga.toBlockMatrix(*split[:2]).transpose().toBreeze(), A.T)
    assert np.array_equal(ga.add(ga.toBlockMatrix(2, 2)).toBreeze(), A + A)


def test_ratio_resplit_branch(M, oracle):
    """BlockMatrix.scala:187-216 — the coarser operand is re-sliced by rows as VIEWS (no copy) and fed to the GEMM."""
    rng = np.random.default_rng(21)
    A, B = rng.random((16, 24)), rng.random((24, 12))
    ga = M.DenseVecMatrix(list(enumerate(A))).toBlockMatrix(2, 4)
    gb = M.DenseVecMatrix(list(enumerate(B))).toBlockMatrix(2, 2)
    oa = oracle.DenseVecMatrix(list(enumerate(A))).to_block_matrix(2, 4)
    ob = oracle.DenseVecMatrix(list(enumerate(B))).to_block_matrix(2, 2)
    assert np.abs(ga.multiply(gb).toBreeze() - oa.multiply(ob).to_breeze()).max() <= 1e-12


def test_random_generators_match_reference_streams(M, oracle):
    """MTUtils.randomDenVecMatrix / randomBlockMatrix fill HBM with the same XORShift streams, partitioning and
    fill order as the restated RandomDenVecRDD / RandomBlockRDD (bit-exact for a given seed)."""
    g = M.MTUtils.randomDenVecMatrix(None, 37, 11, numPartitions=4, seed=2024)
    o = oracle.random_den_vec_matrix(37, 11, 4, seed=2024)
    assert np.array_equal(g.toBreeze(), o.to_breeze())
    gb = M.MTUtils.randomBlockMatrix(None, 23, 17, 3, 2, seed=7)
    ob = oracle.random_block_matrix(23, 17, 3, 2, seed=7)
    assert (gb.numBlksByRow(), gb.numBlksByCol()) == (ob.num_blks_by_row(), ob.num_blks_by_col())
    for key, arr in ob.blocks:
        assert np.array_equal(blocks_of(gb)[key], arr)
    assert M.MTUtils.splitMethod(16384, 16384, 16384, 8) == (2, 2, 2)


def test_save_and_load_block_format(M, tmp_path):
    ma = blk(M)
    ma.saveToFileSystem(str(tmp_path / "blk"), "blockmatrix")
    text = (tmp_path / "blk" / "part-00000").read_text().splitlines()
    assert text[0] == "0-0-2-2:0.0,2.0,1.0,3.0"                         # BlockMatrix.scala:553-554, column-major
    again = M.MTUtils.loadBlockMatrixFile(None, "file://" + str(tmp_path / "blk"))
    assert np.array_equal(again.toBreeze(), mc.EXPECTED_DENSE)
    dvm(M).saveToFileSystem(str(tmp_path / "rows"))
    assert np.array_equal(M.MTUtils.loadMatrixFile(None, str(tmp_path / "rows")).toBreeze(), mc.EXPECTED_DENSE)


def test_multiply_by_local_matrix(M):
    """BlockMatrix.multiplyBy(B: BDM) (BlockMatrix.scala:309-319), the small-A arm of the chooser (:114-115,
    DenseVecMatrix.scala:223-224): one block row, B * blk per block."""
    rng = np.random.default_rng(31)
    Bl = rng.integers(-3, 4, size=(5, 6)).astype(float)
    X = rng.integers(-3, 4, size=(6, 8)).astype(float)
    bm = M.DenseVecMatrix(list(enumerate(X))).toBlockMatrix(1, 2)
    got = bm.multiplyBy(Bl)
    assert np.array_equal(got.toBreeze(), Bl @ X)
    # chooser: this (small DenseVecMatrix, 5x6) times a BlockMatrix too big to broadcast -> that.multiplyBy(this.toBreeze())
    Xw = rng.integers(-3, 4, size=(6, 30000)).astype(float)         # 180000 elements > 1 MB / 8 = 131072
    wide = M.DenseVecMatrix(list(enumerate(Xw))).toBlockMatrix(1, 2)
    small = M.DenseVecMatrix(list(enumerate(Bl)))
    res = small.multiply(wide, 2, 1)                   # broadcastThreshold = 1 MB: B too big to broadcast, A small
    assert isinstance(res, M.BlockMatrix)
    assert np.array_equal(res.toBreeze(), Bl @ Xw)


def test_example_drivers(M, capsys):
    """The reference's multiply drivers (examples/MatrixMultiply, BLAS3, RMMcompare) with their own command lines."""
    from marlin_b200.examples import blas3 as BLAS3, matrix_multiply, rmm_compare as RMMcompare
    matrix_multiply.main(["300", "200", "100", "8"])
    BLAS3.main(["256", "128", "64", "2", "2"])
    BLAS3.main(["256", "128", "64", "3", "2", "2", "2"])
    RMMcompare.main(["256", "256", "256", "2", "2", "2", "2"])
    out = capsys.readouterr().out
    assert "Result RDD counts" in out and "mode 3 used time" in out and "RMMv2 in mode 2 used time" in out


# ------------------------------------------------------------------ LU / Cholesky / inverse
def test_inverse_suite_case(M):                  # DistributedMatrixSuite.scala:340-352
    rows = [(0, np.array([0.0, 0.0, 1.0])), (1, np.array([0.0, 1.0, 0.0])), (2, np.array([1.0, 0.0, 0.0]))]
    inv = M.DenseVecMatrix(rows).inverse()
    assert isinstance(inv, M.BlockMatrix)
    assert np.array_equal(inv.toBreeze(), np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]]))


@pytest.mark.parametrize("mode,base", [("breeze", None), ("dist", 7), ("dist", 40), ("auto", None)])
def test_lu_cholesky_inverse_against_the_restated_reference(M, oracle, mode, base):
    """The three block algorithms against the oracle's restatement of matrix/DenseVecMatrix.scala:283-764 (LAPACK through
    scipy for the Breeze calls), both in "breeze" mode and in "dist" mode with small base sizes (ragged last block), to
    1e-10 relative; plus the defining properties P A = L U, L L^T = A, A A^-1 = I."""
    n = 93
    rng = np.random.default_rng(7)
    A = rng.random((n, n)) - 0.5 + 3.0 * np.eye(n)
    rows = list(enumerate(A))
    g, o = M.DenseVecMatrix(rows), oracle.DenseVecMatrix(rows)
    kw = {} if base is None else {"baseSize": base}
    okw = {} if base is None else {"base": base}
    lu, perm = g.luDecompose(mode, **kw)
    olu, operm = o.lu_decompose(mode, keep_unfactored_diagonal=False, **okw)
    assert list(perm) == list(operm)
    got = lu.toBreeze()
    assert np.abs(got - olu.to_breeze()).max() <= 1e-10 * np.abs(got).max()
    L, U = np.tril(got, -1) + np.eye(n), np.triu(got)
    assert np.abs(A[perm] - L @ U).max() <= 1e-12
    if mode == "dist":                           # the reference as written keeps the unfactored diagonal blocks (:355)
        q, _ = g.luDecompose(mode, keepUnfactoredDiagonal=True, **kw)
        oq, _ = o.lu_decompose(mode, keep_unfactored_diagonal=True, **okw)
        assert np.abs(q.toBreeze() - oq.to_breeze()).max() <= 1e-10 * np.abs(got).max()
    inv = g.inverse(mode, **kw).toBreeze()
    assert np.abs(inv - o.inverse(mode, **okw).to_breeze()).max() <= 1e-10 * np.abs(inv).max()
    assert np.abs(inv @ A - np.eye(n)).max() <= 1e-12
    S = A @ A.T
    srows = list(enumerate(S))
    Lc = M.DenseVecMatrix(srows).choleskyDecompose(mode, **kw).toBreeze()
    assert np.abs(Lc - oracle.DenseVecMatrix(srows).cholesky_decompose(mode, **okw).to_breeze()).max() <= 1e-10 * np.abs(Lc).max()
    assert np.array_equal(np.triu(Lc, 1), np.zeros((n, n))) and np.abs(Lc @ Lc.T - S).max() <= 1e-11 * np.abs(S).max()


def test_factorization_argument_errors(M):
    d = M.DenseVecMatrix([(0, np.array([1.0, 2.0, 3.0])), (1, np.array([4.0, 5.0, 6.0]))])
    for f in (d.luDecompose, d.choleskyDecompose, d.inverse):
        with pytest.raises(ValueError):
            f()
    sq = M.DenseVecMatrix([(0, np.array([2.0, 0.0])), (1, np.array([0.0, 2.0]))])
    with pytest.raises(ValueError):
        sq.inverse("spark")
    assert np.array_equal(M.BlockMatrix([(M.BlockID(0, 0), M.SubMatrix(np.array([[2.0, 0.0], [0.0, 4.0]])))]).inverse().toBreeze(),
                          np.array([[0.5, 0.0], [0.0, 0.25]]))
