"""bf16 tcgen05 GEMM (BASELINE config 5 path) against an fp64 oracle on the SAME bf16-rounded inputs.

Stated tolerances: (i) vs. the fp64 oracle fed the bf16-rounded inputs, norm-wise <= 1e-4 for fp32
accumulation (here K <= 4096 gives ~1e-6); (ii) a bf16 result block is within 1 bf16 ulp (2^-8 relative) of (i);
transpose of a bf16 block is bit-exact."""
import ctypes as C

import numpy as np
import pytest

from marlin_b200 import _native as nat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import marlin_b200 as mb
    mb.Runtime.get()
    return mb


def rounded(M, arr):
    """bf16-round on the device (RNE from fp64), return (device bf16 block, the rounded values as fp64 ndarray)."""
    blk = M.SubMatrix(arr).copy(nat.MB_BF16)
    return blk, blk.toBreeze()


@pytest.mark.parametrize("shape", [(128, 256, 64), (128, 256, 128), (256, 512, 256), (64, 40, 72), (200, 136, 1000), (384, 768, 4096)])
@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_bf16_gemm_vs_fp64_oracle(M, oracle, shape, ta, tb):
    m, n, k = shape
    rng = np.random.default_rng(m + n + k + 2 * ta + tb)
    A, B = rng.random((m, k)), rng.random((k, n))
    if ta:
        at, Ar = rounded(M, A.T)
        a_blk, Ar = at.t, Ar.T
    else:
        a_blk, Ar = rounded(M, A)
    if tb:
        bt, Br = rounded(M, B.T)
        b_blk, Br = bt.t, Br.T
    else:
        b_blk, Br = rounded(M, B)
    ref = oracle.block_multiply(np.asfortranarray(Ar), np.asfortranarray(Br), "blas")
    c32 = a_blk.multiply(b_blk)                                   # fp32 result block
    assert c32.dtype == nat.MB_F32
    got = c32.toBreeze()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-4
    assert (np.abs(got - ref) / ref).max() <= 1e-4               # U[0,1) inputs: no cancellation
    c16 = a_blk.multiply(b_blk, out_dtype=nat.MB_BF16)
    assert (np.abs(c16.toBreeze() - ref) / ref).max() <= 2.0 ** -8 + 1e-4
    # accumulate: C += A*B (the k-way sum of BlockMatrix.scala:177 kept in the epilogue)
    a_blk.multiply(b_blk, out=c32, accumulate=True)
    assert (np.abs(c32.toBreeze() - 2 * ref) / ref).max() <= 2e-4


def test_bf16_transpose_bit_exact_and_add_within_ulp(M):
    rng = np.random.default_rng(3)
    A = rng.random((96, 160))
    blk, Ar = rounded(M, A)
    t = blk.transpose()
    assert t.dtype == nat.MB_BF16 and np.array_equal(t.toBreeze(), Ar.T)
    # add is defined on fp64 blocks; a bf16 block is widened exactly, added in fp64, and re-rounded: <= 1 bf16 ulp
    s = blk.copy(nat.MB_F64).add(blk.copy(nat.MB_F64)).copy(nat.MB_BF16).toBreeze()
    assert (np.abs(s - 2 * Ar) / (2 * Ar)).max() <= 2.0 ** -8


def test_bf16_block_matrix_multiply_4x4_grid_sampled(M, oracle):
    """Config-5 shape in miniature: 4x4 block grid, (4,4,4) split, bf16 tiles, fp32 C tiles; checked per C tile."""
    n = 1024
    A = M.MTUtils.randomBlockMatrix(None, n, n, 4, 4, seed=5, dtype=nat.MB_BF16)
    B = M.MTUtils.randomBlockMatrix(None, n, n, 4, 4, seed=6, dtype=nat.MB_BF16)
    Cm = A.multiply(B)
    assert (Cm.numBlksByRow(), Cm.numBlksByCol()) == (4, 4)
    Ah, Bh = A.toBreeze(), B.toBreeze()
    ref = Ah @ Bh
    got = Cm.toBreeze()
    assert (np.abs(got - ref) / ref).max() <= 1e-4
