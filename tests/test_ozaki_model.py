"""The int8-split fp64 GEMM as an algorithm (oracle/ozaki_model.py, exact integer arithmetic on the CPU): digit ranges,
exact reconstruction, and the error of the result against the exactly rounded product, for the digit configurations the
library offers.  The kernel itself is compared with this model bit for bit in tests/test_gpu_ozaki.py."""
import numpy as np
import pytest

from oracle import ozaki_model as om


def _exact(A, B):
    return (A.astype(np.longdouble) @ B.astype(np.longdouble))


@pytest.mark.parametrize("s,bits", [(5, 7), (6, 7), (7, 7), (4, 8), (5, 8), (6, 8)])
def test_digits_fit_int8_and_reconstruct(s, bits):
    rng = np.random.default_rng(s * 10 + bits)
    A = (rng.random((37, 90)) - 0.5) * np.exp2(rng.integers(-20, 20, size=(37, 1)))     # rows on very different scales
    A[3, :] = 0.0
    A[5, 7] = np.abs(A[5]).max() * 0.99999999                                          # just below the row's power of two
    e = om.scale_exponents(np.abs(A).max(axis=1))
    assert e[3] == 0 and all(np.abs(A[i]).max() < np.exp2(float(e[i])) for i in range(37))
    assert all(np.abs(A[i]).max() >= np.exp2(float(e[i] - 1)) for i in range(37) if i != 3)
    X, digs = om.split(A, e, s, bits, 0)
    lim = 1 << (bits - 1)
    for t, d in enumerate(digs):
        assert d.min() >= -128 and d.max() <= 127                                       # int8
        assert np.abs(d).max() <= (lim + 1 if t == 0 else lim)
    back = sum(d * (1 << (bits * (s - 1 - t))) for t, d in enumerate(digs))
    assert np.array_equal(back, X)
    P = om.frac_bits(s, bits)
    assert np.abs(np.ldexp(X.astype(np.float64), (e - P).reshape(-1, 1).astype(np.int32)) - A).max() <= np.exp2(float(e.max() - P - 1))


@pytest.mark.parametrize("s,bits", [(5, 7), (6, 7), (7, 7), (4, 8), (5, 8), (6, 8)])
def test_error_bound_relative_to_row_and_column_maxima(s, bits):
    """|C - AB|_ij <= 4 * 2^-P * K * rowmax_i(A) * colmax_j(B), P = bits*s - 1 (7-bit) or 8s - 2 (8-bit): each operand is
    rounded to P bits below its row / column power of two (<= 2 * max), the dropped digit pairs weigh less than that
    rounding.  Observed: about 2^-P / 4.  This — not a bound relative to (|A||B|)_ij — is what the mode guarantees, which
    is why it is opt-in."""
    bound = 4.0 * 2.0 ** -om.frac_bits(s, bits)
    rng = np.random.default_rng(s + bits)
    K = 512
    A = rng.standard_normal((40, K)) * np.exp2(rng.integers(-8, 8, size=(40, 1)))
    B = rng.standard_normal((K, 48)) * np.exp2(rng.integers(-8, 8, size=(1, 48)))
    C, _ = om.gemm(A, B, s, bits)
    scale = K * np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=0)[None, :]
    err = np.abs(C.astype(np.longdouble) - _exact(A, B)).astype(np.float64) / scale
    assert err.max() <= bound, err.max()


def test_uniform_inputs_meet_the_multiply_tolerance():
    """U[0,1) operands (the benchmark's inputs): 8-bit x 5 planes stays far inside the 1e-10 tolerance of the north_star
    measured against (|A||B|)_ij; accumulate adds onto an existing C in the same group order."""
    rng = np.random.default_rng(1)
    A, B = rng.random((64, 1024)), rng.random((1024, 80))
    C, info = om.gemm(A, B, 5, 8)
    ref = _exact(A, B)
    assert (np.abs(C.astype(np.longdouble) - ref) / ref).max() <= 1e-11
    C2, _ = om.gemm(A, B, 5, 8, C0=C)
    assert (np.abs(C2 - 2 * C) / C).max() <= 1e-11
    assert info["P"] == 38 and len(info["dA"]) == 5
