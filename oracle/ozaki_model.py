"""CPU model of the opt-in int8-split fp64 GEMM (marlin_b200/csrc/gemm_ozaki.cu) — TEST INFRASTRUCTURE ONLY.

The reference has no such mode (its multiply is IEEE fp64 dgemm); this restates OUR algorithm in exact integer
arithmetic so that its error bound can be checked on the CPU and the kernel can be compared against it bit for bit:
row / column scaling by powers of two, rounding to P fractional bits, balanced digits of `bits` bits, exact integer
products per digit pair, pairs grouped by d = t + u, groups with d > s + 1 dropped, groups folded into C in fp64 from the
least significant kept group (d = s + 1) up to d = 2.
"""
from __future__ import annotations

import numpy as np


def frac_bits(s: int, bits: int) -> int:
    return bits * s - (2 if bits == 8 else 1)


def scale_exponents(absmax: np.ndarray) -> np.ndarray:
    """smallest e with r < 2^e (0 for an all-zero row / column)"""
    e = np.zeros(absmax.shape, dtype=np.int64)
    nz = absmax > 0
    e[nz] = np.floor(np.log2(absmax[nz])).astype(np.int64) + 1
    # log2 of values just below a power of two may round up: fix with exact comparisons
    e[nz] -= (np.ldexp(1.0, (e[nz] - 1).astype(np.int32)) > absmax[nz]).astype(np.int64)
    e[nz] += (np.ldexp(1.0, e[nz].astype(np.int32)) <= absmax[nz]).astype(np.int64)
    return e


def split(x: np.ndarray, e: np.ndarray, s: int, bits: int, axis: int):
    """digits[t] (t = 0 most significant) of X = rint(x * 2^(P - e)) with e broadcast along `axis`."""
    P = frac_bits(s, bits)
    shape = [1, 1]
    shape[axis] = -1
    X = np.rint(np.ldexp(x, (P - e).reshape(shape).astype(np.int32))).astype(np.int64)
    half, mask = 1 << (bits - 1), (1 << bits) - 1
    digs = [None] * s
    R = X.copy()
    for t in range(s - 1, 0, -1):
        d = ((R + half) & mask) - half
        digs[t] = d
        R = (R - d) >> bits
    digs[0] = R
    return X, digs


def gemm(A: np.ndarray, B: np.ndarray, s: int, bits: int, C0: np.ndarray | None = None):
    """The kernel's result, operation for operation.  Returns (C, info)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    P = frac_bits(s, bits)
    eA = scale_exponents(np.abs(A).max(axis=1))
    eB = scale_exponents(np.abs(B).max(axis=0))
    XA, dA = split(A, eA, s, bits, 0)
    XB, dB = split(B, eB, s, bits, 1)
    C = None if C0 is None else np.array(C0, dtype=np.float64)
    for d in range(s + 1, 1, -1):                       # d = t + u with 1-based digit indices
        acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for t in range(1, s + 1):
            u = d - t
            if 1 <= u <= s:
                acc += dA[t - 1] @ dB[u - 1]
        assert np.abs(acc).max(initial=0) < 2 ** 31      # the TMEM accumulator is int32
        ex = (eA[:, None] + eB[None, :] - 2 * P + bits * (2 * s - d)).astype(np.int32)
        term = np.ldexp(acc.astype(np.float64), ex)      # exact: |acc| < 2^31
        C = term if C is None else C + term
    return C, {"P": P, "eA": eA, "eB": eB, "XA": XA, "XB": XB, "dA": dA, "dB": dB}
