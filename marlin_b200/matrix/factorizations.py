"""Block LU / Cholesky / inverse of DenseVecMatrix on the GPU (matrix/DenseVecMatrix.scala:283-764).

"breeze" mode (the reference collects the matrix and calls Breeze -> LAPACK) is one call into the recursive device
factorizations (`mb_block_lu / cholesky / inverse`, csrc/factor.cu).  "dist" mode is the reference's right-looking
block algorithm over a ceil(n / basesize)^2 grid; every block operation is a kernel of libmarlin_b200.so:

  LU        P_i A_ii = L_ii U_ii (pivoting inside the diagonal block only),  U_ic = L_ii^-1 (P_i A_ic),  L_ri = A_ri U_ii^-1,
            A_rc -= L_ri U_ic;  finally block row r of L is permuted by P_r                                   (:310-466)
  Cholesky  L_ii = chol(A_ii),  L_ri = A_ri L_ii^-T,  A_rc -= L_ri L_ci^T  (lower blocks only)                 (:497-556)
  inverse   elimination with inv(A_ii), then back substitution from the last block                            (:589-760)

The reference's permutation-matrix products (`(l \\ permutation) * block`, `permutation * blk`) are kept as products
with an uploaded 0/1 matrix (exact), its `block * brzInv(u)` / `blk * inv(l.t)` are triangular solves on transposed
views.  With several ranks every rank runs the factorization on the gathered matrix (replicas: the algorithm is a
chain of dependent panel steps) and keeps the result blocks that MatrixElemOpPartitioner order places on it.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from .. import _native as nat
from .block import BlockID
from .sub_matrix import SubMatrix

Blocks = Dict[Tuple[int, int], SubMatrix]


def mode_is_dist(mode: str, n: int) -> bool:
    """:284-299 — "auto" switches to the distributed algorithm above 6000 rows."""
    if mode == "auto":
        return n > 6000
    if mode == "breeze":
        return False
    if mode == "dist":
        return True
    raise nat.MarlinArgumentError(nat.MB_ERR_INVALID_ARG, f"Do not support mode {mode}.")


def _perm_matrix(perm, device) -> SubMatrix:
    """permutation(j, p(j)) = 1 (:362-366): (P x)[j] = x[p[j]]."""
    n = len(perm)
    pm = np.zeros((n, n))
    pm[np.arange(n), np.asarray(perm)] = 1.0
    return SubMatrix(pm, device=device)


def _right_solve_upper(blk: SubMatrix, u: SubMatrix) -> SubMatrix:
    """X U = blk  <=>  U^T X^T = blk^T (U^T is lower triangular): `block * brzInv(u)` without forming the inverse."""
    return u.t.solveTriangular(blk.t, lower=True, unit=False).transpose()


def lu_blocks(blocks: Blocks, nb: int, sub: int, n: int, keep_unfactored_diagonal: bool = False):
    dev = next(iter(blocks.values())).buf.device
    cur = dict(blocks)
    p_array = [0] * n
    done: Blocks = {}
    for i in range(nb):
        first = cur[(i, i)]
        packed, perm = first.lu()
        for j, pj in enumerate(perm):
            p_array[i * sub + j] = i * sub + int(pj)
        if i == nb - 1:
            done[(i, i)] = packed
            break
        second = {k: v for k, v in cur.items() if k[0] == i and k[1] > i}
        third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
        forth = {k: v for k, v in cur.items() if k[0] > i and k[1] > i}
        pm = _perm_matrix(perm, dev)
        done[(i, i)] = first if keep_unfactored_diagonal else packed
        u_row = {k: packed.solveTriangular(pm.multiply(v), lower=True, unit=True) for k, v in second.items()}
        l_col = {k: _right_solve_upper(v, packed) for k, v in third.items()}
        done.update(u_row)
        done.update(l_col)
        cur = {(r, c): forth[(r, c)].subtract(l_col[(r, i)].multiply(u_row[(i, c)])) for (r, c) in forth}
    out: Blocks = {}
    for (r, c), blk in done.items():
        if r > c:
            arr = p_array[sub * r: (n if r == nb - 1 else sub * r + sub)]
            blk = _perm_matrix([a - sub * r for a in arr], dev).multiply(blk)
        out[(r, c)] = blk
    return out, p_array


def cholesky_blocks(blocks: Blocks, nb: int) -> Blocks:
    cur = dict(blocks)
    done: Blocks = {}
    for i in range(nb):
        l = cur[(i, i)].cholesky()          # reads the lower triangle (the reference first mirrors the upper one into it)
        done[(i, i)] = l
        if i == nb - 1:
            break
        third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
        forth = {k: v for k, v in cur.items() if k[1] > i and k[0] >= k[1]}
        # L_ri = A_ri L^-T  <=>  L L_ri^T = A_ri^T
        l_col = {k: l.solveTriangular(v.t, lower=True, unit=False).transpose() for k, v in third.items()}
        done.update(l_col)
        cur = {(r, c): forth[(r, c)].subtract(l_col[(r, i)].multiply(l_col[(c, i)].t)) for (r, c) in forth}
    return done


def inverse_blocks(blocks: Blocks, nb: int) -> Blocks:
    cur = dict(blocks)
    sc = {}
    for i in range(nb):
        if i == nb - 1:
            cur = {k: v.inverse() for k, v in cur.items()}
            break
        second = {k: v for k, v in cur.items() if k[0] == i and k[1] > i}
        third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
        forth = {k: v for k, v in cur.items() if k[0] > i and k[1] > i}
        inv = cur[(i, i)].inverse()
        sc[(i, 0)] = inv
        sc[(i, 1)] = {k: inv.multiply(v).multiply(-1.0) for k, v in second.items()}
        sc[(i, 2)] = {k: v.multiply(inv).multiply(-1.0) for k, v in third.items()}
        cur = {(r, c): forth[(r, c)].subtract(third[(r, i)].multiply(inv).multiply(second[(i, c)])) for (r, c) in forth}
    for i in range(nb - 2, -1, -1):
        second_mat, third_mat = sc[(i, 1)], sc[(i, 2)]
        rng = range(i + 1, nb)

        def total(pairs):
            acc = None
            for a, b in pairs:
                acc = a.multiply(b) if acc is None else a.multiply(b, out=acc, accumulate=True)
            return acc

        mult_third = {(r, i): total((cur[(r, c)], third_mat[(c, i)]) for c in rng) for r in rng}
        mult_second = {(i, c): total((second_mat[(i, r)], cur[(r, c)]) for r in rng) for c in rng}
        first = total((second_mat[(i, c)], mult_third[(c, i)]) for c in rng).add(sc[(i, 0)])
        cur.update(mult_second)
        cur.update(mult_third)
        cur[(i, i)] = first
    return cur


def grid(n: int, base: int):
    """:312-314 — numBlksByRow = ceil(n / basesize), block length = ceil(n / numBlksByRow)."""
    nb = int(math.ceil(n / float(base)))
    return nb, int(math.ceil(n / float(nb)))


def block_pairs(blocks: Blocks, owner, rank: int):
    return [(BlockID(r, c), b) for (r, c), b in sorted(blocks.items()) if owner(r, c) == rank]
