"""Property tests (hypothesis, CPU) of the integer logic either side of the kernels: the split chooser behind the C ABI
against the oracle's restatement on random shapes, the product -> rank plan (every product once, every operand tile
delivered, every partial reduced onto its owner), the re-grid split tables, ceil block sizing, and the DistributedVector
re-split of the reference suite generalised to random lengths.  No GPU, no arithmetic."""
import ctypes as C
import math

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from marlin_b200 import _native as nat
from marlin_b200 import comm


@pytest.fixture(scope="module")
def lib():
    return nat.load()


@settings(max_examples=300, deadline=None)
@given(m=st.integers(1, 2_000_000), k=st.integers(1, 2_000_000), n=st.integers(1, 2_000_000), cores=st.integers(1, 512))
def test_choose_split_equals_oracle_and_is_a_power_of_two_cut(lib, oracle, m, k, n, cores):
    out = (C.c_int32 * 3)()
    assert lib.mb_choose_split(m, k, n, cores, out) == 0
    got = tuple(out)
    assert got == oracle.split_method(m, k, n, cores)              # utils/MTUtils.scala:150-175
    # the chooser halves the largest dimension per step while cores > 1 and no dimension has shrunk to 1: the grid is
    # 2^floor(log2(cores)) products unless a dimension ran out first
    assert all(g >= 1 and (g & (g - 1)) == 0 for g in got)
    bound = 2 ** int(math.floor(math.log2(cores)))
    prod = got[0] * got[1] * got[2]
    assert prod <= bound
    if prod < bound:
        assert min(m // got[0], k // got[1], n // got[2]) <= 1


@settings(max_examples=200, deadline=None)
@given(total=st.integers(1, 10 ** 7), parts=st.integers(1, 4096))
def test_block_len_is_scala_ceil(lib, total, parts):
    bl, ap = C.c_int32(), C.c_int32()
    assert lib.mb_block_len(total, parts, C.byref(bl), C.byref(ap)) == 0
    want = int(math.ceil(float(total) / float(parts)))             # math.ceil(x.toDouble / y.toDouble).toInt
    assert bl.value == want and ap.value == int(math.ceil(float(total) / float(want)))
    assert (ap.value - 1) * bl.value < total <= ap.value * bl.value


@settings(max_examples=200, deadline=None)
@given(m=st.integers(1, 6), k=st.integers(1, 6), n=st.integers(1, 6), world=st.integers(1, 8), shift_a=st.integers(0, 7),
       shift_b=st.integers(0, 7))
def test_plan_multiply_invariants(lib, m, k, n, world, shift_a, shift_b):
    a_owner = lambda r, c: (comm.elem_owner(r, c, k, world) + shift_a) % world
    b_owner = lambda r, c: (comm.elem_owner(r, c, n, world) + shift_b) % world
    plan = comm.plan_multiply(m, k, n, world, a_owner, b_owner)
    prods = [p for r in sorted(plan.products) for p in plan.products[r]]
    assert sorted(prods) == sorted((i, j, kk) for i in range(m) for j in range(n) for kk in range(k))     # each once
    where = {p: r for r, ps in plan.products.items() for p in ps}
    for r, ps in plan.products.items():
        seqs = [lib.mb_mult_partition(i, j, kk, m, k, n) for (i, j, kk) in ps]
        assert seqs == sorted(seqs) and seqs == list(range(seqs[0], seqs[0] + len(seqs)))               # contiguous seq range
        for (i, j, kk) in ps:
            assert a_owner(i, kk) == r or (a_owner(i, kk), r, (i, kk)) in plan.a_sends
            assert b_owner(kk, j) == r or (b_owner(kk, j), r, (kk, j)) in plan.b_sends
    assert all(s != d for s, d, _ in plan.a_sends + plan.b_sends + plan.c_reduces)
    assert len(set(plan.a_sends)) == len(plan.a_sends) and len(set(plan.b_sends)) == len(plan.b_sends)
    for (i, j), owner in plan.c_owner.items():
        holders = {where[(i, j, kk)] for kk in range(k)}
        assert owner == where[(i, j, 0)]                                       # the rank of the kk = 0 partial keeps the tile
        assert {s for s, d, key in plan.c_reduces if key == (i, j)} == holders - {owner}
        assert all(d == owner for s, d, key in plan.c_reduces if key == (i, j))
    if (m * n) % world == 0:                      # whole C tiles per rank: C-stationary, no reduce traffic (configs 3 at G<=4, 5)
        assert plan.c_reduces == []
    loads = [len(plan.products.get(r, [])) for r in range(world)]
    if m * k * n >= world:                        # products are dealt evenly even when that splits a kk-group (one reduce
        assert max(loads) - min(loads) <= 1       # of a C tile costs less than an idle GPU)


@settings(max_examples=200, deadline=None)
@given(data=st.data())
def test_regrid_split_tables_tile_the_ranges(oracle, data):
    """MTUtils.splitMethod(oldRange, newSubBlk) (utils/MTUtils.scala:182-202): the pieces of every old block are
    contiguous, cover it exactly, stay inside one new block each and land at the right offset there."""
    nblk = data.draw(st.integers(1, 8))
    old_len = data.draw(st.integers(1, 50))
    new_len = data.draw(st.integers(1, 50))
    total = data.draw(st.integers((nblk - 1) * old_len + 1, nblk * old_len))
    ranges = [(b * old_len, min((b + 1) * old_len, total) - 1) for b in range(nblk)]
    status = oracle.regrid_split_method(ranges, new_len)
    assert len(status) == nblk
    for (start, end), pieces in zip(ranges, status):
        pos = 0
        for new_id, (o0, o1), (n0, n1) in pieces:
            assert o0 == pos and o1 >= o0 and o1 - o0 == n1 - n0
            g0, g1 = start + o0, start + o1                                     # global rows of the piece
            assert g0 // new_len == new_id == g1 // new_len and g0 % new_len == n0 and g1 % new_len == n1
            pos = o1 + 1
        assert pos == end - start + 1


@settings(max_examples=100, deadline=None)
@given(length=st.integers(1, 400), old_splits=st.integers(1, 7), new_splits=st.integers(1, 9))
def test_distributed_vector_resplit_roundtrip(oracle, length, old_splits, new_splits):
    """DistributedVector.toDisVector (matrix/DistributedVector.scala:84-107) driven by the re-grid tables, as the
    reference's callers do (examples/NeuralNetwork.scala:78-79): the re-split vector has the same elements."""
    v = np.arange(float(length))
    old_len = int(math.ceil(length / old_splits))
    if (old_splits - 1) * old_len >= length:
        return                                           # fromVector would produce an empty trailing piece
    dv = oracle.DistributedVector.from_vector(v, old_splits)
    new_len = int(math.ceil(length / new_splits))
    ranges = [(i * old_len, min((i + 1) * old_len, length) - 1) for i in range(old_splits)]
    status = oracle.regrid_split_method(ranges, new_len)
    out = dv.to_dis_vector(status, new_splits)
    got = np.concatenate([p for _, p in sorted(out.vectors, key=lambda t: t[0])])
    assert np.array_equal(got, v)
    assert all(p.shape[0] <= new_len for _, p in out.vectors)
