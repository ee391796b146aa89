"""CPU-side tests: the C-ABI library loads and exports every symbol of include/marlin_b200.h, the pure integer
host logic matches the oracle, and compute entries fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from marlin_b200 import _native as nat

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    return nat.load()


def test_header_symbols_all_exported(lib):
    header = (ROOT / "include" / "marlin_b200.h").read_text()
    declared = set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", header))
    declared -= {"mb_ctx", "mb_block"}
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in marlin_b200.h but not exported by libmarlin_b200.so"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)


def test_no_torch_or_oracle_in_library():
    out = Path(nat.lib_path()).read_bytes()
    assert b"marlin_oracle" not in out and b"libtorch" not in out and b"libc10" not in out


def test_product_never_imports_oracle():
    for py in (ROOT / "marlin_b200").rglob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py


def test_version(lib):
    assert b"sm_100a" in lib.mb_version()


def test_init_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = nat.c_ctx()
    rc = lib.mb_init(0, C.byref(ctx))
    assert rc == nat.MB_ERR_CUDA
    assert b"no CPU fallback" in lib.mb_last_error()
    with pytest.raises(nat.MarlinError):
        nat.check(rc)


def test_host_api_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from marlin_b200 import SubMatrix
    a = SubMatrix(np.eye(2))                      # host-resident plumbing object
    with pytest.raises(nat.MarlinError):
        a.multiply(a)


@pytest.mark.parametrize("mkn_cores", [(16384, 16384, 16384, 8), (1048576, 1024, 1024, 8), (100, 100, 100, 1),
                                       (10, 1000, 10, 4), (7, 5, 3, 64), (1, 10, 10, 8), (65536, 65536, 65536, 64)])
def test_choose_split_matches_oracle(lib, oracle, mkn_cores):
    m, k, n, cores = mkn_cores
    out = (C.c_int32 * 3)()
    assert lib.mb_choose_split(m, k, n, cores, out) == 0
    assert tuple(out) == oracle.split_method(m, k, n, cores)


def test_choose_strategy(lib):
    def run(M, K, N, cores, thr=300, blk=0):
        s = C.c_int32()
        mkn = (C.c_int32 * 3)()
        assert lib.mb_choose_strategy(M, K, N, cores, thr, blk, C.byref(s), mkn) == 0
        return s.value, tuple(mkn)
    assert run(4, 4, 4, 2)[0] == 0                                   # DMS.scala:225-234 broadcast branch
    assert run(4096, 4096, 4096, 8)[0] == 0                          # cfg2: 16.8M elements <= 39,321,600
    assert run(16384, 16384, 16384, 8) == (2, (2, 2, 2))             # cfg3: square-ish shortcut floor(24^(1/3)) = 2
    assert run(16384, 16384, 16384, 8, blk=1) == (2, (2, 2, 2))      # BlockMatrix arm: splitMethod
    assert run(65536, 65536, 65536, 27) == (2, (4, 4, 4))            # cfg5: floor(81^(1/3)) = 4
    assert run(1048576, 1024, 1024, 8)[0] == 0                       # cfg4: B is 1M elements -> broadcast
    assert run(100, 100000, 100000, 8)[0] == 1                       # small A: the reference's quirk arm
    assert run(100000, 100000, 20000, 8) == (2, (4, 2, 1))           # not square-ish -> splitMethod


def test_partition_and_block_len(lib):
    assert lib.mb_mult_partition(1, 0, 1, 2, 2, 2) == 5
    assert lib.mb_elem_partition(1, 1, 2) == 3
    bl, ap = C.c_int32(), C.c_int32()
    assert lib.mb_block_len(100, 3, C.byref(bl), C.byref(ap)) == 0 and (bl.value, ap.value) == (34, 3)
    assert lib.mb_block_len(4, 3, C.byref(bl), C.byref(ap)) == 0 and (bl.value, ap.value) == (2, 2)
    assert lib.mb_block_len(0, 3, C.byref(bl), C.byref(ap)) == nat.MB_ERR_INVALID_ARG


def test_seed_logic_matches_oracle(lib, oracle):
    for seed in (0, 1, 42, -1, 2 ** 62 + 12345, -(2 ** 63)):
        assert lib.mb_hash_seed(seed) == oracle.hash_seed(seed)
    out = (C.c_int64 * 6)()
    assert lib.mb_partition_seeds(42, 6, out) == 0
    assert list(out) == oracle.java_random_longs(42, 6)
    assert out[0] == -5025562857975149833                    # java.util.Random(42).nextLong()


def test_blockid_and_partitioners():
    from marlin_b200 import BlockID, MatrixElemOpPartitioner, MatrixMultPartitioner
    assert BlockID(1, 2) == BlockID(1, 2, 0) and BlockID(1, 2, 3) != BlockID(1, 2, 0)     # Block.scala:39-43
    assert BlockID(1, 2, 3).hashCode() == 1 * 31 + 2 + 3                                  # Block.scala:45-47
    p = MatrixMultPartitioner(2, 2, 2)
    assert p.numPartitions == 8 and p.getPartition(BlockID(1, 0, 5)) == 5 and p.seq(1, 0, 1) == 5
    with pytest.raises(ValueError):
        p.getPartition("x")
    e = MatrixElemOpPartitioner(2, 3)
    assert e.numPartitions == 6 and e.getPartition(BlockID(1, 2)) == 5
    assert p == MatrixMultPartitioner(2, 2, 2) and e != MatrixElemOpPartitioner(3, 2)


def test_plan_multiply_cfg3_and_cfg5():
    from marlin_b200 import comm
    # cfg3: (2,2,2) on 8 ranks -> one product per rank (partition == GPU), pairs reduce onto the kk=0 rank
    plan = comm.plan_multiply(2, 2, 2, 8, lambda r, c: comm.elem_owner(r, c, 2, 8), lambda r, c: comm.elem_owner(r, c, 2, 8))
    assert all(len(v) == 1 for v in plan.products.values()) and len(plan.products) == 8
    assert plan.products[5] == [(1, 0, 1)]
    assert sorted(plan.c_reduces) == [(1, 0, (0, 0)), (3, 2, (0, 1)), (5, 4, (1, 0)), (7, 6, (1, 1))]
    # every product's operands arrive (or are local)
    for r, prods in plan.products.items():
        for (i, j, kk) in prods:
            assert comm.elem_owner(i, kk, 2, 8) == r or (comm.elem_owner(i, kk, 2, 8), r, (i, kk)) in plan.a_sends
            assert comm.elem_owner(kk, j, 2, 8) == r or (comm.elem_owner(kk, j, 2, 8), r, (kk, j)) in plan.b_sends
    # folding onto 4 / 2 / 1 ranks keeps the k-sum local (no reduce traffic)
    for ws in (4, 2, 1):
        p = comm.plan_multiply(2, 2, 2, ws, lambda r, c: comm.elem_owner(r, c, 2, ws), lambda r, c: comm.elem_owner(r, c, 2, ws))
        assert p.c_reduces == [] and sum(len(v) for v in p.products.values()) == 8
    # cfg5: (4,4,4) on 8 ranks is C-stationary: 2 C tiles per rank, all kk local
    p5 = comm.plan_multiply(4, 4, 4, 8, lambda r, c: comm.elem_owner(r, c, 4, 8), lambda r, c: comm.elem_owner(r, c, 4, 8))
    assert p5.c_reduces == [] and all(len(v) == 8 for v in p5.products.values())
    assert {(i, j) for (i, j, kk) in p5.products[3]} == {(1, 2), (1, 3)}


def test_dist_plan_matches_python_plan():
    """mb_dist_plan (csrc/dist.cu) and comm.plan_multiply deal the m*k*n products (seq = i*n*k + j*k + kk,
    BlockMatrix.scala:163,168) to ranks identically, and agree on the owner of every C tile."""
    import ctypes as C
    from marlin_b200 import comm
    lib = nat.load()
    for (m, k, n) in [(2, 2, 2), (1, 8, 1), (4, 4, 4), (3, 2, 2), (2, 3, 1), (1, 1, 1), (5, 1, 3)]:
        for world in (1, 2, 3, 4, 8):
            pr = (C.c_int32 * (m * k * n))()
            co = (C.c_int32 * (m * n))()
            assert lib.mb_dist_plan(m, k, n, world, pr, co) == 0
            plan = comm.plan_multiply(m, k, n, world, lambda r, c: 0, lambda r, c: 0)
            for r, prods in plan.products.items():
                for (i, j, kk) in prods:
                    assert pr[i * n * k + j * k + kk] == r
            for (i, j), o in plan.c_owner.items():
                assert co[i * n + j] == o


def test_dist_host_homes_is_a_balanced_matching():
    """mb_dist_host_homes: every input tile is uploaded by one of the ranks that multiply with it, and the uploads are
    spread as evenly over the ranks (PCIe links) as those constraints allow — e.g. the headline (2,2,2) grid on 8 GPUs
    gives every rank exactly one of the eight tiles."""
    import ctypes as C
    lib = nat.load()
    for (m, k, n) in [(2, 2, 2), (1, 2, 1), (4, 4, 4), (3, 2, 2), (1, 8, 1), (2, 1, 3)]:
        for world in (1, 2, 4, 8):
            pr = (C.c_int32 * (m * k * n))()
            assert lib.mb_dist_plan(m, k, n, world, pr, None) == 0
            ah = (C.c_int32 * (m * k))()
            bh = (C.c_int32 * (k * n))()
            assert lib.mb_dist_host_homes(m, k, n, world, ah, bh) == 0
            need_a = {t: set() for t in range(m * k)}
            need_b = {t: set() for t in range(k * n)}
            for i in range(m):
                for j in range(n):
                    for kk in range(k):
                        r = pr[i * n * k + j * k + kk]
                        need_a[i * k + kk].add(r)
                        need_b[kk * n + j].add(r)
            loads = [0] * world
            for t in range(m * k):
                assert ah[t] in need_a[t]
                loads[ah[t]] += 1
            for t in range(k * n):
                assert bh[t] in need_b[t]
                loads[bh[t]] += 1
            active = {r for s_ in (need_a, need_b) for v in s_.values() for r in v}
            assert max(loads) <= -(-(m * k + k * n) // len(active)) + 1, (m, k, n, world, loads)
    ah = (C.c_int32 * 4)()
    bh = (C.c_int32 * 4)()
    assert lib.mb_dist_host_homes(2, 2, 2, 8, ah, bh) == 0
    assert sorted(list(ah) + list(bh)) == list(range(8))         # one tile per GPU
