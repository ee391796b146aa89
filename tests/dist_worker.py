"""Worker for the world_size-2 gloo test (CPU): exercises the multi-rank host logic of the multiply path —
placement, the exchange plan, grouped P2P transfers and metadata gathers — with host-resident blocks.
No arithmetic happens here (the kernels are GPU-only); payloads are checked for identity."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import marlin_b200 as mb
from marlin_b200 import comm


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    assert ws == 2
    m, k, n = 2, 2, 2
    M = K = N = 8
    rng = np.random.default_rng(0)                       # same on both ranks
    A, B = rng.random((M, K)), rng.random((K, N))
    tile = lambda X, r, c: X[r * 4:(r + 1) * 4, c * 4:(c + 1) * 4]
    cpu = torch.device("cpu")
    a_blocks = [(mb.BlockID(r, c), mb.SubMatrix(tile(A, r, c), device=cpu)) for r in range(2) for c in range(2)
                if comm.elem_owner(r, c, 2, ws) == rank]
    b_blocks = [(mb.BlockID(r, c), mb.SubMatrix(tile(B, r, c), device=cpu)) for r in range(2) for c in range(2)
                if comm.elem_owner(r, c, 2, ws) == rank]
    am = mb.BlockMatrix(a_blocks)
    bm = mb.BlockMatrix(b_blocks)
    # lazily derived dims need a metadata gather across ranks (BlockMatrix.scala:36-65)
    assert (am.numRows(), am.numCols(), am.numBlksByRow(), am.numBlksByCol()) == (8, 8, 2, 2)
    assert am.elementsCount() == 4
    assert np.array_equal(am.toBreeze(), A) and np.array_equal(bm.toBreeze(), B)
    # the plan: 8 products over 2 ranks in contiguous seq ranges, k-sum local => no reduce traffic
    plan = comm.plan_multiply(m, k, n, ws, am.owner, bm.owner)
    assert sorted(plan.products) == [0, 1] and all(len(v) == 4 for v in plan.products.values())
    assert plan.products[0] == [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)] and plan.c_reduces == []
    # run the tile replication exactly as BlockMatrix._multiply_same_grid does, on CPU tensors over gloo
    a_local = {(b.row, b.column): s for b, s in am.blocks}
    b_local = {(b.row, b.column): s for b, s in bm.blocks}
    sends = [(s, d, ("A",) + key) for s, d, key in plan.a_sends] + [(s, d, ("B",) + key) for s, d, key in plan.b_sends]
    bufs = {}
    for s, d, key in sends:
        if s == rank:
            src = (a_local if key[0] == "A" else b_local)[key[1:]]
            bufs[key] = src.buf[:16]
    got = comm.exchange(sends, bufs, lambda key: torch.empty(16, dtype=torch.float64), rank)
    tiles_a = {k_: v.toBreeze() for k_, v in a_local.items()}
    tiles_b = {k_: v.toBreeze() for k_, v in b_local.items()}
    for key, buf in got.items():
        arr = buf.numpy().reshape((4, 4), order="F")
        (tiles_a if key[0] == "A" else tiles_b)[key[1:]] = arr
    for (i, j, kk) in plan.products[rank]:
        assert np.array_equal(tiles_a[(i, kk)], tile(A, i, kk))
        assert np.array_equal(tiles_b[(kk, j)], tile(B, kk, j))
    # a product that needs a k-split across ranks: (1,2,1) on 2 ranks => partial of rank 1 reduces onto rank 0
    p2 = comm.plan_multiply(1, 2, 1, 2, lambda r, c: c % 2, lambda r, c: r % 2)
    assert p2.c_reduces == [(1, 0, (0, 0))] and p2.c_owner == {(0, 0): 0} and p2.a_sends == [] and p2.b_sends == []
    # DenseVecMatrix metadata across ranks
    rows = [(i, A[i]) for i in range(M) if i % 2 == rank]
    dv = mb.DenseVecMatrix(rows, device=cpu)
    assert (dv.numRows(), dv.numCols()) == (8, 8)
    assert np.array_equal(dv.toBreeze(), A)
    # DistributedVector: pieces live on rank id mod G; metadata gathers and the replication before a matrix x vector
    v = np.arange(10.0)
    dvec = mb.DistributedVector.fromVector(None, v, 3)                    # pieces 4, 4, 2
    assert [i for i, _ in dvec.vectors] == [i for i in range(3) if i % 2 == rank]
    assert dvec.length == 10 and dvec.splitNum == 3 and dvec.owner(2) == 0
    lazy = mb.DistributedVector(dvec.vectors)                             # length / splitNum derived by a gather (:31-43)
    assert lazy.length == 10 and lazy.splitNum == 3
    rep = dvec._replicated()
    assert sorted(rep) == [0, 1, 2] and [rep[i].rows for i in range(3)] == [4, 4, 2]
    assert np.array_equal(np.concatenate([rep[i].toBreeze().reshape(-1) for i in range(3)]), v)
    assert np.array_equal(mb.DistributedVector.fromVector(None, np.arange(12.0), 3).toBreeze(), np.arange(12.0))
    # compute entries still refuse to run on host memory
    try:
        am.blocks[0][1].multiply(bm.blocks[0][1])
        raise SystemExit("expected MarlinError: no CPU fallback")
    except mb.MarlinError:
        pass
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
