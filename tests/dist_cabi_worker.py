"""One rank of the torch-free multi-process test of the C-ABI multi-GPU multiply (mb_comm_* / mb_matmul_blocked_dist).
Only ctypes + numpy + the CPU oracle: no torch, no NCCL.   argv: rank world session [devices]
Ranks map to device rank % devices, so the protocol (shared-memory rendezvous, CUDA IPC pulls, band flags, fused
reduce-scatter, staged adds, FREE/DONE handshakes) is exercised on a one-GPU box too (two processes time-slicing it)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import importlib.util                            # noqa: E402
# the ctypes declarations only — loaded by path, because importing the marlin_b200 PACKAGE would pull in torch
_spec = importlib.util.spec_from_file_location("mb_native_standalone", ROOT / "marlin_b200" / "_native.py")
nat = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(nat)
from oracle import reference_model as rm        # noqa: E402


def main():
    rank, world, session = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    ndev = int(sys.argv[4]) if len(sys.argv) > 4 else world
    lib = nat.load(build_if_missing=False)
    ctx = nat.c_ctx()
    nat.check(lib.mb_init(rank % ndev, C.byref(ctx)))
    comm = C.c_void_p()
    nat.check(lib.mb_comm_init(ctx, rank, world, session.encode(), C.byref(comm)))
    assert lib.mb_comm_rank(comm) == rank and lib.mb_comm_world(comm) == world

    def upload(mat, view=None):
        """A device block holding `mat`; view = 'T' stores the transpose and hands out a transposed view, 'pad' a
        slice of a taller allocation (odd leading dimension handling is the library's business)."""
        f = np.asfortranarray(mat, dtype=np.float64)
        h = nat.c_blk()
        if view == "T":
            ft = np.asfortranarray(f.T)
            nat.check(lib.mb_block_upload(ctx, ft.ctypes.data_as(C.c_void_p), 0, ft.shape[0], ft.shape[1], max(1, ft.shape[0]), 0,
                                          nat.MB_F64, C.byref(h)))
            v = nat.c_blk()
            nat.check(lib.mb_block_view_t(ctx, h, C.byref(v)))
            return v
        if view == "pad":
            big = nat.c_blk()
            nat.check(lib.mb_block_alloc(ctx, f.shape[0] + 3, f.shape[1], nat.MB_F64, C.byref(big)))
            v = nat.c_blk()
            nat.check(lib.mb_block_slice(ctx, big, 1, 1 + f.shape[0], 0, f.shape[1], C.byref(v)))
            nat.check(lib.mb_block_upload(ctx, f.ctypes.data_as(C.c_void_p), 0, f.shape[0], f.shape[1], max(1, f.shape[0]), 0,
                                          nat.MB_F64, C.byref(h)))
            nat.check(lib.mb_block_copy(ctx, h, v))
            return v
        nat.check(lib.mb_block_upload(ctx, f.ctypes.data_as(C.c_void_p), 0, f.shape[0], f.shape[1], max(1, f.shape[0]), 0, nat.MB_F64,
                                      C.byref(h)))
        return h

    rng = np.random.default_rng(2026)               # the same stream on every rank
    # alternating plans: fused pairs, staged reduces with different slot layouts (the round-1 staging race: a (2,2,2)
    # multiply followed by a (1,8,1) one), C-stationary plans, ragged blocks, views; `place` shifts the owner maps.
    cases = [(256, 384, 640, 2, 2, 2, 0, None), (96, 1024, 64, 1, 8, 1, 1, None), (300, 260, 212, 2, 2, 2, 3, None),
             (1300, 520, 1400, 1, 2, 1, 0, None), (130, 120, 110, 3, 2, 2, 2, None), (256, 256, 256, 4, 4, 4, 1, None),
             (200, 300, 160, 2, 3, 1, 0, "T"), (1300, 520, 1400, 1, 2, 1, 1, None), (260, 520, 300, 1, 2, 2, 5, "pad"),
             (64, 512, 64, 1, 8, 1, 0, None), (1100, 260, 1200, 2, 1, 2, 1, None)]
    worst = 0.0
    if os.environ.get("MB_SKIP_SMALL") == "1":
        cases = cases[:1]
    for rep in range(2):
        for (M, K, N, m, k, n, place, view) in cases:
            A, B = rng.random((M, K)) - 0.5, rng.random((K, N)) - 0.5
            oa = rm.DenseVecMatrix(list(enumerate(A))).to_block_matrix(m, k)
            ob = rm.DenseVecMatrix(list(enumerate(B))).to_block_matrix(k, n)
            ta, tb = dict(oa.blocks), dict(ob.blocks)
            ref = dict(oa.multiply(ob, gemm="blas").blocks)
            a_own = [(i * k + kk + place) % world for i in range(m) for kk in range(k)]
            b_own = [(kk * n + j + 2 * place) % world for kk in range(k) for j in range(n)]
            a_h = (nat.c_blk * (m * k))()
            b_h = (nat.c_blk * (k * n))()
            c_h = (nat.c_blk * (m * n))()
            for t in range(m * k):
                if a_own[t] == rank:
                    a_h[t] = upload(ta[(t // k, t % k)], view if t % 2 == 0 else None)
            for t in range(k * n):
                if b_own[t] == rank:
                    b_h[t] = upload(tb[(t // n, t % n)], view if t % 2 == 1 else None)
            pr = (C.c_int32 * (m * k * n))()
            co = (C.c_int32 * (m * n))()
            nat.check(lib.mb_dist_plan(m, k, n, world, pr, co))
            row_len = (C.c_int32 * m)(*[ta[(i, 0)].shape[0] for i in range(m)])
            k_len = (C.c_int32 * k)(*[ta[(0, kk)].shape[1] for kk in range(k)])
            col_len = (C.c_int32 * n)(*[tb[(0, j)].shape[1] for j in range(n)])
            mine = [t for t in range(m * n) if co[t] == rank]
            for t in mine:
                h = nat.c_blk()
                nat.check(lib.mb_block_alloc(ctx, row_len[t // n], col_len[t % n], nat.MB_F64, C.byref(h)))
                nat.check(lib.mb_block_fill(ctx, h, float("nan")))
                c_h[t] = h
            for _ in range(2):                                   # back-to-back epochs reuse staging slots and flags
                nat.check(lib.mb_matmul_blocked_dist(comm, a_h, (C.c_int32 * (m * k))(*a_own), b_h, (C.c_int32 * (k * n))(*b_own),
                                                     m, k, n, row_len, k_len, col_len, nat.MB_F64, c_h))
            for t in mine:
                i, j = divmod(t, n)
                got = np.empty((row_len[i], col_len[j]), order="F")
                nat.check(lib.mb_block_download(ctx, c_h[t], got.ctypes.data_as(C.c_void_p), max(1, row_len[i])))
                Ai = np.hstack([ta[(i, kk)] for kk in range(k)])
                Bj = np.vstack([tb[(kk, j)] for kk in range(k)])
                err = (np.abs(got - ref[(i, j)]) / (np.abs(Ai) @ np.abs(Bj))).max()
                assert err <= 1e-10, (rep, M, K, N, m, k, n, place, view, i, j, err)
                worst = max(worst, err)
            nat.check(lib.mb_comm_check(comm))
            for arr in (a_h, b_h, c_h):
                for h in arr:
                    if h:
                        nat.check(lib.mb_block_free(ctx, h))
    # ---- the end-to-end entry: HOST tiles in, HOST tiles out (mb_matmul_blocked_dist_host), C tiles in shared pinned memory ----
    host_cases = [(1300, 520, 1400, 1, 2, 1), (300, 260, 212, 2, 2, 2), (1100, 2100, 1200, 2, 2, 2), (2304, 1040, 1100, 1, 2, 1)]
    for ci, (M, K, N, m, k, n) in enumerate(host_cases):
        A, B = rng.random((M, K)) - 0.5, rng.random((K, N)) - 0.5
        oa = rm.DenseVecMatrix(list(enumerate(A))).to_block_matrix(m, k)
        ob = rm.DenseVecMatrix(list(enumerate(B))).to_block_matrix(k, n)
        ta, tb = dict(oa.blocks), dict(ob.blocks)
        ref = dict(oa.multiply(ob, gemm="blas").blocks)
        a_home = (C.c_int32 * (m * k))()
        b_home = (C.c_int32 * (k * n))()
        nat.check(lib.mb_dist_host_homes(m, k, n, world, a_home, b_home))
        loads = [0] * world
        for h in list(a_home) + list(b_home):
            loads[h] += 1
        assert max(loads) - min(loads) <= 1 or world > m * k + k * n, loads          # uploads spread over the PCIe links
        pr = (C.c_int32 * (m * k * n))()
        co = (C.c_int32 * (m * n))()
        nat.check(lib.mb_dist_plan(m, k, n, world, pr, co))
        a_np = {t: np.asfortranarray(ta[(t // k, t % k)]) for t in range(m * k) if a_home[t] == rank}
        b_np = {t: np.asfortranarray(tb[(t // n, t % n)]) for t in range(k * n) if b_home[t] == rank}
        pa = (C.c_void_p * (m * k))(*[a_np[t].ctypes.data if t in a_np else None for t in range(m * k)])
        pb = (C.c_void_p * (k * n))(*[b_np[t].ctypes.data if t in b_np else None for t in range(k * n)])
        row_len = (C.c_int32 * m)(*[ta[(i, 0)].shape[0] for i in range(m)])
        k_len = (C.c_int32 * k)(*[ta[(0, kk)].shape[1] for kk in range(k)])
        col_len = (C.c_int32 * n)(*[tb[(0, j)].shape[1] for j in range(n)])
        mine = sorted({(s // k) // n * n + (s // k) % n for s in range(m * k * n) if pr[s] == rank})     # C tiles I hold a partial of
        pc = (C.c_void_p * (m * n))()
        shared = {}
        for t in mine:
            ptr = C.c_void_p()
            nbytes = int(row_len[t // n]) * int(col_len[t % n]) * 8
            nat.check(lib.mb_host_alloc_shared(f"{session}_{ci}_{t}".encode(), nbytes, C.byref(ptr)))
            shared[t] = (ptr, nbytes)
            pc[t] = ptr
        for rep in range(2):
            for t in mine:
                if co[t] == rank:
                    C.memset(shared[t][0], 0xFF, shared[t][1])           # NaN pattern; every byte must be overwritten
            nat.check(lib.mb_comm_barrier(comm))
            nat.check(lib.mb_matmul_blocked_dist_host(comm, pa, a_home, pb, b_home, m, k, n, row_len, k_len, col_len, pc))
            nat.check(lib.mb_comm_barrier(comm))                         # both holders have written their sub-blocks
            for t in mine:
                i, j = divmod(t, n)
                got = np.ctypeslib.as_array(C.cast(shared[t][0], C.POINTER(C.c_double)), shape=(col_len[j], row_len[i])).T
                Ai = np.hstack([ta[(i, kk)] for kk in range(k)])
                Bj = np.vstack([tb[(kk, j)] for kk in range(k)])
                err = (np.abs(got - ref[(i, j)]) / (np.abs(Ai) @ np.abs(Bj))).max()
                assert err <= 1e-10, ("host", M, K, N, m, k, n, i, j, rep, err)
                worst = max(worst, err)
            nat.check(lib.mb_comm_barrier(comm))
        for t in mine:
            nat.check(lib.mb_host_free_shared(f"{session}_{ci}_{t}".encode(), shared[t][0], shared[t][1], 1 if co[t] == rank else 0))
    # ---- optional: the fused GEMM + reduce-scatter at full size, back to back (MB_BIG_FUSED=n): an n x n product with the
    #      contraction split over the two ranks ((1,2,1) grid), six multiplies queued without a host sync in between (flags,
    #      staging slots and the remotely written C tile are reused from call to call), then Freivalds on the last result ----
    bigf = int(os.environ.get("MB_BIG_FUSED", "0"))
    if bigf and world == 2:
        nn, kh = bigf, bigf // 2
        a_own, b_own = [0, 1], [0, 1]
        a_h = (nat.c_blk * 2)()
        b_h = (nat.c_blk * 2)()
        c_h = (nat.c_blk * 1)()
        h = nat.c_blk()
        nat.check(lib.mb_block_alloc(ctx, nn, kh, nat.MB_F64, C.byref(h)))
        nat.check(lib.mb_fill_uniform(ctx, h, 1000 + rank, 0, 0.0, 1.0, 0))
        a_h[rank] = h
        h2 = nat.c_blk()
        nat.check(lib.mb_block_alloc(ctx, kh, nn, nat.MB_F64, C.byref(h2)))
        nat.check(lib.mb_fill_uniform(ctx, h2, 2000 + rank, 0, 0.0, 1.0, 0))
        b_h[rank] = h2
        co = (C.c_int32 * 1)()
        nat.check(lib.mb_dist_plan(1, 2, 1, world, None, co))
        cs = []
        for rep in range(6):
            if co[0] == rank:
                hc = nat.c_blk()
                nat.check(lib.mb_block_alloc(ctx, nn, nn, nat.MB_F64, C.byref(hc)))
                cs.append(hc)
                c_h[0] = hc
            nat.check(lib.mb_matmul_blocked_dist(comm, a_h, (C.c_int32 * 2)(*a_own), b_h, (C.c_int32 * 2)(*b_own), 1, 2, 1,
                                                 (C.c_int32 * 1)(nn), (C.c_int32 * 2)(kh, kh), (C.c_int32 * 1)(nn), nat.MB_F64, c_h))
        nat.check(lib.mb_synchronize(ctx))
        nat.check(lib.mb_comm_check(comm))
        # gather x-products on the host: y = C x on the owner, z = sum_kk A_kk (B_kk x) from both ranks through shared memory
        x = np.random.default_rng(11).random(nn)
        Ak = np.empty((nn, kh), order="F")
        Bk = np.empty((kh, nn), order="F")
        nat.check(lib.mb_block_download(ctx, a_h[rank], Ak.ctypes.data_as(C.c_void_p), nn))
        nat.check(lib.mb_block_download(ctx, b_h[rank], Bk.ctypes.data_as(C.c_void_p), kh))
        ptr = C.c_void_p()
        nat.check(lib.mb_host_alloc_shared(f"{session}_z".encode(), 2 * nn * 8, C.byref(ptr)))
        zsh = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2, nn))
        zsh[rank, :] = Ak @ (Bk @ x)
        nat.check(lib.mb_comm_barrier(comm))
        # exact element-wise reference for a diagnosis: the owner gathers the partner's operand tiles through shared memory
        ptrA = C.c_void_p(); ptrB = C.c_void_p()
        nat.check(lib.mb_host_alloc_shared(f"{session}_gA".encode(), 2 * nn * kh * 8, C.byref(ptrA)))
        nat.check(lib.mb_host_alloc_shared(f"{session}_gB".encode(), 2 * nn * kh * 8, C.byref(ptrB)))
        gA = np.ctypeslib.as_array(C.cast(ptrA, C.POINTER(C.c_double)), shape=(2, kh, nn))
        gB = np.ctypeslib.as_array(C.cast(ptrB, C.POINTER(C.c_double)), shape=(2, nn, kh))
        gA[rank] = Ak.T
        gB[rank] = Bk.T
        nat.check(lib.mb_comm_barrier(comm))
        if co[0] == rank:
            z = zsh[0] + zsh[1]
            worst_idx = None
            for idx in range(len(cs)):
                Cm = np.empty((nn, nn), order="F")
                nat.check(lib.mb_block_download(ctx, cs[idx], Cm.ctypes.data_as(C.c_void_p), nn))
                err = (np.abs(Cm @ x - z) / z).max()
                print(f"big fused {nn}^2 call {idx}: Freivalds scaled error {err:.2e}", flush=True)
                if err > 1e-10 and worst_idx is None:
                    worst_idx = idx
                    ref = gA[0].T @ gB[0].T + gA[1].T @ gB[1].T
                    rel = np.abs(Cm - ref) / ref
                    bad = np.argwhere(rel > 1e-12)
                    print(f"  diagnosis: {len(bad)} elements off; max rel {rel.max():.3e}; rows {bad[:, 0].min()}..{bad[:, 0].max()} cols "
                          f"{bad[:, 1].min()}..{bad[:, 1].max()}; left half {int((bad[:, 1] < nn // 2).sum())}, right half {int((bad[:, 1] >= nn // 2).sum())}", flush=True)
                    for (r_, c_) in bad[:12]:
                        p0 = gA[0].T[r_] @ gB[0].T[:, c_]
                        p1 = gA[1].T[r_] @ gB[1].T[:, c_]
                        print(f"    ({r_},{c_}) got {Cm[r_, c_]:.6f} ref {ref[r_, c_]:.6f} P0 {p0:.6f} P1 {p1:.6f}", flush=True)
                    tiles_bad = sorted({(int(r_) // 128, int(c_) // 128) for r_, c_ in bad})
                    print(f"    128x128 tiles touched: {len(tiles_bad)} e.g. {tiles_bad[:16]}", flush=True)
            assert worst_idx is None, ("big fused", worst_idx)
        nat.check(lib.mb_comm_barrier(comm))

    # ---- optional: the bench-sized end-to-end case (MB_BIG=n): pinned shared host tiles on BOTH sides, so every copy is
    #      truly asynchronous and the grouped GEMM is resident long before its operands arrive; checked with Freivalds ----
    big = int(os.environ.get("MB_BIG", "0"))
    for grid in ([tuple(int(v) for v in g_.split(",")) for g_ in os.environ.get("MB_BIG_GRIDS", "2,2,2;1,2,1").split(";")] if big else []):
        gm, gk, gn = grid
        rl, kl, cl = big // gm, big // gk, big // gn
        a_home = (C.c_int32 * (gm * gk))()
        b_home = (C.c_int32 * (gk * gn))()
        nat.check(lib.mb_dist_host_homes(gm, gk, gn, world, a_home, b_home))
        pr = (C.c_int32 * (gm * gk * gn))()
        co = (C.c_int32 * (gm * gn))()
        nat.check(lib.mb_dist_plan(gm, gk, gn, world, pr, co))
        tagg = f"{gm}{gk}{gn}"

        def shared(tag, t, rows, cols):
            ptr = C.c_void_p()
            nat.check(lib.mb_host_alloc_shared(f"{session}_{tagg}{tag}{t}".encode(), rows * cols * 8, C.byref(ptr)))
            return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(cols, rows)).T      # column-major view

        tiles = {}
        for t in range(gm * gk):
            tiles[("A", t)] = shared("A", t, rl, kl)            # every rank maps every input tile (rank 0 fills them all)
        for t in range(gk * gn):
            tiles[("B", t)] = shared("B", t, kl, cl)
        if rank == 0:
            r2 = np.random.default_rng(99)
            for (tag, t), (_, view) in tiles.items():
                view[...] = r2.random(view.shape)
        ctiles = {t: shared("C", t, rl, cl) for t in range(gm * gn)}
        nat.check(lib.mb_comm_barrier(comm))
        pa = (C.c_void_p * (gm * gk))(*[tiles[("A", t)][0] if a_home[t] == rank else None for t in range(gm * gk)])
        pb = (C.c_void_p * (gk * gn))(*[tiles[("B", t)][0] if b_home[t] == rank else None for t in range(gk * gn)])
        mine = sorted({s_ // gk for s_ in range(gm * gk * gn) if pr[s_] == rank})
        pc = (C.c_void_p * (gm * gn))(*[ctiles[t][0] if t in mine else None for t in range(gm * gn)])
        import time
        for rep in range(4):
            if rank == 0:
                for t in range(gm * gn):
                    ctiles[t][1][...] = np.nan
            nat.check(lib.mb_comm_barrier(comm))
            t0 = time.perf_counter()
            nat.check(lib.mb_matmul_blocked_dist_host(comm, pa, a_home, pb, b_home, gm, gk, gn, (C.c_int32 * gm)(*([rl] * gm)),
                                                      (C.c_int32 * gk)(*([kl] * gk)), (C.c_int32 * gn)(*([cl] * gn)), pc))
            nat.check(lib.mb_comm_barrier(comm))
            dt = time.perf_counter() - t0
            if rank == 0:
                print(f"big e2e {big}^2 grid {grid} on {world} ranks / {ndev} GPU(s): {dt * 1e3:.1f} ms = {2.0 * big ** 3 / dt / 1e12:.1f} TFLOP/s", flush=True)
                Afull = np.block([[tiles[("A", i * gk + kk)][1] for kk in range(gk)] for i in range(gm)])
                Bfull = np.block([[tiles[("B", kk * gn + j)][1] for j in range(gn)] for kk in range(gk)])
                Cfull = np.block([[ctiles[i * gn + j][1] for j in range(gn)] for i in range(gm)])
                x = np.random.default_rng(3 + rep).random(big)
                lhs, rhs = Cfull @ x, Afull @ (Bfull @ x)
                err = (np.abs(lhs - rhs) / rhs).max()
                assert err <= 1e-10, ("big", grid, rep, err)
                print(f"big e2e parity grid {grid} rep {rep} (Freivalds) {err:.2e}", flush=True)
        nat.check(lib.mb_comm_barrier(comm))
    nat.check(lib.mb_comm_barrier(comm))
    nat.check(lib.mb_comm_destroy(comm))
    nat.check(lib.mb_shutdown(ctx))
    assert "torch" not in sys.modules, "this worker must stay torch-free"
    print(f"cabi rank {rank}/{world} ok worst_scaled_err={worst:.2e} slow={os.environ.get('MARLIN_B200_DIST_SLOW', '0')}", flush=True)


if __name__ == "__main__":
    main()
