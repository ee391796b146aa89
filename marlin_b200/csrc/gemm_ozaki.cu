// fp64 block GEMM on the int8 tensor cores (tcgen05.mma.kind::i8) — the literal "fp64 multiply as a tcgen05 kernel".
//
// tcgen05 has no fp64 kind, and DMMA tops out at 37 TFLOP/s on B200; the int8 pipe is ~100x wider.  This path splits
// each operand into `s` signed 7-bit digit planes after exact power-of-two row/column scaling (an Ozaki split):
//     a_ik = 2^(eA_i - P) * X_ik,   X_ik = sum_t dA_t(i,k) * 128^(s-t),   |dA_t| <= 64,   P = 7s - 1
//     b_kj = 2^(eB_j - P) * Y_kj    likewise
// so that   sum_k a_ik b_kj = 2^(eA_i + eB_j - 2P) * sum_{t,u} 128^(2s-t-u) * [ sum_k dA_t(i,k) dB_u(k,j) ] .
// Every bracket is an int8 GEMM accumulated EXACTLY in int32 (TMEM).  Pairs with t + u > s + 1 are dropped (their
// weight is below the rounding of the operands), the brackets of one group d = t + u share an accumulator, and the
// epilogue adds the groups into C in fp64 from the least significant group up.  With s = 7 the result is within a few
// 1e-14 of the exact product relative to (|A||B|)_ij for well-scaled rows; s = 6 stays below 1e-11.
// The error is relative to rowmax_i(A)*colmax_j(B)*K, not to (|A||B|)_ij, so this mode is opt-in (mb_set_fp64_mode)
// and the DMMA kernel remains the default.
//
// Kernel structure = the bf16 tcgen05 kernel (TMA producer warp, single-thread MMA issuer, 4 epilogue warps,
// double-buffered TMEM accumulators), with 3-D TMA maps over the digit planes (A planes MN-major, B planes K-major)
// and an fp64 read-modify-write epilogue.
#include "gemm_ozaki.h"
#include "ptx.cuh"
#include <atomic>
#include <cstdlib>

namespace mb {

namespace {

constexpr int BM = 128, BN = 256, BK = 128;       // BK int8 = 128 B = one swizzle row
constexpr int A_BYTES = BM * BK;                   // 16 KiB
constexpr int B_BYTES = BN * BK;                   // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_STAGES = 4;
constexpr int NUM_ACC = 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + 1024 + 256;
constexpr int BAND = 16;

struct Params {
    int M, N, K, s;
    int bits;          // digit width: 7 (|digit| <= 64) or 8 (|digit| <= 128, top digit <= 64)
    int P;             // fractional bits of the scaled operands: bits*s - (bits == 8 ? 2 : 1)
    double* C;
    long long ldc;
    const int* eA;
    const int* eB;
    int tiles_m, tiles_n;
    int accumulate;
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
    const int band_tiles = BAND * tiles_n;
    const int band = t / band_tiles;
    const int r = t - band * band_tiles;
    const int rows_in_band = min(BAND, tiles_m - band * BAND);
    tn = r / rows_in_band;
    tm = band * BAND + (r - tn * rows_in_band);
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;     // SWIZZLE_128B
    return d;
}
// s8 x s8 -> s32, A MN-major, B K-major, M = 128, N = 256
__host__ __device__ constexpr uint32_t instr_desc_i8() {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ double pow2_scale(int ex) {
    if (ex >= -1022 && ex <= 1023) return __hiloint2double((ex + 1023) << 20, 0);
    return ldexp(1.0, ex);
}

// One 32-column chunk of one digit group: C[m][n0..n0+31] (+)= acc * 2^(ex_m + eB[n]).  All 32 old values of C are
// loaded before any store so the read-modify-write costs ONE memory latency per chunk instead of 32 dependent ones.
__device__ __forceinline__ void fold_group_chunk(const uint32_t (&v)[32], double* crow, long long ldc, int n_first, int N,
                                                 const int* __restrict__ eB, int ex_m, bool init) {
    double old[32];
    if (!init) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int n = n_first + j;
            old[j] = (n < N) ? __ldcg(crow + (long long)n * ldc) : 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = n_first + j;
        if (n < N) {
            const double term = (double)(int)v[j] * pow2_scale(ex_m + __ldg(eB + n));
            crow[(long long)n * ldc] = init ? term : (old[j] + term);
        }
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_ozaki_i8_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NUM_STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar_base;
    const uint32_t bar_empty = bar_full + 8 * NUM_STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * NUM_STAGES;
    const uint32_t bar_tempty = bar_tfull + 8 * NUM_ACC;
    const uint32_t tmem_slot = bar_tempty + 8 * NUM_ACC;
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int num_kb = (p.K + BK - 1) / BK;
    const int s = p.s;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int a = 0; a < NUM_ACC; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
        fence_barrier_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, NUM_ACC * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int d = s + 1; d >= 2; --d) {
                    const int t_lo = max(1, d - s), t_hi = min(s, d - 1);
                    for (int ta = t_lo; ta <= t_hi; ++ta) {
                        const int ub = d - ta;
                        for (int kb = 0; kb < num_kb; ++kb) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t full = bar_full + 8 * stage;
                            const uint32_t sA = smem_base + stage * STAGE_BYTES;
                            mbar_arrive_expect_tx(full, STAGE_BYTES);
                            tma_load_3d(sA, &mapA, full, m0, kb * BK, ta - 1);
                            tma_load_3d(sA + A_BYTES, &mapB, full, kb * BK, n0, ub - 1);
                            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_i8();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                for (int d = s + 1; d >= 2; --d) {
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    const int npairs = min(s, d - 1) - max(1, d - s) + 1;
                    uint32_t first = 1;
                    for (int pr = 0; pr < npairs; ++pr) {
                        for (int kb = 0; kb < num_kb; ++kb) {
                            mbar_wait(bar_full + 8 * stage, phase);
                            tc_fence_after();
                            const uint32_t sA = smem_base + stage * STAGE_BYTES;
                            const uint32_t sB = sA + A_BYTES;
#pragma unroll
                            for (int k = 0; k < BK / 32; ++k) {
                                // A planes are MN-major: 32 k-rows of 128 B per UMMA_K; B planes K-major: 32 B per UMMA_K
                                const uint64_t adesc = smem_desc(sA + k * 4096, 0, 1024);
                                const uint64_t bdesc = smem_desc(sB + k * 32, 0, 1024);
                                umma_i8(d_tmem, adesc, bdesc, idesc, first ? 0u : 1u);
                                first = 0;
                            }
                            umma_commit(bar_empty + 8 * stage);
                            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    umma_commit(bar_tfull + 8 * acc);
                    if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int P2 = 2 * p.P;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int tm, tn;
            tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
            const int m = tm * BM + quarter * 32 + lane;
            const int n0 = tn * BN;
            const bool m_ok = m < p.M;
            const int ea = m_ok ? __ldg(p.eA + m) : 0;
            double* crow = p.C + m;
            for (int d = s + 1; d >= 2; --d) {
                const bool init = (d == s + 1) && !p.accumulate;
                const int ex_m = ea - P2 + p.bits * (2 * s - d);
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    }
                    if (m_ok) fold_group_chunk(v, crow, p.ldc, n0 + c * 32, p.N, p.eB, ex_m, init);
                }
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, NUM_ACC * BN);
    }
}

// -------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA stages its own
// 128 rows of the A plane and HALF of the B plane (128 of the 256 columns); one tcgen05.mma.cta_group::2 issued by the
// leader CTA drives both SMs' tensor cores, reading B from both CTAs' shared memory.  Per SM and k-slab that is
// 32 KiB from L2 instead of 48 KiB for the same 128 x 256 x 128 MACs — the int8 path is L2 -> SM bandwidth bound,
// so this is worth ~1.5x.  Barriers: `full` and `tmem_empty` live in the leader CTA (remote arrives / TMA
// complete_tx from the peer), `empty` and `tmem_full` are multicast by tcgen05.commit to both CTAs.
// -------------------------------------------------------------------------------------------
constexpr int STAGE2_BYTES = 2 * A_BYTES;          // A 128x128 + B half 128x128 = 32 KiB
constexpr int NUM_STAGES2 = 6;                     // 192 KiB
constexpr int SMEM2_BYTES = NUM_STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address (-> leader CTA)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst_smem, const CUtensorMap* map, uint32_t leader_bar, int32_t c0,
                                                int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// arrive (+ expect_tx) on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t bar, uint32_t cta, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %2;\n\t}"
        ::"r"(bar), "r"(cta), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// s8 x s8 -> s32, A MN-major, B K-major, M = 256 (two CTAs), N = 256
__host__ __device__ constexpr uint32_t instr_desc_i8_2sm() {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_ozaki_i8_2cta_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NUM_STAGES2 * STAGE2_BYTES;
    const uint32_t bar_full = bar_base;                        // used in the leader CTA only
    const uint32_t bar_empty = bar_full + 8 * NUM_STAGES2;     // per CTA
    const uint32_t bar_tfull = bar_empty + 8 * NUM_STAGES2;    // per CTA
    const uint32_t bar_tempty = bar_tfull + 8 * NUM_ACC;       // used in the leader CTA only
    const uint32_t tmem_slot = bar_tempty + 8 * NUM_ACC;
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = p.tiles_m * p.tiles_n;               // 256 x 256 cluster tiles
    const int num_kb = (p.K + BK - 1) / BK;
    const int s = p.s;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NUM_STAGES2; ++i) { mbar_init(bar_full + 8 * i, 2); mbar_init(bar_empty + 8 * i, 1); }
        for (int a = 0; a < NUM_ACC; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8); }
        fence_barrier_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) {
        tmem_alloc_2sm(tmem_slot, NUM_ACC * BN);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                int tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
                const int m0 = tm * 256 + 128 * (int)rank, n0 = tn * 256 + 128 * (int)rank;
                for (int d = s + 1; d >= 2; --d) {
                    const int t_lo = max(1, d - s), t_hi = min(s, d - 1);
                    for (int ta = t_lo; ta <= t_hi; ++ta) {
                        const int ub = d - ta;
                        for (int kb = 0; kb < num_kb; ++kb) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t full_local = bar_full + 8 * stage;
                            const uint32_t sA = smem_base + stage * STAGE2_BYTES;
                            mbar_arrive_expect_tx_remote(full_local, 0, STAGE2_BYTES);
                            tma_load_3d_2sm(sA, &mapA, full_local & PEER_MASK, m0, kb * BK, ta - 1);
                            tma_load_3d_2sm(sA + A_BYTES, &mapB, full_local & PEER_MASK, kb * BK, n0, ub - 1);
                            if (++stage == NUM_STAGES2) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = instr_desc_i8_2sm();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                for (int d = s + 1; d >= 2; --d) {
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    const int npairs = min(s, d - 1) - max(1, d - s) + 1;
                    uint32_t first = 1;
                    for (int pr = 0; pr < npairs; ++pr) {
                        for (int kb = 0; kb < num_kb; ++kb) {
                            mbar_wait(bar_full + 8 * stage, phase);
                            tc_fence_after();
                            const uint32_t sA = smem_base + stage * STAGE2_BYTES;
                            const uint32_t sB = sA + A_BYTES;
#pragma unroll
                            for (int k = 0; k < BK / 32; ++k) {
                                const uint64_t adesc = smem_desc(sA + k * 4096, 0, 1024);
                                const uint64_t bdesc = smem_desc(sB + k * 32, 0, 1024);
                                umma_i8_2sm(d_tmem, adesc, bdesc, idesc, first ? 0u : 1u);
                                first = 0;
                            }
                            umma_commit_2sm(bar_empty + 8 * stage);      // frees this stage in BOTH CTAs
                            if (++stage == NUM_STAGES2) { stage = 0; phase ^= 1; }
                        }
                    }
                    umma_commit_2sm(bar_tfull + 8 * acc);                 // accumulator ready in BOTH CTAs
                    if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int P2 = 2 * p.P;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            int tm, tn;
            tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
            const int m = tm * 256 + 128 * (int)rank + quarter * 32 + lane;
            const int n0 = tn * 256;
            const bool m_ok = m < p.M;
            const int ea = m_ok ? __ldg(p.eA + m) : 0;
            double* crow = p.C + m;
            for (int d = s + 1; d >= 2; --d) {
                const bool init = (d == s + 1) && !p.accumulate;
                const int ex_m = ea - P2 + p.bits * (2 * s - d);
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * acc, 0);   // leader's barrier: 4 warps x 2 CTAs
                    }
                    if (m_ok) fold_group_chunk(v, crow, p.ldc, n0 + c * 32, p.N, p.eB, ex_m, init);
                }
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();          // nobody leaves (or frees TMEM) while the pair may still touch its smem / barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, NUM_ACC * BN);
    }
}

// ------------------------------------------------------------------------------------------- operand splitting
__global__ void ozaki_rowmax_kernel(const double* __restrict__ A, long long lda, int M, int K, int kchunk,
                                    unsigned long long* __restrict__ rowmax_bits) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int k0 = blockIdx.y * kchunk, k1 = min(K, k0 + kchunk);
    double r = 0.0;
    for (int k = k0; k < k1; ++k) r = fmax(r, fabs(A[m + (long long)k * lda]));
    // non-negative doubles order like their bit patterns; NaN/Inf propagate as a huge max (caller's data problem)
    atomicMax(rowmax_bits + m, (unsigned long long)__double_as_longlong(r));
}

__device__ __forceinline__ int scale_exponent(double r) {
    // smallest e with r < 2^e (r > 0); 0 for an all-zero row/column
    if (!(r > 0.0)) return 0;
    return ilogb(r) + 1;
}

// digits of X = rint(a * 2^(P - e)), most significant first: out[t] for t = 0..s-1, each in [-64, 64]
__device__ __forceinline__ void split_digits(double a, int e, int s, int bits, int P, signed char* digs) {
    long long X = __double2ll_rn(ldexp(a, P - e));
    const long long half = 1ll << (bits - 1), mask = (1ll << bits) - 1;
    for (int t = s - 1; t >= 1; --t) {
        const long long dgt = ((X + half) & mask) - half;
        digs[t] = (signed char)dgt;
        X = (X - dgt) >> bits;
    }
    digs[0] = (signed char)X;          // |X| <= 64 by the choice of P
}

// A (M x K, column-major) -> planes A8[t][k * ld8 + m]
__global__ void ozaki_split_a_kernel(const double* __restrict__ A, long long lda, int M, int K, int s, int bits, int P,
                                     const unsigned long long* __restrict__ rowmax_bits, signed char* __restrict__ A8,
                                     long long ld8, long long plane, int* __restrict__ eA) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int e = scale_exponent(__longlong_as_double((long long)rowmax_bits[m]));
    if (blockIdx.y == 0) eA[m] = e;
    const int kchunk = (K + gridDim.y - 1) / gridDim.y;
    const int k0 = blockIdx.y * kchunk, k1 = min(K, k0 + kchunk);
    signed char digs[8];
    for (int k = k0; k < k1; ++k) {
        split_digits(A[m + (long long)k * lda], e, s, bits, P, digs);
        for (int t = 0; t < s; ++t) A8[t * plane + (long long)k * ld8 + m] = digs[t];
    }
}

// B (K x N, column-major): one block per column -> planes B8[u][j * ld8 + k] (k contiguous)
__global__ void __launch_bounds__(256) ozaki_split_b_kernel(const double* __restrict__ B, long long ldb, int K, int N, int s,
                                                           int bits, int P,
                                                           signed char* __restrict__ B8, long long ld8, long long plane,
                                                           int* __restrict__ eB) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    const double* col = B + (long long)j * ldb;
    double r = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) r = fmax(r, fabs(col[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = r;
    __syncthreads();
    r = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = fmax(r, red[w]);
    const int e = scale_exponent(r);
    if (threadIdx.x == 0) eB[j] = e;
    signed char digs[8];
    for (int k = threadIdx.x; k < K; k += 256) {
        split_digits(col[k], e, s, bits, P, digs);
        for (int t = 0; t < s; ++t) B8[t * plane + (long long)j * ld8 + k] = digs[t];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

bool make_map_i8_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                    uint64_t stride2, uint32_t b0, uint32_t b1) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1, stride2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

size_t ozaki_workspace_bytes(int M, int N, int K, int s) {
    const size_t ldA = round_up((size_t)M, 16), ldB = round_up((size_t)K, 16);
    return round_up(ldA * K * s, 256) + round_up(ldB * N * s, 256) + round_up((size_t)M * 8, 256) + round_up((size_t)M * 4, 256) +
           round_up((size_t)N * 4, 256);
}

bool ozaki_supported(int M, int N, int K, int s, int bits) {
    // int32 accumulation of up to s pairs of K products of magnitude <= 2^(2*(bits-1)) must stay below 2^31
    return s >= 2 && s <= 8 && (bits == 7 || bits == 8) && M > 0 && N > 0 && K > 0 &&
           (long long)K * s * (1ll << (2 * (bits - 1))) < (1ll << 31) && get_encode_fn() != nullptr;
}

cudaError_t gemm_f64_ozaki(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
                           long long ldc, bool accumulate, int s, int bits, void* workspace, int num_sms, cudaStream_t stream,
                           int* launches) {
    if (!ozaki_supported(M, N, K, s, bits)) return cudaErrorNotSupported;
    const int P = bits * s - (bits == 8 ? 2 : 1);
    const size_t ldA = round_up((size_t)M, 16), ldB = round_up((size_t)K, 16);
    const size_t planeA = ldA * K, planeB = ldB * N;
    char* w = static_cast<char*>(workspace);
    signed char* A8 = reinterpret_cast<signed char*>(w);
    w += round_up(planeA * s, 256);
    signed char* B8 = reinterpret_cast<signed char*>(w);
    w += round_up(planeB * s, 256);
    unsigned long long* rowmax = reinterpret_cast<unsigned long long*>(w);
    w += round_up((size_t)M * 8, 256);
    int* eA = reinterpret_cast<int*>(w);
    w += round_up((size_t)M * 4, 256);
    int* eB = reinterpret_cast<int*>(w);

    cudaError_t e = cudaMemsetAsync(rowmax, 0, (size_t)M * 8, stream);
    if (e != cudaSuccess) return e;
    const int ksplit = max(1, min(64, K / 256));
    const int kchunk = (K + ksplit - 1) / ksplit;
    dim3 grid_a((M + 127) / 128, ksplit);
    ozaki_rowmax_kernel<<<grid_a, 128, 0, stream>>>(A, lda, M, K, kchunk, rowmax);
    ozaki_split_a_kernel<<<grid_a, 128, 0, stream>>>(A, lda, M, K, s, bits, P, rowmax, A8, (long long)ldA, (long long)planeA, eA);
    ozaki_split_b_kernel<<<N, 256, 0, stream>>>(B, ldb, K, N, s, bits, P, B8, (long long)ldB, (long long)planeB, eB);
    if (launches) *launches += 3;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;

    CUtensorMap mA, mB;
    if (!make_map_i8_3d(&mA, A8, M, K, s, ldA, planeA, 128, 128)) return cudaErrorNotSupported;
    if (!make_map_i8_3d(&mB, B8, K, N, s, ldB, planeB, 128, 256)) return cudaErrorNotSupported;
    Params p;
    p.M = M; p.N = N; p.K = K; p.s = s; p.bits = bits; p.P = P; p.C = C; p.ldc = ldc; p.eA = eA; p.eB = eB;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    p.accumulate = accumulate ? 1 : 0;
    static int use_2cta = -1;
    if (use_2cta < 0) {
        const char* env = getenv("MARLIN_B200_TC_2CTA");
        use_2cta = (env && env[0] == '1') ? 1 : 0;     // measured: no gain (not L2-bound), so opt-in
    }
    if (use_2cta && M > 128 && N > 128) {
        // cluster pairs: 256 x 256 tiles, B split across the two CTAs (box 128 k x 128 n)
        if (!make_map_i8_3d(&mB, B8, K, N, s, ldB, planeB, 128, 128)) return cudaErrorNotSupported;
        p.tiles_m = (M + 255) / 256;
        p.tiles_n = (N + 255) / 256;
        static std::atomic<bool> attr2_done{false};   // idempotent attribute set; atomic so threads sharing a context may race here
        if (!attr2_done.load(std::memory_order_acquire)) {
            e = cudaFuncSetAttribute(gemm_ozaki_i8_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
            if (e != cudaSuccess) return e;
            attr2_done.store(true, std::memory_order_release);
        }
        const int clusters = min(p.tiles_m * p.tiles_n, num_sms / 2);
        gemm_ozaki_i8_2cta_kernel<<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(mA, mB, p);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    static std::atomic<bool> attr_done{false};   // idempotent attribute set; atomic so threads sharing a context may race here
    if (!attr_done.load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(gemm_ozaki_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done.store(true, std::memory_order_release);
    }
    const int grid = min(p.tiles_m * p.tiles_n, num_sms);
    gemm_ozaki_i8_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(mA, mB, p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace mb
