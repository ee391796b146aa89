// bf16 block GEMM on the 5th-gen tensor cores — the bf16 flavour of SubMatrix.multiply
// (reference: matrix/SubMatrix.scala:87-91) for BASELINE config 5 (65536^2 bf16, 4x4 grid).
//
//   C (M x N, column-major, fp32 or bf16) = op(A) * op(B) [+ C],  A/B bf16, fp32 accumulation in TMEM.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0        TMA producer : cp.async.bulk.tensor.2d tiles (SWIZZLE_128B) into a 4-stage smem ring
//   warp 1        MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x 256 x 16),
//                                tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..5    epilogue     : tcgen05.ld (32 lanes x 32 columns) TMEM -> registers -> coalesced global stores
// The accumulator is double buffered in TMEM (2 x 256 columns = all 512), so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//
// Column-major operands map onto UMMA majors without any data movement:
//   A 'N' (M x K, m contiguous)  -> MN-major A : smem atoms [64 k-rows][64 m] (TMA box 64x64), LBO = 8 KiB, SBO = 1 KiB
//   A 'T' (K x M, k contiguous)  -> K-major  A : smem [128 m-rows][64 k]      (TMA box 64x128),            SBO = 1 KiB
//   B 'N' (K x N, k contiguous)  -> K-major  B : smem [256 n-rows][64 k]      (TMA box 64x256)
//   B 'T' (N x K, n contiguous)  -> MN-major B : 4 atoms [64 k-rows][64 n]
#include "gemm_bf16.h"
#include "ptx.cuh"
#include <atomic>
#include <cuda_bf16.h>

namespace mb {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2;             // 16 KiB
constexpr int B_BYTES = BN * BK * 2;             // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KiB
constexpr int NUM_STAGES = 4;                    // 192 KiB
constexpr int NUM_ACC = 2;                       // TMEM accumulator stages (2 x 256 columns)
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + 1024 + 256;
constexpr int BAND = 16;

constexpr int MAX_SEG = 8;
// K segments: C = sum over seg of op(A_seg) * op(B_seg).  One segment for a plain GEMM; k of them when the kk-sum of a
// blocked multiply (matrix/BlockMatrix.scala:177) is folded into ONE launch, so the fp32 accumulator never leaves TMEM.
struct SegMaps {
    CUtensorMap a[MAX_SEG];
    CUtensorMap b[MAX_SEG];
    int num_kb[MAX_SEG];
    int nseg;
};

struct Params {
    int M, N, K;
    void* C;
    long long ldc;
    int tiles_m, tiles_n;
    int c_is_f32;
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

// ---- tcgen05 wrappers ----
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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;     // LayoutType::SWIZZLE_128B
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t instr_desc(bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((uint32_t)(BN >> 3) << 17)     // n_dim
         | ((uint32_t)(BM >> 4) << 24);    // m_dim
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ SegMaps maps, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + NUM_STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar_base;                       // [NUM_STAGES]
    const uint32_t bar_empty = bar_full + 8 * NUM_STAGES;     // [NUM_STAGES]
    const uint32_t bar_tfull = bar_empty + 8 * NUM_STAGES;    // [NUM_ACC]
    const uint32_t bar_tempty = bar_tfull + 8 * NUM_ACC;      // [NUM_ACC]
    const uint32_t tmem_slot = bar_tempty + 8 * NUM_ACC;      // u32
    const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    int num_kb = 0;                     // k-slabs of all segments
    for (int sg = 0; sg < maps.nseg; ++sg) num_kb += maps.num_kb[sg];

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < NUM_ACC; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
        fence_barrier_init();
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b[0]);
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int sg = 0; sg < maps.nseg; ++sg) {
                    const CUtensorMap* mapA = &maps.a[sg];
                    const CUtensorMap* mapB = &maps.b[sg];
                    const int nkb = maps.num_kb[sg];
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t full = bar_full + 8 * stage;
                        const uint32_t sA = smem_base + stage * STAGE_BYTES;
                        const uint32_t sB = sA + A_BYTES;
                        mbar_arrive_expect_tx(full, STAGE_BYTES);
                        const int k0 = kb * BK;
                        if (!TA) {
                            tma_load_2d(sA, mapA, full, m0, k0);
                            tma_load_2d(sA + 8192, mapA, full, m0 + 64, k0);
                        } else {
                            tma_load_2d(sA, mapA, full, k0, m0);
                        }
                        if (!TB) {
                            tma_load_2d(sB, mapB, full, k0, n0);
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; ++b) tma_load_2d(sB + b * 8192, mapB, full, n0 + 64 * b, k0);
                        }
                        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(!TA, TB);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sA = smem_base + stage * STAGE_BYTES;
                    const uint32_t sB = sA + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // MN-major: 16 k-rows of 128 B per UMMA_K; K-major: 32 B along the swizzled 128 B row
                        const uint64_t adesc = !TA ? smem_desc(sA + k * 2048, 8192, 1024) : smem_desc(sA + k * 32, 0, 1024);
                        const uint64_t bdesc = TB ? smem_desc(sB + k * 2048, 8192, 1024) : smem_desc(sB + k * 32, 0, 1024);
                        umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(bar_empty + 8 * stage);             // smem slot free once these MMAs retire
                    if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * acc);                   // accumulator complete
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (TMEM lanes 32*(warp%4) .. +31) =====================
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int tm, tn;
            tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
            const int m = tm * BM + quarter * 32 + lane;
            const int n0 = tn * BN;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
            const bool m_ok = m < p.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                if (c == BN / 32 - 1) {
                    // all TMEM reads of this accumulator are done: hand it back before the global stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                }
                if (m_ok) {
                    if (p.c_is_f32) {
                        float* crow = static_cast<float*>(p.C) + m;
                        float old[32];
                        if (p.accumulate) {      // all 32 loads first: one memory latency per chunk, not 32 dependent ones
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = n0 + c * 32 + j;
                                old[j] = (n < p.N) ? __ldcg(crow + (long long)n * p.ldc) : 0.f;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + c * 32 + j;
                            if (n < p.N) {
                                float x = __uint_as_float(v[j]);
                                if (p.accumulate) x += old[j];
                                crow[(long long)n * p.ldc] = x;
                            }
                        }
                    } else {
                        __nv_bfloat16* crow = static_cast<__nv_bfloat16*>(p.C) + m;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + c * 32 + j;
                            if (n < p.N) {
                                float x = __uint_as_float(v[j]);
                                __nv_bfloat16* dst = crow + (long long)n * p.ldc;
                                if (p.accumulate) x += __bfloat162float(*dst);
                                *dst = __float2bfloat16_rn(x);
                            }
                        }
                    }
                }
            }
            if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, NUM_ACC * BN);
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

bool make_map_bf16(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0,
                   uint32_t box1) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool TA, bool TB>
cudaError_t launch(const SegMaps& maps, const Params& p, int num_sms, cudaStream_t stream) {
    auto kern = gemm_bf16_tcgen05_kernel<TA, TB>;
    static std::atomic<bool> attr_done{false};   // idempotent attribute set; atomic so threads sharing a context may race here
    if (!attr_done.load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done.store(true, std::memory_order_release);
    }
    const int grid = min(p.tiles_m * p.tiles_n, num_sms);
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(maps, p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t gemm_bf16_segments(bool transA, bool transB, int M, int N, int nseg, const int* Kseg, const void* const* A,
                               const long long* lda, const void* const* B, const long long* ldb, void* C, long long ldc,
                               bool c_is_f32, bool accumulate, int num_sms, cudaStream_t stream, int* launches) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (nseg <= 0 || nseg > MAX_SEG || !get_encode_fn()) return cudaErrorNotSupported;
    static thread_local SegMaps maps;
    maps.nseg = nseg;
    int Ktot = 0;
    for (int sg = 0; sg < nseg; ++sg) {
        const int K = Kseg[sg];
        if (K <= 0) return cudaErrorNotSupported;
        const bool aligned = ((reinterpret_cast<uintptr_t>(A[sg]) | reinterpret_cast<uintptr_t>(B[sg])) & 15) == 0 &&
                             (lda[sg] % 8 == 0) && (ldb[sg] % 8 == 0);
        if (!aligned) return cudaErrorNotSupported;
        bool ok = true;
        ok = ok && (!transA ? make_map_bf16(&maps.a[sg], A[sg], M, K, lda[sg], 64, 64) : make_map_bf16(&maps.a[sg], A[sg], K, M, lda[sg], 64, 128));
        ok = ok && (!transB ? make_map_bf16(&maps.b[sg], B[sg], K, N, ldb[sg], 64, 256) : make_map_bf16(&maps.b[sg], B[sg], N, K, ldb[sg], 64, 64));
        if (!ok) return cudaErrorNotSupported;
        maps.num_kb[sg] = (K + BK - 1) / BK;
        Ktot += K;
    }
    Params p;
    p.M = M; p.N = N; p.K = Ktot; p.C = C; p.ldc = ldc;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    p.c_is_f32 = c_is_f32 ? 1 : 0;
    p.accumulate = accumulate ? 1 : 0;
    if (launches) ++*launches;
    if (!transA && !transB) return launch<false, false>(maps, p, num_sms, stream);
    if (transA && !transB) return launch<true, false>(maps, p, num_sms, stream);
    if (!transA && transB) return launch<false, true>(maps, p, num_sms, stream);
    return launch<true, true>(maps, p, num_sms, stream);
}

cudaError_t gemm_bf16(bool transA, bool transB, int M, int N, int K, const void* A, long long lda, const void* B,
                      long long ldb, void* C, long long ldc, bool c_is_f32, bool accumulate, int num_sms,
                      cudaStream_t stream, int* launches) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (K <= 0) return accumulate ? cudaSuccess : cudaErrorNotSupported;
    return gemm_bf16_segments(transA, transB, M, N, 1, &K, &A, &lda, &B, &ldb, C, ldc, c_is_f32, accumulate, num_sms, stream,
                              launches);
}

}  // namespace mb
