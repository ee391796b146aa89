// Microbenchmark: issue rate of tcgen05.mma kind::f16 (bf16) vs kind::i8 on the attached GPU, operands resident in
// shared memory (random bytes), no TMA traffic.  One CTA per SM, 128x256xK per instruction, 4 descriptors per "slab".
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND, int AMN>   // KIND 0 = f16(bf16), 1 = i8 ; AMN = 1: A operand MN-major
__global__ void __launch_bounds__(128, 1) umma_kernel(int iters, const uint8_t* seed) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = seed[(i * 7 + blockIdx.x * 13) & 65535];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 32) {
        // K-major operands: A [128 rows][128 B], B [256 rows][128 B], SW128
        const uint32_t idesc = (KIND == 0 ? ((1u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24))
                                          : ((2u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24))) | ((uint32_t)AMN << 15);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // MN-major A: 128 B rows along M, UMMA_K k-rows per instruction (16 for bf16 -> 2 KiB, 32 for i8 -> 4 KiB)
                const uint64_t ad = AMN ? smem_desc(base + k * (KIND == 0 ? 2048 : 4096), 8192, 1024) : smem_desc(base + k * 32, 0, 1024);
                const uint64_t bd = smem_desc(base + 16384 + k * 32, 0, 1024);
                const uint32_t acc = (it | k) ? 1u : 0u;
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint8_t* seed;
    cudaMalloc(&seed, 65536);
    uint8_t* h = (uint8_t*)malloc(65536);
    for (int i = 0; i < 65536; ++i) h[i] = (uint8_t)(rand() & 0x3f);      // small positive bf16 / int8 values, no inf/nan
    cudaMemcpy(seed, h, 65536, cudaMemcpyHostToDevice);
    const int smem = 50 * 1024;
    cudaFuncSetAttribute(umma_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(umma_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(umma_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(umma_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int kk = 0; kk < 4; ++kk) {
        const int kind = kk & 1, amn = kk >> 1;
        for (int iters : {20000, 200000}) {
            cudaEventRecord(e0);
            if (kk == 0) umma_kernel<0, 0><<<sms, 128, smem>>>(iters, seed);
            else if (kk == 1) umma_kernel<1, 0><<<sms, 128, smem>>>(iters, seed);
            else if (kk == 2) umma_kernel<0, 1><<<sms, 128, smem>>>(iters, seed);
            else umma_kernel<1, 1><<<sms, 128, smem>>>(iters, seed);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double kdim = kind == 0 ? 16 : 32;
            const double ops = 2.0 * 128 * 256 * kdim * 4.0 * iters * sms;
            printf("%s A=%s iters %6d: %8.3f ms  %8.1f T%s/s  (%.1f ns per MMA)  err=%s\n", kind == 0 ? "bf16" : "int8", amn ? "MN-major" : "K-major", iters, ms,
                   ops / ms / 1e9, kind == 0 ? "FLOP" : "OP", ms * 1e6 / (4.0 * iters), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
