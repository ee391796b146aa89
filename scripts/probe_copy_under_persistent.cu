// Which copies make progress while a persistent kernel holds every SM (1 CTA / SM, ~197 KB shared memory, ~64.5 K
// registers, like the grouped DMMA GEMM) and spins on a flag that only the work queued BEHIND the copy can set?
// The answer decides how the multi-GPU engine may move operand bands while its GEMM is already resident (csrc/dist.cu).
//   nvcc -arch=sm_100a -O3 -o scripts/bin/probe_copy scripts/probe_copy_under_persistent.cu && scripts/bin/probe_copy
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(384, 1) spinner(volatile unsigned long long* flag, unsigned long long want, long long timeout_ns,
                                                  int* timed_out, double* sink) {
    extern __shared__ unsigned char smem[];
    // keep ~150 registers alive so the CTA's register footprint resembles the GEMM's
    double acc[72];
#pragma unroll
    for (int i = 0; i < 72; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int to = 0;
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= want) break;
#pragma unroll
        for (int i = 0; i < 72; ++i) acc[i] = acc[i] * 1.0000001 + 1e-9;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if ((long long)(now - t0) > timeout_ns) { to = 1; break; }
        __nanosleep(200);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 72; ++i) s += acc[i];
    if (s == 12345.678) sink[0] = s + smem[threadIdx.x];
    if (to && threadIdx.x == 0) atomicAdd(timed_out, 1);
}
__global__ void set_flag(unsigned long long* flag, unsigned long long v) {
    __threadfence_system();
    *flag = v;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t rows = 8192, cols = 8192, band = 2048;            // a 512 MiB fp64 tile, 128 MiB row / column bands
    double *h = nullptr, *d0 = nullptr, *d1 = nullptr, *sink = nullptr;
    CK(cudaMallocHost(&h, rows * cols * 8));
    memset(h, 0, rows * cols * 8);
    CK(cudaMalloc(&d0, rows * cols * 8));
    CK(cudaMalloc(&d1, rows * cols * 8));
    CK(cudaMalloc(&sink, 8));
    unsigned long long* flag = nullptr;
    int* timed_out = nullptr;
    CK(cudaMalloc(&flag, 8));
    CK(cudaMalloc(&timed_out, 4));
    cudaStream_t s0, s1;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    const int smem = 197 * 1024;
    CK(cudaFuncSetAttribute(spinner, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, spinner));
    printf("spinner: %d regs/thread, %d SMs, %d KB dynamic smem\n", fa.numRegs, sms, smem / 1024);
    const char* names[] = {"flag kernel only", "H2D 1D pinned (column band)", "H2D 2D pinned (row band, 16 KiB runs)", "D2D 1D", "D2D 2D (row band)",
                           "D2H 2D (sub-block)", "D2H 1D", "memset 1D"};
    unsigned long long epoch = 0;
    for (int t = 0; t < 8; ++t) {
        ++epoch;
        CK(cudaMemset(timed_out, 0, 4));
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        spinner<<<sms, 384, smem, s0>>>(flag, epoch, 3000000000ll, timed_out, sink);      // 3 s budget
        CK(cudaGetLastError());
        CK(cudaEventRecord(e0, s1));
        switch (t) {
            case 0: break;
            case 1: CK(cudaMemcpyAsync(d0, h, band * rows * 8, cudaMemcpyHostToDevice, s1)); break;
            case 2: CK(cudaMemcpy2DAsync(d0, rows * 8, h, rows * 8, band * 8, cols, cudaMemcpyHostToDevice, s1)); break;
            case 3: CK(cudaMemcpyAsync(d1, d0, band * rows * 8, cudaMemcpyDeviceToDevice, s1)); break;
            case 4: CK(cudaMemcpy2DAsync(d1, rows * 8, d0, rows * 8, band * 8, cols, cudaMemcpyDeviceToDevice, s1)); break;
            case 5: CK(cudaMemcpy2DAsync(h, rows * 8, d0, rows * 8, band * 8, band, cudaMemcpyDeviceToHost, s1)); break;
            case 6: CK(cudaMemcpyAsync(h, d0, band * rows * 8, cudaMemcpyDeviceToHost, s1)); break;
            case 7: CK(cudaMemsetAsync(d1, 0, band * rows * 8, s1)); break;
        }
        set_flag<<<1, 1, 0, s1>>>(flag, epoch);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
        int to = 0;
        float ms = 0;
        CK(cudaMemcpy(&to, timed_out, 4, cudaMemcpyDeviceToHost));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-40s : %s  (copy + flag %.2f ms%s)\n", names[t], to ? "BLOCKED until the persistent kernel gave up" : "ran concurrently", ms,
               t >= 1 && t <= 7 && !to ? "" : "");
    }
    return 0;
}
