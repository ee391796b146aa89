// Microbenchmark: peak DMMA (mma.sync.m8n8k4.f64) and DFMA issue rate per SM on the attached GPU.
// Used once to pin the fp64 roofline denominator; not part of the product.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = a0 + threadIdx.x, b = b0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = a0 + threadIdx.x, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {1, 2, 4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<16><<<sms, warps * 32>>>(out, iters, 1.0, 2.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)sms * warps * iters * 16 * 256.0;
            if (rep) printf("DMMA warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/ns/SM)\n", warps, ms, 2 * fma / ms / 1e9, fma / ms / 1e6 / sms);
        }
    }
    for (int warps : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<16><<<sms, warps * 32>>>(out, iters, 1.0, 2.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)sms * warps * iters * 16 * 32.0;
            if (rep) printf("DFMA warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/ns/SM)\n", warps, ms, 2 * fma / ms / 1e9, fma / ms / 1e6 / sms);
        }
    }
    // sustained: ~3 s of DMMA to see the power-capped clock
    cudaEventRecord(e0);
    for (int i = 0; i < 40; ++i) dmma_kernel<16><<<sms, 8 * 32>>>(out, iters * 4, 1.0, 2.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = 40.0 * sms * 8 * iters * 4 * 16 * 256.0;
    printf("DMMA sustained (%.0f ms): %.2f TFLOP/s\n", ms, 2 * fma / ms / 1e9);
    return 0;
}
