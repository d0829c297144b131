// microbench.cu -- measured FP64 peaks of the device the library runs on.
// MEASURED_PEAKS.json carries HBM and bf16 figures only; the roofline of this path
// is FP64 (DFMA for the descriptor kernels, DMMA for the GPR contraction), so the
// denominators are measured here with dependent-chain-free register kernels.
#include <cuda_runtime.h>

#include "launch.cuh"

namespace gapcu {

__global__ void __launch_bounds__(512) k_peak_dfma(double *out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 1.0 + 1e-9 * (threadIdx.x + q);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = fma(a[q], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += a[q];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) k_peak_dmma(double *out, int iters, double av, double bv) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; q++) c[q][0] = c[q][1] = 0.0;
    const double a = av + 1e-9 * threadIdx.x, b = bv;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += c[q][0] + c[q][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

void launch_fp64_peaks(cudaStream_t st, double *dfma_tflops, double *dmma_tflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 4, threads = 512, iters = 8192;
    double *out = nullptr;
    cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0.f;
    double best_f = 0.0, best_m = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, st);
        k_peak_dfma<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep && tf > best_f) best_f = tf;
        cudaEventRecord(e0, st);
        k_peak_dmma<<<blocks, threads, 0, st>>>(out, iters, 1.0, 1e-3);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tm = 512.0 * 8.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (rep && tm > best_m) best_m = tm;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *dfma_tflops = best_f;
    *dmma_tflops = best_m;
}

}  // namespace gapcu
