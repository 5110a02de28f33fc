// Diagnostics: the FP64 pipe peaks of the device the benchmark runs on, measured live so that bench.py's roofline
// denominator is this box's, not a number carried over from another run.  Two dependent-chain micro-kernels, both
// saturating the SM's FP64 unit: DMMA.8x8x4 (mma.sync.m8n8k4.f64, the instruction the EM / K-means kernels issue)
// and DFMA.  8 independent accumulator chains per warp, 4 CTAs of 256 threads per SM.
#include <algorithm>

#include "internal.h"

namespace mlb {

constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; }
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = threadIdx.x * 1e-9 + j;
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 123.456) out[0] = s;
}

}  // namespace mlb

using namespace mlb;

extern "C" int mlb_selftest_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops)
{
    DeviceRestore restore;
    MLB_REQUIRE(dmma_tflops && dfma_tflops, "mlb_selftest_fp64_peak: null argument");
    MLB_CUDA(cudaSetDevice(device));
    int sms = 0;
    MLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    MLB_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    MLB_CUDA(cudaEventCreate(&e0));
    MLB_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 4, threads = 256;
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {   // the first repetition is the warm-up
            MLB_CUDA(cudaEventRecord(e0));
            if (which == 0) peak_dmma_kernel<<<blocks, threads>>>(out, 1.0000001, 0.9999999);
            else peak_dfma_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
            MLB_CUDA(cudaGetLastError());
            MLB_CUDA(cudaEventRecord(e1));
            MLB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            MLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double warps = static_cast<double>(blocks) * threads / 32;
            const double flops = which == 0 ? warps * 8.0 * kPeakIters * 512.0 : static_cast<double>(blocks) * threads * 16.0 * kPeakIters * 2.0;
            if (rep > 0) best[which] = std::max(best[which], flops / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *dmma_tflops = best[0];
    *dfma_tflops = best[1];
    return MLB_OK;
}
