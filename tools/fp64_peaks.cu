// FP64 pipe micro-benchmarks for B200 (sm_100a): DFMA chains, mma.sync f64 (DMMA) in the
// m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 shapes, FP64 exp(), and a read-only HBM stream.
// The numbers are the roofline denominators for the clustering kernels:
// MEASURED_PEAKS.json has no FP64 figure, so we measure it on the box, in the same gpurun
// call as the result it normalises.  Output: one JSON object on stdout.
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = threadIdx.x * 1e-9 + j;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per lane. 256 FMA per warp instruction.
__global__ void k_dmma884(double* out, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

// m16n8k4: A 2, B 1, C 4. 512 FMA.
__global__ void k_dmma1684(double* out, double a, double b)
{
    double c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k8: A 4, B 2, C 4. 1024 FMA.
__global__ void k_dmma1688(double* out, double a, double b)
{
    double c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8, B 4, C 4. 2048 FMA.
__global__ void k_dmma16816(double* out, double a, double b)
{
    double c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                           "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 123.456) out[0] = s;
}

__global__ void k_exp(double* out, double a)
{
    double x[4] = {-0.001 * threadIdx.x, -0.002 * threadIdx.x - 1, -3.0, -20.0 - 0.01 * threadIdx.x};
    double s = 0;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += exp(x[j]); x[j] = x[j] * a - 1e-3; }
    }
    if (s == 123.456) out[0] = s;
}

__global__ void k_read(const double2* __restrict__ in, size_t n2, double* out)
{
    double s = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v = in[i];
        s += v.x + v.y;
    }
    if (s == 123.456) out[0] = s;
}

template <class F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    return best;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    const int threads = 256;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int bps : {1, 2, 4}) {
        const int blocks = sms * bps;
        const double nthreads = (double)blocks * threads;
        const double nwarps = nthreads / 32;
        float ms;
        ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dfma_tflops_bps%d\": %.3f", bps, nthreads * 16.0 * ITERS * 2 / ms / 1e9);
        ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma_m8n8k4_tflops_bps%d\": %.3f", bps, nwarps * 8.0 * ITERS * 512 / ms / 1e9);
        ms = time_ms([&] { k_dmma1684<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma_m16n8k4_tflops_bps%d\": %.3f", bps, nwarps * 8.0 * ITERS * 1024 / ms / 1e9);
        ms = time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma_m16n8k8_tflops_bps%d\": %.3f", bps, nwarps * 8.0 * ITERS * 2048 / ms / 1e9);
        ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma_m16n8k16_tflops_bps%d\": %.3f", bps, nwarps * 8.0 * (ITERS / 2) * 4096 / ms / 1e9);
        ms = time_ms([&] { k_exp<<<blocks, threads>>>(out, 0.9999); });
        printf(", \"exp_gops_bps%d\": %.3f", bps, nthreads * ITERS / ms / 1e6);
    }
    {
        size_t bytes = (size_t)4 << 30;
        double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        float ms = time_ms([&] { k_read<<<sms * 8, 512>>>(buf, bytes / 16, out); });
        printf(", \"hbm_read_gbs\": %.1f", bytes / ms / 1e6);
        CK(cudaFree(buf));
    }
    printf("}\n");
    return 0;
}
