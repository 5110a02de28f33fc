// FP64 pipe latency / occupancy micro-benchmarks for B200 (sm_100a): how many independent DMMA
// chains per SM sub-partition are needed to saturate the FP64 unit, the dependent-issue latency of
// DMMA.8x8x4 and DFMA, and the cost of the table-driven exp used by the EM kernels next to the CUDA
// library exp().  One JSON object on stdout.  Design input for DESIGN.md "EM kernels".
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "../ml_b200/csrc/fastmath.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ITERS = 2048;

template <int CHAINS>
__global__ void k_dmma_chains(double* out, long long* cycles, double a, double b)
{
    double c[CHAINS][2];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; }
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS>
__global__ void k_dfma_chains(double* out, long long* cycles, double a, double b)
{
    double c[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) c[j] = threadIdx.x + j;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) c[j] = fma(c[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += c[j];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int WHICH>
__global__ void k_exp(double* out, double a)
{
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = -0.37 * j - 0.001 * threadIdx.x;
    double s = 0;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s += WHICH == 0 ? exp(x[j]) : mlb::exp_nonpositive(x[j]);
            x[j] = x[j] * a - 1e-3;
        }
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
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    long long* cyc; CK(cudaMalloc(&cyc, 64));
    long long h = 0;
    printf("{\"gpu\": \"%s\"", p.name);
    // dependent-issue latency: one warp, one chain
    k_dmma_chains<1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dmma_m8n8k4_dependent_cycles\": %.2f", double(h) / ITERS);
    k_dfma_chains<1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dfma_dependent_cycles\": %.2f", double(h) / ITERS);
    // one warp per sub-partition, C chains: cycles per DMMA when issue-limited by one warp
#define ONE(C) k_dmma_chains<C><<<1, 32>>>(out, cyc, 1.0000001, 1e-9); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
    printf(", \"dmma_1warp_%dchains_cycles_per_mma\": %.2f", C, double(h) / ITERS / C);
    ONE(2) ONE(4) ONE(8)
    // throughput: W warps per SM sub-partition (blocks of 128 threads = 1 warp per sub-partition each), C chains per warp
#define TP(W, C) { float ms = time_ms([&] { k_dmma_chains<C><<<sms * W, 128>>>(out, cyc, 1.0000001, 1e-9); }); \
    printf(", \"dmma_tflops_%dwarps_%dchains\": %.2f", W, C, double(sms) * W * 4 * C * ITERS * 512 / ms / 1e9); }
    TP(1, 1) TP(1, 2) TP(1, 4) TP(1, 8) TP(2, 1) TP(2, 2) TP(2, 4) TP(3, 1) TP(3, 2) TP(3, 4) TP(4, 1) TP(4, 2) TP(4, 4) TP(6, 1) TP(6, 2) TP(8, 1) TP(8, 2)
    {
        float ms = time_ms([&] { k_exp<0><<<sms * 4, 256>>>(out, 0.9999); });
        printf(", \"exp_cuda_gops\": %.1f", double(sms) * 4 * 256 * ITERS / ms / 1e6);
        ms = time_ms([&] { k_exp<1><<<sms * 4, 256>>>(out, 0.9999); });
        printf(", \"exp_table_gops\": %.1f", double(sms) * 4 * 256 * ITERS / ms / 1e6);
    }
    printf("}\n");
    return 0;
}
