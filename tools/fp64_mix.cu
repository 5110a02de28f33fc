// Does the B200 FP64 pipe lose throughput when DMMA and scalar FP64 instructions are mixed, as the EM kernels mix them
// (C2: 92 DMMA.8x8x4 and 167 DFMA/DMUL/DADD per 16-point sub-tile and warp)?  Each kernel issues, per loop iteration and
// warp, ND DMMAs on 4 independent accumulator chains and NF DFMAs on 8 independent chains; the pipe time such a mix
// needs with no penalty is 16 ND + 2 NF cycles per sub-partition and warp (256 and 32 FMAs at 16 FMA per cycle).
// Output: one JSON object; "eff" = needed pipe cycles / measured cycles.
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mix tools/fp64_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ITERS = 2048;

template <int ND, int NF, bool SPLIT_WARPS>
__global__ void k_mix(double* out, double a, double b)
{
    double c[4][2], f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = threadIdx.x * 1e-9 + j;
    const int warp = threadIdx.x >> 5;
    // SPLIT_WARPS: even warps issue only the DMMAs (twice as many), odd warps only the DFMAs (twice as many): the same
    // totals per sub-partition, but no warp mixes the two kinds.
    const bool do_d = !SPLIT_WARPS || (warp & 4) == 0, do_f = !SPLIT_WARPS || (warp & 4) != 0;
    const int rep = SPLIT_WARPS ? 2 : 1;
    for (int it = 0; it < ITERS; ++it) {
        for (int r = 0; r < rep; ++r) {
            if (do_d) {
#pragma unroll
                for (int j = 0; j < ND; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j & 3][0]), "+d"(c[j & 3][1]) : "d"(a), "d"(b));
            }
            if (do_f) {
#pragma unroll
                for (int j = 0; j < NF; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j & 7]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456) out[0] = s;
}

template <int ND, int NF, bool SPLIT>
static void run(const char* name, int sms, double clock_ghz, bool last)
{
    double* out;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int threads = 512, blocks = sms;   // 16 warps per SM, 4 per sub-partition: the occupancy of em_small_kernel
    k_mix<ND, NF, SPLIT><<<blocks, threads>>>(out, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        k_mix<ND, NF, SPLIT><<<blocks, threads>>>(out, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    // per sub-partition: 4 warps x ITERS x (16 ND + 2 NF) pipe cycles
    const double need_cycles = 4.0 * ITERS * (16.0 * ND + 2.0 * NF);
    const double got_cycles = best * 1e-3 * clock_ghz * 1e9;
    const double fma = static_cast<double>(sms) * 16 * ITERS * (256.0 * ND + 32.0 * NF);
    printf("\"%s\": {\"nd\": %d, \"nf\": %d, \"split_warps\": %s, \"ms\": %.4f, \"tflops\": %.2f, \"eff\": %.3f}%s\n", name, ND, NF, SPLIT ? "true" : "false", best,
           2 * fma / (best * 1e-3) / 1e12, need_cycles / got_cycles, last ? "" : ",");
    CK(cudaFree(out));
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f,\n", prop.name, prop.multiProcessorCount, ghz);
    run<4, 0, false>("dmma_only", prop.multiProcessorCount, ghz, false);
    run<0, 32, false>("dfma_only", prop.multiProcessorCount, ghz, false);
    run<4, 8, false>("mix_4_8", prop.multiProcessorCount, ghz, false);       // the C2 ratio (92 : 167)
    run<4, 8, true>("mix_4_8_split_warps", prop.multiProcessorCount, ghz, false);
    run<4, 2, false>("mix_4_2", prop.multiProcessorCount, ghz, false);       // the C3 ratio (DMMA 80 %, other 7 %)
    run<4, 16, false>("mix_4_16", prop.multiProcessorCount, ghz, false);
    run<1, 2, false>("mix_1_2_fine", prop.multiProcessorCount, ghz, true);   // alternating at the finest grain
    printf("}\n");
    return 0;
}
