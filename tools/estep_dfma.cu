// Head-to-head for the E-step formulation (BASELINE.json north_star): the north star describes an
// E-step that keeps per-component Cholesky factors in shared memory and evaluates every point's log-density with scalar
// FMAs (w = x - mu_k, y = L_k^-1 w by a triangular product, q = |y|^2, thread-local log-sum-exp).  The library instead
// evaluates phi(z) . theta_k on the FP64 tensor pipe (DMMA) for every shape.  This tool times the scalar formulation as
// well as it can reasonably be written (two points per thread in registers, the factor rows read as 16-byte broadcasts,
// log-densities parked in shared memory for a two-pass log-sum-exp), at C3's shape, so that the two can be compared on
// the same device: run `tools/estep_compare.py` for the DMMA side (the library's E-step-only kernel under ncu).
//
// Work per point-component pair, both formulations: D(D+1)/2 + D multiply-adds for the quadratic form (152 at D = 16).
// What differs is where the operands come from: a DFMA needs its factor entry from shared memory (one 8-byte broadcast
// per FMA and point pair here), a DMMA.8x8x4 gets 256 multiply-adds out of two registers.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o estep_dfma estep_dfma.cu ; prints one JSON line.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int D = 16, K = 32, PT = 2, THREADS = 128, TILE = THREADS * PT;
constexpr int TRI = D * (D + 1) / 2;
constexpr int CS = TRI + D + 8;   // per component: packed rows of L^-1 (row j: entries 0..j), mu, log c, padding to 16 bytes

__global__ void __launch_bounds__(THREADS, 2) estep_dfma_kernel(const double* __restrict__ x, long long n, const double* __restrict__ comp, double* __restrict__ ll_out)
{
    extern __shared__ __align__(16) double sm[];
    double* C = sm;                    // [K][CS]
    double* Q = C + K * CS;            // [K][TILE] log-densities of the tile, component-major (conflict-free per thread column)
    double* X = Q;                     // [TILE][D + 1] staging of the point tile (dead before Q is written)
    __shared__ double red[THREADS / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < K * CS; i += THREADS) C[i] = comp[i];
    double ll_acc = 0.0;
    const long long ntiles = (n + TILE - 1) / TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * TILE;
        const int nvalid = static_cast<int>(n - p0 < TILE ? n - p0 : TILE);
        __syncthreads();
        for (int e = tid; e < TILE * D; e += THREADS) {
            const int pt = e / D, dm = e % D;
            X[pt * (D + 1) + dm] = pt < nvalid ? x[p0 * D + e] : 0.0;
        }
        __syncthreads();
        double xr[PT][D];
#pragma unroll
        for (int u = 0; u < PT; ++u)
#pragma unroll
            for (int a = 0; a < D; ++a) xr[u][a] = X[(tid + u * THREADS) * (D + 1) + a];
        __syncthreads();   // X and Q share storage
        for (int k = 0; k < K; ++k) {
            const double* ck = C + k * CS;
            double w[PT][D];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const double m = ck[TRI + a];
#pragma unroll
                for (int u = 0; u < PT; ++u) w[u][a] = xr[u][a] - m;
            }
            double q[PT];
#pragma unroll
            for (int u = 0; u < PT; ++u) q[u] = 0.0;
            int off = 0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double y[PT];
#pragma unroll
                for (int u = 0; u < PT; ++u) y[u] = 0.0;
#pragma unroll
                for (int t = 0; t <= j; ++t) {
                    const double l = ck[off + t];
#pragma unroll
                    for (int u = 0; u < PT; ++u) y[u] = fma(l, w[u][t], y[u]);
                }
                off += j + 1;
#pragma unroll
                for (int u = 0; u < PT; ++u) q[u] = fma(y[u], y[u], q[u]);
            }
            const double logc = ck[TRI + D];
#pragma unroll
            for (int u = 0; u < PT; ++u) Q[k * TILE + tid + u * THREADS] = fma(-0.5, q[u], logc);
        }
        // two-pass log-sum-exp over the thread's own columns
#pragma unroll
        for (int u = 0; u < PT; ++u) {
            const int col = tid + u * THREADS;
            double mx = -INFINITY;
            for (int k = 0; k < K; ++k) mx = fmax(mx, Q[k * TILE + col]);
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += exp(Q[k * TILE + col] - mx);
            if (col < nvalid) ll_acc += mx + log(s);
        }
    }
    for (int off = 16; off >= 1; off >>= 1) ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, off);
    if ((tid & 31) == 0) red[tid >> 5] = ll_acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) s += red[w];
        ll_out[blockIdx.x] = s;
    }
}

__global__ void fill_kernel(double* x, long long n, unsigned long long seed)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    x[i] = (static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 20.0;
}

int main(int argc, char** argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 12500000ll;
    double *x = nullptr, *comp = nullptr, *ll = nullptr;
    CK(cudaMalloc(&x, sizeof(double) * n * D));
    fill_kernel<<<static_cast<unsigned>((n * D + 255) / 256), 256>>>(x, n * D, 12345ull);
    // components: means on a grid in [-10, 10], L^-1 = a well-conditioned lower-triangular matrix
    std::vector<double> host(static_cast<size_t>(K) * CS, 0.0);
    for (int k = 0; k < K; ++k) {
        double* ck = host.data() + static_cast<size_t>(k) * CS;
        int off = 0;
        for (int j = 0; j < D; ++j) {
            for (int t = 0; t <= j; ++t) ck[off + t] = t == j ? 0.8 + 0.01 * ((k + j) % 7) : 0.02 * (((k + 3 * j + 5 * t) % 11) - 5);
            off += j + 1;
        }
        for (int a = 0; a < D; ++a) ck[TRI + a] = -10.0 + 20.0 * (((k * 7 + a * 3) % 32) / 31.0);
        ck[TRI + D] = std::log(1.0 / K);
    }
    CK(cudaMalloc(&comp, sizeof(double) * host.size()));
    CK(cudaMemcpy(comp, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double) * (static_cast<size_t>(K) * CS + static_cast<size_t>(K) * TILE);   // K * TILE >= TILE * (D + 1)
    CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(estep_dfma_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(estep_dfma_kernel), THREADS, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = per_sm * sms;
    CK(cudaMalloc(&ll, sizeof(double) * grid));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        estep_dfma_kernel<<<grid, THREADS, smem>>>(x, n, comp, ll);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<double> parts(grid);
    CK(cudaMemcpy(parts.data(), ll, sizeof(double) * grid, cudaMemcpyDeviceToHost));
    double total = 0.0;
    for (double v : parts) total += v;
    const double fma_per_pair = TRI + 2.0 * D;   // triangular product, squares, differences
    printf("{\"kernel\": \"estep_dfma (Cholesky rows in shared memory, 2 points per thread)\", \"n\": %lld, \"d\": %d, \"k\": %d, \"ctas_per_sm\": %d, "
           "\"smem_bytes\": %zu, \"ms\": %.4f, \"gpoint_comp_per_s\": %.2f, \"executed_tflops\": %.2f, \"mean_log_likelihood\": %.12f}\n",
           n, D, K, per_sm, smem, best, n * static_cast<double>(K) / (best * 1e-3) / 1e9, n * static_cast<double>(K) * fma_per_pair * 2 / (best * 1e-3) / 1e12,
           total / static_cast<double>(n));
    return 0;
}
