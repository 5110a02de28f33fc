// The distance pass of the K-means++ initialiser (KPP::init, ML/Clustering.cpp:39-59) on the resident points.
//
// The reference recomputes, for every new centroid n, the minimum over the n centroids chosen so far of
// (x_i - c_k).squaredNorm() for every point: O(N n D) per draw, O(N K^2 D) per initialisation.  A minimum does not
// depend on the order of its arguments, so the same values come out of keeping `nearest_i` and folding in only the
// newest centroid: O(N D) per draw.  The kernel is a stream: 8 D bytes of points and 16 bytes of `nearest` per point
// against 3 D flops, i.e. HBM-bound (DESIGN.md "Initialiser passes").  The draw itself stays on the host.
#include <algorithm>

#include "internal.h"

namespace mlb {

constexpr int kSeedTile = 128;     // points per tile, one per thread
constexpr int kSeedThreads = 128;

__device__ __forceinline__ void seed_cp_async8(void* smem, const void* gmem)
{
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a), "l"(gmem));
}

inline size_t seed_smem_bytes(int d) { return sizeof(double) * (static_cast<size_t>(kSeedTile) * (d | 1) + d); }

// nearest_i <- min(nearest_i, |x_i - c|^2), the squared norm summed over the dimensions in order with fused
// multiply-adds, which is how the reference's scalar loop is compiled (Clustering.cpp:47; the K-means refinement in
// kmeans.cu evaluates KMeans.cpp:158 the same way).  A tile of 128 points is staged through shared memory with
// coalesced asynchronous copies (rows padded to an odd stride, so the per-thread row walks hit distinct banks).
__global__ void __launch_bounds__(kSeedThreads) kpp_nearest_kernel(const double* __restrict__ x, long long n, int d, const double* __restrict__ centroid,
                                                                   double* __restrict__ nearest, int first)
{
    extern __shared__ __align__(16) double sm[];
    const int XS = d | 1;
    double* X = sm;                    // [kSeedTile][XS]
    double* c = X + kSeedTile * XS;    // [d]
    const int tid = threadIdx.x;
    for (int l = tid; l < d; l += kSeedThreads) c[l] = centroid[l];
    const FastDiv by_d(d);
    const long long ntiles = (n + kSeedTile - 1) / kSeedTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * kSeedTile;
        const int nvalid = static_cast<int>(n - p0 < kSeedTile ? n - p0 : kSeedTile);
        const int nel = nvalid * d;
        const double* xg = x + p0 * d;
        const double old = (!first && tid < nvalid) ? nearest[p0 + tid] : INFINITY;
        __syncthreads();   // the previous tile has been consumed (first pass: c is visible)
        // The whole tile is requested at once (asynchronous 8-byte copies, no register staging): one memory latency per
        // tile instead of one per batch of loads, which is what bounded the first version at D = 64 (49 % of the HBM peak).
        for (int e = tid; e < nel; e += kSeedThreads) {
            const int pt = by_d.div(e);
            seed_cp_async8(X + pt * XS + (e - pt * d), xg + e);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (tid < nvalid) {
            const double* row = X + tid * XS;
            double s = 0.0;
#pragma unroll 8
            for (int l = 0; l < d; ++l) {
                const double t = row[l] - c[l];
                s = fma(t, t, s);
            }
            nearest[p0 + tid] = s < old ? s : old;   // std::min(nearest, s)
        }
    }
}

}  // namespace mlb

using namespace mlb;

extern "C" {

int mlb_data_kpp_update(mlb_data* data, const double* centroid, int first, double* nearest_out)
{
    MLB_ENTER(data ? data->ctx : nullptr);
    MLB_REQUIRE(data && centroid, "mlb_data_kpp_update: null argument");
    const int d = data->d;
    MLB_REQUIRE(d <= 128, "mlb_data_kpp_update: D=%d not supported by this build (D <= 128)", d);
    mlb_ctx* ctx = data->ctx;
    const size_t smem = seed_smem_bytes(d);
    MLB_TRY(for_each_gpu(ctx, [&](int g, Gpu& gpu) -> int {
        DataShard& sh = data->shards[g];
        if (!sh.nearest) {
            MLB_REQUIRE(first, "mlb_data_kpp_update: the first call on a data set must pass first != 0");
            MLB_CUDA(cudaMallocFromPoolAsync(&sh.nearest, sizeof(double) * std::max<int64_t>(1, sh.n()), gpu.pool, gpu.stream));
            MLB_CUDA(cudaMallocFromPoolAsync(&sh.seed_centroid, sizeof(double) * d, gpu.pool, gpu.stream));
            MLB_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(kpp_nearest_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        }
        MLB_CUDA(cudaMemcpyAsync(sh.seed_centroid, centroid, sizeof(double) * d, cudaMemcpyHostToDevice, gpu.stream));
        if (sh.n() > 0) {
            int per_sm = 0, sms = 0;
            MLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(kpp_nearest_kernel), kSeedThreads, smem));
            MLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gpu.device));
            MLB_REQUIRE(per_sm >= 1, "mlb_data_kpp_update: kernel does not fit on an SM (D=%d)", d);
            const long long ntiles = (sh.n() + kSeedTile - 1) / kSeedTile;
            const unsigned grid = static_cast<unsigned>(std::min<long long>(ntiles, static_cast<long long>(per_sm) * sms));
            kpp_nearest_kernel<<<grid, kSeedThreads, smem, gpu.stream>>>(sh.x, sh.n(), d, sh.seed_centroid, sh.nearest, first ? 1 : 0);
            MLB_CUDA(cudaGetLastError());
            ++data->launches;
        }
        return MLB_OK;
    }));
    if (!nearest_out) return MLB_OK;   // stream-ordered: the pass is complete before any later download or pass
    const int64_t host_begin = ctx->rank_mode ? data->shards[0].begin : 0;
    MLB_TRY(for_each_gpu(ctx, [&](int g, Gpu& gpu) -> int {
        const DataShard& sh = data->shards[g];
        if (sh.n() > 0) MLB_TRY(staged_d2h(gpu, nearest_out + (sh.begin - host_begin), sh.nearest, sizeof(double) * sh.n()));
        return MLB_OK;
    }));
    return mlb_ctx_synchronize(ctx);
}

int mlb_data_launch_count(const mlb_data* data, int64_t* launches)
{
    MLB_REQUIRE(data && launches, "mlb_data_launch_count: null argument");
    *launches = data->launches;
    return MLB_OK;
}

}  // extern "C"
