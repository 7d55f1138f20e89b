// f4 -- similarity-search distance engine.
//
// Reference: similaritySearch_calc.runEuclideanDistance (similaritySearch_calc.py:67-123).  For one region of interest
// (ROI: nS reduced bins x K states) and every window w of the reduced genome (G x K):
//      D[w] = sum_{j < nS} ed[w + j][j],   ed[a][b] = max(0, ((-2 X[a].Y[b]) + XX[a]) + YY[b])
// -- sklearn's euclidean_distances(X, Y, squared=True) gathered along diagonals and summed left to right (np.sum of
// nS < 8 values) -- then half the MODE of D is the acceptance threshold and the windows are visited in increasing D.
//
//   epi_simsearch_row_norms   XX[a] = sum_k X[a][k]^2 (once per genome)
//   epi_simsearch_distances   D for a batch of ROIs: one thread per window, the window's rows are read once and reused
//                             for every ROI of the batch (terms keep the reference's order of operations; the dot
//                             product is a sequential FMA chain where the reference calls a BLAS dgemm, so values can
//                             differ from the reference's in the last bits)
//   epi_simsearch_mode_sorted the most frequent value of an ASCENDING array, smallest value among ties (scipy.stats.mode):
//                             every run end finds its run start by binary search, run lengths meet in one 64-bit atomicMax
#include "common.cuh"

namespace epi {

constexpr int SS_MAX_K = 64;          // states per reduced bin
constexpr int SS_MAX_NS = 32;         // reduced bins per window
constexpr int SS_ROI_BATCH = 8;

__global__ void __launch_bounds__(256) ss_row_norms_kernel(const double* __restrict__ x, long long rows, int K,
                                                           double* __restrict__ xx) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = fma(x[r * K + k], x[r * K + k], s);
        xx[r] = s;
    }
}

__global__ void __launch_bounds__(128) ss_distances_kernel(const double* __restrict__ genome, const double* __restrict__ xx,
                                                           long long G, int K, const double* __restrict__ rois, int R, int nS,
                                                           double* __restrict__ dist) {
    extern __shared__ double sm[];                    // [RB][nS][K] ROI values, then [RB][nS] their squared norms
    const long long W = G - nS + 1;
    const int r0 = blockIdx.y * SS_ROI_BATCH;
    const int rb = (R - r0) < SS_ROI_BATCH ? (R - r0) : SS_ROI_BATCH;
    double* yy = sm + SS_ROI_BATCH * nS * K;
    for (int i = threadIdx.x; i < rb * nS * K; i += blockDim.x) sm[i] = rois[(long long)r0 * nS * K + i];
    __syncthreads();
    for (int i = threadIdx.x; i < rb * nS; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = fma(sm[i * K + k], sm[i * K + k], s);
        yy[i] = s;
    }
    __syncthreads();
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < W; w += (long long)gridDim.x * blockDim.x) {
        double acc[SS_ROI_BATCH];
#pragma unroll
        for (int q = 0; q < SS_ROI_BATCH; ++q) acc[q] = 0.0;
        for (int j = 0; j < nS; ++j) {
            const double* g = genome + (w + j) * K;
            const double xxa = xx[w + j];
            double dot[SS_ROI_BATCH];
#pragma unroll
            for (int q = 0; q < SS_ROI_BATCH; ++q) dot[q] = 0.0;
            for (int k = 0; k < K; ++k) {
                const double gv = g[k];
#pragma unroll
                for (int q = 0; q < SS_ROI_BATCH; ++q)
                    if (q < rb) dot[q] = fma(gv, sm[(q * nS + j) * K + k], dot[q]);
            }
#pragma unroll
            for (int q = 0; q < SS_ROI_BATCH; ++q) {
                if (q < rb) {
                    double e = (-2.0 * dot[q] + xxa) + yy[q * nS + j];
                    e = e > 0.0 ? e : 0.0;                             // np.maximum(distances, 0)
                    acc[q] += e;                                      // np.sum(axis=1) of nS < 8 values: left to right
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SS_ROI_BATCH; ++q)
            if (q < rb) dist[(long long)(r0 + q) * W + w] = acc[q];
    }
}

// sorted[r][0..W) ascending.  best[r] = max over runs of (length << 32 | ~start): longest run, smallest value among ties.
__global__ void __launch_bounds__(256) ss_mode_kernel(const double* __restrict__ sorted, long long W,
                                                      unsigned long long* __restrict__ best) {
    const double* v = sorted + (long long)blockIdx.y * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < W; i += (long long)gridDim.x * blockDim.x) {
        const double x = v[i];
        if (i + 1 < W && v[i + 1] == x) continue;               // not the end of a run
        long long lo = i, hi = i;
        if (i > 0 && v[i - 1] == x) lo = 0;                      // a run of more than one value: search its start
        // first index with v[idx] == x  (v ascending); runs of one (almost every value of real distances) skip the search
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (v[mid] < x) lo = mid + 1;
            else hi = mid;
        }
        const unsigned long long len = (unsigned long long)(i - lo + 1);
        const unsigned long long key = (len << 32) | (0xffffffffull - (unsigned long long)lo);
        // almost every run is a single value and loses against the current best: look before the atomic (3 M atomics on
        // one address cost 2 ms per ROI; the racy read only ever skips keys that could not have won)
        if (key > *reinterpret_cast<volatile unsigned long long*>(&best[blockIdx.y])) atomicMax(&best[blockIdx.y], key);
    }
}

__global__ void ss_mode_value_kernel(const double* __restrict__ sorted, long long W, const unsigned long long* __restrict__ best,
                                     int R, double* __restrict__ mode, long long* __restrict__ count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const unsigned long long b = best[r];
    const long long start = (long long)(0xffffffffull - (b & 0xffffffffull));
    mode[r] = sorted[(long long)r * W + start];
    if (count != nullptr) count[r] = (long long)(b >> 32);
}

// similaritySearch_calc.py:103-123 -- one warp per ROI walks its windows in increasing distance (svals / sidx: the sorted
// distances and their window indices) and picks up to n_desired windows that do not overlap the ROI itself or an earlier
// pick (|hit - a| < nS  <=>  np.any(overlapArr[hit : hit + nS])); the first admissible window farther than mode / 2 ends
// the list with -1.  The output row starts as zeros, as the reference's array does.
__global__ void __launch_bounds__(128) ss_greedy_kernel(const double* __restrict__ svals, const long long* __restrict__ sidx,
                                                        long long W, const double* __restrict__ mode,
                                                        const long long* __restrict__ region_start, int nS, int n_desired,
                                                        int R, int* __restrict__ out) {
    const int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const double* v = svals + (long long)r * W;
    const long long* ix = sidx + (long long)r * W;
    int* row = out + (long long)r * n_desired;
    for (int i = lane; i < n_desired; i += 32) row[i] = 0;
    __syncwarp();
    const double half_mode = mode[r] / 2;
    const long long self = region_start[r];
    int found = 0;
    for (long long pos0 = 0; pos0 < W; pos0 += 32) {
        // 32 candidates per coalesced load, handed round by shuffles (the walk itself is sequential: a pick blocks later ones)
        const long long my_hit = pos0 + lane < W ? ix[pos0 + lane] : 0;
        const double my_d = pos0 + lane < W ? v[pos0 + lane] : 0.0;
        const int n = (W - pos0) < 32 ? (int)(W - pos0) : 32;
        for (int j = 0; j < n; ++j) {
            const long long hit = __shfl_sync(0xffffffffu, my_hit, j);
            const double d = __shfl_sync(0xffffffffu, my_d, j);
            bool clash = false;
            if (lane == 0) clash = llabs(hit - self) < nS;
            for (int i = lane; i < found; i += 32) clash |= llabs(hit - (long long)row[i]) < nS;
            if (__any_sync(0xffffffffu, clash)) continue;
            if (d > half_mode) {
                for (int i = found + lane; i < n_desired; i += 32) row[i] = -1;
                return;
            }
            if (lane == 0) row[found] = (int)hit;
            __syncwarp();
            if (++found >= n_desired) return;
        }
    }
}

}  // namespace epi

using namespace epi;

extern "C" int epi_simsearch_pick(const double* sorted_dev, const int64_t* index_dev, int32_t R, int64_t W,
                                  const double* mode_dev, const int64_t* region_start_dev, int32_t nS, int32_t n_desired,
                                  int32_t* out_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(R >= 0 && W >= 1 && nS >= 1 && n_desired >= 1, "bad shape");
    if (R == 0) return 0;
    EPI_REQUIRE(sorted_dev != nullptr && index_dev != nullptr && mode_dev != nullptr && region_start_dev != nullptr &&
                    out_dev != nullptr, "null pointer argument");
    ss_greedy_kernel<<<(R * 32 + 127) / 128, 128, 0, st>>>(sorted_dev, reinterpret_cast<const long long*>(index_dev), W, mode_dev,
                                                          reinterpret_cast<const long long*>(region_start_dev), nS, n_desired, R,
                                                          out_dev);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int epi_simsearch_row_norms(const double* genome_dev, int64_t G, int32_t K, double* xx_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(G >= 0 && K >= 1 && K <= SS_MAX_K, "bad reduced-genome shape %lld x %d", (long long)G, K);
    if (G == 0) return 0;
    EPI_REQUIRE(genome_dev != nullptr && xx_dev != nullptr, "null pointer argument");
    long long blocks = (G + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    ss_row_norms_kernel<<<(unsigned)blocks, 256, 0, st>>>(genome_dev, G, K, xx_dev);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int epi_simsearch_distances(const double* genome_dev, const double* xx_dev, int64_t G, int32_t K,
                                       const double* rois_dev, int32_t R, int32_t nS, double* dist_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(K >= 1 && K <= SS_MAX_K && nS >= 1 && nS <= SS_MAX_NS, "bad window shape: %d bins x %d states", nS, K);
    EPI_REQUIRE(G >= nS && G < (1ll << 32), "reduced genome of %lld bins is shorter than the window or too long", (long long)G);
    if (R == 0) return 0;
    EPI_REQUIRE(R > 0 && genome_dev != nullptr && xx_dev != nullptr && rois_dev != nullptr && dist_dev != nullptr,
                "null pointer argument");
    const long long W = G - nS + 1;
    long long bx = (W + 127) / 128;
    const long long cap = (long long)sm_count() * 8;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)((R + SS_ROI_BATCH - 1) / SS_ROI_BATCH));
    const size_t smem = (size_t)SS_ROI_BATCH * nS * (K + 1) * sizeof(double);
    EPI_CUDA(cudaFuncSetAttribute(ss_distances_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ss_distances_kernel<<<grid, 128, smem, st>>>(genome_dev, xx_dev, G, K, rois_dev, R, nS, dist_dev);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int epi_simsearch_mode_sorted(const double* sorted_dev, int32_t R, int64_t W, double* mode_dev,
                                         int64_t* count_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(R >= 0 && W >= 1 && W < (1ll << 32), "bad shape %d x %lld", R, (long long)W);
    if (R == 0) return 0;
    EPI_REQUIRE(sorted_dev != nullptr && mode_dev != nullptr, "null pointer argument");
    unsigned long long* best = nullptr;
    EPI_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&best), (size_t)R * 8, st));
    EPI_CUDA(cudaMemsetAsync(best, 0, (size_t)R * 8, st));
    long long bx = (W + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (bx > cap) bx = cap;
    ss_mode_kernel<<<dim3((unsigned)bx, (unsigned)R), 256, 0, st>>>(sorted_dev, W, best);
    ss_mode_value_kernel<<<(R + 127) / 128, 128, 0, st>>>(sorted_dev, W, best, R, mode_dev,
                                                          reinterpret_cast<long long*>(count_dev));
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(best, st);
    EPI_CUDA(e);
    return 0;
}
