// K2 (S1/S2 expected count tables), K4 (normalise), K5 (S1/S2 KL scores) -- all functions of the per-bin
// uint16 counts produced by K1 (36 B/bin at K=18 instead of the 833 B/bin label row).
#include "common.cuh"

namespace epi {

// ================================================================================================
// K2: n1[s] += sum_b c[b][s];  n2[s][t] += sum_b c[b][s]*c[b][t] - [s==t] c[b][s]
//     (expected.py:106-113 s1Calc, expected.py:146-158 s2Calc)
// Each thread owns a TS x TS register tile of the K x K table and walks a private stream of bins of a
// shared-memory count tile; 32-bit partial sums over G bins, widened to 64 bits per thread, then one
// shared and one global atomic per entry per CTA.  Integer sums are order independent, so the result is bit-exact for any grid.
// ================================================================================================
constexpr int K2_THREADS = 256;
constexpr int K2_TILE = 512;

template <int TS, int G>
__global__ void __launch_bounds__(K2_THREADS) k2_expected_kernel(const uint16_t* __restrict__ cnt, long long bins,
                                                                 int K, unsigned long long* __restrict__ n1,
                                                                 unsigned long long* __restrict__ n2) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw);                               // K2_TILE * K
    unsigned long long* s2 = reinterpret_cast<unsigned long long*>(smem_raw + ((K2_TILE * K * 2 + 15) & ~15));
    unsigned long long* s1 = s2 + K * K;

    const int tid = threadIdx.x;
    const int NB = (K + TS - 1) / TS;
    const int per_stream = NB * NB;
    const int nstreams = K2_THREADS / per_stream;
    const int stream_id = tid / per_stream;
    const int pair = tid - stream_id * per_stream;
    const bool active = stream_id < nstreams;
    const int bi = pair / NB, bj = pair - (pair / NB) * NB;

    for (int i = tid; i < K * K + K; i += K2_THREADS) s2[i] = 0ull;   // s1 follows s2 contiguously

    unsigned long long acc[TS][TS];
    unsigned long long accd[TS];
#pragma unroll
    for (int i = 0; i < TS; ++i) {
        accd[i] = 0ull;
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = 0ull;
    }

    const long long ntiles = (bins + K2_TILE - 1) / K2_TILE;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long bin0 = t * K2_TILE;
        const int nb = (int)((bins - bin0) < K2_TILE ? (bins - bin0) : K2_TILE);
        const int nelem = nb * K;
        __syncthreads();
        {
            const uint16_t* src = cnt + bin0 * K;
            const int nvec = nelem >> 3;     // tile base is 16-byte aligned (K2_TILE*K*2 % 16 == 0)
            const int4* src16 = reinterpret_cast<const int4*>(src);
            int4* dst16 = reinterpret_cast<int4*>(tile);
            for (int i = tid; i < nvec; i += K2_THREADS) dst16[i] = __ldg(src16 + i);
            for (int i = (nvec << 3) + tid; i < nelem; i += K2_THREADS) tile[i] = src[i];
        }
        __syncthreads();
        if (active) {
            for (int b0 = stream_id; b0 < nb; b0 += nstreams * G) {
                uint32_t part[TS][TS];
                uint32_t partd[TS];
#pragma unroll
                for (int i = 0; i < TS; ++i) {
                    partd[i] = 0;
#pragma unroll
                    for (int j = 0; j < TS; ++j) part[i][j] = 0;
                }
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int b = b0 + u * nstreams;
                    if (b < nb) {
                        const uint16_t* row = tile + b * K;
                        uint32_t a[TS], c[TS];
#pragma unroll
                        for (int i = 0; i < TS; ++i) {
                            const int s = bi * TS + i, q = bj * TS + i;
                            a[i] = s < K ? row[s] : 0u;
                            c[i] = q < K ? row[q] : 0u;
                        }
#pragma unroll
                        for (int i = 0; i < TS; ++i) {
                            partd[i] += a[i];
#pragma unroll
                            for (int j = 0; j < TS; ++j) part[i][j] += a[i] * c[j];
                        }
                    }
                }
                // G products of at most width^2 each fit 32 bits (G = 8 for width <= 23170, else 1)
#pragma unroll
                for (int i = 0; i < TS; ++i) {
                    accd[i] += partd[i];
#pragma unroll
                    for (int j = 0; j < TS; ++j) acc[i][j] += part[i][j];
                }
            }
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            const int s = bi * TS + i;
#pragma unroll
            for (int j = 0; j < TS; ++j) {
                const int q = bj * TS + j;
                if (s < K && q < K) {
                    unsigned long long v = acc[i][j];
                    if (s == q) v -= accd[i];                       // c*(c-1) on the diagonal
                    atomicAdd(&s2[s * K + q], v);
                }
            }
            if (bi == bj && s < K) atomicAdd(&s1[s], accd[i]);
        }
    }
    __syncthreads();
    for (int i = tid; i < K * K; i += K2_THREADS)
        if (n2 != nullptr && s2[i] != 0ull) atomicAdd(&n2[i], s2[i]);
    for (int i = tid; i < K; i += K2_THREADS)
        if (n1 != nullptr && s1[i] != 0ull) atomicAdd(&n1[i], s1[i]);
}

// ================================================================================================
// K4: out = float( double(N) / double(sum N) )      (expectedCombination.py:42)
// ================================================================================================
__global__ void k4_sum_kernel(const long long* __restrict__ counts, long long n, unsigned long long* total) {
    long long local = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        local += counts[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ long long wsum[32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && v != 0) atomicAdd(total, (unsigned long long)v);
    }
}

__global__ void k4_divide_kernel(const long long* __restrict__ counts, long long n,
                                 const unsigned long long* __restrict__ total, float* __restrict__ out) {
    const double denom = (double)(long long)(*total);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)((double)counts[i] / denom);
}

// ================================================================================================
// K5: scores
// ================================================================================================
constexpr int K5_THREADS = 256;

template <class T>
__device__ __forceinline__ T* align16(const void* p) {
    return reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(p) + 15) & ~static_cast<uintptr_t>(15));
}

__device__ __forceinline__ double kl_direct(double obs, double e) {
    // klScoreND (scores.py:550): 0 where E == 0 (masked divide) or obs/E <= 0 (masked log2)
    if (e == 0.0 || obs == 0.0) return 0.0;
    return obs * log2(obs / e);
}

// cooperative, coalesced copy of a [nb][K] uint16 count tile into shared memory
__device__ __forceinline__ void load_count_tile(const uint16_t* __restrict__ src, uint16_t* tile, int nelem, int tid,
                                                int nthreads) {
    const int nvec = nelem >> 3;
    const int4* src16 = reinterpret_cast<const int4*>(src);
    int4* dst16 = reinterpret_cast<int4*>(tile);
    for (int i = tid; i < nvec; i += nthreads) dst16[i] = __ldg(src16 + i);
    for (int i = (nvec << 3) + tid; i < nelem; i += nthreads) tile[i] = src[i];
}

// scores staged in shared memory as [nb][K] doubles are written out as coalesced float / double rows
__device__ __forceinline__ void store_score_tile(const double* stage, int nelem, float* out32, double* out64,
                                                 long long base, int tid, int nthreads) {
    if (out32 != nullptr)
        for (int i = tid; i < nelem; i += nthreads) out32[base + i] = (float)stage[i];
    if (out64 != nullptr)
        for (int i = tid; i < nelem; i += nthreads) out64[base + i] = stage[i];
}

// ---- S1 ----------------------------------------------------------------------------------------
// score[s] = o * log2(o / E[s]),  o = c / width       (scores.py:339-344, 317)
// TABLE mode:  o from a table of correctly rounded c/width, log2(o/E) = lc[c] - log2(width) - log2(E[s]).
template <bool DIRECT, bool USE_LC>
__global__ void __launch_bounds__(K5_THREADS) k5_s1_kernel(const uint16_t* __restrict__ cnt, long long bins, int K,
                                                           int width, const float* __restrict__ exp1,
                                                           float* __restrict__ out32, double* __restrict__ out64) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);                          // K5_THREADS * K
    double* e_s = stage + K5_THREADS * K;                                         // K   (E as double)
    double* le_s = e_s + K;                                                       // K   log2(width * E)
    double* lc = le_s + K;                                                        // width + 1 (USE_LC)
    double* ot = lc + (USE_LC ? width + 1 : 0);                                   // width + 1 (USE_LC)
    uint16_t* tile = align16<uint16_t>(ot + (USE_LC ? width + 1 : 0));            // K5_THREADS * K

    const int tid = threadIdx.x;
    const double dw = (double)width;
    __shared__ int any_zero;
    if (tid == 0) any_zero = 0;
    __syncthreads();
    for (int i = tid; i < K; i += K5_THREADS) {
        const double e = (double)exp1[i];
        e_s[i] = e;
        le_s[i] = e > 0.0 ? log2(dw) + log2(e) : 0.0;
        if (!(e > 0.0)) any_zero = 1;
    }
    if (USE_LC) {
        for (int i = tid; i <= width; i += K5_THREADS) {
            lc[i] = i > 0 ? log2((double)i) : 0.0;
            ot[i] = (double)i / dw;
        }
    }
    __syncthreads();
    const bool direct = DIRECT || any_zero;

    const long long ntiles = (bins + K5_THREADS - 1) / K5_THREADS;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long bin0 = t * K5_THREADS;
        const int nb = (int)((bins - bin0) < K5_THREADS ? (bins - bin0) : K5_THREADS);
        __syncthreads();
        load_count_tile(cnt + bin0 * K, tile, nb * K, tid, K5_THREADS);
        __syncthreads();
        if (tid < nb) {
            const uint16_t* row = tile + tid * K;
            double* orow = stage + tid * K;
            for (int s = 0; s < K; ++s) {
                const int c = row[s];
                double v;
                if (direct) {
                    v = kl_direct((double)c / dw, e_s[s]);
                } else if (USE_LC) {
                    v = ot[c] * (lc[c] - le_s[s]);
                } else {
                    v = c > 0 ? ((double)c / dw) * (log2((double)c) - le_s[s]) : 0.0;
                }
                orow[s] = v;
            }
        }
        __syncthreads();
        store_score_tile(stage, nb * K, out32, out64, bin0 * K, tid, K5_THREADS);
    }
}

// ---- S2 ----------------------------------------------------------------------------------------
// score[t] = sum_s o_st log2(o_st / E[s][t]),  o_st = (c_s c_t - [s==t] c_s) / P   (scores.py:443-451, 412)
//
// TABLE mode factors the sum (L_x = log2 x, m_st = log2(P * E[s][t])):
//   off-diagonal term  = (c_s c_t / P) (L c_s + L c_t - m_st)
//   score[t] = (c_t / P) * { [A + W L c_t - (M c)_t]  -  c_t (2 L c_t - m_tt)  +  (c_t - 1)(L c_t + L(c_t - 1) - m_tt) }
//   with A = sum_s c_s L c_s, W = sum_s c_s, (M c)_t = sum_s c_s m_st -- an 18x18 mat-vec in DFMAs instead of
//   171 divide+log2 evaluations per bin.  Differs from the term-by-term float64 evaluation by rounding only
//   (|diff| <= ~1e-13 absolute for counts <= 65535; tests hold it to 1e-9 relative + 1e-12 absolute).
// DIRECT mode (and automatically whenever E has a zero entry, where the masked-term semantics matter)
//   evaluates every term with a correctly rounded divide and adds them in the reference's order s = 0..K-1.
template <int KT, bool DIRECT, bool USE_LC>
__global__ void __launch_bounds__(K5_THREADS) k5_s2_kernel(const uint16_t* __restrict__ cnt, long long bins, int K,
                                                           int width, double perms, const float* __restrict__ exp2,
                                                           float* __restrict__ out32, double* __restrict__ out64) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);                          // K5_THREADS * K
    double* m_s = stage + K5_THREADS * K;                                         // KT * KT  log2(P*E), [s][t]
    double* e_s = m_s + KT * KT;                                                  // K * K    E as double
    double* lc = e_s + K * K;                                                     // width + 1
    uint16_t* tile = align16<uint16_t>(lc + (USE_LC ? width + 1 : 0));            // K5_THREADS * K

    const int tid = threadIdx.x;
    __shared__ int any_zero;
    if (tid == 0) any_zero = 0;
    __syncthreads();
    const double lp = log2(perms);
    for (int i = tid; i < KT * KT; i += K5_THREADS) {
        const int s = i / KT, q = i - (i / KT) * KT;
        double mv = 0.0;
        if (s < K && q < K) {
            const double e = (double)exp2[s * K + q];
            e_s[s * K + q] = e;
            if (e > 0.0) mv = lp + log2(e);
            else any_zero = 1;
        }
        m_s[i] = mv;
    }
    if (USE_LC)
        for (int i = tid; i <= width; i += K5_THREADS) lc[i] = i > 0 ? log2((double)i) : 0.0;
    __syncthreads();
    const bool direct = DIRECT || any_zero;

    const long long ntiles = (bins + K5_THREADS - 1) / K5_THREADS;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long bin0 = t * K5_THREADS;
        const int nb = (int)((bins - bin0) < K5_THREADS ? (bins - bin0) : K5_THREADS);
        __syncthreads();
        load_count_tile(cnt + bin0 * K, tile, nb * K, tid, K5_THREADS);
        __syncthreads();
        if (tid < nb) {
            const uint16_t* row = tile + tid * K;
            double* orow = stage + tid * K;
            if (direct) {
                for (int q = 0; q < K; ++q) {
                    const long long cq = row[q];
                    double sum = 0.0;
                    if (cq != 0) {
                        for (int s = 0; s < K; ++s) {
                            const long long cs = row[s];
                            const long long prod = (s == q) ? cs * (cs - 1) : cs * cq;
                            if (prod != 0) sum += kl_direct((double)prod / perms, e_s[s * K + q]);
                        }
                    }
                    orow[q] = sum;
                }
            } else {
                double cd[KT], lcs[KT], mc[KT];
                double a = 0.0, w = 0.0;
#pragma unroll
                for (int s = 0; s < KT; ++s) {
                    const int c = s < K ? row[s] : 0;
                    cd[s] = (double)c;
                    lcs[s] = USE_LC ? lc[c] : (c > 0 ? log2((double)c) : 0.0);
                    a = fma(cd[s], lcs[s], a);
                    w += cd[s];
                    mc[s] = 0.0;
                }
#pragma unroll
                for (int s = 0; s < KT; ++s) {
#pragma unroll
                    for (int q = 0; q < KT; ++q) mc[q] = fma(cd[s], m_s[s * KT + q], mc[q]);
                }
#pragma unroll
                for (int q = 0; q < KT; ++q) {
                    if (q < K) {
                        const int c = row[q];
                        double v = 0.0;
                        if (c > 0) {
                            const double mqq = m_s[q * KT + q];
                            const double l1 = USE_LC ? lc[c - 1] : (c > 1 ? log2((double)(c - 1)) : 0.0);
                            double br = (a + w * lcs[q]) - mc[q];
                            br -= cd[q] * (2.0 * lcs[q] - mqq);
                            br += (cd[q] - 1.0) * (lcs[q] + l1 - mqq);
                            v = (cd[q] / perms) * br;
                        }
                        orow[q] = v;
                    }
                }
            }
        }
        __syncthreads();
        store_score_tile(stage, nb * K, out32, out64, bin0 * K, tid, K5_THREADS);
    }
}

static int persistent_grid(int64_t ntiles, int per_sm) {
    int64_t g = (int64_t)sm_count() * per_sm;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    return (int)g;
}

constexpr int LC_MAX_WIDTH = 2047;    // log2(count) table lives in shared memory up to this row width

template <bool DIRECT, bool USE_LC>
static int launch_s1(const uint16_t* cnt, int64_t bins, int K, int width, const float* e, float* o32, double* o64,
                     cudaStream_t st) {
    auto kern = k5_s1_kernel<DIRECT, USE_LC>;
    size_t smem = (size_t)K5_THREADS * K * 8 + 2 * (size_t)K * 8 + (USE_LC ? 2 * (size_t)(width + 1) * 8 : 0) +
                  (size_t)K5_THREADS * K * 2 + 16;
    EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (bins + K5_THREADS - 1) / K5_THREADS;
    kern<<<persistent_grid(ntiles, 2), K5_THREADS, smem, st>>>(cnt, bins, K, width, e, o32, o64);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

template <int KT, bool DIRECT, bool USE_LC>
static int launch_s2(const uint16_t* cnt, int64_t bins, int K, int width, int64_t perms, const float* e, float* o32,
                     double* o64, cudaStream_t st) {
    auto kern = k5_s2_kernel<KT, DIRECT, USE_LC>;
    size_t smem = (size_t)K5_THREADS * K * 8 + (size_t)KT * KT * 8 + (size_t)K * K * 8 +
                  (USE_LC ? (size_t)(width + 1) * 8 : 0) + (size_t)K5_THREADS * K * 2 + 16;
    EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (bins + K5_THREADS - 1) / K5_THREADS;
    kern<<<persistent_grid(ntiles, 2), K5_THREADS, smem, st>>>(cnt, bins, K, width, (double)perms, e, o32, o64);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

template <int KT>
static int dispatch_s2(const uint16_t* cnt, int64_t bins, int K, int width, int64_t perms, const float* e, float* o32,
                       double* o64, int mode, cudaStream_t st) {
    const bool lc = width <= LC_MAX_WIDTH;
    if (mode == EPI_SCORE_DIRECT) return launch_s2<KT, true, false>(cnt, bins, K, width, perms, e, o32, o64, st);
    if (lc) return launch_s2<KT, false, true>(cnt, bins, K, width, perms, e, o32, o64, st);
    return launch_s2<KT, false, false>(cnt, bins, K, width, perms, e, o32, o64, st);
}

}  // namespace epi

using namespace epi;

static int check_common(const void* cnt, int64_t bins, int K) {
    if (check_device()) return 3;
    EPI_REQUIRE(bins >= 0 && bins < (1ll << 31), "bins=%lld out of range", (long long)bins);
    EPI_REQUIRE(K >= 1 && K <= EPI_MAX_STATES, "num_states=%d out of range [1, %d]", K, EPI_MAX_STATES);
    EPI_REQUIRE(cnt != nullptr || bins == 0, "null count pointer");
    EPI_REQUIRE((reinterpret_cast<uintptr_t>(cnt) & 15) == 0, "cnt_dev must be 16-byte aligned");
    return 0;
}

template <int TS, int G>
static int launch_k2(const uint16_t* cnt, int64_t bins, int K, int64_t* n1, int64_t* n2, cudaStream_t st) {
    const int64_t ntiles = (bins + K2_TILE - 1) / K2_TILE;
    const size_t smem = ((size_t)(K2_TILE * K * 2 + 15) & ~(size_t)15) + (size_t)(K * K + K) * 8;
    auto kern = k2_expected_kernel<TS, G>;
    EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<persistent_grid(ntiles, 4), K2_THREADS, smem, st>>>(cnt, bins, K, reinterpret_cast<unsigned long long*>(n1),
                                                               reinterpret_cast<unsigned long long*>(n2));
    EPI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int epi_expected_s1s2(const uint16_t* cnt_dev, int64_t bins, int32_t K, int32_t width, int64_t* n1_dev,
                                 int64_t* n2_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (int rc = check_common(cnt_dev, bins, K)) return rc;
    EPI_REQUIRE(width >= 1 && width <= 65535, "width=%d out of range [1, 65535]", width);
    if (bins == 0 || (n1_dev == nullptr && n2_dev == nullptr)) return 0;
    const bool narrow = width <= 23170;     // 8 * width^2 < 2^32
    if (K <= 18) {
        return narrow ? launch_k2<3, 8>(cnt_dev, bins, K, n1_dev, n2_dev, st)
                      : launch_k2<3, 1>(cnt_dev, bins, K, n1_dev, n2_dev, st);
    }
    return narrow ? launch_k2<6, 8>(cnt_dev, bins, K, n1_dev, n2_dev, st)
                  : launch_k2<6, 1>(cnt_dev, bins, K, n1_dev, n2_dev, st);
}

extern "C" int epi_normalize_i64(const int64_t* counts_dev, int64_t n, float* out_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(n >= 0, "n=%lld negative", (long long)n);
    EPI_REQUIRE(counts_dev != nullptr && out_dev != nullptr, "null pointer argument");
    if (n == 0) return 0;
    unsigned long long* total = nullptr;
    EPI_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&total), 8, st));
    EPI_CUDA(cudaMemsetAsync(total, 0, 8, st));
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k4_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(counts_dev), n, total);
    k4_divide_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(counts_dev), n, total,
                                                       out_dev);
    EPI_CUDA(cudaGetLastError());
    EPI_CUDA(cudaFreeAsync(total, st));
    return 0;
}

extern "C" int epi_scores_s1(const uint16_t* cnt_dev, int64_t bins, int32_t K, int32_t width, const float* exp1_dev,
                             float* out32_dev, double* out64_dev, int32_t mode, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (int rc = check_common(cnt_dev, bins, K)) return rc;
    EPI_REQUIRE(width >= 1 && width <= 65535, "width=%d out of range [1, 65535]", width);
    EPI_REQUIRE(exp1_dev != nullptr, "null expected-table pointer");
    EPI_REQUIRE(mode == EPI_SCORE_TABLE || mode == EPI_SCORE_DIRECT, "unknown score mode %d", mode);
    if (bins == 0 || (out32_dev == nullptr && out64_dev == nullptr)) return 0;
    if (mode == EPI_SCORE_DIRECT) return launch_s1<true, false>(cnt_dev, bins, K, width, exp1_dev, out32_dev, out64_dev, st);
    if (width <= LC_MAX_WIDTH) return launch_s1<false, true>(cnt_dev, bins, K, width, exp1_dev, out32_dev, out64_dev, st);
    return launch_s1<false, false>(cnt_dev, bins, K, width, exp1_dev, out32_dev, out64_dev, st);
}

extern "C" int epi_scores_s2(const uint16_t* cnt_dev, int64_t bins, int32_t K, int32_t width, int64_t perms,
                             const float* exp2_dev, float* out32_dev, double* out64_dev, int32_t mode, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (int rc = check_common(cnt_dev, bins, K)) return rc;
    EPI_REQUIRE(width >= 1 && width <= 65535, "width=%d out of range [1, 65535]", width);
    EPI_REQUIRE(perms >= 1, "perms=%lld must be positive (needs at least 2 biosamples)", (long long)perms);
    EPI_REQUIRE(exp2_dev != nullptr, "null expected-table pointer");
    EPI_REQUIRE(mode == EPI_SCORE_TABLE || mode == EPI_SCORE_DIRECT, "unknown score mode %d", mode);
    if (bins == 0 || (out32_dev == nullptr && out64_dev == nullptr)) return 0;
    if (K <= 16) return dispatch_s2<16>(cnt_dev, bins, K, width, perms, exp2_dev, out32_dev, out64_dev, mode, st);
    if (K <= 18) return dispatch_s2<18>(cnt_dev, bins, K, width, perms, exp2_dev, out32_dev, out64_dev, mode, st);
    return dispatch_s2<32>(cnt_dev, bins, K, width, perms, exp2_dev, out32_dev, out64_dev, mode, st);
}
