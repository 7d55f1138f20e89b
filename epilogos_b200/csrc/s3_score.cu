// K6 -- S3 scores.
//
// Reference (scores.py:455-506, s3Score): scoreArrOnes = klScoreND(ones/(C(C-1)), E3) is a [C][C][K][K] table of
// pair terms T[i][j][a][c] = q log2(q / E3[i][j][a][c]), q = 1/(C(C-1)) (0 where E3 == 0), and per bin
//      score[b][s] = sum over ordered pairs i != j with x[b][j] == s of T[i][j][x[b][i]][x[b][j]]
// (np.add.at buckets by the state of the SECOND biosample of the pair).  The reference evaluates the table and
// the 693 k-term sums in float32; this kernel evaluates both in float64 -- it is held to the float64
// restatement (oracle s3_scores_f64) at 1e-9 and agrees with the reference's float32 result to the
// reference's own accumulation noise.
//
//   epi_s3_terms   E3 (float32) -> T (float64), one thread per entry.
//   epi_scores_s3  one thread per bin; the CTA's label rows sit in shared memory (row stride an odd number of
//                  words, so the per-thread row walks are bank-conflict free); j is register-blocked by 4 so
//                  each label read feeds 4 table look-ups; T is read through L1/L2 (each [i][j] block of K*K
//                  doubles is shared by all bins of the CTA).
#include "common.cuh"

namespace epi {

__global__ void __launch_bounds__(256) s3_terms_kernel(const float* __restrict__ e3, long long n, double q,
                                                       double* __restrict__ terms) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double e = (double)e3[i];
        terms[i] = (e == 0.0) ? 0.0 : q * log2(q / e);       // klScoreND masks (scores.py:550); q > 0 always
    }
}

constexpr int S3S_JB = 4;

template <int BINS>
__global__ void __launch_bounds__(BINS) s3_score_kernel(const int8_t* __restrict__ x, long long bins, int cols,
                                                        long long pitch, int K, int stride_words,
                                                        const double* __restrict__ terms, float* __restrict__ out32,
                                                        double* __restrict__ out64) {
    extern __shared__ __align__(16) uint32_t xs[];              // [BINS][stride_words]
    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * BINS;
    const int nb = (int)((bins - b0) < BINS ? (bins - b0) : BINS);
    // stage the label rows (coalesced along the row); rows beyond nb and bytes beyond cols are never read
    {
        uint8_t* xb = reinterpret_cast<uint8_t*>(xs);
        const int row_bytes = stride_words * 4;
        for (int r = 0; r < nb; ++r) {
            const int8_t* src = x + (b0 + r) * pitch;
            for (int j = tid; j < cols; j += BINS) xb[r * row_bytes + j] = (uint8_t)src[j];
        }
    }
    __syncthreads();
    if (tid >= nb) return;

    const uint32_t* row = xs + (size_t)tid * stride_words;
    const uint8_t* rowb = reinterpret_cast<const uint8_t*>(row);
    double sc[EPI_MAX_STATES];
#pragma unroll
    for (int s = 0; s < EPI_MAX_STATES; ++s) sc[s] = 0.0;
    const long long kk = (long long)K * K;

    for (int j0 = 0; j0 < cols; j0 += S3S_JB) {
        int c[S3S_JB];
        double acc[S3S_JB];
        const double* tj[S3S_JB];
#pragma unroll
        for (int jj = 0; jj < S3S_JB; ++jj) {
            const int j = j0 + jj < cols ? j0 + jj : cols - 1;
            c[jj] = rowb[j];
            acc[jj] = 0.0;
            tj[jj] = terms + (long long)j * kk + c[jj];               // + i*cols*kk + a*K below
        }
        for (int i = 0; i < cols; ++i) {
            const int a = rowb[i];
            const long long off = (long long)i * cols * kk + (long long)a * K;
#pragma unroll
            for (int jj = 0; jj < S3S_JB; ++jj) {
                const double t = __ldg(tj[jj] + off);
                if (i != j0 + jj) acc[jj] += t;
            }
        }
#pragma unroll
        for (int jj = 0; jj < S3S_JB; ++jj)
            if (j0 + jj < cols) sc[c[jj]] += acc[jj];
    }
    const long long b = b0 + tid;
    for (int s = 0; s < K; ++s) {
        if (out32 != nullptr) out32[b * K + s] = (float)sc[s];
        if (out64 != nullptr) out64[b * K + s] = sc[s];
    }
}

}  // namespace epi

using namespace epi;

extern "C" int epi_s3_terms(const float* exp3_dev, int32_t cols, int32_t K, double* terms_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(cols >= 2 && K >= 1 && K <= EPI_MAX_STATES, "bad S3 shape (needs at least 2 biosamples)");
    EPI_REQUIRE(exp3_dev != nullptr && terms_dev != nullptr, "null pointer argument");
    const long long n = (long long)cols * cols * K * K;
    const double q = 1.0 / ((double)cols * (double)(cols - 1));
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    s3_terms_kernel<<<(unsigned)blocks, 256, 0, st>>>(exp3_dev, n, q, terms_dev);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int epi_scores_s3(const int8_t* x_dev, int64_t bins, int32_t cols, int64_t pitch, int32_t K,
                             const double* terms_dev, float* out32_dev, double* out64_dev, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    EPI_REQUIRE(bins >= 0 && cols >= 2 && pitch >= cols && K >= 1 && K <= EPI_MAX_STATES, "bad S3 shape");
    if (bins == 0 || (out32_dev == nullptr && out64_dev == nullptr)) return 0;
    EPI_REQUIRE(x_dev != nullptr && terms_dev != nullptr, "null pointer argument");
    int stride_words = (cols + 3) / 4;
    if ((stride_words & 1) == 0) ++stride_words;                 // odd word stride: conflict-free row walks
    const size_t row_bytes = (size_t)stride_words * 4;
    if (row_bytes * 256 <= 220 * 1024) {
        auto kern = s3_score_kernel<256>;
        const size_t smem = row_bytes * 256;
        EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((bins + 255) / 256), 256, smem, st>>>(x_dev, bins, cols, pitch, K, stride_words, terms_dev,
                                                                out32_dev, out64_dev);
    } else {
        EPI_REQUIRE(row_bytes * 64 <= 220 * 1024, "too many biosamples (%d) for the S3 score kernel", cols);
        auto kern = s3_score_kernel<64>;
        const size_t smem = row_bytes * 64;
        EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((bins + 63) / 64), 64, smem, st>>>(x_dev, bins, cols, pitch, K, stride_words, terms_dev,
                                                             out32_dev, out64_dev);
    }
    EPI_CUDA(cudaGetLastError());
    return 0;
}
