// epi_single_host: the whole S1/S2 path for one in-memory matrix held in HOST memory -- what
// expected.main -> expectedCombination.main -> scores.main compute (run.py:196, 231, 246), without the
// TSV / gzip I/O.  H2D copies of the matrix are chunked on a copy stream and overlapped with the count
// kernel; the score kernel runs chunk by chunk so its D2H copies overlap as well.
#include <mutex>

#include "common.cuh"

namespace epi {

struct HostPipe {
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, counted[2] = {nullptr, nullptr}, scored[2] = {nullptr, nullptr};
    int8_t* xbuf[2] = {nullptr, nullptr};
    size_t xbuf_bytes = 0;
    uint16_t* cnt = nullptr;
    size_t cnt_bytes = 0;
    float* scores[2] = {nullptr, nullptr};
    size_t score_bytes = 0;
    int64_t* tables = nullptr;     // n (int64 K*K) followed by exp (float K*K)
    int device = -1;
};

static HostPipe g_pipe;
static std::mutex g_pipe_mutex;

static int ensure(void** p, size_t* have, size_t want) {
    if (*have >= want) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    EPI_CUDA(cudaMalloc(p, want));
    *have = want;
    return 0;
}

static int pipe_init(HostPipe& hp) {
    int dev = 0;
    EPI_CUDA(cudaGetDevice(&dev));
    if (hp.device == dev) return 0;
    EPI_REQUIRE(hp.device == -1, "epi_single_host was first used on device %d; it keeps one workspace per process",
                hp.device);
    EPI_CUDA(cudaStreamCreateWithFlags(&hp.copy, cudaStreamNonBlocking));
    EPI_CUDA(cudaStreamCreateWithFlags(&hp.compute, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        EPI_CUDA(cudaEventCreateWithFlags(&hp.copied[i], cudaEventDisableTiming));
        EPI_CUDA(cudaEventCreateWithFlags(&hp.counted[i], cudaEventDisableTiming));
        EPI_CUDA(cudaEventCreateWithFlags(&hp.scored[i], cudaEventDisableTiming));
    }
    EPI_CUDA(cudaMalloc(reinterpret_cast<void**>(&hp.tables),
                        (size_t)EPI_MAX_STATES * EPI_MAX_STATES * (sizeof(int64_t) + sizeof(float))));
    hp.device = dev;
    return 0;
}

}  // namespace epi

using namespace epi;

extern "C" int epi_single_host(const int8_t* x_host, int64_t bins, int32_t cols, int64_t pitch, int32_t K,
                               int32_t saliency, int64_t* counts_host, float* exp_host, float* scores_host) {
    if (check_device()) return 3;
    EPI_REQUIRE(saliency == 1 || saliency == 2, "epi_single_host handles saliency 1 and 2 (got %d)", saliency);
    EPI_REQUIRE(bins >= 1 && bins < (1ll << 31), "bins=%lld out of range", (long long)bins);
    EPI_REQUIRE(cols >= (saliency == 2 ? 2 : 1) && cols <= 65535, "cols=%d out of range", cols);
    EPI_REQUIRE(pitch >= cols, "pitch=%lld smaller than cols=%d", (long long)pitch, cols);
    EPI_REQUIRE(K >= 1 && K <= EPI_MAX_STATES, "num_states=%d out of range [1, %d]", K, EPI_MAX_STATES);
    EPI_REQUIRE(x_host != nullptr, "null matrix pointer");

    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    HostPipe& hp = g_pipe;
    if (int rc = pipe_init(hp)) return rc;

    const int64_t dpitch = ((int64_t)cols + 15) & ~15ll;
    int64_t chunk = (int64_t)(192ll << 20) / dpitch;           // ~192 MB of matrix per chunk
    chunk = (chunk / 4096) * 4096;
    if (chunk < 4096) chunk = 4096;
    if (chunk > bins) chunk = bins;
    const int64_t nchunks = (bins + chunk - 1) / chunk;
    const int ntab = saliency == 1 ? K : K * K;

    size_t xb = hp.xbuf_bytes;
    if (int rc = ensure(reinterpret_cast<void**>(&hp.xbuf[0]), &xb, (size_t)(chunk * dpitch))) return rc;
    if (int rc = ensure(reinterpret_cast<void**>(&hp.xbuf[1]), &hp.xbuf_bytes, (size_t)(chunk * dpitch))) return rc;
    if (int rc = ensure(reinterpret_cast<void**>(&hp.cnt), &hp.cnt_bytes, (size_t)bins * K * 2 + 16)) return rc;
    if (scores_host != nullptr) {
        size_t sb = hp.score_bytes;
        if (int rc = ensure(reinterpret_cast<void**>(&hp.scores[0]), &sb, (size_t)chunk * K * 4)) return rc;
        if (int rc = ensure(reinterpret_cast<void**>(&hp.scores[1]), &hp.score_bytes, (size_t)chunk * K * 4)) return rc;
    }
    int64_t* n_dev = hp.tables;
    float* e_dev = reinterpret_cast<float*>(hp.tables + EPI_MAX_STATES * EPI_MAX_STATES);

    // ---- pass 1: stream the matrix in, count ----
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const int64_t lo = c * chunk;
        const int64_t nb = (bins - lo) < chunk ? (bins - lo) : chunk;
        if (c >= 2) EPI_CUDA(cudaStreamWaitEvent(hp.copy, hp.counted[b], 0));
        if (pitch == dpitch)   // already in the device layout: one contiguous DMA (pad bytes travel, never read)
            EPI_CUDA(cudaMemcpyAsync(hp.xbuf[b], x_host + lo * pitch, (size_t)(nb * pitch), cudaMemcpyHostToDevice,
                                     hp.copy));
        else
            EPI_CUDA(cudaMemcpy2DAsync(hp.xbuf[b], (size_t)dpitch, x_host + lo * pitch, (size_t)pitch, (size_t)cols,
                                       (size_t)nb, cudaMemcpyHostToDevice, hp.copy));
        EPI_CUDA(cudaEventRecord(hp.copied[b], hp.copy));
        EPI_CUDA(cudaStreamWaitEvent(hp.compute, hp.copied[b], 0));
        if (int rc = epi_bin_counts(hp.xbuf[b], nb, cols, dpitch, K, hp.cnt + lo * K, hp.compute)) return rc;
        EPI_CUDA(cudaEventRecord(hp.counted[b], hp.compute));
    }
    // ---- expected table ----
    EPI_CUDA(cudaMemsetAsync(n_dev, 0, (size_t)ntab * 8, hp.compute));
    if (int rc = epi_expected_s1s2(hp.cnt, bins, K, cols, saliency == 1 ? n_dev : nullptr,
                                   saliency == 2 ? n_dev : nullptr, hp.compute))
        return rc;
    if (int rc = epi_normalize_i64(n_dev, ntab, e_dev, hp.compute)) return rc;
    if (counts_host)
        EPI_CUDA(cudaMemcpyAsync(counts_host, n_dev, (size_t)ntab * 8, cudaMemcpyDeviceToHost, hp.compute));
    if (exp_host) EPI_CUDA(cudaMemcpyAsync(exp_host, e_dev, (size_t)ntab * 4, cudaMemcpyDeviceToHost, hp.compute));
    // ---- pass 2: scores, D2H overlapped ----
    if (scores_host != nullptr) {
        for (int64_t c = 0; c < nchunks; ++c) {
            const int b = (int)(c & 1);
            const int64_t lo = c * chunk;
            const int64_t nb = (bins - lo) < chunk ? (bins - lo) : chunk;
            if (c >= 2) EPI_CUDA(cudaStreamWaitEvent(hp.compute, hp.copied[b], 0));
            int rc;
            if (saliency == 1)
                rc = epi_scores_s1(hp.cnt + lo * K, nb, K, cols, e_dev, hp.scores[b], nullptr, EPI_SCORE_TABLE, hp.compute);
            else
                rc = epi_scores_s2(hp.cnt + lo * K, nb, K, cols, (int64_t)cols * (cols - 1), e_dev, hp.scores[b],
                                   nullptr, EPI_SCORE_TABLE, hp.compute);
            if (rc) return rc;
            EPI_CUDA(cudaEventRecord(hp.scored[b], hp.compute));
            EPI_CUDA(cudaStreamWaitEvent(hp.copy, hp.scored[b], 0));
            EPI_CUDA(cudaMemcpyAsync(scores_host + lo * K, hp.scores[b], (size_t)nb * K * 4, cudaMemcpyDeviceToHost,
                                     hp.copy));
            EPI_CUDA(cudaEventRecord(hp.copied[b], hp.copy));
        }
    }
    EPI_CUDA(cudaStreamSynchronize(hp.compute));
    EPI_CUDA(cudaStreamSynchronize(hp.copy));
    return 0;
}
