// Bit-packed transport layout of the state matrix, for the host <-> device leg of the path.
//
// The hot path computes on the int8 bins x biosamples matrix (helpers.readStates' array narrowed to one byte per label),
// but a label of an 18-state model carries 5 bits (4 for <= 16 states): moving int8 rows over PCIe moves 848 bytes per bin
// where 528 suffice, and the end-to-end run of epi_single_host is PCIe-bound (13.1 GB at ~52 GB/s for a genome).  The
// packed layout is produced on the host by the packer (epi_pack_states_host, or by the TSV reader directly), shipped, and
// expanded on the device into the int8 pitched matrix right before the count kernel reads it (epi_unpack_states:
// ~8 GB in + 13 GB out per genome at HBM speed, a few ms against ~100 ms of PCIe time saved).
//
// Layout: a row is ceil(cols / 8) groups; a group holds 8 labels in `bits` bytes, label j in bits [bits*j, bits*(j+1)) of
// the group read as a little-endian integer; rows are `packed_pitch` bytes apart (epi_packed_pitch: groups * bits rounded
// up to 16).  Labels beyond `cols` in the last group are 0.  bits = 4 (labels < 16) or 5 (labels < 32).
#include <stdlib.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace epi {

constexpr int PB_ROWS = 32;           // rows per CTA
constexpr int PB_THREADS = 256;

__device__ __forceinline__ uint2 unpack_group(const uint8_t* __restrict__ src, int bits) {
    // src: first byte of the group in shared memory (any alignment); two aligned words cover its <= 5 bytes
    const uint32_t a = (uint32_t)(uintptr_t)src;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src - (a & 3u));
    const uint32_t sh = (a & 3u) * 8u;
    const unsigned long long v = (((unsigned long long)w[1] << 32) | w[0]) >> sh;
    uint2 o;
    if (bits == 4) {
        const uint32_t x = (uint32_t)v;
        const uint32_t lo = x & 0x0f0f0f0fu, hi = (x >> 4) & 0x0f0f0f0fu;
        o.x = __byte_perm(lo, hi, 0x5140);      // bytes: lo0 hi0 lo1 hi1
        o.y = __byte_perm(lo, hi, 0x7362);      //        lo2 hi2 lo3 hi3
    } else {
        const uint32_t x0 = (uint32_t)v & 0xfffffu, x1 = (uint32_t)(v >> 20) & 0xfffffu;
        o.x = (x0 & 0x1fu) | ((x0 & 0x3e0u) << 3) | ((x0 & 0x7c00u) << 6) | ((x0 & 0xf8000u) << 9);
        o.y = (x1 & 0x1fu) | ((x1 & 0x3e0u) << 3) | ((x1 & 0x7c00u) << 6) | ((x1 & 0xf8000u) << 9);
    }
    return o;
}

// packed [bins][ppitch] -> int8 [bins][pitch] (pitch % 16 == 0, pitch >= 8 * groups).  A CTA stages PB_ROWS packed rows
// (contiguous in memory) in shared memory with 16-byte loads, then every thread expands groups: 8-byte coalesced stores.
__global__ void __launch_bounds__(PB_THREADS) unpack_states_kernel(const uint8_t* __restrict__ packed, long long bins,
                                                                    int groups, int bits, long long ppitch,
                                                                    int8_t* __restrict__ out, long long pitch) {
    extern __shared__ __align__(16) uint8_t pb_smem[];
    const long long ntiles = (bins + PB_ROWS - 1) / PB_ROWS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * PB_ROWS;
        const int nr = (int)((bins - r0) < PB_ROWS ? (bins - r0) : PB_ROWS);
        const int nvec = (int)((nr * ppitch) >> 4);
        const uint4* src = reinterpret_cast<const uint4*>(packed + r0 * ppitch);
        uint4* dst = reinterpret_cast<uint4*>(pb_smem);
        for (int i = threadIdx.x; i < nvec; i += PB_THREADS) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) *reinterpret_cast<uint4*>(pb_smem + (size_t)nr * ppitch) = make_uint4(0u, 0u, 0u, 0u);   // slack read by the last group
        __syncthreads();
        const int total = nr * groups;
        for (int i = threadIdx.x; i < total; i += PB_THREADS) {
            const int r = i / groups, g = i - r * groups;
            const uint2 o = unpack_group(pb_smem + (size_t)r * ppitch + g * bits, bits);
            *reinterpret_cast<uint2*>(out + (r0 + r) * pitch + 8 * g) = o;
        }
        __syncthreads();
    }
}

// int8 [bins][pitch] -> packed [bins][ppitch]; one thread per group (setup / test path: not tuned)
__global__ void __launch_bounds__(256) pack_states_kernel(const int8_t* __restrict__ x, long long bins, int cols,
                                                          long long pitch, int groups, int bits, long long ppitch,
                                                          uint8_t* __restrict__ packed) {
    const long long total = bins * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int g = (int)(i - r * groups);
        unsigned long long v = 0;
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * g + j;
            const unsigned long long lab = c < cols ? (unsigned long long)(uint8_t)x[r * pitch + c] : 0ull;
            v |= (lab & ((1ull << bits) - 1)) << (bits * j);
        }
        uint8_t* dst = packed + r * ppitch + (long long)g * bits;
        for (int b = 0; b < bits; ++b) dst[b] = (uint8_t)(v >> (8 * b));
        if (g == groups - 1)
            for (long long b = (long long)groups * bits; b < ppitch; ++b) packed[r * ppitch + b] = 0;
    }
}

int unpack_states_device(const uint8_t* packed, int64_t bins, int cols, int bits, int64_t ppitch, int8_t* out,
                         int64_t pitch, cudaStream_t st) {
    const int groups = (cols + 7) / 8;
    const size_t smem = (size_t)PB_ROWS * ppitch + 16;
    EPI_REQUIRE(smem <= 200 * 1024, "packed rows of %lld bytes are too long for the unpack kernel", (long long)ppitch);
    EPI_CUDA(cudaFuncSetAttribute(unpack_states_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (bins + PB_ROWS - 1) / PB_ROWS;
    int per_sm = (int)((size_t)200 * 1024 / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    unpack_states_kernel<<<persistent_grid(ntiles, per_sm), PB_THREADS, smem, st>>>(packed, (long long)bins, groups, bits,
                                                                                   (long long)ppitch, out, (long long)pitch);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

static void pack_rows_host(const int8_t* x, int64_t lo, int64_t hi, int cols, int64_t pitch, int bits, uint8_t* out,
                           int64_t ppitch) {
    const int groups = (cols + 7) / 8;
    const int full = cols / 8;
    for (int64_t r = lo; r < hi; ++r) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(x) + r * pitch;
        uint8_t* dst = out + r * ppitch;
        for (int g = 0; g < groups; ++g) {
            unsigned long long v = 0;
            if (g < full) {
                unsigned long long q;
                memcpy(&q, src + 8 * g, 8);                       // eight labels, one per byte
                if (bits == 4) {
                    for (int j = 0; j < 8; ++j) v |= ((q >> (8 * j)) & 15ull) << (4 * j);
                } else {
                    for (int j = 0; j < 8; ++j) v |= ((q >> (8 * j)) & 31ull) << (5 * j);
                }
            } else {
                for (int j = 0; 8 * g + j < cols; ++j) v |= ((unsigned long long)src[8 * g + j] & ((1ull << bits) - 1)) << (bits * j);
            }
            for (int b = 0; b < bits; ++b) dst[g * bits + b] = (uint8_t)(v >> (8 * b));
        }
        for (int64_t b = (int64_t)groups * bits; b < ppitch; ++b) dst[b] = 0;
    }
}

}  // namespace epi

using namespace epi;

extern "C" int epi_packed_bits(int32_t num_states) { return num_states <= 16 ? 4 : 5; }

extern "C" int64_t epi_packed_pitch(int32_t cols, int32_t bits) {
    const int64_t row = (int64_t)((cols + 7) / 8) * bits;
    return (row + 15) & ~15ll;
}

static int check_pack_args(int64_t bins, int32_t cols, int32_t bits, int64_t pitch, int64_t ppitch) {
    EPI_REQUIRE(bins >= 0 && bins < (1ll << 31), "bins=%lld out of range", (long long)bins);
    EPI_REQUIRE(cols >= 1 && cols <= 65535, "cols=%d out of range [1, 65535]", cols);
    EPI_REQUIRE(bits == 4 || bits == 5, "bits=%d: the packed layout holds 4 or 5 bits per label", bits);
    EPI_REQUIRE(pitch >= (int64_t)((cols + 7) / 8) * 8 || pitch >= cols, "pitch=%lld smaller than cols=%d", (long long)pitch, cols);
    EPI_REQUIRE(ppitch >= (int64_t)((cols + 7) / 8) * bits, "packed pitch %lld too small for %d labels of %d bits",
                (long long)ppitch, cols, bits);
    return 0;
}

extern "C" int epi_unpack_states(const uint8_t* packed_dev, int64_t bins, int32_t cols, int32_t bits, int64_t packed_pitch,
                                 int8_t* x_dev, int64_t pitch, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    if (int rc = check_pack_args(bins, cols, bits, pitch, packed_pitch)) return rc;
    if (bins == 0) return 0;
    EPI_REQUIRE(packed_dev != nullptr && x_dev != nullptr, "null pointer argument");
    EPI_REQUIRE((packed_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(packed_dev) & 15) == 0,
                "the packed matrix must be 16-byte aligned with a pitch that is a multiple of 16 (epi_packed_pitch)");
    EPI_REQUIRE((pitch & 7) == 0 && (reinterpret_cast<uintptr_t>(x_dev) & 7) == 0 && pitch >= (int64_t)((cols + 7) / 8) * 8,
                "the int8 matrix must be 8-byte aligned with pitch %% 8 == 0 and room for whole groups of 8 labels");
    return unpack_states_device(packed_dev, bins, cols, bits, packed_pitch, x_dev, pitch, st);
}

extern "C" int epi_pack_states(const int8_t* x_dev, int64_t bins, int32_t cols, int64_t pitch, int32_t bits,
                               uint8_t* packed_dev, int64_t packed_pitch, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (check_device()) return 3;
    if (int rc = check_pack_args(bins, cols, bits, pitch, packed_pitch)) return rc;
    EPI_REQUIRE(pitch >= cols, "pitch=%lld smaller than cols=%d", (long long)pitch, cols);
    if (bins == 0) return 0;
    EPI_REQUIRE(packed_dev != nullptr && x_dev != nullptr, "null pointer argument");
    const int groups = (cols + 7) / 8;
    int64_t blocks = (bins * groups + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    pack_states_kernel<<<(unsigned)blocks, 256, 0, st>>>(x_dev, (long long)bins, cols, (long long)pitch, groups, bits,
                                                         (long long)packed_pitch, packed_dev);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

// Host packer (no GPU needed): what the reader side runs to produce the transport layout; `threads` <= 0 = all cores.
extern "C" int epi_pack_states_host(const int8_t* x_host, int64_t bins, int32_t cols, int64_t pitch, int32_t bits,
                                    uint8_t* packed_host, int64_t packed_pitch, int32_t threads) {
    if (int rc = check_pack_args(bins, cols, bits, pitch, packed_pitch)) return rc;
    EPI_REQUIRE(pitch >= cols, "pitch=%lld smaller than cols=%d", (long long)pitch, cols);
    if (bins == 0) return 0;
    EPI_REQUIRE(packed_host != nullptr && x_host != nullptr, "null pointer argument");
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((int64_t)nt > bins) nt = (int)bins;
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = bins * t / nt, hi = bins * (t + 1) / nt;
        pool.emplace_back(pack_rows_host, x_host, lo, hi, (int)cols, pitch, (int)bits, packed_host, packed_pitch);
    }
    for (auto& th : pool) th.join();
    return 0;
}
