// K5 (S2 scores, scores.py:404-421, 443-451) -- the K x K mat-vec of the TABLE evaluation on the tensor cores in
// kind::f16, second generation of the kernel in tc_tables.cu (which remains the path for 1023 < width <= 2047).
//
// What changed against the kind::i8 kernel, and why it is still EXACT integer arithmetic:
//   * A operand.  A uint16 count c <= 1023 read as an IEEE half is the subnormal c * 2^-24: the count row of a bin IS an
//     fp16 operand row, bit for bit -- no conversion, no low/high byte split (the i8 kernel needs two operand bytes per
//     count and therefore a carry column per state).
//   * B operand.  M_st = -log2 E_st in 55-bit fixed point is cut into FIVE 11-bit digits (an fp16 significand holds 11
//     bits); B[n][s] = half(16 * digit_d(M_st)) <= 32752, exact.  One more contraction row multiplies a constant 1.0 of
//     the operand row with 8.0.
//   * Accumulator.  D[b][n] = 8 + 2^-20 * N,  N = sum_s c_bs digit_d(M_st) <= 1023 * 2047 < 2^21.  Every addend is a
//     multiple of 2^-20 and every partial sum is below 16, i.e. fits the 24-bit significand of the fp32 accumulator:
//     nothing is ever rounded, and the accumulator's BIT PATTERN is 0x41000000 + N.  The tensor core delivers the
//     integer N already biased for the integer -> double "magic number" conversion: the epilogue assembles
//         L = N0 + 2^11 N1 + 2^22 N2,  H = N3 + 2^11 N4,   sum_s c_s M_st = L + 2^33 H
//     with five IMAD.WIDE per state straight from the tcgen05.ld registers (the constant parts of the five biased words
//     are folded into the 64-bit addend), against 4 IMAD + 2 IMAD.WIDE + 4 IADD3 for the eight int32 columns before.
//     (tests/test_gpu_parity.py::test_k5_tensor_core_matvec_is_exact compares L and H with integer arithmetic.)
//   * 5 accumulator columns per state instead of 8: six states share one 32-column tcgen05.ld (30 + 2 unused), a tile is
//     96 TMEM columns instead of 160, and FOUR warpgroups (tiles in flight per SM) fit instead of three.
// Everything after the mat-vec is the float64 epilogue of the i8 kernel (same tables, same order of operations).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "count_tile.cuh"

namespace epi {

constexpr int H5_BINS = 128;
constexpr int H5_A_BYTES = 128 * 128;
constexpr int H5_MAX_WIDTH = 1023;                                  // counts must be fp16 subnormals
constexpr unsigned long long H5_MAGIC = 0x4338000000000000ull;      // bits of 2^52 + 2^51
constexpr double H5_MAGIC_VALUE = 6755399441055744.0;               // 2^52 + 2^51
constexpr unsigned long long H5_BIAS = 0x41000000ull;               // bits of 8.0f: accumulator = 8 + N 2^-20

__host__ __device__ constexpr int h5_col(int t, int d) { return 32 * (t / 6) + 5 * (t % 6) + d; }
__host__ __device__ constexpr int h5_npad(int kt) { return 32 * ((kt + 5) / 6); }

// instruction descriptor, kind::f16: D = F32 (1 << 4), A = B = F16 (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_f16_idesc(uint32_t m, uint32_t n) {
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ unsigned long long h5_mad_wide(uint32_t a, uint32_t b, unsigned long long c) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

struct K5HConsts {
    double dg[EPI_MAX_STATES];                  // M_tt 2^-F / perms / 8: the [s==t] term of y_t
    double scale;                               // 2^-F / perms
    double cst;                                 // (2^52 + 2^51) * scale
    unsigned long long kl, kh;                  // 64-bit addends of the digit assembly (magic bits - folded accumulator biases)
    int has_zero;
    int debug;                                  // WANT64 kernels only: 1 -> out64 = L, 2 -> out64 = H (exactness tests)
    uint32_t m1, m11, m22;                      // 1, 2^11, 2^22 (opaque to the compiler: they must stay IMAD.WIDE operands)
    int fbits;
};
__constant__ K5HConsts c_k5h;

// one CTA: float32 expected table -> 55-bit fixed point of -log2 E, its 11-bit digits as the 128-byte-swizzled fp16 B
// operand (row n = h5_col(t, d), element s), the constant row, per-state constants, zero flag; mfix_out (optional)
// receives the fixed-point table itself (diagnostic / tests).
__global__ void k5h_prepare_kernel(const float* __restrict__ e, int K, double perms, int nrows, int debug,
                                   uint8_t* __restrict__ b_image, K5HConsts* __restrict__ out, int* __restrict__ zero_flag,
                                   unsigned long long* __restrict__ mfix_out) {
    __shared__ unsigned long long mfix[EPI_MAX_STATES * EPI_MAX_STATES];
    __shared__ double mval[EPI_MAX_STATES * EPI_MAX_STATES];
    __shared__ unsigned long long maxbits;
    __shared__ int zero, fbits;
    if (threadIdx.x == 0) {
        maxbits = 0ull;
        zero = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
        const double ev = (double)e[i];
        double m = 0.0;
        if (ev > 0.0) m = -log2(ev);
        else zero = 1;
        if (m < 0.0) m = 0.0;                      // E <= 1 always; guards -0.0
        mval[i] = m;
        atomicMax(&maxbits, (unsigned long long)__double_as_longlong(m));      // non-negative doubles order like integers
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double mx = __longlong_as_double((long long)maxbits);
        int ib = 1;
        while (ib < 12 && (double)(1ull << ib) <= mx) ++ib;       // mx < 2^ib
        fbits = 55 - ib;
    }
    __syncthreads();
    const int F = fbits;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
        unsigned long long v = (unsigned long long)llrint(ldexp(mval[i], F));
        if (v >= (1ull << 55)) v = (1ull << 55) - 1;
        mfix[i] = v;
        if (mfix_out != nullptr) mfix_out[i] = v;
    }
    for (int i = threadIdx.x; i < nrows * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(b_image)[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < K * K * 5; i += blockDim.x) {
        const int d = i % 5, st = i / 5;
        const int s = st / K, t = st - s * K;
        const uint32_t dig = (uint32_t)((mfix[st] >> (11 * d)) & 2047ull);
        const __half h = __float2half_rn((float)(dig * 16u));                   // <= 32752, 11 significant bits: exact
        *reinterpret_cast<unsigned short*>(b_image + sw128_offset((uint32_t)h5_col(t, d), (uint32_t)(2 * s))) = __half_as_ushort(h);
    }
    for (int n = threadIdx.x; n < nrows; n += blockDim.x)                       // contraction row K: 1.0 (operand) x 8.0
        *reinterpret_cast<unsigned short*>(b_image + sw128_offset((uint32_t)n, (uint32_t)(2 * K))) = 0x4800;
    if (threadIdx.x < K) out->dg[threadIdx.x] = ldexp((double)mfix[threadIdx.x * K + threadIdx.x], -F - 3) / perms;
    if (threadIdx.x == 0) {
        const double scale = ldexp(1.0, -F) / perms;
        out->scale = scale;
        out->cst = H5_MAGIC_VALUE * scale;
        out->kl = H5_MAGIC - H5_BIAS * (1ull + (1ull << 11) + (1ull << 22));    // modulo 2^64
        out->kh = H5_MAGIC - H5_BIAS * (1ull + (1ull << 11));
        out->has_zero = zero;
        out->debug = debug;
        out->m1 = 1u;
        out->m11 = 1u << 11;
        out->m22 = 1u << 22;
        out->fbits = F;
        *zero_flag = zero;
    }
}

// KT: even unroll bound on the state index; KR: the state count when known at compile time (0 = runtime K <= KT);
// NWG: warpgroups = accumulator regions in TMEM; WANT64: also write the unrounded float64 scores (tests).
template <int KT, int KR, int NWG, bool WANT64>
__global__ void __launch_bounds__(NWG * 128 + 64, 1)
k5_s2_h_kernel(const uint16_t* __restrict__ cnt, long long bins, int Krt, int width, double perms,
               const uint8_t* __restrict__ b_image, float* __restrict__ out32, double* __restrict__ out64) {
    if (c_k5h.has_zero) return;          // masked terms: the DIRECT kernel launched behind this one does the work

    static_assert(KT % 2 == 0, "KT must be even");
    constexpr int NPAD = h5_npad(KT);                        // TMEM columns per warpgroup = UMMA N
    constexpr int NCH = NPAD / 32;                           // 32-column chunks (6 states each)
    static_assert(NPAD * NWG <= 512, "accumulator regions exceed TMEM");
    const int K = KR ? KR : Krt;
    const int ksteps = (K + 16) / 16;                        // contraction rows: K counts + the constant, 16 per UMMA

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int slab_bytes = H5_BINS * K * 2;                                  // multiple of 256
    uint8_t* a_ops = smem;                                                   // NWG x 16 KB
    uint8_t* b_op = a_ops + NWG * H5_A_BYTES;                                // NPAD x 128 B
    uint8_t* slabs = b_op + NPAD * 128;                                      // NWG x 2 x slab_bytes
    float* stage = reinterpret_cast<float*>(slabs + NWG * 2 * slab_bytes);   // NWG x 128 x K
    double* g1 = reinterpret_cast<double*>(stage + NWG * H5_BINS * K);       // width + 1 entries c HG[c] (+1 pad)
    double* f1 = g1 + ((width + 2) & ~1);                                    // width + 1 entries F[c] (+1 pad)
    uint64_t* ld_full = reinterpret_cast<uint64_t*>(f1 + ((width + 2) & ~1)); // [NWG][2]
    uint64_t* ld_empty = ld_full + NWG * 2;
    uint64_t* a_full = ld_empty + NWG * 2;                                   // [NWG]
    uint64_t* mma_done = a_full + NWG;                                       // [NWG]
    uint64_t* b_full = mma_done + NWG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NT = NWG * 128 + 64;
    const long long ntiles = (bins + H5_BINS - 1) / H5_BINS;
    const long long nfull = bins / H5_BINS;

    for (int i = tid; i < NWG * H5_A_BYTES / 16; i += NT) reinterpret_cast<uint4*>(a_ops)[i] = make_uint4(0u, 0u, 0u, 0u);
    {
        // F[c] = c log2(c) / P,  HG[c] = ((log2 c - log2 P)(W - 1) - c log2 c + (c-1) log2(c-1)) / P      (scores.cu header)
        // score_t = c_t (base_t + HG[c_t]) = fma(c_t, base_t, c_t HG[c_t]): the second table holds c HG[c]
        const double lp = log2(perms), invp = 1.0 / perms, wm1 = (double)width - 1.0;
        for (int c = tid; c <= width; c += NT) {
            const double l = c > 0 ? log2((double)c) : 0.0;
            const double l1 = c > 1 ? log2((double)(c - 1)) : 0.0;
            const double cl = (double)c * l;
            f1[c] = cl * invp;
            g1[c] = (double)c * ((fma(l - lp, wm1, -cl) + (double)(c > 0 ? c - 1 : 0) * l1) * invp);
        }
    }
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int i = 0; i < NWG * 2; ++i) {
            mbar_init(&ld_full[i], 1);
            mbar_init(&ld_empty[i], 4);
        }
        for (int g = 0; g < NWG; ++g) {
            mbar_init(&a_full[g], 4);
            mbar_init(&mma_done[g], 1);
        }
        mbar_init(b_full, 1);
        mbar_fence_init();
    }
    if (warp == 4 * NWG + 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * NWG) {
        // ------------------------------ bulk-copy producer ------------------------------
        if (lane == 0) {
            mbar_expect_tx(b_full, NPAD * 128);
            bulk_load_1d(b_op, b_image, NPAD * 128, b_full);
            int j = 0;
            for (long long t = blockIdx.x; t < nfull; t += gridDim.x, ++j) {
                const int g = j % NWG, u = j / NWG, b = u & 1;
                mbar_wait_wd(&ld_empty[g * 2 + b], (((uint32_t)(u >> 1)) & 1u) ^ 1u);
                mbar_expect_tx(&ld_full[g * 2 + b], (uint32_t)slab_bytes);
                bulk_load_1d(slabs + (g * 2 + b) * slab_bytes, cnt + t * (long long)H5_BINS * K, (uint32_t)slab_bytes,
                             &ld_full[g * 2 + b]);
            }
        }
    } else if (warp == 4 * NWG + 1) {
        // ------------------------------ MMA issuer (one thread) ------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_f16_idesc(128, NPAD);
            mbar_wait_wd(b_full, 0);
            const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(b_op));
            // warpgroup g owns tiles blockIdx + (u NWG + g) gridDim, u = 0, 1, ...; the operands are served in the order in
            // which they become ready (no head-of-line blocking behind a slower warpgroup)
            int u_of[NWG];
            long long left = 0;
            for (int g = 0; g < NWG; ++g) {
                u_of[g] = 0;
                const long long first = blockIdx.x + (long long)g * gridDim.x;
                if (first < ntiles) left += (ntiles - first + (long long)NWG * gridDim.x - 1) / ((long long)NWG * gridDim.x);
            }
            uint32_t spins = 0;
            while (left > 0) {
#pragma unroll
                for (int g = 0; g < NWG; ++g) {
                    const long long t = blockIdx.x + ((long long)u_of[g] * NWG + g) * gridDim.x;
                    if (t >= ntiles || !mbar_try_wait(&a_full[g], ((uint32_t)u_of[g]) & 1u)) continue;
                    tc_fence_after();
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(a_ops + g * H5_A_BYTES));
                    for (int k = 0; k < ksteps; ++k)      // 16 halfs = 32 bytes along K inside the swizzle atom: +2
                        umma_f16(tmem_base + (uint32_t)(g * NPAD), a_desc + 2 * k, b_desc + 2 * k, idesc, k ? 1u : 0u);
                    umma_commit(&mma_done[g]);
                    ++u_of[g];
                    --left;
                    spins = 0;
                }
                if (++spins > (1u << 28)) __trap();
            }
        }
    } else {
        // ------------------------------ warpgroups: one bin per thread ------------------------------
        const int g = warp >> 2, r = tid & 127;
        uint8_t* a_row = a_ops + g * H5_A_BYTES + r * 128;
        float* mystage = stage + g * H5_BINS * K;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * NPAD);
        // everything that multiplies the count is kept divided by 8: the count enters as the double 8 c (its table byte offset)
        const double scale = c_k5h.scale * 0.125, cst = c_k5h.cst, scale33 = c_k5h.scale * (8589934592.0 * 0.125);
        int u = 0;
        for (long long t = blockIdx.x + (long long)g * gridDim.x; t < ntiles; t += (long long)NWG * gridDim.x, ++u) {
            const long long bin0 = t * H5_BINS;
            const int b = u & 1;
            uint32_t cw[KT / 2];                       // the count row as uint16 pairs (state 2i in the low half)
            if (t < nfull) {
                mbar_wait_wd(&ld_full[g * 2 + b], ((uint32_t)(u >> 1)) & 1u);
                load_count_row<KT, KR>(reinterpret_cast<const uint16_t*>(slabs + (g * 2 + b) * slab_bytes) + r * K, K, cw);
            } else {
                load_count_row_guarded<KT>(cnt + (bin0 + r) * K, K, bin0 + r < bins, cw);
            }
            // the count row's bytes ARE the fp16 (subnormal) A operand row; element K is the constant 1.0 (0x3C00)
            store_operand_row<KT>(a_row, r, cw, K / 2, 0x3C00u << (16 * (K & 1)));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[g]);
                // released only after the operand stores have CONSUMED every loaded count (see tc_tables.cu)
                if (t < nfull) mbar_arrive(&ld_empty[g * 2 + b]);
            }

            // work that does not need the accumulators: A = sum_s F[c_s]      (c = 0 for s >= K: F[0] = 0)
            double a4[4] = {0.0, 0.0, 0.0, 0.0};       // four partial sums: the adds are dependent fp64 operations
#pragma unroll
            for (int s = 0; s < KT; ++s) {
                const uint32_t off = (s & 1) ? ((cw[s >> 1] >> 13) & 0x7fff8u) : ((cw[s >> 1] & 0xffffu) << 3);
                a4[s & 3] += *reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(f1) + off);
            }
            const double a_adj = (((a4[0] + a4[1]) + (a4[2] + a4[3])) - cst) * 0.125;

            // the warp's 32 rows of the previous tile have left the staging buffer
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            mbar_wait_wd(&mma_done[g], ((uint32_t)u) & 1u);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + (uint32_t)(32 * ch), v);
                float f[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const int tt = 6 * ch + i;
                    f[i] = 0.0f;
                    if (tt < KT && (KR == 0 || tt < KR)) {
                        const uint32_t off = (tt & 1) ? ((cw[tt >> 1] >> 13) & 0x7fff8u) : ((cw[tt >> 1] & 0xffffu) << 3);
                        const double chg = *reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(g1) + off);
                        const double cd = __hiloint2double(0x43300000, (int)off) - 4503599627370496.0;      // 8 c, exactly
                        // accumulator word d = 0x41000000 + N_d; the biases are folded into kl / kh:
                        // bits(2^52 + 2^51) + L  and  bits(2^52 + 2^51) + H, five IMAD.WIDE
                        unsigned long long bl = h5_mad_wide(v[5 * i], c_k5h.m1, c_k5h.kl);
                        unsigned long long bh = h5_mad_wide(v[5 * i + 3], c_k5h.m1, c_k5h.kh);
                        bl = h5_mad_wide(v[5 * i + 1], c_k5h.m11, bl);
                        bh = h5_mad_wide(v[5 * i + 4], c_k5h.m11, bh);
                        bl = h5_mad_wide(v[5 * i + 2], c_k5h.m22, bl);
                        const double dl = __longlong_as_double((long long)bl);                          // L + magic
                        // H exactly: its magic offset cannot be folded into a_adj like that of L -- base is multiplied by 8 c and
                        // must be good to ~1e-16, while magic * scale33 ~ 2e4 would round the small terms away at 4e-12
                        const double dh = __longlong_as_double((long long)bh) - H5_MAGIC_VALUE;         // H
                        // base = (H 2^33 + L) scale + A - [s==t] term
                        const double base = fma(dh, scale33, fma(dl, scale, a_adj)) - c_k5h.dg[tt];
                        double val = fma(cd, base, chg);      // absent state: 0 * base + (+0) = +0.0 as in the reference
                        if (WANT64) {
                            if (c_k5h.debug == 1) val = dl - H5_MAGIC_VALUE;
                            if (c_k5h.debug == 2) val = dh;
                            if ((KR != 0 || tt < K) && bin0 + r < bins) out64[(bin0 + r) * K + tt] = val;
                        }
                        f[i] = (float)val;
                    }
                }
                if constexpr (KR != 0 && KR % 2 == 0) {
                    // even row length: 8-byte stores
                    float2* dst = reinterpret_cast<float2*>(mystage + r * KR + 6 * ch);
                    if (6 * ch < KR) dst[0] = make_float2(f[0], f[1]);
                    if (6 * ch + 2 < KR) dst[1] = make_float2(f[2], f[3]);
                    if (6 * ch + 4 < KR) dst[2] = make_float2(f[4], f[5]);
                } else {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
                        if (6 * ch + i < K) mystage[r * K + 6 * ch + i] = f[i];
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (out32 != nullptr) {
                const int wq = warp & 3;
                const long long wbin0 = bin0 + 32 * wq;
                if (t < nfull) {
                    if (lane == 0) {
                        bulk_store_1d(out32 + wbin0 * K, mystage + 32 * wq * K, (uint32_t)(32 * K * 4));
                        bulk_commit();
                    }
                } else if (wbin0 < bins) {
                    const int n = (int)((bins - wbin0) < 32 ? (bins - wbin0) : 32) * K;
                    for (int i = lane; i < n; i += 32) out32[wbin0 * K + i] = mystage[32 * wq * K + i];
                }
            }
        }
        if (lane == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4 * NWG + 1) tmem_dealloc(tmem_base, 512);
}

static uint8_t* k5h_workspace() {
    static uint8_t* base[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (base[dev] == nullptr && cudaMalloc(reinterpret_cast<void**>(&base[dev]), 256 * 128 + 1024) != cudaSuccess) return nullptr;
    return base[dev];
}

static int k5h_debug_mode() {
    const char* e = getenv("EPI_K5_DEBUG");
    return e ? atoi(e) : 0;
}

template <int KT, int KR, int NWG, bool WANT64>
static int launch_k5_h2(const uint16_t* cnt, int64_t bins, int K, int width, int64_t perms, const uint8_t* ws, float* o32,
                        double* o64, cudaStream_t st) {
    constexpr int NPAD = h5_npad(KT);
    if (int rc = apply_wait_hint(st)) return rc;
    auto kern = k5_s2_h_kernel<KT, KR, NWG, WANT64>;
    const size_t smem = 1024 + (size_t)NWG * H5_A_BYTES + (size_t)NPAD * 128 + (size_t)NWG * 2 * H5_BINS * K * 2 +
                        (size_t)NWG * H5_BINS * K * 4 + (size_t)(width + 2) * 16 +
                        (size_t)(6 * NWG + 1) * 8 + 16;
    EPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (bins + H5_BINS - 1) / H5_BINS;
    kern<<<persistent_grid(ntiles, 1), NWG * 128 + 64, smem, st>>>(cnt, (long long)bins, K, width, (double)perms, ws, o32, o64);
    EPI_CUDA(cudaGetLastError());
    return 0;
}

template <int KT, int KR, int NWG>
static int launch_k5_h(const uint16_t* cnt, int64_t bins, int K, int width, int64_t perms, const float* e, int* zero_flag,
                       float* o32, double* o64, cudaStream_t st) {
    constexpr int NPAD = h5_npad(KT);
    uint8_t* ws = k5h_workspace();
    EPI_REQUIRE(ws != nullptr, "could not allocate the score-table workspace");
    K5HConsts* consts = reinterpret_cast<K5HConsts*>(ws + 256 * 128);
    k5h_prepare_kernel<<<1, 256, 0, st>>>(e, K, (double)perms, NPAD, k5h_debug_mode(), ws, consts, zero_flag, nullptr);
    EPI_CUDA(cudaGetLastError());
    EPI_CUDA(cudaMemcpyToSymbolAsync(c_k5h, consts, sizeof(K5HConsts), 0, cudaMemcpyDeviceToDevice, st));
    if (o64 != nullptr) return launch_k5_h2<KT, KR, NWG, true>(cnt, bins, K, width, perms, ws, o32, o64, st);
    return launch_k5_h2<KT, KR, NWG, false>(cnt, bins, K, width, perms, ws, o32, o64, st);
}

// Opt-in (EPI_K5_F16=1): measured on B200 at 15.5 M bins this kernel runs at the speed of the kind::i8 kernel for 18 states
// (0.54 ms both) and slower for 15 states (0.50 vs 0.47 ms) -- see DESIGN.md: both are bound by the ~650 instructions per
// 32-bin warp tile of the float64 epilogue, not by TMEM capacity or resident warps (3 / 4 / 5 warpgroups: 0.59 / 0.55 / 0.54 ms).
bool scores_s2_h_eligible(int width) { return width <= H5_MAX_WIDTH && getenv("EPI_K5_F16") != nullptr; }

// TABLE evaluation of the S2 scores on the tensor cores (kind::f16 form).  Same contract as scores_s2_tc (tc_tables.cu).
int scores_s2_h(const uint16_t* cnt, int64_t bins, int K, int width, int64_t perms, const float* e, int* zero_flag,
                float* o32, double* o64, cudaStream_t st) {
    const char* wg = getenv("EPI_K5_WG");                 // tuning knob: warpgroups (= tiles in flight) per SM
    const int nwg = wg ? atoi(wg) : 4;
    if (K == 18 && nwg == 5) return launch_k5_h<18, 18, 5>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K == 18 && nwg == 3) return launch_k5_h<18, 18, 3>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K == 15 && nwg == 5) return launch_k5_h<16, 15, 5>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K == 18) return launch_k5_h<18, 18, 4>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K == 15) return launch_k5_h<16, 15, 4>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K <= 16) return launch_k5_h<16, 0, 4>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    if (K <= 18) return launch_k5_h<18, 0, 4>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
    return launch_k5_h<32, 0, 2>(cnt, bins, K, width, perms, e, zero_flag, o32, o64, st);
}

// the fixed-point image of -log2 E the kernel multiplies with (diagnostic; used by the exactness test)
int scores_s2_h_fixed_point(const float* e, int K, int64_t perms, unsigned long long* mfix_dev, int* fbits_host,
                            cudaStream_t st) {
    uint8_t* ws = k5h_workspace();
    EPI_REQUIRE(ws != nullptr, "could not allocate the score-table workspace");
    K5HConsts* consts = reinterpret_cast<K5HConsts*>(ws + 256 * 128);
    int* zero_flag = reinterpret_cast<int*>(ws + 256 * 128 + 768);
    k5h_prepare_kernel<<<1, 256, 0, st>>>(e, K, (double)perms, h5_npad((K + 1) & ~1), 0, ws, consts, zero_flag, mfix_dev);
    EPI_CUDA(cudaGetLastError());
    K5HConsts host;
    EPI_CUDA(cudaMemcpyAsync(&host, consts, sizeof(host), cudaMemcpyDeviceToHost, st));
    EPI_CUDA(cudaStreamSynchronize(st));
    if (fbits_host) *fbits_host = host.fbits;
    return 0;
}

}  // namespace epi
