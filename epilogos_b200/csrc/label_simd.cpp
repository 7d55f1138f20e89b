// See label_simd.h.  Why: with the gzip stream decoded on several threads the row parser is the slowest stage of reading a
// matrix (4-6 ns per label in the branchy scalar loop: the label width, one or two digits, is a coin the branch predictor
// cannot call).  Here a 16-byte window is classified at once -- tab positions, the digit in front of each tab, the digit
// (or tab) two in front -- the values label-1 are formed for all 16 positions, and the ones at tab positions are packed
// to the front with a byte shuffle taken from a 256-entry table, eight positions at a time: ~1.4 ns per label.
#include "label_simd.h"

#include <string.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define EPI_LABEL_SIMD 1
#else
#define EPI_LABEL_SIMD 0
#endif

namespace epi {

#if EPI_LABEL_SIMD

namespace {
struct CompressLut {
    alignas(16) uint8_t idx[256][8];      // byte indices of the set bits of m, packed to the front; 0x80 (-> zero) behind
    uint8_t cnt[256];
    CompressLut() {
        for (int m = 0; m < 256; ++m) {
            int n = 0;
            for (int b = 0; b < 8; ++b)
                if (m & (1 << b)) idx[m][n++] = (uint8_t)b;
            cnt[m] = (uint8_t)n;
            for (; n < 8; ++n) idx[m][n] = 0x80;
        }
    }
};
const CompressLut g_lut;
}  // namespace

bool label_simd_available() {
    static const bool ok = __builtin_cpu_supports("ssse3") != 0;
    return ok;
}

__attribute__((target("ssse3")))
int parse_labels_simd(const char* p, const char* e, int want, int num_states, int8_t* dst, const char** resume) {
    const __m128i tab = _mm_set1_epi8('\t');
    const __m128i zero = _mm_set1_epi8('0');
    const __m128i nine = _mm_set1_epi8(9);
    const __m128i ones_all = _mm_set1_epi8(-1);
    const __m128i kmax = _mm_set1_epi8((char)(num_states - 1));
    __m128i bad = _mm_setzero_si128();
    int j = 0;
    const char* w = p;
    const char* after = p;
    while (e - w >= 16 && want - j >= 16) {
        const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(w));
        const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(w - 1));
        const __m128i v2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(w - 2));
        const __m128i v3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(w - 3));
        const __m128i t0 = _mm_cmpeq_epi8(v0, tab);
        // every byte of the window is a digit or a tab
        const __m128i d0 = _mm_sub_epi8(v0, zero);
        const __m128i isdig0 = _mm_cmpeq_epi8(_mm_min_epu8(d0, nine), d0);
        bad = _mm_or_si128(bad, _mm_andnot_si128(_mm_or_si128(isdig0, t0), ones_all));
        // seen from a tab at position i: the units digit at i - 1, the tens digit (or the previous tab) at i - 2
        const __m128i units = _mm_sub_epi8(v1, zero);
        const __m128i isdig1 = _mm_cmpeq_epi8(_mm_min_epu8(units, nine), units);
        const __m128i d2 = _mm_sub_epi8(v2, zero);
        const __m128i isdig2 = _mm_cmpeq_epi8(_mm_min_epu8(d2, nine), d2);
        const __m128i t3 = _mm_cmpeq_epi8(v3, tab);
        const __m128i tens = _mm_and_si128(d2, isdig2);                         // 0..9, so 8 * tens stays inside its byte
        const __m128i tens10 = _mm_add_epi8(_mm_slli_epi16(tens, 3), _mm_add_epi8(tens, tens));
        const __m128i val = _mm_sub_epi8(_mm_add_epi8(units, tens10), _mm_set1_epi8(1));      // label - 1
        // at a tab: a digit in front of it; two digits in front need a tab in front of them; the label is in range
        __m128i wrong = _mm_andnot_si128(isdig1, ones_all);
        wrong = _mm_or_si128(wrong, _mm_andnot_si128(t3, isdig2));
        wrong = _mm_or_si128(wrong, _mm_andnot_si128(_mm_cmpeq_epi8(_mm_min_epu8(val, kmax), val), ones_all));
        bad = _mm_or_si128(bad, _mm_and_si128(wrong, t0));
        // pack the values at tab positions to the front, eight positions at a time
        const unsigned mask = (unsigned)_mm_movemask_epi8(t0);
        const unsigned lo = mask & 0xffu, hi = mask >> 8;
        const __m128i outlo = _mm_shuffle_epi8(val, _mm_loadl_epi64(reinterpret_cast<const __m128i*>(g_lut.idx[lo])));
        const __m128i outhi = _mm_shuffle_epi8(_mm_srli_si128(val, 8), _mm_loadl_epi64(reinterpret_cast<const __m128i*>(g_lut.idx[hi])));
        _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + j), outlo);      // 8 bytes, cnt[lo] of them meaningful
        j += g_lut.cnt[lo];
        _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + j), outhi);      // stays inside the row: want - j >= 16 on entry
        j += g_lut.cnt[hi];
        if (mask) after = w + (32 - __builtin_clz(mask));
        w += 16;
    }
    if (_mm_movemask_epi8(_mm_cmpeq_epi8(bad, _mm_setzero_si128())) != 0xffff) return -1;
    *resume = after;
    return j;
}

#else

bool label_simd_available() { return false; }
int parse_labels_simd(const char*, const char*, int, int, int8_t*, const char**) { return -1; }

#endif

}  // namespace epi
