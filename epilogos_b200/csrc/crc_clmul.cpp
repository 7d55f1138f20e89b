// See crc_clmul.h.
#include "crc_clmul.h"
#include <zlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#define EPI_CLMUL 1
#else
#define EPI_CLMUL 0
#endif
namespace epi {
#if EPI_CLMUL
bool crc32_clmul_available() {
    static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}
// CRC-32 (reflected 0xEDB88320) by carry-less multiplication: the message is folded 64 bytes at a time into four 128-bit
// registers (x^(512+-32) mod P), these into one (x^(128+-32) mod P), and the last 16 bytes + tail go through zlib's table
// code: folding keeps "register ++ rest of the message" congruent to the original message, so no Barrett step is needed.
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(uint32_t crc, const uint8_t* p, size_t n) {
    const __m128i k1k2 = _mm_set_epi64x(0x00000001c6e41596ll, 0x0000000154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x00000000ccaa009ell, 0x00000001751997d0ll);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0));
    __m128i x2 = _mm_loadu_si128((const __m128i*)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32));
    __m128i x4 = _mm_loadu_si128((const __m128i*)(p + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)~crc));
    p += 64;
    n -= 64;
    while (n >= 64) {
        __m128i t1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), t2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        __m128i t3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), t4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, t1), _mm_loadu_si128((const __m128i*)(p + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, t2), _mm_loadu_si128((const __m128i*)(p + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, t3), _mm_loadu_si128((const __m128i*)(p + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, t4), _mm_loadu_si128((const __m128i*)(p + 48)));
        p += 64;
        n -= 64;
    }
    // four registers -> one
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x2);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x3);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x4);
    while (n >= 16) {
        t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, t), _mm_loadu_si128((const __m128i*)p));
        p += 16;
        n -= 16;
    }
    uint8_t last[16];
    _mm_storeu_si128((__m128i*)last, x1);
    uLong c = crc32(0xFFFFFFFFul, last, 16);      // raw register 0 in, the register's bytes, then the tail
    if (n) c = crc32(c, p, (uInt)n);
    return (uint32_t)c;
}
uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n) {
    if (n >= 256 && crc32_clmul_available()) {
        while (n > (1u << 30)) {
            crc = crc32_clmul(crc, p, 1u << 30);
            p += 1u << 30;
            n -= 1u << 30;
        }
        if (n >= 64) return crc32_clmul(crc, p, n);
    }
    while (n > (1u << 30)) {
        crc = (uint32_t)crc32(crc, p, 1u << 30);
        p += 1u << 30;
        n -= 1u << 30;
    }
    return (uint32_t)crc32(crc, p, (uInt)n);
}
#else
bool crc32_clmul_available() { return false; }
uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n) {
    while (n > (1u << 30)) {
        crc = (uint32_t)crc32(crc, p, 1u << 30);
        p += 1u << 30;
        n -= 1u << 30;
    }
    return (uint32_t)crc32(crc, p, (uInt)n);
}
#endif
}
