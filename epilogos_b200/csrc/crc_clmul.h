// CRC-32 of the gzip trailer check by carry-less multiplication (csrc/crc_clmul.cpp; host code, x86-64 PCLMULQDQ with a
// run-time check and zlib's table code as the fall-back): 12.9 GB/s against zlib's 1.8 on the authoring host.  With the
// stream decoded on several threads, the CRC of every byte was a fifth of the reader's CPU time.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace epi {

bool crc32_clmul_available();
// the value zlib's crc32(crc, p, n) returns
uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n);

}  // namespace epi
