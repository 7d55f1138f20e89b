// Parallel decoding of ONE gzip stream for the host reader (csrc/hostio.cu).  Host code.
//
// Why: after the row parser was spread over several threads, reading a gzipped state matrix is bound by the single
// DEFLATE stream (fast_inflate.h: ~0.45-0.8 GB/s of text on one core), and that read IS a file-to-file run (the kernels
// take milliseconds).  A DEFLATE stream can be decoded from any block boundary if the 32 KiB of history in front of it
// are treated as unknowns:
//   1. the compressed file is cut into chunks of equal size; for every chunk a finder looks for the first position that
//      parses as the start of a dynamic-Huffman block (complete code-length / literal / distance codes, a first block
//      that decodes without error) or of a gzip member (bgzip files and the files this library writes are multi-member);
//   2. every chunk is decoded from its start to the start of the next chunk into 16-bit symbols: a literal is its byte
//      value, a back-reference that reaches into the unknown history copies a MARKER 0x8000 + (index into that window);
//   3. the chunks are chained in order: a chunk is accepted only if its start is EXACTLY the position at which the
//      previous accepted chunk stopped (a block boundary reached by decoding from a known-good position), so a false
//      positive of the finder costs time, never correctness; the accepted chunk's window is the resolved tail of its
//      predecessor, its markers are replaced through a 64 Ki-entry table, and per-member CRC-32 / ISIZE are checked from
//      per-piece CRCs joined with crc32_combine.
// Anything unexpected (no block starts found, a decode error, a CRC mismatch) makes the caller fall back to the
// sequential decoders, which report corrupt input exactly as before.  Same idea as pugz / rapidgzip; written for this reader.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "crc_clmul.h"
#include "fast_inflate.h"

namespace epi {

constexpr uint32_t PI_WINDOW = 32768;
constexpr uint16_t PI_MARK = 0x8000;

// growable array of 16-bit symbols with PI_WINDOW marker entries in front (no zero fill on growth)
struct SymbolBuffer {
    uint16_t* p = nullptr;       // p[0 .. PI_WINDOW) = markers, symbols follow
    size_t cap = 0;              // symbols that fit behind the prefix (incl. slack)
    SymbolBuffer() = default;
    SymbolBuffer(const SymbolBuffer&) = delete;
    SymbolBuffer& operator=(const SymbolBuffer&) = delete;
    ~SymbolBuffer() { free(p); }
    bool reserve(size_t want) {
        if (want <= cap && p != nullptr) return true;
        const bool fresh = p == nullptr;
        uint16_t* q = static_cast<uint16_t*>(realloc(p, (PI_WINDOW + want) * sizeof(uint16_t)));
        if (q == nullptr) return false;
        p = q;
        cap = want;
        if (fresh)
            for (uint32_t w = 0; w < PI_WINDOW; ++w) p[w] = (uint16_t)(PI_MARK + w);
        return true;
    }
    uint16_t* syms() { return p + PI_WINDOW; }
    void release() {
        free(p);
        p = nullptr;
        cap = 0;
    }
};

class MarkerDecoder {
public:
    enum Kind : uint8_t { NONE = 0, BLOCK = 1, MEMBER = 2 };
    enum Status { STOPPED = 0, END_OF_STREAM = 1, FAILED = 2, GARBAGE = 3, ABORTED = 4 };
    struct Start {
        uint64_t pos = 0;        // bit position in the compressed file
        Kind kind = NONE;
    };
    struct MemberEnd {
        uint64_t off;            // symbols this chunk had produced when the member ended
        uint32_t crc, isize;
    };

    // [data, data + size) must be followed by readable zero bytes (the reader appends 64: a refill reads 8 bytes ahead and a
    // block header takes a few refills between its checks)
    MarkerDecoder(const uint8_t* data, size_t size) : base_(data), end_(data + size), lim_(data + size + 8) {}
    const char* error() const { return err_; }

    // RFC 1952 member header at p: its length, or -1
    static long gzip_header_len(const uint8_t* p, const uint8_t* end) {
        const uint8_t* const p0 = p;
        if (end - p < 18) return -1;
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return -1;
        const int flg = p[3];
        if (flg & 0xe0) return -1;
        p += 10;
        if (flg & 4) {
            if (end - p < 2) return -1;
            const size_t xlen = (size_t)p[0] | ((size_t)p[1] << 8);
            p += 2;
            if ((size_t)(end - p) < xlen) return -1;
            p += xlen;
        }
        for (int f = 8; f <= 16; f <<= 1)
            if (flg & f) {
                while (p < end && *p) ++p;
                if (p >= end) return -1;
                ++p;
            }
        if (flg & 2) p += 2;
        if (p > end) return -1;
        return (long)(p - p0);
    }

    // First plausible start in bytes [lo, hi): a gzip member header or a non-final dynamic-Huffman block header whose
    // first block decodes cleanly.  `scratch` is reused between calls.
    Start find_start(size_t lo, size_t hi, SymbolBuffer& scratch) {
        Start none;
        const size_t size = (size_t)(end_ - base_);
        if (hi > size) hi = size;
        for (size_t b = lo; b < hi; ++b) {
            const uint8_t* q = base_ + b;
            if (q[0] == 0x1f && q[1] == 0x8b && q[2] == 8 && !(q[3] & 0xe0) && plausible_member(b, scratch)) {
                Start s;
                s.pos = (uint64_t)b * 8;
                s.kind = MEMBER;
                return s;
            }
            uint64_t w;
            memcpy(&w, q, 8);
            for (int k = 0; k < 8; ++k) {
                const uint32_t v = (uint32_t)(w >> k);
                if ((v & 7u) != 4u) continue;                               // BFINAL = 0, BTYPE = 2 (dynamic)
                if (((v >> 3) & 31u) > 29u || ((v >> 8) & 31u) > 29u) continue;   // HLIT <= 286, HDIST <= 30
                {
                    // the code-length code right behind (HCLEN x 3 bits from bit 17) must be a complete prefix code: the
                    // Kraft sum over its first 18 lengths settles that for all but one candidate in thousands, from the
                    // bits at hand (the full header parse of plausible_block costs ~150 ns)
                    const unsigned hclen = ((v >> 13) & 15u) + 4u;
                    uint64_t w2;
                    memcpy(&w2, q + 2, 8);
                    uint64_t cl = w2 >> (k + 1);                          // bit 17 of the candidate onwards, >= 56 bits
                    const unsigned m = hclen < 18u ? hclen : 18u;
                    unsigned sum = 0;
                    for (unsigned i = 0; i < m; ++i) {
                        const unsigned len = (unsigned)(cl & 7u);
                        cl >>= 3;
                        sum += len ? (128u >> len) : 0u;
                    }
                    if (sum > 128u || (hclen <= 18u && sum != 128u)) continue;
                }
                if (plausible_block((uint64_t)b * 8 + (uint64_t)k, scratch)) {
                    Start s;
                    s.pos = (uint64_t)b * 8 + (uint64_t)k;
                    s.kind = BLOCK;
                    return s;
                }
            }
        }
        return none;
    }

    // Decode from `start` until stop(position, kind) says that another chunk takes over there (asked at every block
    // boundary after the first block and at every member boundary), the stream ends, or an error.  Symbols go to out
    // (count in *n_out); every member that ends inside the chunk is recorded with its trailer.
    template <class StopFn>
    Status decode(const Start& start, StopFn stop, SymbolBuffer& out, size_t* n_out, std::vector<MemberEnd>& ends, Start* end_at,
                  const std::atomic<bool>* abort, size_t expect_symbols, size_t max_symbols) {
        size_t n = 0;
        *n_out = 0;
        err_ = "";
        if (!out.reserve(expect_symbols + 65536)) return fail_st("out of memory");
        if (start.kind == MEMBER) {
            in_ = base_ + (start.pos >> 3);
            bitbuf_ = 0;
            bitcnt_ = 0;
            const long h = gzip_header_len(in_, end_);
            if (h < 0) return fail_st("not a gzip member");
            in_ += h;
        } else {
            seek(start.pos);
        }
        bool first = true;
        for (;;) {
            const uint64_t p = tell();
            if (!first && stop(p, BLOCK)) {
                end_at->pos = p;
                end_at->kind = BLOCK;
                *n_out = n;
                return STOPPED;
            }
            first = false;
            if (abort != nullptr && abort->load(std::memory_order_relaxed)) return ABORTED;
            const int type = block_header(false);
            if (type < 0) return FAILED;
            for (;;) {                                       // the block's data, growing the buffer as needed
                if (n > max_symbols) return fail_st("a chunk expands beyond the parallel decoder's memory budget");
                if (out.cap < n + 66000 && !out.reserve(out.cap + out.cap / 2 + 66000)) return fail_st("out of memory");
                uint16_t* o = out.syms() + n;
                uint16_t* const soft_end = out.syms() + out.cap - 280;
                const int r = (type == 0) ? stored16(o, soft_end) : huffman16(o, soft_end);
                n = (size_t)(o - out.syms());
                if (r == 2) return FAILED;
                if (r == 1) break;
            }
            if (!last_block_) continue;
            // member trailer (byte aligned): CRC-32, ISIZE
            take(bitcnt_ & 7);
            in_ -= bitcnt_ >> 3;
            bitbuf_ = 0;
            bitcnt_ = 0;
            if (end_ - in_ < 8) return fail_st("truncated gzip trailer");
            MemberEnd me;
            me.off = n;
            me.crc = (uint32_t)in_[0] | ((uint32_t)in_[1] << 8) | ((uint32_t)in_[2] << 16) | ((uint32_t)in_[3] << 24);
            me.isize = (uint32_t)in_[4] | ((uint32_t)in_[5] << 8) | ((uint32_t)in_[6] << 16) | ((uint32_t)in_[7] << 24);
            in_ += 8;
            ends.push_back(me);
            *n_out = n;
            const uint64_t q = (uint64_t)(in_ - base_) * 8;
            if (stop(q, MEMBER)) {
                end_at->pos = q;
                end_at->kind = MEMBER;
                return STOPPED;
            }
            bool padding = true;
            for (const uint8_t* z = in_; z < end_; ++z)
                if (*z) {
                    padding = false;
                    break;
                }
            if (padding) return END_OF_STREAM;               // nothing, or zero padding, after the last member
            const long h = gzip_header_len(in_, end_);
            if (h < 0) {
                err_ = "trailing garbage after the last gzip member";
                return GARBAGE;
            }
            in_ += h;
        }
    }

private:
    enum { K_LIT = FastInflate::K_LIT, K_BASE = FastInflate::K_BASE, K_EOB = FastInflate::K_EOB, K_SUB = FastInflate::K_SUB };
    static constexpr int LBITS = FastInflate::LBITS, DBITS = FastInflate::DBITS;

    const uint8_t *base_, *end_, *lim_;
    const uint8_t* in_ = nullptr;
    uint64_t bitbuf_ = 0;
    int bitcnt_ = 0;
    bool last_block_ = false;
    uint32_t stored_left_ = 0;
    const char* err_ = "";
    uint32_t lit_[(1 << LBITS) + 1024];
    uint32_t dist_[(1 << DBITS) + 1024];

    Status fail_st(const char* why) {
        err_ = why;
        return FAILED;
    }
    int fail(const char* why) {
        err_ = why;
        return -1;
    }
    inline void refill() {
        uint64_t w;
        memcpy(&w, in_, 8);
        bitbuf_ |= w << bitcnt_;
        in_ += (63 - bitcnt_) >> 3;
        bitcnt_ |= 56;
    }
    inline uint32_t take(int n) {
        const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1));
        bitbuf_ >>= n;
        bitcnt_ -= n;
        return v;
    }
    void seek(uint64_t pos) {
        in_ = base_ + (pos >> 3);
        bitbuf_ = 0;
        bitcnt_ = 0;
        refill();
        take((int)(pos & 7));
    }
    uint64_t tell() const { return (uint64_t)(in_ - base_) * 8 - (uint64_t)bitcnt_; }

    // 0 = the code is complete, > 0 incomplete, < 0 over-subscribed
    static long kraft_left(const uint8_t* lens, int n) {
        long left = 1l << 15;
        for (int s = 0; s < n; ++s)
            if (lens[s]) left -= 1l << (15 - lens[s]);
        return left;
    }

    // Block header at the current position: returns the block type (0 stored, 1 fixed, 2 dynamic) with the tables built,
    // or -1.  strict: what zlib-family compressors always emit -- complete code-length and literal/length codes, a
    // complete distance code or a single distance code; used to reject candidates of the finder early.
    int block_header(bool strict) {
        if (in_ > lim_) return fail("truncated deflate stream");
        refill();
        last_block_ = take(1) != 0;
        const uint32_t type = take(2);
        if (type == 0) {
            take(bitcnt_ & 7);
            in_ -= bitcnt_ >> 3;
            bitbuf_ = 0;
            bitcnt_ = 0;
            if (end_ - in_ < 4) return fail("truncated stored block");
            const uint32_t len = (uint32_t)in_[0] | ((uint32_t)in_[1] << 8), nlen = (uint32_t)in_[2] | ((uint32_t)in_[3] << 8);
            if ((len ^ nlen) != 0xffffu) return fail("stored block length check failed");
            in_ += 4;
            stored_left_ = len;
            return 0;
        }
        if (type == 1) {
            uint8_t lens[288 + 32];
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            if (!FastInflate::build(lens, 288, LBITS, lit_, (int)(sizeof(lit_) / 4), 1)) return fail("bad fixed code");
            if (!FastInflate::build(lens + 288, 32, DBITS, dist_, (int)(sizeof(dist_) / 4), 2)) return fail("bad fixed code");
            return 1;
        }
        if (type == 3) return fail("reserved block type");
        const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
        if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        refill();
        for (int i = 0; i < hclen; ++i) {
            if (bitcnt_ < 3) refill();
            cl[order[i]] = (uint8_t)take(3);
        }
        if (strict) {
            long left = 1l << 7;
            for (int s = 0; s < 19; ++s)
                if (cl[s]) left -= 1l << (7 - cl[s]);
            if (left != 0) return fail("incomplete code-length code");
        }
        uint32_t cltab[(1 << 7) + 8];
        if (!FastInflate::build(cl, 19, 7, cltab, (int)(sizeof(cltab) / 4), 0)) return fail("bad code-length code");
        uint8_t lens[286 + 30 + 138];
        int n = 0;
        while (n < hlit + hdist) {
            refill();
            if (in_ > lim_) return fail("truncated dynamic block header");
            const uint32_t e = cltab[bitbuf_ & 127u];
            if (((e >> 8) & 15u) != (uint32_t)K_LIT) return fail("invalid code-length symbol");
            take((int)(e & 15u));
            const uint32_t sym = e >> 16;
            if (sym < 16) lens[n++] = (uint8_t)sym;
            else {
                int rep;
                uint8_t v = 0;
                if (sym == 16) {
                    if (n == 0) return fail("repeat with no previous length");
                    v = lens[n - 1];
                    rep = 3 + (int)take(2);
                } else if (sym == 17) rep = 3 + (int)take(3);
                else rep = 11 + (int)take(7);
                if (n + rep > hlit + hdist) return fail("code lengths overflow");
                while (rep--) lens[n++] = v;
            }
        }
        if (lens[256] == 0) return fail("no end-of-block code");
        if (strict) {
            if (kraft_left(lens, hlit) != 0) return fail("incomplete literal/length code");
            int used = 0;
            for (int s = 0; s < hdist; ++s) used += lens[hlit + s] != 0;
            if (kraft_left(lens + hlit, hdist) != 0 && used > 1) return fail("incomplete distance code");
        }
        if (!FastInflate::build(lens, hlit, LBITS, lit_, (int)(sizeof(lit_) / 4), 1)) return fail("bad literal/length code");
        if (!FastInflate::build(lens + hlit, hdist, DBITS, dist_, (int)(sizeof(dist_) / 4), 2)) return fail("bad distance code");
        return 2;
    }

    static inline void copy16(uint16_t* dst, uint32_t dist, uint32_t len) {
        // may write up to 15 symbols past dst + len (soft_end leaves 280 symbols of room for a match of at most 258)
        const uint16_t* src = dst - dist;
        if (dist >= 16) {                          // 32 bytes per step; may write up to 15 symbols past dst + len
            uint16_t* const end = dst + len;
            do {
                memcpy(dst, src, 32);
                src += 16;
                dst += 16;
            } while (dst < end);
            return;
        }
        if (dist >= 4) {
            uint16_t* const end = dst + len;
            do {
                uint64_t w;
                memcpy(&w, src, 8);
                memcpy(dst, &w, 8);
                src += 4;
                dst += 4;
            } while (dst < end);
            return;
        }
        if (dist == 1) {
            const uint16_t v = *src;
            for (uint32_t i = 0; i < len; ++i) dst[i] = v;
            return;
        }
        for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
    }

    // stored block data -> symbols.  1 = block finished, 0 = out of room, 2 = error
    int stored16(uint16_t*& outp, uint16_t* const soft_end) {
        uint16_t* out = outp;
        const size_t room = out < soft_end ? (size_t)(soft_end - out) : 0;
        const size_t n = stored_left_ < room ? stored_left_ : room;
        if ((size_t)(end_ - in_) < n) {
            err_ = "truncated stored block";
            return 2;
        }
        for (size_t i = 0; i < n; ++i) out[i] = in_[i];
        in_ += n;
        out += n;
        stored_left_ -= (uint32_t)n;
        outp = out;
        return stored_left_ == 0 ? 1 : 0;
    }

    // One Huffman block into 16-bit symbols.  1 = end of block, 0 = out of room, 2 = error
    int huffman16(uint16_t*& outp, uint16_t* const soft_end) {
        const uint32_t lmask = (1u << LBITS) - 1, dmask = (1u << DBITS) - 1;
        uint16_t* out = outp;
        for (;;) {
            if (out >= soft_end) {
                outp = out;
                return 0;
            }
            if (in_ > lim_) {
                err_ = "truncated deflate stream";
                outp = out;
                return 2;
            }
            refill();
            uint32_t e = lit_[bitbuf_ & lmask];
            if (((e >> 8) & 15u) == (uint32_t)K_SUB) {
                const uint32_t sb = (e >> 4) & 15u;
                e = lit_[(e >> 16) + ((uint32_t)(bitbuf_ >> LBITS) & ((1u << sb) - 1))];
                bitbuf_ >>= LBITS;
                bitcnt_ -= LBITS;
            }
            const uint32_t kind = (e >> 8) & 15u;
            bitbuf_ >>= (e & 15u);
            bitcnt_ -= (int)(e & 15u);
            if (kind == (uint32_t)K_LIT) {
                *out++ = (uint16_t)(e >> 16);
                continue;
            }
            if (kind == (uint32_t)K_BASE) {
                const uint32_t len = (e >> 16) + take((int)((e >> 4) & 15u));
                uint32_t d = dist_[bitbuf_ & dmask];
                if (((d >> 8) & 15u) == (uint32_t)K_SUB) {
                    const uint32_t sb = (d >> 4) & 15u;
                    d = dist_[(d >> 16) + ((uint32_t)(bitbuf_ >> DBITS) & ((1u << sb) - 1))];
                    bitbuf_ >>= DBITS;
                    bitcnt_ -= DBITS;
                }
                if (((d >> 8) & 15u) != (uint32_t)K_BASE) {
                    err_ = "invalid distance code";
                    outp = out;
                    return 2;
                }
                bitbuf_ >>= (d & 15u);
                bitcnt_ -= (int)(d & 15u);
                const uint32_t dist = (d >> 16) + take((int)((d >> 4) & 15u));
                if (bitcnt_ < 0 || dist > PI_WINDOW) {
                    err_ = "truncated or invalid deflate stream";
                    outp = out;
                    return 2;
                }
                copy16(out, dist, len);              // out - dist >= start of the marker prefix: dist <= PI_WINDOW
                out += len;
                continue;
            }
            if (kind == (uint32_t)K_EOB) {
                outp = out;
                if (bitcnt_ < 0) {
                    err_ = "truncated deflate stream";
                    return 2;
                }
                return 1;
            }
            err_ = "invalid literal/length code";
            outp = out;
            return 2;
        }
    }

    // trial decode of the block whose header was just parsed (type 1 / 2), bounded; true if no error shows up
    bool trial_block(int type, SymbolBuffer& scratch) {
        if (!scratch.reserve(1u << 20)) return false;
        if (type == 0) return true;
        uint16_t* o = scratch.syms();
        const int r = huffman16(o, scratch.syms() + scratch.cap - 280);
        if (r == 2) return false;
        if (r == 0) return true;                     // a very long block: accept, the chain check has the last word
        if (last_block_) return true;
        // what follows must look like a block header again
        if (in_ > lim_) return false;
        refill();
        const uint32_t v = (uint32_t)bitbuf_;
        const uint32_t nt = (v >> 1) & 3u;
        if (nt == 3) return false;
        if (nt == 2) return block_header(true) == 2;
        if (nt == 0) return block_header(false) == 0;
        return true;
    }
    bool plausible_block(uint64_t pos, SymbolBuffer& scratch) {
        seek(pos);
        if (block_header(true) != 2 || last_block_) return false;
        return trial_block(2, scratch);
    }
    bool plausible_member(size_t byte, SymbolBuffer& scratch) {
        const long h = gzip_header_len(base_ + byte, end_);
        if (h < 0) return false;
        in_ = base_ + byte + h;
        bitbuf_ = 0;
        bitcnt_ = 0;
        const int type = block_header(true);
        if (type < 0) return false;
        return trial_block(type, scratch);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// The pipeline: a pool of threads decodes chunks, chains them and resolves their markers; next() hands the resolved text
// out in order and checks the member trailers.
// ------------------------------------------------------------------------------------------------------------------
class ParallelInflate {
public:
    ~ParallelInflate() { stop(); }

    // Cuts the file into chunks (small ones first, so that the first text is there early) and looks for the starts of the
    // first few of them; the others are found by the pool when it gets to them.  Returns false when parallel decoding is
    // not worthwhile (no second start among the probed chunks: stored or fixed blocks only, one giant block, ...);
    // nothing is running then.
    bool start(const uint8_t* data, size_t size, int threads, size_t chunk_bytes, size_t hist) {
        data_ = data;
        size_ = size;
        hist_ = hist;
        nthreads_ = threads < 1 ? 1 : threads;
        if (chunk_bytes < 1024) chunk_bytes = 1024;
        chunk_bytes_ = chunk_bytes;
        offsets_.clear();
        {
            size_t at = 0, step = chunk_bytes / 8 < 1024 ? 1024 : chunk_bytes / 8;
            while (at < size) {
                offsets_.push_back(at);
                at += step;
                if (offsets_.size() % (size_t)nthreads_ == 0 && step < chunk_bytes) step = step * 2 < chunk_bytes ? step * 2 : chunk_bytes;
            }
            offsets_.push_back(size);
        }
        const size_t n = offsets_.size() - 1;
        if (n < 2) return false;
        nchunks_ = n;
        chunks_.reset(new Chunk[n]);
        links_.clear();
        links_.resize(n + 1);
        const size_t probe = n < (size_t)nthreads_ * 2 ? n : (size_t)nthreads_ * 2;
        {
            std::atomic<size_t> next(0);
            auto finder = [&]() {
                MarkerDecoder dec(data_, size_);
                SymbolBuffer scratch;
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= probe) break;
                    chunk_start(i, dec, scratch);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads_ && (size_t)t < probe; ++t) pool.emplace_back(finder);
            for (auto& th : pool) th.join();
        }
        size_t found = 0;
        for (size_t i = 0; i < probe; ++i) found += chunks_[i].start.kind != MarkerDecoder::NONE;
        if (found < 2) {
            chunks_.reset();
            nchunks_ = 0;
            links_.clear();
            return false;
        }
        links_[0].ready = true;                     // the stream starts with a member header at bit 0, empty window
        links_[0].end_pos = 0;
        links_[0].kind = MarkerDecoder::MEMBER;
        max_inflight_ = (size_t)nthreads_ * 2 + 2;
        for (int t = 0; t < nthreads_; ++t) pool_.emplace_back([this]() { worker(); });
        return true;
    }

    // Next piece of text in order: 1 = `bytes` holds hist + n bytes (text at offset hist), 0 = clean end of the stream,
    // -1 = failed (error() says why; the caller falls back to a sequential decoder), -2 = trailing garbage after the last
    // member (an error the sequential zlib path would not report).
    int next(std::vector<char>& bytes, size_t* n) {
        for (;;) {
            if (failed_) return -1;
            if (consumed_ == nchunks_) {
                if (!ended_) return set_failed("the chunk chain did not reach the end of the stream");
                return 0;
            }
            Chunk& c = chunks_[consumed_];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return c.done; });
            }
            const size_t idx = consumed_;
            {
                std::lock_guard<std::mutex> lk(mu_);
                ++consumed_;
            }
            cv_.notify_all();
            (void)idx;
            if (c.discarded) continue;
            if (c.status == MarkerDecoder::GARBAGE) {
                error_ = c.err;
                failed_ = true;
                return -2;
            }
            if (c.status == MarkerDecoder::FAILED || c.status == MarkerDecoder::ABORTED) return set_failed(c.err.c_str());
            // member trailers: pieces of this chunk joined to the running CRC of the member
            size_t piece = 0, at = 0;
            for (const MarkerDecoder::MemberEnd& me : c.ends) {
                const size_t len = (size_t)me.off - at;
                crc_run_ = (uint32_t)crc32_combine(crc_run_, c.piece_crc[piece], (z_off_t)len);
                len_run_ += len;
                if (crc_run_ != me.crc || (uint32_t)len_run_ != me.isize) return set_failed("gzip member fails its CRC-32 / length check");
                crc_run_ = 0;
                len_run_ = 0;
                at = (size_t)me.off;
                ++piece;
            }
            if (at < c.n) {
                crc_run_ = (uint32_t)crc32_combine(crc_run_, c.piece_crc[piece], (z_off_t)(c.n - at));
                len_run_ += c.n - at;
            }
            if (c.status == MarkerDecoder::END_OF_STREAM) {
                ended_ = true;
                if (len_run_ != 0) return set_failed("the stream ends inside a gzip member");
            }
            if (c.n == 0) continue;
            bytes.swap(c.bytes);                 // the caller's old buffer goes back to the pool
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!c.bytes.empty() && spare_.size() < max_inflight_ + 2) spare_.emplace_back(std::move(c.bytes));
            }
            std::vector<char>().swap(c.bytes);
            *n = c.n;
            return 1;
        }
    }
    const std::string& error() const { return error_; }
    // diagnostics: chunks the file was cut into, chunks with a start, chunks accepted into the chain so far
    void stats(size_t* chunks, size_t* with_start, size_t* accepted) {
        std::lock_guard<std::mutex> lk(mu_);
        *chunks = nchunks_;
        *with_start = *accepted = 0;
        for (size_t i = 0; i < nchunks_; ++i) {
            const Chunk& c = chunks_[i];
            *with_start += c.start_known.load() == 2 && c.start.kind != MarkerDecoder::NONE;
            *accepted += c.done && !c.discarded && c.status <= MarkerDecoder::END_OF_STREAM;
        }
    }
    // may be called from several threads (the reader's worker on a failure, the reader's owner on destruction)
    void stop() {
        std::lock_guard<std::mutex> guard(stop_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            abort_.store(true);
        }
        cv_.notify_all();
        for (auto& th : pool_)
            if (th.joinable()) th.join();
        pool_.clear();
    }

private:
    struct Chunk {
        MarkerDecoder::Start start, end_at;
        std::atomic<int> start_known{0};          // 2 once `start` holds the finder's answer
        std::mutex start_mu;
        int status = MarkerDecoder::FAILED;
        std::vector<MarkerDecoder::MemberEnd> ends;
        std::vector<uint32_t> piece_crc;          // one per piece between member ends (the last one may be open)
        std::vector<char> bytes;                  // [hist bytes unused | resolved text | slack]
        size_t n = 0;
        bool discarded = false, done = false;
        std::string err;
    };
    // what a chunk hands to its successor: where the accepted chain stands and the last 32 KiB of text before it
    struct Link {
        bool ready = false, failed = false, ended = false;
        uint64_t end_pos = 0;
        MarkerDecoder::Kind kind = MarkerDecoder::NONE;
        uint32_t wlen = 0;                        // valid bytes, right aligned in window
        std::vector<uint8_t> window;              // PI_WINDOW bytes once wlen > 0
    };

    const uint8_t* data_ = nullptr;
    size_t size_ = 0, hist_ = 0, chunk_bytes_ = 0, max_inflight_ = 4;
    int nthreads_ = 1;
    std::unique_ptr<Chunk[]> chunks_;             // not movable (mutex): a plain array
    size_t nchunks_ = 0;
    std::vector<size_t> offsets_;                 // chunk i covers bytes [offsets_[i], offsets_[i + 1])
    std::vector<Link> links_;
    std::vector<std::thread> pool_;
    std::vector<std::vector<char>> spare_;        // text buffers handed back by next(), guarded by mu_
    std::mutex mu_, stop_mu_;
    std::condition_variable cv_;
    std::atomic<bool> abort_{false};
    size_t next_chunk_ = 0;                       // guarded by mu_
    size_t consumed_ = 0;                         // guarded by mu_ (written by next())
    bool failed_ = false, ended_ = false;
    uint32_t crc_run_ = 0;
    uint64_t len_run_ = 0;
    std::string error_;

public:
    std::atomic<uint64_t> us_decode{0}, us_wait{0}, us_resolve{0};     // summed over the pool (diagnostics; resolve includes CRC-32)
private:
    static uint64_t now_us() {
        return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    int set_failed(const char* why) {
        error_ = why;
        failed_ = true;
        return -1;
    }

    // The start of chunk i, found on first use by whoever needs it first (the chunk's own worker, or the worker of an
    // earlier chunk that has to know where to stop).  `finder` must not be the decoder that is in the middle of a chunk.
    const MarkerDecoder::Start& chunk_start(size_t i, MarkerDecoder& finder, SymbolBuffer& scratch) {
        Chunk& c = chunks_[i];
        if (c.start_known.load(std::memory_order_acquire) == 2) return c.start;
        std::lock_guard<std::mutex> lk(c.start_mu);
        if (c.start_known.load(std::memory_order_relaxed) != 2) {
            if (i == 0) {
                c.start.pos = 0;
                c.start.kind = MarkerDecoder::MEMBER;
            } else {
                c.start = finder.find_start(offsets_[i], offsets_[i + 1], scratch);
            }
            c.start_known.store(2, std::memory_order_release);
        }
        return c.start;
    }

    void worker() {
        MarkerDecoder dec(data_, size_), finder(data_, size_);
        SymbolBuffer syms, scratch;
        const size_t n = nchunks_;
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (next_chunk_ >= n) return;
                i = next_chunk_++;
                cv_.wait(lk, [&] { return i < consumed_ + max_inflight_ || abort_.load(); });
            }
            process(i, dec, finder, syms, scratch);
        }
    }

    void publish_link(size_t i) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            links_[i].ready = true;
        }
        cv_.notify_all();
    }
    void publish_done(Chunk& c) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            c.done = true;
        }
        cv_.notify_all();
    }

    void process(size_t i, MarkerDecoder& dec, MarkerDecoder& finder, SymbolBuffer& syms, SymbolBuffer& scratch) {
        Chunk& c = chunks_[i];
        const size_t nchunks = nchunks_;
        size_t nsym = 0;
        uint64_t t0 = now_us();
        if (!abort_.load()) chunk_start(i, finder, scratch);
        if (c.start_known.load() == 2 && c.start.kind != MarkerDecoder::NONE && !abort_.load()) {
            size_t nxt = i + 1;
            auto stop_at = [&](uint64_t p, MarkerDecoder::Kind k) {
                while (nxt < nchunks) {
                    const MarkerDecoder::Start& st = chunk_start(nxt, finder, scratch);
                    if (st.kind != MarkerDecoder::NONE && st.pos >= p) return st.pos == p && st.kind == k;
                    ++nxt;
                }
                return false;
            };
            c.status = dec.decode(c.start, stop_at, syms, &nsym, c.ends, &c.end_at, &abort_, chunk_bytes_ * 10,
                                  std::max<size_t>(64u << 20, chunk_bytes_ * 128));
            if (c.status == MarkerDecoder::FAILED || c.status == MarkerDecoder::GARBAGE) c.err = dec.error();
        } else {
            c.status = MarkerDecoder::ABORTED;
        }
        // ---- the chain: wait for the predecessor's link
        uint64_t t1 = now_us();
        us_decode += t1 - t0;
        bool have_link;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return links_[i].ready || abort_.load(); });
            have_link = links_[i].ready;
        }
        t0 = now_us();
        us_wait += t0 - t1;
        Link& in = links_[i];
        Link& out = links_[i + 1];
        if (!have_link) {                        // woken by stop(): the predecessor never got here
            c.status = MarkerDecoder::ABORTED;
            c.err = "aborted";
            out.failed = true;
            publish_link(i + 1);
            syms.release();
            publish_done(c);
            return;
        }
        const bool accept = c.start.kind != MarkerDecoder::NONE && !in.failed && !in.ended && in.end_pos == c.start.pos &&
                            in.kind == c.start.kind;
        if (!accept) {
            c.discarded = !in.failed;            // a failed chain is reported by the chunk that failed
            if (in.failed) {
                c.status = MarkerDecoder::ABORTED;
                c.err = "an earlier chunk failed";
                c.discarded = true;
            }
            out.failed = in.failed;
            out.ended = in.ended;
            out.end_pos = in.end_pos;
            out.kind = in.kind;
            out.wlen = in.wlen;
            out.window = in.window;
            publish_link(i + 1);
            syms.release();
            publish_done(c);
            return;
        }
        if (c.status != MarkerDecoder::STOPPED && c.status != MarkerDecoder::END_OF_STREAM) {
            if (c.status == MarkerDecoder::ABORTED) c.err = "aborted";
            out.failed = true;
            publish_link(i + 1);
            syms.release();
            publish_done(c);
            return;
        }
        // ---- marker table from the predecessor's window; a chunk that starts a member has no history at all
        const uint32_t wlen = c.start.kind == MarkerDecoder::MEMBER ? 0u : in.wlen;
        std::vector<uint8_t> lut(65536, 0);
        for (int v = 0; v < 256; ++v) lut[(size_t)v] = (uint8_t)v;
        if (wlen) memcpy(lut.data() + PI_MARK, in.window.data(), PI_WINDOW);
        const uint16_t* s = syms.syms();
        if (wlen < PI_WINDOW) {
            // markers below this index point in front of the data: the stream is invalid there
            const uint16_t lowest_ok = (uint16_t)(PI_MARK + (PI_WINDOW - wlen));
            bool bad = false;
            if (wlen == 0) {
                uint16_t acc = 0;
                for (size_t j = 0; j < nsym; ++j) acc |= s[j];
                bad = (acc & PI_MARK) != 0;
            } else {
                for (size_t j = 0; j < nsym && !bad; ++j) bad = s[j] >= PI_MARK && s[j] < lowest_ok;
            }
            if (bad) {
                c.status = MarkerDecoder::FAILED;
                c.err = "distance reaches before the start of the data";
                out.failed = true;
                publish_link(i + 1);
                syms.release();
                publish_done(c);
                return;
            }
        }
        // ---- tail first: the successor's window
        out.window.resize(PI_WINDOW);
        if (nsym >= PI_WINDOW) {
            const uint16_t* t = s + (nsym - PI_WINDOW);
            for (uint32_t j = 0; j < PI_WINDOW; ++j) out.window[j] = lut[t[j]];
            out.wlen = PI_WINDOW;
        } else {
            const uint32_t keep = (uint32_t)((size_t)wlen < PI_WINDOW - nsym ? (size_t)wlen : PI_WINDOW - nsym);
            // [ kept tail of the old window | this chunk's text ] right aligned
            if (keep) memmove(out.window.data() + (PI_WINDOW - nsym - keep), in.window.data() + (PI_WINDOW - keep), keep);
            for (size_t j = 0; j < nsym; ++j) out.window[PI_WINDOW - nsym + j] = lut[s[j]];
            out.wlen = keep + (uint32_t)nsym;
        }
        out.end_pos = c.end_at.pos;
        out.kind = c.end_at.kind;
        out.ended = c.status == MarkerDecoder::END_OF_STREAM;
        publish_link(i + 1);
        std::vector<uint8_t>().swap(in.window);              // only this chunk read it (the table holds a copy): 32 KiB per chunk add up
        if (const char* dbg = getenv("EPI_INFLATE_DEBUG"); dbg != nullptr && dbg[0] == '2') {
            size_t marks = 0, last = 0;
            for (size_t j = 0; j < nsym; ++j)
                if (s[j] >= PI_MARK) {
                    ++marks;
                    last = j;
                }
            fprintf(stderr, "[epi reader]   chunk %zu: %zu symbols, %zu markers (%.2f %%), last at %zu\n", i, nsym, marks,
                    100.0 * marks / (nsym ? nsym : 1), last);
        }
        // ---- the bulk: symbols -> bytes, per-piece CRC-32
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!spare_.empty()) {               // recycled text buffers: no fresh pages, no zero fill
                c.bytes.swap(spare_.back());
                spare_.pop_back();
            }
        }
        if (c.bytes.size() < hist_ + nsym + 64) c.bytes.resize(hist_ + nsym + 64 + nsym / 16);
        uint8_t* o = reinterpret_cast<uint8_t*>(c.bytes.data()) + hist_;
        // in blocks of 128 KiB, so that the CRC reads the text the table pass has just written while it is still in cache
        {
            const uint8_t* const t = lut.data();
            constexpr size_t BLK = 128u << 10;
            size_t next_end = 0;                                 // index into c.ends
            uint32_t crc = 0;                                    // = crc32(0, NULL, 0)
            for (size_t j0 = 0; j0 < nsym || next_end < c.ends.size(); j0 += BLK) {
                const size_t j1 = j0 + BLK < nsym ? j0 + BLK : nsym;
                size_t j = j0;
                for (; j + 8 <= j1; j += 8) {
                    const uint64_t w = (uint64_t)t[s[j]] | ((uint64_t)t[s[j + 1]] << 8) | ((uint64_t)t[s[j + 2]] << 16) |
                                       ((uint64_t)t[s[j + 3]] << 24) | ((uint64_t)t[s[j + 4]] << 32) | ((uint64_t)t[s[j + 5]] << 40) |
                                       ((uint64_t)t[s[j + 6]] << 48) | ((uint64_t)t[s[j + 7]] << 56);
                    memcpy(o + j, &w, 8);
                }
                for (; j < j1; ++j) o[j] = t[s[j]];
                // CRC of [j0, j1), cut at the member ends that fall inside (an end at j1 belongs to this block)
                size_t at = j0;
                while (next_end < c.ends.size() && (size_t)c.ends[next_end].off <= j1) {
                    const size_t e = (size_t)c.ends[next_end].off;
                    if (e > at) crc = crc32_fast(crc, o + at, e - at);
                    c.piece_crc.push_back(crc);
                    crc = 0;
                    at = e;
                    ++next_end;
                }
                if (j1 > at) crc = crc32_fast(crc, o + at, j1 - at);
                if (j1 >= nsym && next_end >= c.ends.size()) break;
            }
            c.piece_crc.push_back(crc);                          // the open piece behind the last member end (may be empty)
        }
        c.n = nsym;
        t1 = now_us();
        us_resolve += t1 - t0;
        if (syms.cap > chunk_bytes_ * 40 + (1u << 22)) syms.release();      // do not keep an outlier's buffer around
        publish_done(c);
    }

};

}  // namespace epi
