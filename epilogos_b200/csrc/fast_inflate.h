// A DEFLATE / gzip decoder for the host reader (csrc/hostio.cu).  Host code.
//
// Why: reading a gzipped state matrix is bounded by zlib's inflate (~290 MB/s of text per file, measured with
// tools/read_bandwidth.py); the parser behind it runs at 450 MB/s and the GPU needs milliseconds.  Label matrices
// compress ~9:1 into long matches, which this decoder copies 8 bytes at a time from a 64-bit bit buffer with one
// refill per symbol (RFC 1951 decoding with two-level canonical Huffman tables; RFC 1952 member framing).
//
// Contract: decode() writes into a caller buffer that keeps the previous 32 KiB of output directly in front of
// `out` (the caller carries that history from block to block), stops exactly at `out_end` (a match or stored run is
// resumed by the next call) and reports the end of every gzip member so that the caller can check CRC-32 and ISIZE.
// Every malformed input is an error return, never undefined behaviour: table construction rejects over-subscribed
// codes, unassigned codes decode to an error entry, distances are checked against the bytes produced so far, and the
// input pointer may overrun only into the zero bytes the caller appends (a refill reads 8 bytes ahead and a block header
// takes a few refills between its checks: 16 bytes are the minimum, the reader appends 64).
#pragma once
#include <stdint.h>
#include <string.h>

namespace epi {

class MarkerDecoder;

class FastInflate {
    friend class MarkerDecoder;          // parallel_inflate.h shares the table builder and the entry encoding
public:
    enum Status { OUT_FULL = 0, MEMBER_END = 1, ERROR = 2 };

    // [data, data + size) must be followed by zero bytes (see above)
    void reset(const uint8_t* data, size_t size) {
        in_ = data;
        in_end_ = data + size;
        bitbuf_ = 0;
        bitcnt_ = 0;
        state_ = ST_HEADER;
        pend_len_ = 0;
        stored_left_ = 0;
        produced_ = 0;
        err_ = "";
    }
    bool at_end_of_input() const { return state_ == ST_HEADER && bitcnt_ < 8 && in_ >= in_end_; }
    // after MEMBER_END: nothing but zero bytes follows (some writers pad the file) -- the end of the stream
    bool only_padding_left() const {
        if (state_ != ST_HEADER) return false;
        for (const uint8_t* p = in_ - (bitcnt_ >> 3); p < in_end_; ++p)
            if (*p) return false;
        return true;
    }
    const char* error() const { return err_; }
    uint32_t member_crc() const { return crc_; }        // trailer fields of the member that just ended
    uint32_t member_isize() const { return isize_; }

    // Decode into [out, out_end).  *out_pos receives the new write position.
    Status decode(uint8_t* out, uint8_t* out_end, uint8_t** out_pos) {
        Status st = run(out, out_end, out_pos);
        return st;
    }

private:
    enum { ST_HEADER, ST_BLOCK_HEADER, ST_STORED, ST_HUFFMAN, ST_TRAILER };
    enum { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4 };
    static constexpr int LBITS = 11, DBITS = 8;

    const uint8_t *in_ = nullptr, *in_end_ = nullptr;
    uint64_t bitbuf_ = 0;
    int bitcnt_ = 0;
    int state_ = ST_HEADER;
    bool last_block_ = false;
    uint32_t pend_len_ = 0, pend_dist_ = 0;      // a match cut by the end of the output buffer
    uint32_t stored_left_ = 0;
    uint64_t produced_ = 0;                       // bytes of the current member so far (distance check)
    uint32_t crc_ = 0, isize_ = 0;
    const char* err_ = "";
    uint32_t lit_[(1 << LBITS) + 1024];           // primary + secondary tables (worst case < 852 + margin entries, as in zlib's ENOUGH)
    uint32_t dist_[(1 << DBITS) + 1024];

    static inline uint32_t entry(uint32_t val, uint32_t kind, uint32_t extra, uint32_t nbits) {
        return (val << 16) | (kind << 8) | (extra << 4) | nbits;
    }
    inline void refill() {
        // branch-free: top up to at least 56 bits with one unaligned 8-byte load (the input is padded)
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
    bool overrun() const { return in_ > in_end_ + 8; }

    static uint32_t reverse_bits(uint32_t code, int len) {
        uint32_t r = 0;
        for (int i = 0; i < len; ++i) {
            r = (r << 1) | (code & 1);
            code >>= 1;
        }
        return r;
    }

    // Canonical Huffman code (RFC 1951 3.2.2) -> two-level decoding table.  what: 0 = code lengths (symbols are values),
    // 1 = literal/length alphabet, 2 = distance alphabet.  Returns false for an over-subscribed code or lengths > 15.
    static bool build(const uint8_t* lens, int n, int tbits, uint32_t* table, int table_cap, int what) {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        int count[16] = {0};
        for (int s = 0; s < n; ++s) {
            if (lens[s] > 15) return false;
            ++count[lens[s]];
        }
        count[0] = 0;
        uint32_t next[16];
        uint32_t code = 0;
        long long left = 1;                                 // Kraft: must not go negative
        for (int l = 1; l <= 15; ++l) {
            left = (left << 1) - count[l];
            if (left < 0) return false;
            code = (code + (uint32_t)count[l - 1]) << 1;
            next[l] = code;
        }
        const int psize = 1 << tbits;
        const uint32_t bad = entry(0, K_BAD, 0, 1);
        for (int i = 0; i < psize; ++i) table[i] = bad;
        // secondary tables: the longest code under each primary prefix decides the size of its table
        uint8_t subbits[1 << LBITS];
        memset(subbits, 0, (size_t)psize);
        uint32_t codes[320];
        {
            uint32_t nx[16];
            memcpy(nx, next, sizeof(nx));
            for (int s = 0; s < n; ++s) {
                const int l = lens[s];
                if (l == 0) continue;
                codes[s] = reverse_bits(nx[l]++, l);
                if (l > tbits) {
                    const uint32_t p = codes[s] & (uint32_t)(psize - 1);
                    if (l - tbits > subbits[p]) subbits[p] = (uint8_t)(l - tbits);
                }
            }
        }
        int used = psize;
        for (int p = 0; p < psize; ++p) {
            if (subbits[p] == 0) continue;
            const int sz = 1 << subbits[p];
            if (used + sz > table_cap) return false;
            for (int i = 0; i < sz; ++i) table[used + i] = bad;
            table[p] = entry((uint32_t)used, K_SUB, subbits[p], (uint32_t)tbits);
            used += sz;
        }
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (l == 0) continue;
            uint32_t e;
            if (what == 0) e = entry((uint32_t)s, K_LIT, 0, 0);
            else if (what == 1) {
                if (s < 256) e = entry((uint32_t)s, K_LIT, 0, 0);
                else if (s == 256) e = entry(0, K_EOB, 0, 0);
                else if (s <= 285) e = entry(lbase[s - 257], K_BASE, lext[s - 257], 0);
                else e = entry(0, K_BAD, 0, 0);                // 286, 287: cannot occur in valid data
            } else {
                if (s < 30) e = entry(dbase[s], K_BASE, dext[s], 0);
                else e = entry(0, K_BAD, 0, 0);
            }
            if (l <= tbits) {
                e |= (uint32_t)l;
                for (uint32_t i = codes[s]; i < (uint32_t)psize; i += 1u << l) table[i] = e;
            } else {
                const uint32_t p = codes[s] & (uint32_t)(psize - 1);
                const uint32_t base = table[p] >> 16, sb = (table[p] >> 4) & 15u;
                e |= (uint32_t)(l - tbits);
                for (uint32_t i = codes[s] >> tbits; i < (1u << sb); i += 1u << (l - tbits)) table[base + i] = e;
            }
        }
        return true;
    }

    bool fail(const char* why) {
        err_ = why;
        return false;
    }

    // gzip member header (RFC 1952 2.3); byte-aligned, bit buffer empty
    bool member_header() {
        // drop whole bytes still in the bit buffer back to the input
        in_ -= bitcnt_ >> 3;
        bitbuf_ = 0;
        bitcnt_ = 0;
        const uint8_t* p = in_;
        if (in_end_ - p < 18) return fail("truncated gzip header");
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return fail("not a gzip member");
        const int flg = p[3];
        if (flg & 0xe0) return fail("reserved gzip flag bits set");
        p += 10;
        if (flg & 4) {
            if (in_end_ - p < 2) return fail("truncated gzip header");
            const size_t xlen = p[0] | (p[1] << 8);
            p += 2;
            if ((size_t)(in_end_ - p) < xlen) return fail("truncated gzip header");
            p += xlen;
        }
        for (int f = 8; f <= 16; f <<= 1)                     // FNAME, FCOMMENT: zero-terminated
            if (flg & f) {
                while (p < in_end_ && *p) ++p;
                if (p >= in_end_) return fail("truncated gzip header");
                ++p;
            }
        if (flg & 2) p += 2;
        if (p > in_end_) return fail("truncated gzip header");
        in_ = p;
        produced_ = 0;
        return true;
    }

    bool block_header() {
        refill();
        last_block_ = take(1) != 0;
        const uint32_t type = take(2);
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN
            take(bitcnt_ & 7);
            in_ -= bitcnt_ >> 3;                                // give whole bytes back
            bitbuf_ = 0;
            bitcnt_ = 0;
            if (in_end_ - in_ < 4) return fail("truncated stored block");
            const uint32_t len = in_[0] | (in_[1] << 8), nlen = in_[2] | (in_[3] << 8);
            if ((len ^ nlen) != 0xffffu) return fail("stored block length check failed");
            in_ += 4;
            stored_left_ = len;
            state_ = ST_STORED;
            return true;
        }
        if (type == 1) {
            uint8_t lens[288 + 32];
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            if (!build(lens, 288, LBITS, lit_, (int)(sizeof(lit_) / 4), 1)) return fail("bad fixed code");
            if (!build(lens + 288, 32, DBITS, dist_, (int)(sizeof(dist_) / 4), 2)) return fail("bad fixed code");
            state_ = ST_HUFFMAN;
            return true;
        }
        if (type == 3) return fail("reserved block type");
        // dynamic Huffman (RFC 1951 3.2.7)
        const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
        if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        refill();
        for (int i = 0; i < hclen; ++i) {
            if (bitcnt_ < 3) refill();
            cl[order[i]] = (uint8_t)take(3);
        }
        uint32_t cltab[(1 << 7) + 8];
        if (!build(cl, 19, 7, cltab, (int)(sizeof(cltab) / 4), 0)) return fail("bad code-length code");
        uint8_t lens[286 + 30 + 138];
        int n = 0;
        while (n < hlit + hdist) {
            refill();
            if (overrun()) return fail("truncated dynamic block header");
            const uint32_t e = cltab[bitbuf_ & 127u];
            if (((e >> 8) & 15u) != K_LIT) return fail("invalid code-length symbol");
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
        if (!build(lens, hlit, LBITS, lit_, (int)(sizeof(lit_) / 4), 1)) return fail("bad literal/length code");
        if (!build(lens + hlit, hdist, DBITS, dist_, (int)(sizeof(dist_) / 4), 2)) return fail("bad distance code");
        state_ = ST_HUFFMAN;
        return true;
    }

    static inline void copy_match(uint8_t* dst, uint32_t dist, uint32_t len, bool wide) {
        // dst - dist .. may overlap dst; may write up to 8 bytes (wide: 32) past dst + len (the caller leaves that slack)
        const uint8_t* src = dst - dist;
        if (dist >= 32 && wide) {                               // 32 bytes per step (label matrices copy whole rows: long matches)
            uint8_t* end = dst + len;
            do {
                memcpy(dst, src, 32);
                src += 32;
                dst += 32;
            } while (dst < end);
            return;
        }
        if (dist >= 8) {
            uint8_t* end = dst + len;
            do {
                uint64_t w;
                memcpy(&w, src, 8);
                memcpy(dst, &w, 8);
                src += 8;
                dst += 8;
            } while (dst < end);
            return;
        }
        if (dist == 1) {
            memset(dst, *src, len);
            return;
        }
        // short period: lay down the first 8 bytes one by one, then copy words from a multiple of the period behind
        uint32_t i = 0;
        for (; i < len && i < 8; ++i) dst[i] = src[i];
        if (i == len) return;
        const uint32_t back = dist * ((7 + dist) / dist);       // >= 8, a multiple of the period
        uint8_t* p = dst + 8;
        // bytes [dst, dst+8) are valid and periodic with `dist`; extend the periodic run
        const uint8_t* q = p - back;
        if (back > 8 + dist) {                                  // q would reach before src: fall back to bytes
            for (; i < len; ++i) dst[i] = dst[i - dist];
            return;
        }
        uint8_t* end = dst + len;
        while (p < end) {
            uint64_t w;
            memcpy(&w, q, 8);
            memcpy(p, &w, 8);
            p += 8;
            q += 8;
        }
    }

    Status run(uint8_t* out, uint8_t* const out_end, uint8_t** out_pos) {
        for (;;) {
            switch (state_) {
            case ST_HEADER:
                if (!member_header()) {
                    *out_pos = out;
                    return ERROR;
                }
                state_ = ST_BLOCK_HEADER;
                break;
            case ST_BLOCK_HEADER:
                if (!block_header() || overrun()) {
                    if (!*err_) err_ = "truncated deflate stream";
                    *out_pos = out;
                    return ERROR;
                }
                break;
            case ST_STORED: {
                const size_t room = (size_t)(out_end - out);
                size_t n = stored_left_ < room ? stored_left_ : room;
                if ((size_t)(in_end_ - in_) < n) {
                    err_ = "truncated stored block";
                    *out_pos = out;
                    return ERROR;
                }
                memcpy(out, in_, n);
                in_ += n;
                out += n;
                produced_ += n;
                stored_left_ -= (uint32_t)n;
                if (stored_left_ != 0) {
                    *out_pos = out;
                    return OUT_FULL;
                }
                state_ = last_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                break;
            }
            case ST_HUFFMAN: {
                if (pend_len_) {                                // finish the match the previous buffer cut
                    const size_t room = (size_t)(out_end - out);
                    const uint32_t n = pend_len_ < room ? pend_len_ : (uint32_t)room;
                    for (uint32_t i = 0; i < n; ++i) out[i] = out[(ptrdiff_t)i - (ptrdiff_t)pend_dist_];
                    out += n;
                    produced_ += n;
                    pend_len_ -= n;
                    if (pend_len_) {
                        *out_pos = out;
                        return OUT_FULL;
                    }
                }
                const int r = huffman(out, out_end, &out);
                if (r == 2) {
                    *out_pos = out;
                    return ERROR;
                }
                if (r == 0) {
                    *out_pos = out;
                    return OUT_FULL;
                }
                state_ = last_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                break;
            }
            case ST_TRAILER: {
                // to the byte boundary, then CRC-32 and ISIZE (little endian)
                take(bitcnt_ & 7);
                in_ -= bitcnt_ >> 3;
                bitbuf_ = 0;
                bitcnt_ = 0;
                if (in_end_ - in_ < 8) {
                    err_ = "truncated gzip trailer";
                    *out_pos = out;
                    return ERROR;
                }
                crc_ = (uint32_t)in_[0] | ((uint32_t)in_[1] << 8) | ((uint32_t)in_[2] << 16) | ((uint32_t)in_[3] << 24);
                isize_ = (uint32_t)in_[4] | ((uint32_t)in_[5] << 8) | ((uint32_t)in_[6] << 16) | ((uint32_t)in_[7] << 24);
                in_ += 8;
                state_ = ST_HEADER;
                *out_pos = out;
                return MEMBER_END;
            }
            }
        }
    }

    // One Huffman block.  Returns 1 at end of block, 0 when the output is full, 2 on error.
    int huffman(uint8_t* out, uint8_t* const out_end, uint8_t** out_pos) {
        const uint32_t lmask = (1u << LBITS) - 1, dmask = (1u << DBITS) - 1;
        uint8_t* const start = out;
        for (;;) {
            if (out >= out_end) {
                produced_ += (uint64_t)(out - start);
                *out_pos = out;
                return 0;
            }
            if (overrun()) {             // a stream cut short: the zero padding may decode as literals for ever
                err_ = "truncated deflate stream";
                *out_pos = out;
                return 2;
            }
            refill();
            uint32_t e = lit_[bitbuf_ & lmask];
            if (((e >> 8) & 15u) == K_SUB) {
                const uint32_t sb = (e >> 4) & 15u;
                e = lit_[(e >> 16) + ((uint32_t)(bitbuf_ >> LBITS) & ((1u << sb) - 1))];
                bitbuf_ >>= LBITS;
                bitcnt_ -= LBITS;
            }
            const uint32_t kind = (e >> 8) & 15u;
            bitbuf_ >>= (e & 15u);
            bitcnt_ -= (int)(e & 15u);
            if (kind == K_LIT) {
                *out++ = (uint8_t)(e >> 16);
                continue;
            }
            if (kind == K_BASE) {
                const uint32_t len = (e >> 16) + take((int)((e >> 4) & 15u));
                uint32_t d = dist_[bitbuf_ & dmask];
                if (((d >> 8) & 15u) == K_SUB) {
                    const uint32_t sb = (d >> 4) & 15u;
                    d = dist_[(d >> 16) + ((uint32_t)(bitbuf_ >> DBITS) & ((1u << sb) - 1))];
                    bitbuf_ >>= DBITS;
                    bitcnt_ -= DBITS;
                }
                if (((d >> 8) & 15u) != K_BASE) {
                    err_ = "invalid distance code";
                    *out_pos = out;
                    return 2;
                }
                bitbuf_ >>= (d & 15u);
                bitcnt_ -= (int)(d & 15u);
                const uint32_t dist = (d >> 16) + take((int)((d >> 4) & 15u));
                if (bitcnt_ < 0 || overrun()) {
                    err_ = "truncated deflate stream";
                    *out_pos = out;
                    return 2;
                }
                const uint64_t have = produced_ + (uint64_t)(out - start);
                if (dist > have || dist > 32768u) {
                    err_ = "distance reaches before the start of the data";
                    *out_pos = out;
                    return 2;
                }
                const size_t room = (size_t)(out_end - out);
                if (len + 8 <= room) {
                    copy_match(out, dist, len, len + 32 <= room);
                    out += len;
                } else {
                    const uint32_t n = len < room ? len : (uint32_t)room;
                    for (uint32_t i = 0; i < n; ++i) out[i] = out[(ptrdiff_t)i - (ptrdiff_t)dist];
                    out += n;
                    if (n < len) {
                        pend_len_ = len - n;
                        pend_dist_ = dist;
                        produced_ += (uint64_t)(out - start);
                        *out_pos = out;
                        return 0;
                    }
                }
                continue;
            }
            if (kind == K_EOB) {
                if (bitcnt_ < 0 || overrun()) {
                    err_ = "truncated deflate stream";
                    *out_pos = out;
                    return 2;
                }
                produced_ += (uint64_t)(out - start);
                *out_pos = out;
                return 1;
            }
            err_ = "invalid literal/length code";
            *out_pos = out;
            return 2;
        }
    }
};

}  // namespace epi
