// Raw DEFLATE decoder for BGZF blocks (RFC 1951), written for the BAM unpacker (csrc/bam_unpack.cpp).
//
// Why not zlib's inflate(): a BGZF block is a complete, independent DEFLATE stream of at most 64 KiB whose output size is
// known in advance (ISIZE), so none of zlib's streaming state machine is needed, and inflate is the largest phase of the
// ingest that replaces `samtools view` / pysam (metamlst.py:96, cmseq/cmseq.py:54).  This decoder
// works on whole buffers: a 64-bit bit buffer refilled eight bytes at a time, an 11-bit first-level table for
// literal/length codes and an 8-bit one for distances (second-level tables for longer codes), several literals per
// refill, 8-byte match copies.  Every read and write is bounds-checked against the block's own input and output
// ranges: a corrupt block yields an error code, never an access outside [in, in+n) or [out, out+cap).
//
// mmlst_inflate_raw is exported for the tests (tests/test_inflate.py: byte-for-byte against zlib on every compression
// level, stored / fixed / dynamic blocks, truncations and bit flips).
#include <stdint.h>
#include <string.h>

#include "../../include/mmlst.h"

namespace {

constexpr int LITLEN_TBITS = 11, OFFSET_TBITS = 8, PRECODE_TBITS = 7;
constexpr int LITLEN_TSIZE = (1 << LITLEN_TBITS) + 512;  // main table + second-level tables (worst case 294 entries)
constexpr int OFFSET_TSIZE = (1 << OFFSET_TBITS) + 256;  // worst case 146
constexpr int PRECODE_TSIZE = 1 << PRECODE_TBITS;

// table entry: bits 0-7 bits to consume, bits 8-12 extra-bit count (or second-level index bits), bit 13 end of block,
// bit 14 pointer to a second-level table, bit 15 literal, bits 16-31 payload (literal / base value / second-level start)
constexpr uint32_t F_LITERAL = 1u << 15, F_SUBTABLE = 1u << 14, F_EOB = 1u << 13;
inline uint32_t mk(uint32_t payload, uint32_t flags, uint32_t extra, uint32_t nbits) { return (payload << 16) | flags | (extra << 8) | nbits; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t bitrev(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

enum Kind { K_LITLEN, K_OFFSET, K_PRECODE };

inline uint32_t symbol_entry(Kind kind, uint32_t sym, uint32_t nbits, bool* ok) {
    if (kind == K_PRECODE) return mk(sym, 0, 0, nbits);
    if (kind == K_OFFSET) {
        if (sym >= 30) { *ok = false; return 0; }  // codes 30, 31 never occur in valid data
        return mk(DIST_BASE[sym], 0, DIST_EXTRA[sym], nbits);
    }
    if (sym < 256) return mk(sym, F_LITERAL, 0, nbits);
    if (sym == 256) return mk(0, F_EOB, 0, nbits);
    if (sym >= 286) { *ok = false; return 0; }
    return mk(LEN_BASE[sym - 257], 0, LEN_EXTRA[sym - 257], nbits);
}

// Canonical Huffman decode table from code lengths.  Unused slots of an incomplete code stay 0 (nbits == 0): hitting one is
// an error at decode time.  Returns false for an over-subscribed code or a table that does not fit.
bool build_table(Kind kind, const uint8_t* lens, int n, uint32_t* table, int tbits, int tsize) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    if (count[0] == n) {  // no codes at all: legal for distances when the block has only literals
        memset(table, 0, sizeof(uint32_t) * (size_t)tsize);
        return true;
    }
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;  // over-subscribed
    }
    uint32_t next[16];
    uint32_t code = 0;
    count[0] = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    memset(table, 0, sizeof(uint32_t) * (size_t)tsize);
    const uint32_t tmask = (1u << tbits) - 1u;
    // pass 1: longest code under every first-level prefix that needs a second level
    uint8_t sub_bits[1 << LITLEN_TBITS];
    memset(sub_bits, 0, (size_t)1 << tbits);
    uint16_t codes[288];
    bool ok = true;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = bitrev(next[l]++, l);
        codes[s] = (uint16_t)rev;
        if (l > tbits) {
            const uint32_t pre = rev & tmask;
            if (l - tbits > sub_bits[pre]) sub_bits[pre] = (uint8_t)(l - tbits);
        }
    }
    int used = 1 << tbits;
    for (uint32_t pre = 0; pre <= tmask; ++pre) {
        if (!sub_bits[pre]) continue;
        const int sz = 1 << sub_bits[pre];
        if (used + sz > tsize) return false;
        table[pre] = mk((uint32_t)used, F_SUBTABLE, sub_bits[pre], (uint32_t)tbits);
        used += sz;
    }
    // pass 2: fill
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = codes[s];
        if (l <= tbits) {
            const uint32_t e = symbol_entry(kind, (uint32_t)s, (uint32_t)l, &ok);
            for (uint32_t i = rev; i <= tmask; i += 1u << l) table[i] = e;
        } else {
            const uint32_t pre = rev & tmask;
            const uint32_t start = table[pre] >> 16, sb = sub_bits[pre];
            const uint32_t e = symbol_entry(kind, (uint32_t)s, (uint32_t)(l - tbits), &ok);
            for (uint32_t i = rev >> tbits; i < (1u << sb); i += 1u << (l - tbits)) table[start + i] = e;
        }
    }
    return ok;
}

struct Bits {
    const uint8_t* in; const uint8_t* end;
    uint64_t buf = 0;
    int cnt = 0;  // valid bits in buf
    inline void refill() {
        if (end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < end) { buf |= (uint64_t)(*in++) << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline bool take(int n) {  // false: the stream ended inside a symbol
        if (n > cnt) return false;
        buf >>= n; cnt -= n;
        return true;
    }
};

struct Fixed {
    uint32_t litlen[LITLEN_TSIZE], offset[OFFSET_TSIZE];
    bool ok;
    Fixed() {
        uint8_t l[288];
        for (int i = 0; i < 144; ++i) l[i] = 8;
        for (int i = 144; i < 256; ++i) l[i] = 9;
        for (int i = 256; i < 280; ++i) l[i] = 7;
        for (int i = 280; i < 288; ++i) l[i] = 8;
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) d[i] = 5;
        // symbols 286, 287 / 30, 31 take part in the code but are invalid: build with all of them, then blank their slots
        bool a = build_fixed(l, 288, litlen, LITLEN_TBITS, LITLEN_TSIZE, K_LITLEN);
        bool b = build_fixed(d, 32, offset, OFFSET_TBITS, OFFSET_TSIZE, K_OFFSET);
        ok = a && b;
    }
    static bool build_fixed(const uint8_t* lens, int n, uint32_t* table, int tbits, int tsize, Kind kind) {
        // same as build_table, but invalid symbols leave an empty (error) slot instead of failing the whole table
        int count[16] = {0};
        for (int i = 0; i < n; ++i) count[lens[i]]++;
        uint32_t next[16];
        uint32_t code = 0;
        count[0] = 0;
        for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
        memset(table, 0, sizeof(uint32_t) * (size_t)tsize);
        const uint32_t tmask = (1u << tbits) - 1u;
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            const uint32_t rev = bitrev(next[l]++, l);
            bool ok = true;
            const uint32_t e = symbol_entry(kind, (uint32_t)s, (uint32_t)l, &ok);
            if (!ok) continue;
            for (uint32_t i = rev; i <= tmask; i += 1u << l) table[i] = e;
        }
        return true;
    }
};

const Fixed& fixed_tables() {
    static const Fixed f;
    return f;
}

int inflate_block_data(Bits& b, const uint32_t* litlen, const uint32_t* offset, uint8_t* out, size_t cap, size_t& pos) {
    // Fast loop: while at least 8 input bytes and 266 output bytes (3 literals, or the longest match + 8 bytes of copy
    // slack) remain, a refill leaves >= 56 real bits -- enough for a whole length/distance pair (15 + 5 + 15 + 13 = 48) --
    // so nothing inside needs an input or output check; only invalid codes and distances beyond the output are tested.
    constexpr uint64_t LMASK = (1u << LITLEN_TBITS) - 1u, OMASK = (1u << OFFSET_TBITS) - 1u;
    while (b.end - b.in >= 8 && cap - pos >= 266) {
        uint64_t w;
        memcpy(&w, b.in, 8);
        b.buf |= w << b.cnt;
        b.in += (63 - b.cnt) >> 3;
        b.cnt |= 56;
        uint32_t e = litlen[b.buf & LMASK];
        if (e & F_LITERAL) {  // up to three literals per refill (<= 33 bits), then back for more bits
            b.buf >>= (e & 0xffu); b.cnt -= (int)(e & 0xffu);
            out[pos++] = (uint8_t)(e >> 16);
            e = litlen[b.buf & LMASK];
            if (e & F_LITERAL) {
                b.buf >>= (e & 0xffu); b.cnt -= (int)(e & 0xffu);
                out[pos++] = (uint8_t)(e >> 16);
                e = litlen[b.buf & LMASK];
                if (e & F_LITERAL) {
                    b.buf >>= (e & 0xffu); b.cnt -= (int)(e & 0xffu);
                    out[pos++] = (uint8_t)(e >> 16);
                }
            }
            continue;
        }
        if (e & F_SUBTABLE) {
            b.buf >>= LITLEN_TBITS; b.cnt -= LITLEN_TBITS;
            e = litlen[(e >> 16) + (uint32_t)(b.buf & ((1u << ((e >> 8) & 31u)) - 1u))];
        }
        const uint32_t nb = e & 0xffu;
        if (nb == 0) return MMLST_E_BAM;  // unused code
        b.buf >>= nb; b.cnt -= (int)nb;
        if (e & F_LITERAL) { out[pos++] = (uint8_t)(e >> 16); continue; }
        if (e & F_EOB) return MMLST_OK;
        const uint32_t lx = (e >> 8) & 31u;
        const uint32_t len = (e >> 16) + (uint32_t)(b.buf & ((1u << lx) - 1u));
        b.buf >>= lx; b.cnt -= (int)lx;
        uint32_t d = offset[b.buf & OMASK];
        if (d & F_SUBTABLE) {
            b.buf >>= OFFSET_TBITS; b.cnt -= OFFSET_TBITS;
            d = offset[(d >> 16) + (uint32_t)(b.buf & ((1u << ((d >> 8) & 31u)) - 1u))];
        }
        const uint32_t db = d & 0xffu;
        if (db == 0) return MMLST_E_BAM;
        b.buf >>= db; b.cnt -= (int)db;
        const uint32_t dx = (d >> 8) & 31u;
        const uint32_t dist = (d >> 16) + (uint32_t)(b.buf & ((1u << dx) - 1u));
        b.buf >>= dx; b.cnt -= (int)dx;
        if (dist > pos) return MMLST_E_BAM;
        uint8_t* dst = out + pos;
        const uint8_t* src = dst - dist;
        pos += len;
        if (dist >= 8) {
            uint8_t* const stop = dst + len;
            do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < stop);
        } else if (dist == 1) {
            memset(dst, *src, len);
        } else {
            for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
        }
    }
    // Careful loop for the ends of the input / output: every step checked.
    for (;;) {
        b.refill();
        uint32_t e = litlen[b.peek(LITLEN_TBITS)];
        if (e & F_SUBTABLE) {
            if (!b.take(LITLEN_TBITS)) return MMLST_E_BAM;
            e = litlen[(e >> 16) + b.peek((int)((e >> 8) & 31u))];
        }
        if ((e & 0xffu) == 0) return MMLST_E_BAM;  // unused code
        if (!b.take((int)(e & 0xffu))) return MMLST_E_BAM;
        if (e & F_LITERAL) {
            if (pos >= cap) return MMLST_E_BAM;
            out[pos++] = (uint8_t)(e >> 16);
            // up to two more literals from the bits already in the buffer (>= 41 left after a full refill)
            for (int k = 0; k < 2 && b.cnt >= 15; ++k) {
                const uint32_t e2 = litlen[b.peek(LITLEN_TBITS)];
                if (!(e2 & F_LITERAL) || pos >= cap) break;
                b.take((int)(e2 & 0xffu));
                out[pos++] = (uint8_t)(e2 >> 16);
            }
            continue;
        }
        if (e & F_EOB) return MMLST_OK;
        // length symbol
        const int lx = (int)((e >> 8) & 31u);
        uint32_t len = e >> 16;
        if (lx) { len += b.peek(lx); if (!b.take(lx)) return MMLST_E_BAM; }
        uint32_t d = offset[b.peek(OFFSET_TBITS)];
        if (d & F_SUBTABLE) {
            if (!b.take(OFFSET_TBITS)) return MMLST_E_BAM;
            d = offset[(d >> 16) + b.peek((int)((d >> 8) & 31u))];
        }
        if ((d & 0xffu) == 0) return MMLST_E_BAM;
        if (!b.take((int)(d & 0xffu))) return MMLST_E_BAM;
        const int dx = (int)((d >> 8) & 31u);
        uint32_t dist = d >> 16;
        if (dx) { dist += b.peek(dx); if (!b.take(dx)) return MMLST_E_BAM; }
        if (dist > pos || len > cap - pos) return MMLST_E_BAM;
        uint8_t* dst = out + pos;
        const uint8_t* src = dst - dist;
        pos += len;
        if (dist >= 8 && cap - pos >= 8) {  // 8 bytes at a time; may write up to 7 bytes past the match, inside the block's output
            uint8_t* const stop = dst + len;
            do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < stop);
        } else if (dist == 1) {
            memset(dst, *src, len);
        } else {
            for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
        }
    }
}

int inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* produced) {
    Bits b{in, in + n};
    size_t pos = 0;
    uint32_t litlen[LITLEN_TSIZE], offset[OFFSET_TSIZE], pre[PRECODE_TSIZE];
    for (;;) {
        b.refill();
        if (b.cnt < 3) return MMLST_E_BAM;
        const uint32_t final_block = b.peek(1);
        const uint32_t type = (b.peek(3) >> 1) & 3u;
        b.take(3);
        if (type == 0) {  // stored: back to a byte boundary; bytes already in the bit buffer are handed back
            b.take(b.cnt & 7);
            b.in -= b.cnt >> 3;
            b.buf = 0; b.cnt = 0;
            if (b.end - b.in < 4) return MMLST_E_BAM;
            const uint32_t len = (uint32_t)b.in[0] | ((uint32_t)b.in[1] << 8), nlen = (uint32_t)b.in[2] | ((uint32_t)b.in[3] << 8);
            b.in += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(b.end - b.in) < len || cap - pos < len) return MMLST_E_BAM;
            memcpy(out + pos, b.in, len);
            pos += len; b.in += len;
        } else if (type == 1) {
            const Fixed& f = fixed_tables();
            const int rc = inflate_block_data(b, f.litlen, f.offset, out, cap, pos);
            if (rc != MMLST_OK) return rc;
        } else if (type == 2) {
            b.refill();
            if (b.cnt < 14) return MMLST_E_BAM;
            const int hlit = (int)b.peek(5) + 257; b.take(5);
            const int hdist = (int)b.peek(5) + 1; b.take(5);
            const int hclen = (int)b.peek(4) + 4; b.take(4);
            if (hlit > 286 || hdist > 30) return MMLST_E_BAM;
            uint8_t pl[19] = {0};
            for (int i = 0; i < hclen; ++i) {
                b.refill();
                pl[PRECODE_ORDER[i]] = (uint8_t)b.peek(3);
                if (!b.take(3)) return MMLST_E_BAM;
            }
            if (!build_table(K_PRECODE, pl, 19, pre, PRECODE_TBITS, PRECODE_TSIZE)) return MMLST_E_BAM;
            uint8_t lens[286 + 30 + 138];
            int i = 0;
            const int total = hlit + hdist;
            while (i < total) {
                b.refill();
                const uint32_t e = pre[b.peek(PRECODE_TBITS)];
                if ((e & 0xffu) == 0 || !b.take((int)(e & 0xffu))) return MMLST_E_BAM;
                const uint32_t sym = e >> 16;
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                uint32_t rep; uint8_t val = 0;
                if (sym == 16) {
                    if (i == 0) return MMLST_E_BAM;
                    val = lens[i - 1];
                    rep = 3 + b.peek(2); if (!b.take(2)) return MMLST_E_BAM;
                } else if (sym == 17) {
                    rep = 3 + b.peek(3); if (!b.take(3)) return MMLST_E_BAM;
                } else {
                    rep = 11 + b.peek(7); if (!b.take(7)) return MMLST_E_BAM;
                }
                if (i + (int)rep > total) return MMLST_E_BAM;
                memset(lens + i, val, rep);
                i += (int)rep;
            }
            if (lens[256] == 0) return MMLST_E_BAM;  // no end-of-block code
            if (!build_table(K_LITLEN, lens, hlit, litlen, LITLEN_TBITS, LITLEN_TSIZE)) return MMLST_E_BAM;
            if (!build_table(K_OFFSET, lens + hlit, hdist, offset, OFFSET_TBITS, OFFSET_TSIZE)) return MMLST_E_BAM;
            const int rc = inflate_block_data(b, litlen, offset, out, cap, pos);
            if (rc != MMLST_OK) return rc;
        } else {
            return MMLST_E_BAM;
        }
        if (final_block) break;
    }
    *produced = pos;
    return MMLST_OK;
}

}  // namespace

// see include/mmlst.h
extern "C" int mmlst_inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t out_capacity, size_t* produced) {
    size_t p = 0;
    if (!in || !out || !produced) return MMLST_E_ARG;
    const int rc = inflate_raw(in, n, out, out_capacity, &p);
    *produced = rc == MMLST_OK ? p : 0;
    return rc;
}
