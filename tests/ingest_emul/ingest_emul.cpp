// TEST SIDE: the per-record text of the device ingest (metamlst_b200/csrc/ingest_core.cuh) compiled for the host, driven the way
// csrc/ingest.cu drives it (one "thread" per BGZF block for the record chain, one per record for parse / pack), so that the chain
// guessing + verification, the field extraction and the plane rows are checked against the C++ unpacker without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "ingest_core.cuh"

using namespace ingest;

extern "C" {

// Record chain over the inflated stream.  uoff[nb+1] = start of every BGZF block in the stream.  Returns the number of records and
// fills roff (capacity cap); *repairs = blocks whose guessed first boundary had to be re-walked; -1 on a malformed chain.
long long emul_chain(const uint8_t* u, uint64_t usize, const uint64_t* uoff, uint32_t nb, uint64_t first_record, int32_t n_ref,
                     const uint32_t* ref_len, uint64_t* roff, uint64_t cap, uint32_t* repairs) {
    std::vector<uint64_t> start(nb), exit_(nb);
    std::vector<uint32_t> bad(nb, 0);
    auto walk = [&](uint32_t b, uint64_t s) {
        const uint64_t bend = (b + 1 < nb) ? uoff[b + 1] : usize;
        uint64_t off = s;
        uint32_t bd = 0;
        while (off < bend) {
            const uint64_t nx = next_record(u, off, usize);
            if (nx == 0) { bd = 1; off = usize; break; }
            off = nx;
        }
        start[b] = s; exit_[b] = off; bad[b] = bd;
    };
    for (uint32_t b = 0; b < nb; ++b) {   // chain_guess_kernel
        uint64_t s;
        const uint64_t bend = (b + 1 < nb) ? uoff[b + 1] : usize;
        if (uoff[b] <= first_record) {
            if (first_record >= bend) { start[b] = bend; exit_[b] = bend; continue; }
            s = first_record;
        } else {
            s = uoff[b];
            const uint64_t lim = (s + (1ull << 20) < usize) ? s + (1ull << 20) : usize;
            while (s < lim && !plausible_record(u, s, usize, n_ref, ref_len)) ++s;
            if (s >= lim) s = usize;
        }
        walk(b, s);
    }
    *repairs = 0;
    for (;;) {   // chain_repair_kernel sweeps
        uint32_t changed = 0;
        for (uint32_t b = 1; b < nb; ++b) {
            if (uoff[b] <= first_record) continue;
            if (start[b] == exit_[b - 1]) continue;
            ++changed;
            walk(b, exit_[b - 1]);
        }
        if (!changed) break;
        *repairs += changed;
    }
    uint64_t n = 0;
    for (uint32_t b = 0; b < nb; ++b) {   // chain_offsets_kernel
        const uint64_t bend = (b + 1 < nb) ? uoff[b + 1] : usize;
        uint64_t off = start[b];
        while (off < bend) {
            const uint64_t nx = next_record(u, off, usize);
            if (nx == 0) return -1;
            if (n < cap) roff[n] = off;
            ++n;
            off = nx;
        }
    }
    return static_cast<long long>(n);
}

// parse_kernel: returns 0 or (index << 8 | code) of the first refusal
uint64_t emul_parse(const uint8_t* u, const uint64_t* roff, uint64_t n, int32_t n_ref, uint64_t* key, uint16_t* reflen, int16_t* as0, int16_t* asn,
                    uint16_t* qlen, uint8_t* xm3, uint8_t* xmn, uint8_t* bits, uint64_t* qh) {
    for (uint64_t i = 0; i < n; ++i) {
        Core c;
        const uint32_t e = parse_record(u, roff[i], n_ref, &c, qh ? qh + 2 * i : nullptr);
        if (e != E_NONE) return (i << 8) | e;
        key[i] = c.key; reflen[i] = static_cast<uint16_t>(c.reflen); as0[i] = c.as0; asn[i] = c.asn; qlen[i] = c.qlen; xm3[i] = c.xm3; xmn[i] = c.xmn; bits[i] = c.bits;
    }
    return 0;
}

// pack_kernel for records list[0..P): rows at rowoff[j]
uint64_t emul_pack(const uint8_t* u, const uint64_t* roff, const uint32_t* list, uint64_t P, const int32_t* pos, const uint16_t* reflen, const uint8_t* bits,
                   int minqual, const uint64_t* rowoff, uint32_t* planes) {
    for (uint64_t j = 0; j < P; ++j) {
        const uint32_t k = list[j];
        const uint32_t rw = row_words(touched_words(static_cast<uint32_t>(pos[k]), reflen[k]));
        const uint32_t e = pack_record(u, roff[k], static_cast<uint32_t>(pos[k]), reflen[k], (bits[k] & 2u) != 0, minqual, planes + rowoff[j], rw);
        if (e != E_NONE) return (static_cast<uint64_t>(k) << 8) | e;
    }
    return 0;
}

// the warp form of pack_kernel: every 32-column word assembled from column_class of its 32 "lanes"
uint64_t emul_pack_columns(const uint8_t* u, const uint64_t* roff, const uint32_t* list, uint64_t P, const int32_t* pos, const uint16_t* reflen, int minqual,
                           const uint64_t* rowoff, uint32_t* planes) {
    for (uint64_t j = 0; j < P; ++j) {
        const uint32_t k = list[j];
        const uint32_t p0 = static_cast<uint32_t>(pos[k]), rl = reflen[k];
        const uint32_t nw = touched_words(p0, rl), rw = row_words(nw);
        uint32_t* row = planes + rowoff[j];
        for (uint32_t w = 0; w < rw; ++w) row[w] = 0;
        if (rl == 0) continue;
        const uint8_t* r = u + roff[k];
        const uint32_t l_name = r[12], n_cig = rd16(r + 16), l_seq = rd32(r + 20);
        const uint8_t* cig = r + 36 + l_name;
        const uint8_t* seq = cig + 4ull * n_cig;
        const uint8_t* qual = seq + (static_cast<uint64_t>(l_seq) + 1) / 2;
        for (uint32_t w = 0; w < nw; ++w)
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t cls = column_class(cig, n_cig, seq, qual, l_seq, p0, rl, ((p0 >> 5) + w) * 32u + lane, minqual);
                if (cls >= 1 && cls <= 4) row[3 * w] |= 1u << lane;
                if (cls == 3 || cls == 4) row[3 * w + 1] |= 1u << lane;
                if (cls == 2 || cls == 4 || cls == 5) row[3 * w + 2] |= 1u << lane;
            }
    }
    return 0;
}

}  // extern "C"
