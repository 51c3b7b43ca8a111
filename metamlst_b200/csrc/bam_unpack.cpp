// BAM -> packed record streams, host side (C++17 threads + zlib).  Replaces the two ingest routes of the reference:
// the `samtools view -h -` text pipe of stage 1 (metamlst.py:96-110) and pysam's AlignmentFile/pileup record access of
// stage 2 (cmseq/cmseq.py:54,527-545), plus `samtools sort` (metaMLST_functions.py:237-247) -- the BAM is inflated and
// walked ONCE and leaves as the structure-of-arrays streams of include/mmlst.h (score stream, pileup stream with the
// CIGAR already projected into 3 bit-planes per 32 reference columns, htslib depth cap applied).
//
//   phase 1  scan BGZF block headers (BSIZE in the BC extra subfield, ISIZE in the trailer)       sequential, bytes
//   phase 2  raw-inflate every block into one contiguous buffer                                    parallel
//   phase 3  BAM header + reference dictionary; chain of record offsets                            sequential, 4 B/rec
//   phase 4  per record: core fields, aux fields by POSITION (1st, 4th) and by NAME (AS, XM),
//            reference span, QNAME hash; every input the reference would crash on is refused       parallel
//   phase 5  `samtools sort` order = stable by (tid, pos, reverse strand) unless already so        bucket + parallel
//   phase 6  depth-cap admission (mmlst_depth_cap), row offsets, plane rows                        parallel
//
// No result of the reference is computed here: this is layout work (what samtools/pysam/htslib do in C for the
// reference); filtering, sums, histograms, consensus and distances all happen in the CUDA kernels.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mmlst.h"

void mmlst_set_error(const char* fmt, ...);

namespace {

struct Buf {  // host array, page-locked when asked for (async DMA source)
    void* p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    bool alloc(size_t n, bool pin) {
        release();
        bytes = n;
        if (n == 0) n = 16;
        if (pin) {
            p = mmlst_pinned_alloc(n);
            pinned = p != nullptr;
        }
        if (!p) {
            if (posix_memalign(&p, 256, n) != 0) p = nullptr;
            pinned = false;
        }
        return p != nullptr;
    }
    void release() {
        if (!p) return;
        if (pinned) mmlst_pinned_free(p); else free(p);
        p = nullptr; bytes = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

template <class F>
void parallel_for(size_t n, int threads, size_t grain, F&& f) {
    if (n == 0) return;
    const size_t nchunks = (n + grain - 1) / grain;
    int nt = (int)std::min<size_t>((size_t)std::max(threads, 1), nchunks);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            f(c * grain, std::min(n, (c + 1) * grain));
        }
    };
    if (nt <= 1) { work(); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }

struct Block { size_t coff; uint32_t clen; uint32_t isize; uint32_t crc; size_t uoff; };

// first error wins; message kept for mmlst_set_error on the calling thread
struct Err {
    std::atomic<int> code{0};
    char msg[400] = "";
    void set(int c, const char* fmt, ...) {
        int expected = 0;
        if (!code.compare_exchange_strong(expected, c)) return;
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(msg, sizeof(msg), fmt, ap);
        va_end(ap);
    }
};

// size in bytes of one aux value of type `t` at p (end = record end); 0 on malformed input
inline size_t aux_size(uint8_t t, const uint8_t* p, const uint8_t* end) {
    switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'Z': case 'H': {
            const void* z = memchr(p, 0, (size_t)(end - p));
            return z ? (size_t)((const uint8_t*)z - p) + 1 : 0;
        }
        case 'B': {
            if (end - p < 5) return 0;
            size_t es;
            switch (p[0]) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break; case 'i': case 'I': case 'f': es = 4; break; default: return 0; }
            return 5 + es * (size_t)rd32(p + 1);
        }
        default: return 0;
    }
}
inline bool aux_int(uint8_t t, const uint8_t* p, int64_t* v) {
    switch (t) {
        case 'c': *v = (int8_t)p[0]; return true;
        case 'C': *v = p[0]; return true;
        case 's': *v = (int16_t)rd16(p); return true;
        case 'S': *v = rd16(p); return true;
        case 'i': *v = rdi32(p); return true;
        case 'I': *v = rd32(p); return true;
        default: return false;
    }
}

inline uint64_t hash64(const uint8_t* s, size_t n) {  // FNV-1a folded through a splitmix finaliser
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= s[i]; h *= 0x100000001b3ull; }
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
    return h;
}
// second, independent 64-bit hash (different multiplier, seed and finaliser): with hash64 it forms the 128-bit name key
inline uint64_t hash64b(const uint8_t* s, size_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xff51afd7ed558ccdull);
    for (size_t i = 0; i < n; ++i) { h = (h ^ s[i]) * 0xc6a4a7935bd1e995ull; h ^= h >> 47; }
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}

inline uint32_t touched_words(uint32_t pos, uint32_t reflen) { return reflen ? (((pos & 31u) + reflen + 31u) >> 5) : 0u; }
inline uint32_t row_words(uint32_t nw) {  // 3 planes, padded to an odd word count
    const uint32_t rw = 3u * nw;
    return rw + ((rw != 0u && (rw & 1u) == 0u) ? 1u : 0u);
}

}  // namespace

struct mmlst_bam {
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_len;
    std::string names_blob;  // '\n'-joined, for one-call transfer to the binding
    std::string header_text;
    Buf tid, as0, xm3, qlen, orig_idx, qhash, p_recs, planes, contig_start, run_tid, run_start, chunk_run, chunk_qlen;
    uint32_t n_runs = 0;
    bool qc = false;  // chunk_qlen valid: every 256-record chunk has one len(SEQ)
    uint64_t n_rec = 0, n_prec = 0, n_plane_words = 0, n_dropped = 0, n_unmapped_flag = 0, n_untagged = 0;
    uint32_t max_row_words = 0;
    int presorted = 0, minqual = 20;
    uint32_t max_depth = 0;
    double t_read = 0, t_inflate = 0, t_parse = 0, t_sort = 0, t_pack = 0;
    ~mmlst_bam() {
        for (Buf* b : {&tid, &as0, &xm3, &qlen, &orig_idx, &qhash, &p_recs, &planes, &contig_start, &run_tid, &run_start, &chunk_run, &chunk_qlen}) b->release();
    }
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" void mmlst_bam_free(mmlst_bam* b) { delete b; }

extern "C" int mmlst_bam_unpack(const char* path, const mmlst_unpack_opts* opts_in, mmlst_bam** out) {
    if (!path || !out) { mmlst_set_error("mmlst_bam_unpack: null argument"); return MMLST_E_ARG; }
    mmlst_unpack_opts o;
    o.minqual = 20; o.max_depth = 8000; o.sentinel_nodes = 1; o.n_threads = 0; o.pinned = 1; o.assume_sorted = 0; o.want_qhash = 1; o.check_crc = 1; o.lenient_tags = 0;
    if (opts_in) o = *opts_in;
    const bool lenient = o.lenient_tags != 0;
    int threads = o.n_threads > 0 ? o.n_threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const bool pin = o.pinned != 0;
    double t0 = now_s();

    // ---- phase 0: file -> memory
    FILE* f = fopen(path, "rb");
    if (!f) { mmlst_set_error("cannot open %s", path); return MMLST_E_IO; }
    fseek(f, 0, SEEK_END);
    const long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (fsz < 0) { fclose(f); mmlst_set_error("cannot stat %s", path); return MMLST_E_IO; }
    std::vector<uint8_t> raw((size_t)fsz);
    if (fsz && fread(raw.data(), 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); mmlst_set_error("short read on %s", path); return MMLST_E_IO; }
    fclose(f);
    double t1 = now_s();

    // ---- phase 1: BGZF block table
    std::vector<Block> blocks;
    size_t usize = 0;
    for (size_t p = 0; p < raw.size();) {
        if (raw.size() - p < 18 || raw[p] != 0x1f || raw[p + 1] != 0x8b || raw[p + 2] != 8 || !(raw[p + 3] & 4)) {
            mmlst_set_error("%s: not a BGZF block at byte %zu (plain gzip / truncated file?)", path, p);
            return MMLST_E_BAM;
        }
        const uint32_t xlen = rd16(&raw[p + 10]);
        if (raw.size() - p < 12 + (size_t)xlen + 8) { mmlst_set_error("%s: truncated BGZF header at %zu", path, p); return MMLST_E_BAM; }
        int64_t bsize = -1;
        for (size_t q = p + 12; q + 4 <= p + 12 + xlen;) {
            const uint32_t slen = rd16(&raw[q + 2]);
            if (raw[q] == 66 && raw[q + 1] == 67 && slen == 2) bsize = rd16(&raw[q + 4]);
            q += 4 + slen;
        }
        if (bsize < 0) { mmlst_set_error("%s: BGZF block without BC subfield at %zu", path, p); return MMLST_E_BAM; }
        const size_t total = (size_t)bsize + 1;
        if (total < 12 + (size_t)xlen + 8 || raw.size() - p < total) { mmlst_set_error("%s: truncated BGZF block at %zu", path, p); return MMLST_E_BAM; }
        Block b;
        b.coff = p + 12 + xlen;
        b.clen = (uint32_t)(total - 12 - xlen - 8);
        b.crc = rd32(&raw[p + total - 8]);
        b.isize = rd32(&raw[p + total - 4]);
        b.uoff = usize;
        if (b.isize > 65536) { mmlst_set_error("%s: BGZF block with ISIZE %u > 64 KiB at %zu", path, b.isize, p); return MMLST_E_BAM; }
        usize += b.isize;
        if (b.isize) blocks.push_back(b);
        p += total;
    }

    // ---- phase 2: parallel raw inflate
    std::vector<uint8_t> u(usize + 8);
    Err err;
    // MMLST_INFLATE=zlib keeps zlib's inflate() for cross-checks; the default is the whole-buffer decoder of inflate_fast.cpp
    const char* inflate_env = getenv("MMLST_INFLATE");
    const bool use_zlib = inflate_env && strcmp(inflate_env, "zlib") == 0;
    parallel_for(blocks.size(), threads, 64, [&](size_t a, size_t e) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { err.set(MMLST_E_NOMEM, "inflateInit2 failed"); return; }
        for (size_t i = a; i < e && !err.code.load(std::memory_order_relaxed); ++i) {
            const Block& b = blocks[i];
            if (!use_zlib) {
                size_t got = 0;
                const int rc = mmlst_inflate_raw(&raw[b.coff], b.clen, &u[b.uoff], b.isize, &got);
                if (rc == MMLST_OK && got == b.isize &&
                    (!o.check_crc || (uint32_t)crc32(crc32(0L, Z_NULL, 0), &u[b.uoff], b.isize) == b.crc)) continue;
                // anything else gets zlib's verdict on the same block: a file zlib can read is never refused
            }
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef*>(&raw[b.coff]);
            zs.avail_in = b.clen;
            zs.next_out = &u[b.uoff];
            zs.avail_out = b.isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) { err.set(MMLST_E_BAM, "%s: inflate failed in BGZF block %zu (rc %d)", path, i, rc); break; }
            if (o.check_crc && (uint32_t)crc32(crc32(0L, Z_NULL, 0), &u[b.uoff], b.isize) != b.crc) { err.set(MMLST_E_BAM, "%s: CRC mismatch in BGZF block %zu", path, i); break; }
        }
        inflateEnd(&zs);
    });
    if (err.code) { mmlst_set_error("%s", err.msg); return err.code; }
    { std::vector<uint8_t>().swap(raw); }
    double t2 = now_s();

    // ---- phase 3: header, reference dictionary, record offsets
    std::unique_ptr<mmlst_bam> B(new mmlst_bam());
    if (usize < 12 || memcmp(u.data(), "BAM\1", 4) != 0) { mmlst_set_error("%s: not a BAM file (magic)", path); return MMLST_E_BAM; }
    size_t p = 4;
    const int32_t l_text = rdi32(&u[p]); p += 4;
    if (l_text < 0 || p + (size_t)l_text + 4 > usize) { mmlst_set_error("%s: bad l_text", path); return MMLST_E_BAM; }
    B->header_text.assign((const char*)&u[p], strnlen((const char*)&u[p], (size_t)l_text));
    p += (size_t)l_text;
    const int32_t n_ref = rdi32(&u[p]); p += 4;
    if (n_ref < 0) { mmlst_set_error("%s: bad n_ref", path); return MMLST_E_BAM; }
    B->ref_names.reserve(n_ref); B->ref_len.reserve(n_ref);
    for (int32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > usize) { mmlst_set_error("%s: truncated reference dictionary", path); return MMLST_E_BAM; }
        const int32_t ln = rdi32(&u[p]); p += 4;
        if (ln < 1 || p + (size_t)ln + 4 > usize) { mmlst_set_error("%s: truncated reference dictionary", path); return MMLST_E_BAM; }
        B->ref_names.emplace_back((const char*)&u[p], strnlen((const char*)&u[p], (size_t)ln));
        p += (size_t)ln;
        B->ref_len.push_back(rd32(&u[p])); p += 4;
    }
    for (int32_t i = 0; i < n_ref; ++i) { if (i) B->names_blob.push_back('\n'); B->names_blob += B->ref_names[i]; }
    std::vector<uint64_t> roff;
    roff.reserve((usize - p) / 200 + 16);
    while (p < usize) {
        if (p + 4 > usize) { mmlst_set_error("%s: truncated record length at byte %zu", path, p); return MMLST_E_BAM; }
        const int32_t bs = rdi32(&u[p]);
        if (bs < 32 || p + 4 + (size_t)bs > usize) { mmlst_set_error("%s: truncated / malformed record at byte %zu", path, p); return MMLST_E_BAM; }
        roff.push_back(p);
        p += 4 + (size_t)bs;
    }
    const size_t n = roff.size();
    if (n >= 0xFFFFFFFFull) { mmlst_set_error("%s: more than 2^32-1 records", path); return MMLST_E_RANGE; }

    // ---- phase 4: per-record fields (file order)
    struct Core { int32_t tid, pos; uint32_t reflen; uint16_t flag; int16_t as0, asn; uint8_t xm3, xmn; uint16_t qlen; uint8_t named_ok; };
    std::vector<Core> core(n);
    std::vector<uint64_t> qh(o.want_qhash ? 2 * n : 0);
    parallel_for(n, threads, 1 << 15, [&](size_t a, size_t e) {
        for (size_t i = a; i < e; ++i) {
            const uint8_t* r = &u[roff[i]];
            const uint32_t bs = rd32(r);
            const uint8_t* end = r + 4 + bs;
            Core c;
            c.tid = rdi32(r + 4); c.pos = rdi32(r + 8);
            const uint32_t l_name = r[12];
            const uint32_t n_cig = rd16(r + 16);
            c.flag = rd16(r + 18);
            const uint32_t l_seq = rd32(r + 20);
            const uint8_t* q = r + 36;
            const uint8_t* cig = q + l_name;
            const uint8_t* seq = cig + 4 * (size_t)n_cig;
            const uint8_t* aux = seq + (l_seq + 1) / 2 + l_seq;
            if (aux > end || l_name == 0) { err.set(MMLST_E_BAM, "%s: record %zu overruns its block_size", path, i); return; }
            if (c.tid < 0 || c.tid >= n_ref) {
                // RNAME '*': `species,gene,allele = read[2].split('_')` raises ValueError (metamlst.py:107)
                err.set(MMLST_E_BAM, "%s: record %zu has no reference (RNAME '*'): the reference crashes at metamlst.py:107", path, i); return;
            }
            if (c.flag & 0x2) { err.set(MMLST_E_PAIRED, "%s: record %zu is a proper-pair mate: htslib overlap handling (H2) is not implemented -- refusing", path, i); return; }
            if (c.pos < 0) { err.set(MMLST_E_BAM, "%s: record %zu has POS 0 on a reference", path, i); return; }
            if (o.want_qhash) { qh[2 * i] = hash64(q, l_name - 1); qh[2 * i + 1] = hash64b(q, l_name - 1); }
            uint64_t rl = 0, qlsum = 0;
            for (uint32_t k = 0; k < n_cig; ++k) {
                const uint32_t cw = rd32(cig + 4 * k), op = cw & 15u, ln = cw >> 4;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += ln;
                if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlsum += ln;
            }
            if (rl > 65535) { err.set(MMLST_E_RANGE, "%s: record %zu spans %llu reference bases (> 65535)", path, i, (unsigned long long)rl); return; }
            if (n_cig && l_seq && qlsum != l_seq) { err.set(MMLST_E_BAM, "%s: record %zu: CIGAR query length %llu != l_seq %u", path, i, (unsigned long long)qlsum, l_seq); return; }
            c.reflen = (uint32_t)rl;
            if (l_seq > 65535u) { err.set(MMLST_E_RANGE, "%s: record %zu: read of %u bases (len(SEQ) is carried as 16 bits)", path, i, l_seq); return; }
            const uint32_t ql = l_seq ? l_seq : 1u;  // SAM prints SEQ '*' when l_seq == 0: len() == 1 (metamlst.py:111,115)
            c.qlen = (uint16_t)std::min<uint32_t>(ql, 65535u);
            // aux walk: 1st and 4th field by POSITION (metamlst.py:109-110), AS / XM by NAME (cmseq/cmseq.py:545)
            int field = 0;
            int64_t v0 = 0, v3 = 0, vas = 0, vxm = 0;
            bool ok0 = false, ok3 = false, okas = false, okxm = false;
            for (const uint8_t* a2 = aux; a2 + 3 <= end; ++field) {
                const uint8_t t = a2[2];
                const uint8_t* val = a2 + 3;
                const size_t sz = aux_size(t, val, end);
                if (sz == 0 || val + sz > end) { err.set(MMLST_E_BAM, "%s: record %zu: malformed aux field %d", path, i, field); return; }
                int64_t v = 0;
                const bool isint = aux_int(t, val, &v);
                if (field == 0) { ok0 = isint; v0 = v; }
                if (field == 3) { ok3 = isint; v3 = v; }
                if (isint && a2[0] == 'A' && a2[1] == 'S' && !okas) { okas = true; vas = v; }
                if (isint && a2[0] == 'X' && a2[1] == 'M' && !okxm) { okxm = true; vxm = v; }
                a2 = val + sz;
            }
            if (lenient && (field < 4 || !ok0 || !ok3 || v0 < -32768 || v0 > 32767 || v3 < 0)) {
                v0 = -32768; v3 = 255;   // a cmseq caller never runs stage 1: values no stage-1 filter passes
            } else {
                if (field < 4 || !ok0 || !ok3) {
                    err.set(MMLST_E_BAM, "%s: record %zu: 1st / 4th aux field missing or not an integer: the reference crashes at metamlst.py:109-110", path, i);
                    return;
                }
                if (v0 < -32768 || v0 > 32767) { err.set(MMLST_E_RANGE, "%s: record %zu: 1st aux field %lld outside int16", path, i, (long long)v0); return; }
                if (v3 < 0) { err.set(MMLST_E_RANGE, "%s: record %zu: negative 4th aux field", path, i); return; }
            }
            c.as0 = (int16_t)v0;
            c.xm3 = (uint8_t)std::min<int64_t>(v3, 255);
            c.named_ok = okas && okxm && vas >= -32768 && vas <= 32767 && vxm >= 0;
            c.asn = c.named_ok ? (int16_t)vas : 0;
            c.xmn = c.named_ok ? (uint8_t)std::min<int64_t>(vxm, 255) : 0;
            core[i] = c;
        }
    });
    if (err.code) { mmlst_set_error("%s", err.msg); return err.code; }
    double t3 = now_s();

    // ---- phase 5: coordinate order
    auto key_of = [&](size_t i) { return ((uint64_t)(uint32_t)core[i].tid << 33) | (((uint64_t)(uint32_t)core[i].pos + 1ull) << 1) | ((core[i].flag >> 4) & 1u); };
    bool sorted = true, coord_sorted = true;
    for (size_t i = 1; i < n; ++i) {
        if (key_of(i) < key_of(i - 1)) sorted = false;
        if ((key_of(i) >> 1) < (key_of(i - 1) >> 1)) { coord_sorted = false; sorted = false; break; }
    }
    std::vector<uint32_t> order;
    if (o.assume_sorted) {
        // --presorted: the reference trusts the file (metamlst.py:236); htslib refuses records out of order
        if (!coord_sorted) { mmlst_set_error("%s: --presorted given but records are not coordinate-sorted (htslib: 'The input is not sorted')", path); return MMLST_E_UNSORTED; }
        sorted = true;
    }
    if (!sorted) {
        order.resize(n);
        std::vector<uint64_t> bucket((size_t)n_ref + 1, 0);
        for (size_t i = 0; i < n; ++i) ++bucket[(size_t)core[i].tid + 1];
        for (int32_t t = 0; t < n_ref; ++t) bucket[t + 1] += bucket[t];
        {
            std::vector<uint64_t> cur(bucket.begin(), bucket.end() - 1);
            for (size_t i = 0; i < n; ++i) order[cur[core[i].tid]++] = (uint32_t)i;  // stable inside a contig
        }
        parallel_for((size_t)n_ref, threads, 256, [&](size_t a, size_t e) {
            for (size_t t = a; t < e; ++t)
                std::stable_sort(order.begin() + bucket[t], order.begin() + bucket[t + 1],
                                 [&](uint32_t x, uint32_t y) { return (key_of(x) & 0x1ffffffffull) < (key_of(y) & 0x1ffffffffull); });
        });
    }
    B->presorted = sorted ? 1 : 0;
    auto src = [&](size_t k) -> size_t { return sorted ? k : order[k]; };
    double t4 = now_s();

    // ---- score stream (coordinate order)
    if (!B->tid.alloc(n * 4, pin) || !B->as0.alloc(n * 2, pin) || !B->xm3.alloc(n, pin) || !B->qlen.alloc(n * 2, pin) ||
        (!sorted && !B->orig_idx.alloc(n * 4, pin)) || (o.want_qhash && !B->qhash.alloc(n * 16, pin)) ||
        !B->contig_start.alloc(((size_t)n_ref + 1) * 8, false)) {
        mmlst_set_error("mmlst_bam_unpack: out of host memory"); return MMLST_E_NOMEM;
    }
    B->n_rec = n;
    parallel_for(n, threads, 1 << 16, [&](size_t a, size_t e) {
        uint32_t* tid = B->tid.as<uint32_t>(); int16_t* as0 = B->as0.as<int16_t>(); uint8_t* xm3 = B->xm3.as<uint8_t>();
        uint16_t* ql = B->qlen.as<uint16_t>(); uint32_t* oi = B->orig_idx.as<uint32_t>(); uint64_t* hq = B->qhash.as<uint64_t>();
        for (size_t k = a; k < e; ++k) {
            const size_t i = src(k);
            tid[k] = (uint32_t)core[i].tid; as0[k] = core[i].as0; xm3[k] = core[i].xm3; ql[k] = core[i].qlen;
            if (!sorted) oi[k] = (uint32_t)i;
            if (o.want_qhash) { hq[2 * k] = qh[2 * i]; hq[2 * k + 1] = qh[2 * i + 1]; }
        }
    });

    // run-length form of the score stream (mmlst_score_runs_dev): tid per run of equal tid, 5 B / record cross PCIe
    if (n && n < 0xffffff00ull) {
        uint32_t nr = 0;
        int rc = mmlst_build_runs(B->tid.as<uint32_t>(), n, nullptr, nullptr, nullptr, &nr);
        if (rc != MMLST_OK) return rc;
        if (!B->run_tid.alloc((size_t)nr * 4, pin) || !B->run_start.alloc(((size_t)nr + 1) * 4, pin) || !B->chunk_run.alloc(((n + 255) / 256) * 4, pin)) {
            mmlst_set_error("mmlst_bam_unpack: out of host memory"); return MMLST_E_NOMEM;
        }
        rc = mmlst_build_runs(B->tid.as<uint32_t>(), n, B->run_tid.as<uint32_t>(), B->run_start.as<uint32_t>(), B->chunk_run.as<uint32_t>(), &nr);
        if (rc != MMLST_OK) return rc;
        B->n_runs = nr;
        // len(SEQ) once per chunk when every chunk is uniform (3 B / record form, mmlst_score_runs_qc_dev)
        if (!B->chunk_qlen.alloc(((n + 255) / 256) * 2, pin)) { mmlst_set_error("mmlst_bam_unpack: out of host memory"); return MMLST_E_NOMEM; }
        int uniform = 0;
        rc = mmlst_chunk_qlen(B->qlen.as<uint16_t>(), n, B->chunk_qlen.as<uint16_t>(), &uniform);
        if (rc != MMLST_OK) return rc;
        B->qc = uniform != 0;
    }

    // ---- phase 6: pileup candidates (mapped flag), depth cap, rows
    std::vector<uint32_t> cand;  // sorted-order indices k
    cand.reserve(n);
    for (size_t k = 0; k < n; ++k) {
        if (core[src(k)].flag & 0x4) { ++B->n_unmapped_flag; continue; }
        cand.push_back((uint32_t)k);
    }
    const size_t nc = cand.size();
    std::vector<uint8_t> admitted(nc, 1);
    if (o.max_depth > 0 && nc) {
        std::vector<uint32_t> ctid(nc), crl(nc);
        std::vector<int32_t> cpos(nc);
        for (size_t j = 0; j < nc; ++j) { const Core& c = core[src(cand[j])]; ctid[j] = (uint32_t)c.tid; cpos[j] = c.pos; crl[j] = c.reflen; }
        const int rc = mmlst_depth_cap(ctid.data(), cpos.data(), crl.data(), nc, o.max_depth, o.sentinel_nodes, admitted.data());
        if (rc != MMLST_OK) return rc;
    }
    std::vector<uint32_t> adm;
    adm.reserve(nc);
    for (size_t j = 0; j < nc; ++j) { if (admitted[j]) adm.push_back(cand[j]); else ++B->n_dropped; }
    const size_t P = adm.size();
    std::vector<uint64_t> rowoff(P + 1, 0);
    uint32_t maxrw = 0;
    for (size_t j = 0; j < P; ++j) {
        const Core& cj = core[src(adm[j])];
        const uint32_t rw = row_words(touched_words((uint32_t)cj.pos, cj.reflen));
        rowoff[j + 1] = rowoff[j] + rw;
        maxrw = std::max(maxrw, rw);
    }
    const uint64_t kSlack = 8;
    if (rowoff[P] + kSlack >= (1ull << 32)) { mmlst_set_error("%s: plane array exceeds 2^32 words", path); return MMLST_E_RANGE; }
    if (!B->p_recs.alloc(P * sizeof(mmlst_prec), pin) || !B->planes.alloc((rowoff[P] + kSlack) * 4, pin)) { mmlst_set_error("mmlst_bam_unpack: out of host memory"); return MMLST_E_NOMEM; }
    B->n_prec = P; B->n_plane_words = rowoff[P] + kSlack; B->max_row_words = maxrw;
    {
        uint64_t* cs = B->contig_start.as<uint64_t>();
        size_t j = 0;
        for (int32_t t = 0; t <= n_ref; ++t) {
            while (j < P && core[src(adm[j])].tid < t) ++j;
            cs[t] = j;
        }
    }
    memset(B->planes.as<uint32_t>() + rowoff[P], 0, kSlack * 4);
    const int minqual = o.minqual;
    std::atomic<uint64_t> n_untagged{0};
    parallel_for(P, threads, 1 << 13, [&](size_t a, size_t e) {
        mmlst_prec* pr = B->p_recs.as<mmlst_prec>();
        uint32_t* planes = B->planes.as<uint32_t>();
        // BAM 4-bit base codes "=ACMGRSVTWYHKDBN": A=1 C=2 G=4 T=8 -> 2-bit code; everything else is a counted non-ACGT base
        static const int8_t code2[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
        for (size_t j = a; j < e; ++j) {
            const size_t i = src(adm[j]);
            const Core& c = core[i];
            const uint8_t* r = &u[roff[i]];
            const uint32_t l_name = r[12], n_cig = rd16(r + 16), l_seq = rd32(r + 20);
            const uint8_t* cig = r + 36 + l_name;
            const uint8_t* seq = cig + 4 * (size_t)n_cig;
            const uint8_t* qual = seq + (l_seq + 1) / 2;
            mmlst_prec m;
            memset(&m, 0, sizeof(m));
            const uint32_t nw = touched_words((uint32_t)c.pos, c.reflen);
            m.pos = c.pos; m.row_off = (uint32_t)rowoff[j]; m.reflen = (uint16_t)c.reflen; m.as_named = c.asn; m.xm_named = c.xmn;
            m.nw = (uint16_t)nw;
            pr[j] = m;
            uint32_t* row = planes + rowoff[j];
            const uint32_t rw = row_words(nw);
            for (uint32_t w = 0; w < rw; ++w) row[w] = 0;
            if (c.reflen == 0) continue;
            if (!c.named_ok) {
                // get_tag('AS') / get_tag('XM') raise KeyError for the first ACGT base of this read (cmseq/cmseq.py:545) -- when a tag filter is given
                if (!lenient) {
                    err.set(MMLST_E_BAM, "%s: record %zu enters the pileup without integer AS:i / XM:i tags (pysam get_tag KeyError, cmseq/cmseq.py:545)", path, i);
                    return;
                }
                n_untagged.fetch_add(1, std::memory_order_relaxed);
            }
            if (l_seq && qual[0] == 0xff) {
                err.set(MMLST_E_BAM, "%s: record %zu has no base qualities: query_qualities is None (TypeError at cmseq/cmseq.py:538)", path, i);
                return;
            }
            uint32_t x = (uint32_t)c.pos & 31u, y = 0;  // column inside the row (rows are aligned to the contig's words), query index
            // the three planes of the 32-column word being filled stay in registers and reach the row once per word
            uint32_t wi = x >> 5, pv = 0, p1 = 0, p0 = 0;
            for (uint32_t k = 0; k < n_cig; ++k) {
                const uint32_t cw = rd32(cig + 4 * k), op = cw & 15u, ln = cw >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    for (uint32_t t = 0; t < ln; ++t, ++x, ++y) {
                        if ((x >> 5) != wi) {
                            uint32_t* w3 = row + 3 * wi;
                            w3[0] |= pv; w3[1] |= p1; w3[2] |= p0;
                            wi = x >> 5; pv = p1 = p0 = 0;
                        }
                        if (y >= l_seq) continue;               // qpos beyond l_qseq: quality 0 (pysam pileup_base_qual_skip)
                        if ((int)qual[y] < minqual) continue;   // H3: not in column.pileups at all
                        const int nib = (seq[y >> 1] >> ((~y & 1u) << 2)) & 15;
                        const int cd = code2[nib];
                        const uint32_t bit = 1u << (x & 31u);
                        if (cd >= 0) { pv |= bit; if (cd & 2) p1 |= bit; if (cd & 1) p0 |= bit; }
                        else p0 |= bit;                         // V=0, B0=1: bin N
                    }
                } else if (op == 1 || op == 4) y += ln;
                else if (op == 2 || op == 3) x += ln;
            }
            if (pv | p1 | p0) {
                uint32_t* w3 = row + 3 * wi;
                w3[0] |= pv; w3[1] |= p1; w3[2] |= p0;
            }
        }
    });
    if (err.code) { mmlst_set_error("%s", err.msg); return err.code; }
    double t5 = now_s();
    B->minqual = o.minqual; B->max_depth = o.max_depth;
    B->n_untagged = n_untagged.load();
    B->t_read = t1 - t0; B->t_inflate = t2 - t1; B->t_parse = t3 - t2; B->t_sort = t4 - t3; B->t_pack = t5 - t4;
    *out = B.release();
    return MMLST_OK;
}

extern "C" int mmlst_bam_info(const mmlst_bam* b, mmlst_bam_info_t* info) {
    if (!b || !info) { mmlst_set_error("mmlst_bam_info: null argument"); return MMLST_E_ARG; }
    memset(info, 0, sizeof(*info));
    info->soa.tid = b->tid.as<uint32_t>(); info->soa.as0 = b->as0.as<int16_t>(); info->soa.xm3 = b->xm3.as<uint8_t>();
    info->soa.qlen = b->qlen.as<uint16_t>(); info->soa.orig_idx = b->presorted ? nullptr : b->orig_idx.as<uint32_t>();
    info->soa.n_rec = b->n_rec;
    info->soa.p_recs = b->p_recs.as<mmlst_prec>(); info->soa.planes = b->planes.as<uint32_t>();
    info->soa.n_prec = b->n_prec; info->soa.n_plane_words = b->n_plane_words; info->soa.max_row_words = b->max_row_words;
    info->soa.contig_start = b->contig_start.as<uint64_t>(); info->soa.n_ref = (uint32_t)b->ref_names.size();
    if (b->n_runs) {
        info->soa.n_runs = b->n_runs; info->soa.run_tid = b->run_tid.as<uint32_t>();
        info->soa.run_start = b->run_start.as<uint32_t>(); info->soa.chunk_run = b->chunk_run.as<uint32_t>();
        if (b->qc) info->soa.chunk_qlen = b->chunk_qlen.as<uint16_t>();
    }
    info->qhash = b->qhash.as<uint64_t>();
    info->ref_len = b->ref_len.data();
    info->ref_names = b->names_blob.c_str();
    info->header_text = b->header_text.c_str();
    info->n_dropped_by_cap = b->n_dropped; info->n_unmapped_flag = b->n_unmapped_flag;
    info->presorted = b->presorted; info->minqual = b->minqual; info->max_depth = b->max_depth;
    info->n_untagged = b->n_untagged;
    info->seconds[0] = b->t_read; info->seconds[1] = b->t_inflate; info->seconds[2] = b->t_parse; info->seconds[3] = b->t_sort; info->seconds[4] = b->t_pack;
    return MMLST_OK;
}
