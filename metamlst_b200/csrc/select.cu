// Best-allele selection ON THE DEVICE (replaces the host round trip between stage 1 and stage 2).
//
// Reference semantics (metamlst.py:133-151, 184-206, 213-220, 244):
//   per locus   maxLen = max hit count; localScore = sum(AS) - (maxLen - n) * penalty; avg = round(localScore / n, 1)
//   per locus   chosen = lowest int(allele) among the alleles whose ROUNDED average equals the locus maximum
//   per species processed only if int(detected_loci / loci_in_db * 100) >= nloci; order of species and of loci inside
//               a species = dict insertion order = ascending first passing record (H5)
// H6: Python's round(x, 1) is the correctly rounded decimal (ties on the exact binary value go to the even digit).
// Two rounded values are equal iff their integer tenths T are equal, and T is computed here exactly from the bits of
// the IEEE double x = (double)localScore / (double)n (same correctly rounded division as Python's float division).
#include "common.cuh"

namespace {

constexpr unsigned long long T_BIAS = 1ull << 62;

// integer nearest to 10*x, ties to even, computed exactly (x finite)
__device__ __forceinline__ long long round_tenths(double x) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(x));
    const bool neg = bits >> 63;
    const int ex = int((bits >> 52) & 0x7ff);
    unsigned long long m = bits & ((1ull << 52) - 1);
    int e;
    if (ex == 0) { e = -1074; } else { m |= 1ull << 52; e = ex - 1075; }
    unsigned long long t;
    const unsigned long long y = m * 10ull;  // < 2^57
    if (e >= 0) {
        t = (e >= 6) ? 0x3fffffffffffffffull : (y << e);  // |x| >= 2^58: clamp (scores never get there)
    } else {
        const int s = -e;
        if (s > 63) {
            t = 0;  // y < 2^57 <= half
        } else {
            const unsigned long long q = y >> s;
            const unsigned long long rem = y & ((1ull << s) - 1ull);
            const unsigned long long half = 1ull << (s - 1);
            t = q + ((rem > half || (rem == half && (q & 1ull))) ? 1ull : 0ull);
        }
    }
    return neg ? -static_cast<long long>(t) : static_cast<long long>(t);
}

struct SelArgs {
    const long long* sum_as; const uint32_t* n_hit; const uint32_t* first_idx;
    const uint32_t* locus_of; const uint32_t* allele_num; uint32_t n_ref;
    const uint32_t* species_of_locus; const uint32_t* genes_in_db; uint32_t n_loci, n_species;
    int penalty, nloci_pct;
    const unsigned long long* contig_start; const uint32_t* ref_len; const unsigned long long* db_off;
    uint32_t chunk_records, target_chunks;
    // scratch
    uint32_t* maxlen; unsigned long long* best_t; uint32_t* lfirst; unsigned long long* chosen_key;
    // outputs
    uint32_t* header;  // [0]=n_chosen [1]=n_chunks [2]=total_cols [3]=error flags [4]=chunk_records used
    uint32_t* chosen_tid; uint32_t* chosen_species; uint32_t* col_off; unsigned long long* db_start;
    mmlst_chunk* chunks; uint32_t max_chunks;
};

__global__ void sel_pass_a(const SelArgs a) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_ref; t += gridDim.x * blockDim.x) {
        const uint32_t n = a.n_hit[t];
        if (!n) continue;
        const uint32_t l = a.locus_of[t];
        atomicMax(a.maxlen + l, n);
        atomicMin(a.lfirst + l, a.first_idx[t]);
    }
}

__device__ __forceinline__ long long tenths_of(const SelArgs& a, uint32_t t, uint32_t n, uint32_t l) {
    long long score = a.sum_as[t];
    const uint32_t mx = a.maxlen[l];
    if (n != mx) score -= static_cast<long long>(mx - n) * a.penalty;
    return round_tenths(__ddiv_rn(static_cast<double>(score), static_cast<double>(n)));
}

__global__ void sel_pass_b(const SelArgs a) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_ref; t += gridDim.x * blockDim.x) {
        const uint32_t n = a.n_hit[t];
        if (!n) continue;
        const uint32_t l = a.locus_of[t];
        atomicMax(a.best_t + l, static_cast<unsigned long long>(tenths_of(a, t, n, l)) + T_BIAS);
    }
}

__global__ void sel_pass_c(const SelArgs a) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_ref; t += gridDim.x * blockDim.x) {
        const uint32_t n = a.n_hit[t];
        if (!n) continue;
        const uint32_t l = a.locus_of[t];
        if (static_cast<unsigned long long>(tenths_of(a, t, n, l)) + T_BIAS == a.best_t[l])
            atomicMin(a.chosen_key + l, (static_cast<unsigned long long>(a.allele_num[t]) << 32) | t);
    }
}

// one CTA: nloci gate, H5 ordering, column offsets, chunk descriptors
__global__ void __launch_bounds__(1024) sel_finalize(const SelArgs a) {
    extern __shared__ uint32_t sm[];
    uint32_t* s_detected = sm;                       // [n_species]
    uint32_t* s_first = sm + a.n_species;            // [n_species]
    uint32_t* s_pass = sm + 2 * a.n_species;         // [n_species]
    __shared__ uint32_t s_n, s_err;
    __shared__ unsigned long long s_totrec;
    for (uint32_t i = threadIdx.x; i < a.n_species; i += blockDim.x) { s_detected[i] = 0; s_first[i] = 0xffffffffu; }
    if (threadIdx.x == 0) { s_n = 0; s_err = 0; s_totrec = 0; }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < a.n_loci; l += blockDim.x) {
        if (a.chosen_key[l] != ~0ull) {
            const uint32_t sp = a.species_of_locus[l];
            atomicAdd(s_detected + sp, 1u);
            atomicMin(s_first + sp, a.lfirst[l]);
        }
    }
    __syncthreads();
    for (uint32_t sp = threadIdx.x; sp < a.n_species; sp += blockDim.x) {
        const uint32_t det = s_detected[sp], tot = a.genes_in_db[sp];
        uint32_t pass = 0;
        if (det) {
            if (tot < det) atomicOr(&s_err, 1u);  // "Database is broken" (metamlst.py:188-190)
            else pass = int((double(det) / double(tot)) * 100.0) >= a.nloci_pct;  // metamlst.py:206
        }
        s_pass[sp] = pass;
    }
    __syncthreads();
    // rank of every kept locus by (species first record, locus first record)
    for (uint32_t l = threadIdx.x; l < a.n_loci; l += blockDim.x) {
        if (a.chosen_key[l] == ~0ull) continue;
        const uint32_t sp = a.species_of_locus[l];
        if (!s_pass[sp]) continue;
        const unsigned long long key = (static_cast<unsigned long long>(s_first[sp]) << 32) | a.lfirst[l];
        uint32_t rank = 0;
        for (uint32_t m = 0; m < a.n_loci; ++m) {
            if (a.chosen_key[m] == ~0ull) continue;
            const uint32_t sp2 = a.species_of_locus[m];
            if (!s_pass[sp2]) continue;
            const unsigned long long k2 = (static_cast<unsigned long long>(s_first[sp2]) << 32) | a.lfirst[m];
            rank += (k2 < key) || (k2 == key && m < l);
        }
        const uint32_t tid = static_cast<uint32_t>(a.chosen_key[l] & 0xffffffffu);
        a.chosen_tid[rank] = tid;
        a.chosen_species[rank] = sp;
        a.db_start[rank] = a.db_off[tid];
        atomicAdd(&s_n, 1u);
        atomicAdd(&s_totrec, a.contig_start[tid + 1] - a.contig_start[tid]);
    }
    __syncthreads();
    // per-locus chunk counts -> exclusive prefix (serial over <= 8192 loci), then all threads write the descriptors
    uint32_t* s_cbase = sm + 3 * a.n_species;  // [n_loci + 1]
    __shared__ uint32_t s_cr, s_nch, s_col;
    if (threadIdx.x == 0) {
        uint32_t cr = a.chunk_records;
        if (cr == 0) {  // same rule as mmlst_chunk_records(): >= 4 chunks per SM, whole 512-record tiles, <= 63 tiles
            const unsigned long long target = a.target_chunks;
            unsigned long long c = (s_totrec + target - 1) / target;
            c = ((c + 511ull) / 512ull) * 512ull;
            if (c < 512ull) c = 512ull;
            if (c > 63ull * 512ull) c = 63ull * 512ull;
            cr = static_cast<uint32_t>(c);
        }
        uint32_t col = 0, nch = 0;
        for (uint32_t i = 0; i < s_n; ++i) {
            const uint32_t tid = a.chosen_tid[i];
            a.col_off[i] = col;
            s_cbase[i] = nch;
            const unsigned long long nrec = a.contig_start[tid + 1] - a.contig_start[tid];
            nch += static_cast<uint32_t>((nrec + cr - 1) / cr);
            col += a.ref_len[tid];
        }
        a.col_off[s_n] = col;
        s_cbase[s_n] = nch;
        s_cr = cr; s_nch = nch; s_col = col;
        if (nch > a.max_chunks) s_err |= 2u;
    }
    __syncthreads();
    const uint32_t nch = min(s_nch, a.max_chunks), cr = s_cr, nsel = s_n;
    for (uint32_t c = threadIdx.x; c < nch; c += blockDim.x) {
        uint32_t lo = 0, hi = nsel;  // last i with s_cbase[i] <= c
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_cbase[mid] <= c) lo = mid; else hi = mid; }
        const uint32_t tid = a.chosen_tid[lo];
        const unsigned long long r0 = a.contig_start[tid], r1 = a.contig_start[tid + 1];
        const unsigned long long b = r0 + static_cast<unsigned long long>(c - s_cbase[lo]) * cr;
        mmlst_chunk ck;
        ck.rec_begin = static_cast<uint32_t>(b);
        ck.rec_end = static_cast<uint32_t>(b + cr < r1 ? b + cr : r1);
        ck.col_base = a.col_off[lo]; ck.contig_len = a.ref_len[tid]; ck.plane_delta = 0;
        ck.reserved[0] = ck.reserved[1] = ck.reserved[2] = 0;
        a.chunks[c] = ck;
    }
    if (threadIdx.x == 0) { a.header[0] = nsel; a.header[1] = nch; a.header[2] = s_col; a.header[3] = s_err; a.header[4] = cr; }
}

}  // namespace

// see include/mmlst.h
extern "C" int mmlst_select_dev(const int64_t* sum_as, const uint32_t* n_hit, const uint32_t* first_idx, const uint32_t* locus_of,
                                const uint32_t* allele_num, uint32_t n_ref, const uint32_t* species_of_locus,
                                const uint32_t* genes_in_db, uint32_t n_loci, uint32_t n_species, int penalty, int nloci_pct,
                                const uint64_t* contig_start, const uint32_t* ref_len, const uint64_t* db_off,
                                uint32_t chunk_records, void* scratch, size_t scratch_bytes, uint32_t* header,
                                uint32_t* chosen_tid, uint32_t* chosen_species, uint32_t* col_off, uint64_t* db_start,
                                mmlst_chunk* chunks, uint32_t max_chunks, void* stream) {
    if (!sum_as || !n_hit || !first_idx || !locus_of || !allele_num || !species_of_locus || !genes_in_db || !contig_start || !ref_len ||
        !db_off || !scratch || !header || !chosen_tid || !chosen_species || !col_off || !db_start || !chunks) {
        mmlst_set_error("mmlst_select_dev: null pointer");
        return MMLST_E_ARG;
    }
    if (n_loci > 8192 || n_species > 4096) { mmlst_set_error("mmlst_select_dev: more than 8192 loci / 4096 species"); return MMLST_E_RANGE; }
    const size_t need = static_cast<size_t>(n_loci) * 24 + 8;
    if (scratch_bytes < need || (reinterpret_cast<uintptr_t>(scratch) & 7)) { mmlst_set_error("mmlst_select_dev: scratch needs %zu bytes, 8-byte aligned", need); return MMLST_E_ARG; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // scratch layout: best_t u64[n_loci] | maxlen u32[n_loci]  (zeroed)  ||  chosen_key u64[n_loci] | lfirst u32[n_loci] (0xff)
    uint8_t* p = static_cast<uint8_t*>(scratch);
    SelArgs a;
    a.sum_as = reinterpret_cast<const long long*>(sum_as); a.n_hit = n_hit; a.first_idx = first_idx; a.locus_of = locus_of;
    a.allele_num = allele_num; a.n_ref = n_ref; a.species_of_locus = species_of_locus; a.genes_in_db = genes_in_db;
    a.n_loci = n_loci; a.n_species = n_species; a.penalty = penalty; a.nloci_pct = nloci_pct;
    a.contig_start = reinterpret_cast<const unsigned long long*>(contig_start); a.ref_len = ref_len;
    a.db_off = reinterpret_cast<const unsigned long long*>(db_off); a.chunk_records = chunk_records;
    a.best_t = reinterpret_cast<unsigned long long*>(p);
    a.maxlen = reinterpret_cast<uint32_t*>(p + static_cast<size_t>(n_loci) * 8);
    a.chosen_key = reinterpret_cast<unsigned long long*>(p + static_cast<size_t>(n_loci) * 12 + ((n_loci & 1) ? 4 : 0));
    a.lfirst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(a.chosen_key) + static_cast<size_t>(n_loci) * 8);
    a.target_chunks = static_cast<uint32_t>(mmlst_num_sms()) * 4u;
    a.header = header; a.chosen_tid = chosen_tid; a.chosen_species = chosen_species; a.col_off = col_off;
    a.db_start = reinterpret_cast<unsigned long long*>(db_start); a.chunks = chunks; a.max_chunks = max_chunks;
    CUDA_TRY(cudaMemsetAsync(a.best_t, 0, static_cast<size_t>(n_loci) * 12, s));
    CUDA_TRY(cudaMemsetAsync(a.chosen_key, 0xff, static_cast<size_t>(n_loci) * 12, s));
    size_t gsz = (static_cast<size_t>(n_ref) + 255) / 256;
    const size_t gcap = static_cast<size_t>(mmlst_num_sms()) * 4;
    if (gsz > gcap) gsz = gcap;
    if (gsz < 1) gsz = 1;
    const unsigned grid = static_cast<unsigned>(gsz);
    sel_pass_a<<<grid, 256, 0, s>>>(a);
    sel_pass_b<<<grid, 256, 0, s>>>(a);
    sel_pass_c<<<grid, 256, 0, s>>>(a);
    sel_finalize<<<1, 1024, sizeof(uint32_t) * (3 * n_species + n_loci + 2), s>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
