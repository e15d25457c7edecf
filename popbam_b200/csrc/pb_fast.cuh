// pb_fast.cuh -- the counting pileup: the default formulation of the pileup / call / site stage.
// Used when the raw-depth cap cannot bind (k_depth_bound), the caller did not ask for the
// per-(site,sample) words, and min_depth / min_snpQ are positive; k_pileup_call otherwise.
//
// Why.  k_pileup_call runs call_base's whole bookkeeping (histogram update, per-base totals, sum of mapq^2) for every
// base of every cell.  But ~99 % of the cells are "easy": every passing base equals the reference base (or all but
// one do) and a few COUNTS prove that call_base would call the cell homozygous reference (pb_need_entry,
// pb_one_stray_entry in pb_walk.cuh), so all the site needs from the cell is qfilter's coverage bit
// (pop_utils.cpp:102-120).  The counts per (position, sample) cell:
//     k     bases that pass call_base's filters (popbam.cpp:268-284)
//     khi   those with quality level >= 30 (baseQ' >= 30 and mapQ >= 30)
//     m     stray bases (passing, different from the reference base), and whether one of them is a high-quality one
//     lowq  a passing base of a read with mapQ < min_rmsQ is present (else rms >= min_rmsQ is guaranteed)
// and the rule, with per-depth tables built once per context (k_fast_tables):
//     easy = k > 0 and no lowq and ( m == 0 and (depth k alone suffices  or  khi >= hneed[k])
//                                 or m == 1 and (one stray base is harmless at depth k, or a low-quality one is) )
//
// The kernel that touches the reads is in pb_pile.cuh: k_pile_reads (the scatter of the reads into byte counters in shared
// memory, all samples of a block of positions in one CTA, reads in file order; the classification of the cells; the base
// codes of the cells that are not settled by counts).  Here: the per-context tables of the rule
// (k_set_levels / k_need_raw / k_fast_tables), the reference code bytes (k_ref_codes), k_hard_cells (the hard cells,
// one per thread, with the exact machinery of pb_cell.cuh / pb_walk.cuh) and k_fast_sites (sites from the two).
// History: round 1 built bit-planes in a separate streaming kernel and walked them per (record, 32 positions) with
// bit-sliced counters (420 M warp instructions per 2.3 Mb shard); the first counting kernel partitioned the reads by
// sample into segment records first (one CTA per sample and position block, 441 M warp instructions, 63 % of them
// outside the scatter loop: per-CTA set-up for ~280 records, per-record staging slots, a strip index).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_kernels.cuh"

#define PB_H_QUALITY 30            // quality value of the khi count

// ---- quality levels of this path.  Its proofs (pb_need_entry, pb_one_stray_entry) hold for any level SET that contains
// every level a passing base can have, so it does not look for the values present (round 1 paid two shared-memory
// stores per base for that): the set is the RANGE [qlo, qhi] with
//     qlo = clamp(min(min_baseQ, min_mapQ), 4, 63)   -- call_base keeps a base only if baseQ' >= min_baseQ and
//                                                       mapQ >= min_mapQ, and codes clamp(min(baseQ', mapQ), 4, 63)
//     qhi = the context's quality ceiling            -- an ASSUMPTION: k_pile_reads checks every quality byte it
//                                                       touches against it (one add and one OR per four bytes) and
//                                                       reports a violation; the host then raises the ceiling and
//                                                       runs the region again.
// beta[q][n][k] grows with q (checked for the reference's tables), so the lower bounds are decided by qlo and the
// one-stray-base rule by qhi; the values in between cost nothing.
__global__ void k_set_levels(PbCounters *ctr, int qlo, int qhi) {
    const int q = threadIdx.x, nl = qhi - qlo + 1;
    ctr->qrank[q] = (unsigned char)(q < qlo ? 0 : q > qhi ? nl - 1 : q - qlo);
    if (q < nl) ctr->qval[q] = (unsigned char)(qlo + q);
    if (q == 0) ctr->n_levels = nl;
}
// need_raw[q][k] = pb_need_entry for the level set {q, ..., 63} (k_hard_cells ranks the levels of every cell itself)
__global__ void __launch_bounds__(256) k_need_raw(const double *__restrict__ fk, const double *__restrict__ beta,
                                                  const double *__restrict__ lhet, uint8_t *__restrict__ need_raw /* [64][256] */) {
    const int q = blockIdx.x, k = threadIdx.x;
    PbIota lv{q};
    need_raw[q * 256 + k] = q >= 1 ? pb_need_entry(0, 64 - q, lv, k, fk, beta, lhet) : 0;
}

// Per-depth tables of the rule above.  The level set is {qlo, ..., 63} (k_set_levels; need_raw[qlo] is pb_need_entry for
// exactly that set), so the proofs about unanimous cells hold whatever qualities occur.  Only the ONE stray base of the
// one-stray-base rule is assumed to have a quality of at most `ceiling`; k_pile_reads looks at the quality of every stray
// base it finds (they are rare) and leaves a cell with a higher one to k_hard_cells.
//   flags[k]  bit 0  need_raw[qlo][k] <= k: k unanimous bases always take the shortcut (every base is at least level qlo)
//             bit 1  one stray base of quality <= ceiling cannot change the homozygous call at depth k (pb_one_stray_entry)
//             bit 2  the same for a stray base below the khi level
//             bit 3  k >= min_depth (qfilter; k <= max_depth holds because the cap cannot bind)
//   hneed[k]  khi >= hneed[k] bases at or above the khi level prove the shortcut (0: never)
//   altok[k]  bit b: k unanimous bases b that differ from the reference base are a derived allele -- segbase keeps the
//             homozygote because the shortcut's snpQ (a function of k and b alone, pb_unanimous_result) reaches min_snpQ
//             (pop_utils.cpp:139-150; below min_snpQ segbase reverts the call with arithmetic on the packed word, SURVEY Q8:
//             those cells go to k_hard_cells)
//             bit 6  qfilter's rms >= min_rmsQ holds for k bases of reads whose mapping quality is in the top class (below)
//   rlo, rhi, rh   three runs of depths [rlo, rhi] (all below 128) in which a cell without stray bases, flags or mapping-quality
//             deficits is settled by packed compares: bits 3 and 6 hold throughout, and bit 0 in run 0 (rh[0] = 0),
//             khi >= rh[i] >= hneed[k] in the others
//   dq[mapq], m0sq, qstep, rthr[k]   qfilter (pop_utils.cpp:102-120) wants rms = (u64)(sqrtf(sum mapq^2 / k) + 0.499) >= min_rmsQ,
//             i.e. sum mapq^2 >= rthr[k] (k_rms_table: every step of that expression is monotone).  A read's mapq^2 is
//             bounded below by m0sq + qstep * (3 - dq[mapq]) -- four classes between min_mapQ^2 and (min_rmsQ + 4)^2 -- and
//             the cells count the DEFICITS dq of their passing bases (reads of the top class, mapq >= min_rmsQ + 4, add
//             nothing: in ordinary data that is nearly every read), so sum mapq^2 >= k * m0sq + qstep * (3 k - D).  A few
//             reads between min_mapQ and min_rmsQ no longer cost their cells the easy path; sound for k <= 84 (D is a
//             byte); deeper cells go to k_hard_cells.
struct PbFastTables {
    uint8_t flags[256]; uint8_t hneed[256]; uint8_t altok[256]; uint8_t dq[256];
    uint8_t rlo[3], rhi[3], rh[3], pad[7];
    int32_t m0sq, qstep, pad2[2];
    int32_t rthr[256];
};
#define PB_RMS_KMAX 84            // 3 * 84 + the carries a neighbouring byte's overflow can add stays below 256
__global__ void __launch_bounds__(256) k_fast_tables(const PbCounters *__restrict__ ctr, const uint8_t *__restrict__ need_raw,
                                                     const double *__restrict__ fk, const double *__restrict__ beta,
                                                     const double *__restrict__ lhet, int min_depth, int min_snpQ, int ceiling, int min_mapQ, int min_rmsQ,
                                                     const int32_t *__restrict__ rms_thr, PbFastTables *__restrict__ tab) {
    const int nl = ctr->n_levels, k = threadIdx.x, qlo = ctr->qval[0];
    const int top = max(0, min(nl, ceiling - qlo + 1)), hi = max(0, min(top, PB_H_QUALITY - qlo));      // stray-base levels [0, top), low-quality ones [0, hi)
    uint32_t f = 0;
    const int nd = k >= 1 ? need_raw[qlo * 256 + k] : 0;
    if (nd && nd <= k) f |= 1u;
    if (top > 0 && pb_one_stray_entry(nl, ctr->qval, k, fk, beta, lhet, 0, top)) f |= 2u;
    if (hi > 0 && pb_one_stray_entry(nl, ctr->qval, k, fk, beta, lhet, 0, hi)) f |= 4u;
    if (k >= min_depth) f |= 8u;
    // mapping-quality classes: lower bounds of mapq^2 in four steps from min_mapQ^2 to mtop^2, mtop = min_rmsQ + 4
    const int m0 = max(0, min(250, min_mapQ)), m0sq = m0 * m0, mtop = max(m0 + 3, min(255, min_rmsQ + 4)), qstep = (mtop * mtop - m0sq) / 3;
    {
        const int mq = k;                                  // (this thread's table entry: mapping quality k)
        const int q = mq >= m0 ? min(3, (mq * mq - m0sq) / qstep) : 0;
        tab->dq[mq] = (uint8_t)(3 - q);
    }
    const int thr = rms_thr[k];
    tab->rthr[k] = thr;
    if (k >= 1 && k <= PB_RMS_KMAX && (long long)k * m0sq + (long long)qstep * 3 * k >= (long long)thr) f |= 0x40u;
    if (k == 0) { tab->m0sq = m0sq; tab->qstep = qstep; }
    tab->flags[k] = (uint8_t)f;
    tab->hneed[k] = k >= 1 ? need_raw[PB_H_QUALITY * 256 + k] : 0;
    uint32_t ok = 0;
    if (k >= 1)
        for (int b = 0; b < 4; ++b) {
            const uint64_t cb = pb_unanimous_result(lhet, k, b, 0);
            if ((cb >> 48) == 0 && (int)((cb >> 32) & 0xffff) >= min_snpQ && pb_unanimous_het(lhet, k, b) > 0.0f) ok |= 1u << b;      // (no carry out of the 16-bit snpQ field)
        }
    tab->altok[k] = (uint8_t)ok;
    __shared__ uint8_t fs[256];
    fs[k] = (uint8_t)f;
    __syncthreads();
    __shared__ uint8_t hs[256];
    hs[k] = tab->hneed[k];
    __syncthreads();
    if (k == 0) {
        int best_lo = 1, best_hi = 0;
        for (int i = 1; i < 128;) {                        // (packed compares: counts below 128)
            if ((fs[i] & 0x49) != 0x49) { ++i; continue; }
            int j = i;
            while (j + 1 < 128 && (fs[j + 1] & 0x49) == 0x49) ++j;
            if (j - i > best_hi - best_lo) { best_lo = i; best_hi = j; }
            i = j + 1;
        }
        tab->rlo[0] = (uint8_t)best_lo; tab->rhi[0] = (uint8_t)best_hi; tab->rh[0] = 0;
        // the depths behind that run, as long as the count of high-quality bases they ask for stays small
        int at = best_hi >= best_lo ? best_hi + 1 : 1;
        const int lim[2] = {8, 16};
        for (int r = 0; r < 2; ++r) {
            int lo = at, hi = at - 1, hmax = 0;
            while (hi + 1 < 128 && (fs[hi + 1] & 0x48) == 0x48 && hs[hi + 1] > 0 && hs[hi + 1] <= lim[r]) { ++hi; hmax = max(hmax, (int)hs[hi]); }
            tab->rlo[r + 1] = (uint8_t)lo; tab->rhi[r + 1] = (uint8_t)hi; tab->rh[r + 1] = (uint8_t)max(hmax, 1);
            at = hi + 1;
        }
    }
}

// Reference codes of a contig, one NIBBLE per position in the code of seq4 (bam.h:245-258: A 1, C 2, G 4, T 8), eight
// positions per 32-bit word, position p in bits 4 (p & 7) of word p >> 3; 0 for anything that is not an upper-case
// A/C/G/T (the only thing a called base can equal, pop_utils.cpp:139 / SURVEY Q7) and for the padding behind the contig.
// k_pile_reads XORs the nibbles of a read against them: a non-zero nibble is a stray base.
#define PB_REFCODE_PAD 8192
__global__ void __launch_bounds__(256) k_ref_codes(const char *__restrict__ ref, int64_t ref_len, uint32_t *__restrict__ code) {
    const int64_t nw = (ref_len + PB_REFCODE_PAD) >> 3;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int j = 0; j < 8; ++j) {
            const int64_t i = 8 * w + j;
            const int c = i < ref_len ? (int)(unsigned char)ref[i] : 'N';
            v |= (uint32_t)(c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0) << (4 * j);
        }
        code[w] = v;
    }
}

struct PbHardArgs {
    const uint4 *cells;                      // directory and base codes written by k_pile_reads
    const uint16_t *codes;
    const char *ref;
    int64_t ref_len;
    int span_beg, span_end;
    const int32_t *win_beg, *win_end;
    int n_windows;
    int n_samples, n_strips;
    int min_depth, max_depth, min_rmsQ, min_snpQ;
    int het_mode;
    const double *fk, *beta, *lhet;
    const PbCounters *ctr;
    const uint8_t *need_raw;                 // [64][256] k_need_raw
    uint64_t *acc_cov;                       // [span] coverage bits: the easy cells' (k_pile_reads), the hard cells' are added here
    uint32_t *acc_cnt4;                      // [span] derived-base counts of the hard cells (zeroed by k_pile_reads)
    uint64_t *site_type;                     // [span] derived-allele bits                   (zeroed by k_pile_reads; the hard cells are the only writers)
    uint8_t *site_flag;
};

#define PB_HARD_THREADS 128
#define PB_HARD_LCAP 32            // distinct quality levels of one cell held in shared memory (more: local-memory path)
static inline size_t pb_hard_smem() { return (size_t)2 * PB_HARD_LCAP * PB_HARD_THREADS * 4 + (size_t)PB_HARD_LCAP * PB_HARD_THREADS + 256 * 8; }

// the cell's own quality levels, one byte column per thread
struct PbCellLevels {
    const uint8_t *col;
    __device__ __forceinline__ int operator[](int L) const { return col[L * PB_HARD_THREADS]; }
};
struct PbLocalLevels {
    const uint8_t *q;
    __device__ __forceinline__ int operator[](int L) const { return q[L]; }
};

// The cells the counts could not settle, one per THREAD, straight from the directory (consecutive threads
// take consecutive cells: neighbouring positions of one sample, contiguous code runs).  A cell's codes are exactly
// what call_base hands to errmod_cal, in no particular order -- errmod_cal sorts them, and the histogram walk of
// pb_walk.cuh only needs the multiset.  The levels are ranked per cell (64-bit mask of the quality values present,
// level = rank of the value), so no per-region level table is needed; the count shortcut uses need_raw.
// What the site needs from the cell (pb_site_sample: coverage bit, derived-allele bit, derived-base counts) goes into
// per-position accumulators with integer atomics, so the result does not depend on the order of the cells.
__global__ void __launch_bounds__(PB_HARD_THREADS) k_hard_cells(const PbHardArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                             // [2 * LCAP][PB_HARD_THREADS]
    uint8_t *qcol = reinterpret_cast<uint8_t *>(hist + (size_t)2 * PB_HARD_LCAP * PB_HARD_THREADS);   // [LCAP][PB_HARD_THREADS]
    double *fk_s = reinterpret_cast<double *>(qcol + (size_t)PB_HARD_LCAP * PB_HARD_THREADS);
    // an overflowed arena holds no usable directory (the host runs the region again with a larger one)
    const unsigned long long total = (a.ctr->arena_overflow || a.ctr->spec_fail) ? 0ULL : a.ctr->n_cells;
    if ((unsigned long long)blockIdx.x * PB_HARD_THREADS >= total) return;
    for (int i = tid; i < 256; i += PB_HARD_THREADS) fk_s[i] = a.fk[i];
    uint32_t *const my_hist = hist + tid;
    for (int lw = 0; lw < 2 * PB_HARD_LCAP; ++lw) my_hist[lw * PB_HARD_THREADS] = 0;
    __syncthreads();
    for (unsigned long long c = (unsigned long long)blockIdx.x * PB_HARD_THREADS + tid; c < total; c += (unsigned long long)gridDim.x * PB_HARD_THREADS) {
        const uint4 cell = __ldg(a.cells + c);
        const int pos = (int)cell.x, smp = (int)(cell.y & 0xffu), k = (int)(cell.y >> 8), rmsq = (int)cell.z;
        if (k == 0) continue;                                  // cannot happen for a listed cell; nothing to fold anyway
        const uint16_t *cd = a.codes + cell.w;
        unsigned long long present = 0;
        uint32_t tot4 = 0;
        for (int i = 0; i < k; ++i) {
            const uint32_t code = __ldg(cd + i);
            present |= 1ULL << (code >> 5);
            tot4 += 1u << ((code & 3u) << 3);
        }
        const int nl = __popcll(present), n_lw = 2 * nl;
        const int rc = (pos >= 0 && pos < a.ref_len) ? (int)(unsigned char)a.ref[pos] : 'N';
        const int bb = (tot4 >> 8 & 255u) ? 1 : (tot4 >> 16 & 255u) ? 2 : (tot4 >> 24) ? 3 : 0;
        uint64_t cb;
        if (nl <= PB_HARD_LCAP) {
            { unsigned long long m = present; for (int L = 0; L < nl; ++L) { qcol[L * PB_HARD_THREADS + tid] = (uint8_t)(__ffsll((long long)m) - 1); m &= m - 1; } }
            for (int i = 0; i < k; ++i) {
                const uint32_t code = __ldg(cd + i);
                const int L = __popcll(present & ((1ULL << (code >> 5)) - 1ULL));
                my_hist[(2 * L + (int)((code >> 4) & 1u)) * PB_HARD_THREADS] += 1u << ((code & 3u) << 3);
            }
            const PbCellLevels qv{qcol + tid};
            if (pb_tot4_unanimous(tot4)) {
                auto peek = [&](int lw) -> uint32_t { return my_hist[lw * PB_HARD_THREADS]; };
                cb = pb_unanimous_by_count_raw(peek, nl, qv, a.need_raw, k, bb) ? pb_unanimous_result(a.lhet, k, bb, rmsq)
                                                                                 : pb_call_unanimous(peek, n_lw, qv, tot4, rmsq, fk_s, a.beta, a.lhet);
                for (int lw = 0; lw < n_lw; ++lw) my_hist[lw * PB_HARD_THREADS] = 0;
            } else {
                auto take = [&](int lw) -> uint32_t { const uint32_t w = my_hist[lw * PB_HARD_THREADS]; my_hist[lw * PB_HARD_THREADS] = 0; return w; };
                cb = pb_call_general(take, n_lw, qv, tot4, rmsq, fk_s, a.beta, a.lhet);
            }
        } else {
            // more distinct quality values in one cell than the shared-memory columns hold: same walk on a local histogram
            uint32_t lh[128];
            uint8_t lq[64];
            for (int lw = 0; lw < 128; ++lw) lh[lw] = 0;
            { unsigned long long m = present; for (int L = 0; L < nl; ++L) { lq[L] = (uint8_t)(__ffsll((long long)m) - 1); m &= m - 1; } }
            for (int i = 0; i < k; ++i) {
                const uint32_t code = __ldg(cd + i);
                const int L = __popcll(present & ((1ULL << (code >> 5)) - 1ULL));
                lh[2 * L + (int)((code >> 4) & 1u)] += 1u << ((code & 3u) << 3);
            }
            const PbLocalLevels qv{lq};
            auto take = [&](int lw) -> uint32_t { const uint32_t w = lh[lw]; lh[lw] = 0; return w; };
            auto peek = [&](int lw) -> uint32_t { return lh[lw]; };
            cb = pb_tot4_unanimous(tot4) ? pb_call_unanimous(peek, n_lw, qv, tot4, rmsq, fk_s, a.beta, a.lhet)
                                         : pb_call_general(take, n_lw, qv, tot4, rmsq, fk_s, a.beta, a.lhet);
        }
        uint32_t d4 = 0;
        bool cv, der;
        (void)pb_site_sample(cb, rc, pb_iupac_rev(rc), a.het_mode, a.min_snpQ, a.min_rmsQ, a.min_depth, a.max_depth, &d4, &cv, &der);
        const int64_t o = (int64_t)pos - a.span_beg;
        if (d4) atomicAdd(a.acc_cnt4 + o, d4);
        if (cv) atomicOr(reinterpret_cast<unsigned long long *>(a.acc_cov + o), 1ULL << smp);
        if (der) atomicOr(reinterpret_cast<unsigned long long *>(a.site_type + o), 1ULL << smp);
    }
}

// The sites (make_X tail, pop_nucdiv.cpp:168-199): coverage by every sample, segbase's value, window membership.
__global__ void __launch_bounds__(256) k_fast_sites(const PbHardArgs a) {
    const int64_t o = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int p = a.span_beg + (int)o;
    if (p >= a.span_end) return;
    const uint64_t cov = a.acc_cov[o];
    const int fq = pb_site_fq(a.acc_cnt4[o]);
    int lo = 0, hi = a.n_windows;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a.win_end + mid) > p) hi = mid; else lo = mid + 1; }
    const bool in_win = lo < a.n_windows && __ldg(a.win_beg + lo) <= p;
    const bool used = in_win && __popcll(cov) == a.n_samples;
    a.site_flag[o] = (uint8_t)((used ? 1 : 0) | ((used && fq > 0) ? 2 : 0));
}
