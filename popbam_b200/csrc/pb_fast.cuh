// pb_fast.cuh -- the bit-sliced pileup: the default formulation of the pileup / call / site stage.
// Used when the raw-depth cap cannot bind (k_depth_bound), the caller did not ask for the
// per-(site,sample) words, and min_depth / min_snpQ are positive; k_pileup_call otherwise.
//
// Why.  k_pileup_call spends ~29 warp instructions per (record, 32 positions): one code load, one
// histogram update, one total per LANE per record.  But ~97 % of the cells are "easy": every passing
// base equals the reference base (or all but one do) and the depth alone proves that call_base would
// call the cell homozygous reference (pb_need_entry, pb_one_stray_entry in pb_walk.cuh), so all the
// site needs from the cell is qfilter's coverage bit (pop_utils.cpp:102-120).  For such cells nothing
// per base has to be looked at one by one: with the bases' properties as BIT-PLANES (one bit per
// base), a single THREAD handles 32 positions of one sample with 32-bit logic operations:
//     planes (built inside k_pile_fast from qual[] / seq4[]):  P passing base, B0/B1 its two base bits, H quality >= 30
//     per record:  window = funnel-shift of the planes to the strip, masked to the segment
//                  stray    = P & ((B0 ^ R0) | (B1 ^ R1))          (R: the reference strip's planes), counted 0 / 1 / 2+
//                  lowq    |= P when the read's mapQ < min_rmsQ    (else rms >= min_rmsQ is guaranteed)
//                  k  += P,  khi += H    as bit-sliced counters (carry-save adders over groups of four records)
//     per strip:   easy = no lowq, and (no stray base, k or khi in a proven range) or (one stray base, k in a proven range)
// The remaining cells (two stray bases, a variant, a low-mapQ read, an odd depth) are written out by the same
// CTA -- position, sample and the cell's base codes exactly as call_base forms them (popbam.cpp:268-284) -- into
// a compact arena, and k_hard_cells calls them one cell per thread with the exact machinery of pb_cell.cuh /
// pb_walk.cuh; k_fast_sites puts the two together into the per-site result.
//
// One pass over the bases.  Round 1 built the planes in a separate streaming kernel (1.23 GB read, 0.36 GB written,
// 0.48 GB re-read per 2.3 Mb shard) and let k_hard_cells gather its bases from qual[] / seq4[] again (0.98 GB of
// 32-byte sectors for ~2 % of the cells).  Now a CTA loads the 32 quality bytes and 16 sequence bytes behind every
// plane element of its records straight into registers, turns them into the four plane words and keeps those in
// shared memory only; the hard cells' bytes are picked up by the same CTA while they are still in L2.
// Kernels: k_ref_planes (per contig), k_set_levels / k_need_raw / k_fast_params (per context), k_strip_index,
// k_pile_fast, k_hard_cells, k_fast_sites.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_kernels.cuh"

#define PB_H_QUALITY 30            // quality value of the H plane

struct PbFastParams {        // derived from the need table (k_fast_params), cached with it
    int k0lo, k0hi;          // for k0lo <= k <= k0hi: need[0][k] != 0 and need[0][k] <= k  (the depth alone suffices)
    int k1lo, k1hi, hmin;    // for k1lo <= k <= k1hi: need[hi][k] != 0 and <= hmin <= 15   (khi >= hmin suffices)
    int hi_level;            // level of the H plane (n_levels: none)
    int k2lo, k2hi;          // for k2lo <= k <= k2hi: one stray base cannot change a homozygous call (pb_one_stray_entry)
    int k3lo, k3hi;          // the same for a stray base below the H plane's level (a wider range of depths)
    // the H-plane rule as depth segments: for seg_lo[i] <= k <= seg_hi[i], khi >= seg_h[i] bases at or above the H plane's
    // level prove the unanimous shortcut (need[hi][k] <= seg_h[i] <= 15 throughout the segment)
    int nseg;
    int seg_lo[8], seg_hi[8], seg_h[8];
};

// ---- quality levels of the bit-sliced path.  Its proofs (pb_need_entry, pb_one_stray_entry) hold for any level SET
// that contains every level a passing base can have, so it does not look for the values present (round 1 paid two
// shared-memory stores per base for that): the set is the RANGE [qlo, qhi] with
//     qlo = clamp(min(min_baseQ, min_mapQ), 4, 63)   -- call_base keeps a base only if baseQ' >= min_baseQ and
//                                                       mapQ >= min_mapQ, and codes clamp(min(baseQ', mapQ), 4, 63)
//     qhi = the context's quality ceiling            -- an ASSUMPTION: k_pile_fast checks every quality byte it
//                                                       converts against it (one add and one OR per four bytes) and
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

// ---- reference planes of a contig: R0/R1 = the two bits of A/C/G/T, RV = the byte is an upper-case A/C/G/T
// (a reference byte that is anything else never matches a called base, pop_utils.cpp:139 / SURVEY Q7)
__global__ void __launch_bounds__(256) k_ref_planes(const char *__restrict__ ref, int64_t ref_len, uint32_t *__restrict__ r0,
                                                    uint32_t *__restrict__ r1, uint32_t *__restrict__ rv) {
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (ref_len + 31) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words + 1; w += warps) {
        const int64_t o = (w << 5) + lane;
        const int c = o < ref_len ? (int)(unsigned char)ref[o] : 'N';
        const int b = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
        const uint32_t v = __ballot_sync(0xffffffffu, b >= 0);
        const uint32_t x0 = __ballot_sync(0xffffffffu, b >= 0 && (b & 1));
        const uint32_t x1 = __ballot_sync(0xffffffffu, b >= 0 && (b & 2));
        if (lane == 0) { rv[w] = v; r0[w] = x0; r1[w] = x1; }
    }
}

// The count tests of k_pile_fast as depth ranges.  Level 0 holds every passing base, so need[0][k] <= k
// means "k unanimous bases always take the shortcut"; the longest such run of depths is [k0lo, k0hi].
// Deeper cells are proven by their count of bases at or above one chosen level (the H plane): the lowest
// level whose entries stay <= 15 (the bit-sliced counter of H saturates at 16) for the depths right above k0hi.
__global__ void __launch_bounds__(64) k_fast_params(const PbCounters *__restrict__ ctr, const uint8_t *__restrict__ need,
                                                    const double *__restrict__ fk, const double *__restrict__ beta,
                                                    const double *__restrict__ lhet, PbFastParams *__restrict__ fp) {
    __shared__ uint8_t stray_ok[64], stray_low_ok[64];
    const int nl = ctr->n_levels;
    stray_ok[threadIdx.x] = pb_one_stray_entry(nl, ctr->qval, (int)threadIdx.x, fk, beta, lhet, 0, nl);     // depth k = thread index
    {
        int hi0 = 0;
        while (hi0 < nl && (int)ctr->qval[hi0] < PB_H_QUALITY) ++hi0;
        stray_low_ok[threadIdx.x] = (hi0 > 0 && hi0 < nl) ? pb_one_stray_entry(nl, ctr->qval, (int)threadIdx.x, fk, beta, lhet, 0, hi0) : 0;
    }
    __syncthreads();
    if (threadIdx.x) return;
    int k2lo = 1, k2hi = 0;
    for (int k = 1, start = 1; k <= 64; ++k) {
        if (k <= 63 && stray_ok[k]) continue;
        if (k - start > k2hi - k2lo + 1) { k2lo = start; k2hi = k - 1; }
        start = k + 1;
    }
    fp->k2lo = k2lo; fp->k2hi = k2hi;
    int k3lo = 1, k3hi = 0;
    for (int k = 1, start = 1; k <= 64; ++k) {
        if (k <= 63 && stray_low_ok[k]) continue;
        if (k - start > k3hi - k3lo + 1) { k3lo = start; k3hi = k - 1; }
        start = k + 1;
    }
    fp->k3lo = k3lo; fp->k3hi = k3hi;
    int k0lo = 1, k0hi = 0;
    for (int k = 1, start = 1; k <= 64; ++k) {
        const int nd = (k <= 63 && nl > 0) ? need[k] : 0;
        if (nd && nd <= k) continue;
        if (k - start > k0hi - k0lo + 1) { k0lo = start; k0hi = k - 1; }
        start = k + 1;
    }
    // the H plane holds "quality >= PB_H_QUALITY" (k_planes runs before the level table exists): its level is the first
    // one at or above that value, usable for the run of depths above k0hi whose entries stay <= 15
    int hi = 0, k1lo = 1, k1hi = 0, hmin = 0;
    while (hi < nl && (int)ctr->qval[hi] < PB_H_QUALITY) ++hi;
    if (hi < nl && k0hi < 63) {
        int lo = k0hi + 1, mx = 0, k = lo;
        for (; k <= 63; ++k) { const int nd = need[hi * 256 + k]; if (!nd || nd > 15) break; mx = max(mx, nd); }
        if (k > lo) {
            while (lo > 1) { const int nd = need[hi * 256 + lo - 1]; if (!nd || nd > mx) break; --lo; }
            k1lo = lo; k1hi = k - 1; hmin = mx;
        }
    }
    fp->k0lo = k0lo; fp->k0hi = k0hi; fp->k1lo = k1lo; fp->k1hi = k1hi; fp->hmin = hmin; fp->hi_level = hi;
    // need[hi][k] grows slowly with k (1 .. 15 over depths 1 .. 39 for the reference's tables): one segment per pair of
    // need values keeps the bit-sliced test within one base of the exact one
    int nseg = 0;
    if (hi < nl) {
        int start = 0, grp = 0, mx = 0;
        for (int k = 1; k <= 64; ++k) {
            const int nd = k <= 63 ? need[hi * 256 + k] : 0;
            const int g2 = (nd && nd <= 15) ? (nd + 1) / 2 : 0;
            if (start && g2 == grp) { mx = max(mx, nd); continue; }
            if (start && nseg < 8) { fp->seg_lo[nseg] = start; fp->seg_hi[nseg] = k - 1; fp->seg_h[nseg] = mx; ++nseg; }
            start = g2 ? k : 0; grp = g2; mx = nd;
        }
    }
    fp->nseg = nseg;
}

// bit-sliced "value <= C" for 6-plane counters (C in 0..63)
__device__ __forceinline__ uint32_t pb_bs_le6(const uint32_t c[6], int C) {
    uint32_t lt = 0, eq = 0xffffffffu;
#pragma unroll
    for (int b = 5; b >= 0; --b) {
        const uint32_t cb = 0u - (uint32_t)((C >> b) & 1);       // all ones when bit b of C is set (branch-free: C is a run-time value)
        lt |= eq & ~c[b] & cb;
        eq &= ~(c[b] ^ cb);
    }
    return lt | eq;
}

// Strip index of the sample-partitioned records: F[s][i] = index of the first record of sample s whose read
// starts at or after  span_beg + 32 * (i - M),  i in [0, NI), M = ceil(max_span / 32).  The records that can
// cover strip t (positions S .. S+31, S = span_beg + 32 t) are then [F[s][t], F[s][t + M + 1]) -- a superset of
// "read start in (S - max_span, S + 31]" found without any search.  One thread per record writes the entries
// whose boundary falls between its predecessor's start and its own.
// Runs before the host has seen max_span: M comes from the device counters and NI is sized for the largest M.
#define PB_SIDX_MMAX 2048          // 65536 / 32: reads span fewer than 65536 reference bases
__global__ void __launch_bounds__(256) k_strip_index(const int4 *__restrict__ srec, const uint32_t *__restrict__ sstart, int n_samples,
                                                     int span_beg, const PbCounters *__restrict__ ctr, int NI, uint32_t *__restrict__ F) {
    const int M = (ctr->max_span + 31) >> 5;
    __shared__ uint32_t ss[PB_MAX_SAMPLES + 1];
    if (threadIdx.x <= n_samples) ss[threadIdx.x] = sstart[threadIdx.x];
    __syncthreads();
    const uint32_t total = ss[n_samples];
    auto cell_of = [&](int x) -> int {                  // last i with span_beg + 32 (i - M) <= x
        const int d = x - span_beg;
        return (d >= 0 ? d >> 5 : -((-d + 31) >> 5)) + M;
    };
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < total; j += gridDim.x * 256u) {
        int s = 0;
        while (ss[s + 1] <= j) ++s;
        const int x = srec[j].x;
        int i_lo = j == ss[s] ? 0 : cell_of(srec[j - 1].x) + 1;
        int i_hi = min(cell_of(x), NI - 1);
        uint32_t *f = F + (size_t)s * NI;
        for (int i = max(i_lo, 0); i <= i_hi; ++i) f[i] = j;
        if (j + 1 == ss[s + 1]) for (int i = max(cell_of(x) + 1, 0); i < NI; ++i) f[i] = j + 1;
    }
    // samples without records
    for (int s = blockIdx.x; s < n_samples; s += gridDim.x)
        if (ss[s] == ss[s + 1]) for (int i = threadIdx.x; i < NI; i += 256) F[(size_t)s * NI + i] = ss[s];
}

struct PbFastArgs {
    const int4 *srec;
    const uint32_t *F;                       // strip index (k_strip_index), [n_samples][NI]
    int NI, RC;                              // RC: records staged per pass
    const uint8_t *qual, *seq4;              // the read batch's bases (16-byte aligned, padded by 64 bytes)
    const uint32_t *r0, *r1, *rv;            // reference planes, bit = absolute position
    int64_t ref_len;
    int span_beg, span_end;
    int n_samples, n_strips;
    int min_depth, min_rmsQ, min_baseQ, illumina;
    int qual_ceiling;                        // assumed largest (adjusted) base quality, checked on every byte converted
    int W;                                   // plane elements per staged record: covers 31 + the longest segment
    PbCounters *ctr;
    const PbFastParams *fp;
    uint32_t *cov32;                         // [n_samples][n_strips] easy and covered
    uint4 *cells;                            // directory of the cells left for k_hard_cells: {pos, sample | k << 8, sum mapq^2, first code}
    uint16_t *codes;                         // their base codes  q << 5 | strand << 4 | base  (popbam.cpp:279-284)
    unsigned long long cell_cap, code_cap;
};

#define PB_FAST_STRIPS 32          // strips of 32 positions per CTA (one sample)
#define PB_FAST_G 4                // threads sharing a strip (each takes every G-th record; partial counters are added)
#define PB_FAST_THREADS (PB_FAST_STRIPS * PB_FAST_G)
#define PB_FAST_RPT 3              // records a thread stages per pass: RC <= PB_FAST_RPT * PB_FAST_THREADS
#define PB_FAST_WMAX 16            // plane elements per record (item ids keep the element index in four bits)
__host__ __device__ static inline int pb_fast_words(int max_span) { return ((31 + (max_span > 0 ? max_span - 1 : 0)) >> 5) + 1; }
static inline size_t pb_fast_rec_bytes(int W) { return 16 + (size_t)18 * W; }     // record + W plane elements + W item ids
static inline int pb_fast_rc(int W) {           // records staged per pass: ~34 KB per CTA (six CTAs per SM), at least 64
    int rc = (int)((34 * 1024) / pb_fast_rec_bytes(W));
    if (rc > PB_FAST_RPT * PB_FAST_THREADS) rc = PB_FAST_RPT * PB_FAST_THREADS;
    return rc < 64 ? 64 : rc;
}
static inline size_t pb_fast_smem(int W) { const size_t b = (size_t)pb_fast_rc(W) * pb_fast_rec_bytes(W); return b < 8192 ? 8192 : b; }

// carry-save adder: (sum, carry) of three bit-planes -- two LOP3
__device__ __forceinline__ void pb_csa(uint32_t &sum, uint32_t &carry, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t ab = a ^ b;
    carry = (a & b) | (ab & c);
    sum = ab ^ c;
}
// flags "byte + add has bit 7 set" of four bytes, gathered into bits 28..31 (bytes < 128: no carry crosses a byte).
// 0x00204081 moves byte i's bit 7 to bit 28 + i; the sixteen partial products land on distinct bits, nothing carries.
__device__ __forceinline__ uint32_t pb_ge4_top(uint32_t w, uint32_t add) { return ((w + add) & 0x80808080u) * 0x00204081u; }
// the same for any byte value (bit 7 of the byte itself counts as "greater")
__device__ __forceinline__ uint32_t pb_ge4_top_any(uint32_t w, uint32_t add) { return ((((w & 0x7f7f7f7fu) + add) | w) & 0x80808080u) * 0x00204081u; }

// One CTA = one sample x PB_FAST_STRIPS strips of 32 positions; PB_FAST_G threads share a strip.
// Per pass over at most RC of the sample's records that can cover the CTA's positions:
//   stage    the records (16 bytes each) into shared memory; a block scan of "plane elements this record needs"
//            gives a dense list of (record, element) items
//   convert  one item per thread and step: 32 quality bytes and 16 packed-sequence bytes -> {P, B0, B1, H}.  Quality
//            thresholds by packed-byte arithmetic (add 128 - T, bit 7 is the flag; the four flags of a word are gathered
//            by one multiply), bases through a 256-entry table per seq4 byte (two bases).  Element e of a record
//            describes the bytes 32 (o / 32 + e) ... + 31 of qual[], o = offset of the segment's first base; bytes
//            of neighbouring reads that share the first / last element are masked off in the walk.
//   walk     every thread walks the records that can cover its strip (strip index: no search): funnel-shift of the
//            record's planes to the strip, stray-base logic against the reference planes, counters by carry-save
//            adders over groups of four records.
// Then the partial counters of the G threads are added, the range tests settle the easy cells (coverage bit to
// cov32), and the CTA writes the remaining cells' base codes to the arena, one warp per cell: the lanes look at the
// strip's candidate records in parallel, test coverage, fetch the base (the byte was loaded by this CTA moments
// ago: L2), and append its code.
__global__ void __launch_bounds__(PB_FAST_THREADS) k_pile_fast(const PbFastArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t seq_s[256];       // seq4 byte -> valid bits 0-1, B0 bits 8-9, B1 bits 16-17 of its two bases (high nibble = first base = lower bit)
    __shared__ unsigned long long s_base[2];
    const int tid = threadIdx.x;
    const int max_span = a.ctr->max_span;
    const int W = a.W, RC = a.RC;
    if (!a.ctr->nocap || pb_fast_words(max_span) > W) {          // launched on an assumption that does not hold: say so, do nothing
        if (tid == 0) a.ctr->spec_fail = 1;
        return;
    }
    int4 *recS = reinterpret_cast<int4 *>(smem_raw);                      // [RC] {seg start, len | mapq<<16 | strand<<24 | bit offset<<26, first 32-byte group, read start}
    uint4 *plS = reinterpret_cast<uint4 *>(recS + RC);                    // [RC][W]
    uint16_t *itemS = reinterpret_cast<uint16_t *>(plS + (size_t)RC * W); // [RC * W]  record << 4 | element
    for (int i = tid; i < 256; i += PB_FAST_THREADS) {
        const uint32_t n0 = (uint32_t)((PB_NT16_NT4_LUT >> ((i >> 4) * 4)) & 0xf), n1 = (uint32_t)((PB_NT16_NT4_LUT >> ((i & 15) * 4)) & 0xf);
        const uint32_t v0 = n0 < 4u, v1 = n1 < 4u;
        seq_s[i] = v0 | v1 << 1 | (v0 & n0 & 1u) << 8 | (v1 & n1 & 1u) << 9 | (v0 & (n0 >> 1) & 1u) << 16 | (v1 & (n1 >> 1) & 1u) << 17;
    }
    const int M = (max_span + 31) >> 5;
    const int sb = (int)(blockIdx.x / a.n_samples), s = (int)(blockIdx.x % a.n_samples);     // the samples of a strip block run together: one pass through L2
    const int strip = sb * PB_FAST_STRIPS + tid / PB_FAST_G, g = tid % PB_FAST_G;
    const uint32_t *Fs = a.F + (size_t)s * a.NI;
    const int t0 = sb * PB_FAST_STRIPS;
    const uint32_t clo = __ldg(Fs + t0), chi = __ldg(Fs + min(t0 + PB_FAST_STRIPS, a.n_strips) + M);
    const int S = a.span_beg + strip * 32;
    const bool live = strip < a.n_strips;
    const int strip_last = min(S + 31, a.span_end - 1);
    const uint32_t my_lo = live ? __ldg(Fs + strip) : 0u, my_hi = live ? __ldg(Fs + strip + M + 1) : 0u;
    uint32_t ck[6] = {0, 0, 0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
    uint32_t kover = 0, hover = 0, mism = 0, mism2 = 0, mismH = 0, lowq = 0;     // mism: at least one stray base, mism2: at least two, mismH: a high-quality one
    uint32_t R0 = 0, R1 = 0, RV = 0;
    if (live && g == 0 && S >= 0 && S < a.ref_len) {                      // reference planes of the strip (bit i = position S + i)
        const int64_t wi = S >> 5; const int sh = S & 31;
        R0 = __funnelshift_r(a.r0[wi], a.r0[wi + 1], sh);
        R1 = __funnelshift_r(a.r1[wi], a.r1[wi + 1], sh);
        RV = __funnelshift_r(a.rv[wi], a.rv[wi + 1], sh);
    }
    R0 = __shfl_sync(0xffffffffu, R0, (tid & 31) & ~(PB_FAST_G - 1));
    R1 = __shfl_sync(0xffffffffu, R1, (tid & 31) & ~(PB_FAST_G - 1));
    RV = __shfl_sync(0xffffffffu, RV, (tid & 31) & ~(PB_FAST_G - 1));
    // raw quality byte thresholds (host: all <= 128): passing, H plane, above the assumed ceiling
    const int qoff = a.illumina ? 31 : 0;
    const int tp = a.min_baseQ <= 0 ? 0 : a.min_baseQ + qoff;
    const int th = min(128, PB_H_QUALITY + qoff);
    const int tc = min(128, a.qual_ceiling + 1 + qoff);
    const uint32_t addP = (uint32_t)(128 - tp) * 0x01010101u, addH = (uint32_t)(128 - th) * 0x01010101u, addC = (uint32_t)(128 - tc) * 0x01010101u;
    for (uint32_t c0 = clo; c0 < chi; c0 += (uint32_t)RC) {
        const int cnt = (int)min((uint32_t)RC, chi - c0);
        __syncthreads();                                                  // the previous pass's walk is done (first pass: seq_s is filled)
        // ---- stage: PB_FAST_RPT consecutive records per thread, so the item list is in record order
        uint32_t need[PB_FAST_RPT], my_items = 0;
#pragma unroll
        for (int i = 0; i < PB_FAST_RPT; ++i) {
            const int r = tid * PB_FAST_RPT + i;
            need[i] = 0;
            if (r < cnt) {
                const int4 rc = __ldg(&a.srec[c0 + r]);
                const uint32_t z = (uint32_t)rc.z;
                const uint64_t gb = ((uint64_t)(z >> 26) << 32) | (uint32_t)rc.w;          // byte offset of the segment's first base
                const uint32_t bo = (uint32_t)gb & 31u;
                need[i] = (bo + (z & 0xffffu) + 31u) >> 5;                                 // <= W (pb_fast_words)
                recS[r] = make_int4(rc.y, (int)((z & 0x03ffffffu) | bo << 26), (int)(uint32_t)(gb >> 5), rc.x);
            }
            my_items += need[i];
        }
        uint32_t n_items;
        uint32_t at = pb_block_exscan(my_items, &n_items);
#pragma unroll
        for (int i = 0; i < PB_FAST_RPT; ++i)
            for (uint32_t e = 0; e < need[i]; ++e) itemS[at++] = (uint16_t)((uint32_t)(tid * PB_FAST_RPT + i) << 4 | e);
        __syncthreads();
        // ---- convert
        for (uint32_t it = tid; it < n_items; it += PB_FAST_THREADS) {
            const uint32_t item = itemS[it];
            const int4 rec = recS[item >> 4];
            const uint64_t grp = (uint64_t)(uint32_t)rec.z + (item & 15u);
            const uint4 *q4 = reinterpret_cast<const uint4 *>(a.qual + (grp << 5));
            const uint4 qa = __ldg(q4), qb = __ldg(q4 + 1);
            const uint4 sv = __ldg(reinterpret_cast<const uint4 *>(a.seq4 + (grp << 4)));
            const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
            uint32_t P = 0, H = 0, over = 0;
#pragma unroll
            for (int w = 7; w >= 0; --w) {                                 // the last word shifted in ends up in the low nibble
                P = __funnelshift_l(pb_ge4_top(qw[w], addP), P, 4);
                H = __funnelshift_l(pb_ge4_top(qw[w], addH), H, 4);
                over |= qw[w] | (qw[w] + addC);
            }
            if (over & 0x80808080u) {
                // a byte >= 128 (the packed compares above need the general form) or above the assumed ceiling (tell the host)
                P = 0; H = 0;
                int mx = 0;
#pragma unroll
                for (int w = 7; w >= 0; --w) {
                    P = __funnelshift_l(pb_ge4_top_any(qw[w], addP), P, 4);
                    H = __funnelshift_l(pb_ge4_top_any(qw[w], addH), H, 4);
#pragma unroll
                    for (int b = 0; b < 4; ++b) mx = max(mx, (int)((qw[w] >> (8 * b)) & 0xffu));
                }
                mx = min(63, a.illumina ? (mx > 31 ? mx - 31 : 0) : mx);            // levels are clamped to 63 (popbam.cpp:281)
                if (mx > a.qual_ceiling) { atomicMax(&a.ctr->qual_max_seen, mx); a.ctr->qual_over = 1; }
            }
            uint32_t V = 0, B0 = 0, B1 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // four lookups per word: 8 bases -> V in bits 0-7, B0 in 8-15, B1 in 16-23
                const uint32_t acc = seq_s[sw[j] & 0xffu] | seq_s[(sw[j] >> 8) & 0xffu] << 2 | seq_s[(sw[j] >> 16) & 0xffu] << 4 | seq_s[sw[j] >> 24] << 6;
                V |= (acc & 0xffu) << (8 * j); B0 |= ((acc >> 8) & 0xffu) << (8 * j); B1 |= ((acc >> 16) & 0xffu) << (8 * j);
            }
            P &= V;
            // H also needs mapQ >= the H plane's quality: the level is clamp(min(baseQ', mapQ), 4, 63)
            if ((((uint32_t)rec.y >> 16) & 0xffu) < (uint32_t)PB_H_QUALITY) H = 0;
            plS[(size_t)(item >> 4) * W + (item & 15u)] = make_uint4(P, B0 & P, B1 & P, H & P);
        }
        __syncthreads();
        // ---- walk
        if (live) {
            int lo = (int)(max(my_lo, c0) - c0);
            const int j1 = (int)min((uint32_t)cnt, max(my_hi, c0) - c0);
            // the strip index is 32 positions coarse: step over the leading records whose read ends before the strip
            while (lo < j1 && recS[lo].w + max_span <= S) ++lo;
            for (int j = lo + g; j < j1; j += 4 * PB_FAST_G) {
                uint32_t xP[4], xH[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int jj = j + q * PB_FAST_G;
                    xP[q] = 0; xH[q] = 0;
                    if (jj < j1) {
                        const int4 r = recS[jj];
                        const uint32_t z = (uint32_t)r.y;
                        const int len = (int)(z & 0xffffu);
                        const int u = S - r.x;                                   // segment-relative index of the strip's first position
                        const int i0 = max(0, -u), i1 = min(32, len - u);
                        if (i1 > i0) {
                            const uint32_t m = (i1 == 32 ? 0xffffffffu : (1u << i1) - 1u) & ~((1u << i0) - 1u);
                            const int t = (int)(z >> 26) + u;                    // bit index in the record's elements (>= -31)
                            const int wi = t >> 5, sh = t & 31;
                            const uint4 *pw = plS + (size_t)jj * W;
                            const uint4 zero = make_uint4(0, 0, 0, 0);
                            const uint4 lo4 = wi >= 0 ? pw[wi] : zero;
                            const uint4 hi4 = pw[min(wi + 1, W - 1)];           // past the record's elements only masked bits are read
                            const uint32_t P = __funnelshift_r(lo4.x, hi4.x, sh) & m;
                            const uint32_t B0 = __funnelshift_r(lo4.y, hi4.y, sh);
                            const uint32_t B1 = __funnelshift_r(lo4.z, hi4.z, sh);
                            const uint32_t H = __funnelshift_r(lo4.w, hi4.w, sh) & m;
                            const uint32_t mm = P & (((B0 ^ R0) | (B1 ^ R1)) | ~RV);
                            mism2 |= mism & mm; mism |= mm; mismH |= mm & H;
                            if ((int)((z >> 16) & 0xffu) < a.min_rmsQ) lowq |= P;
                            xP[q] = P; xH[q] = H;
                        }
                    }
                }
                // k += xP[0..3], khi += xH[0..3]: two carry-save levels, then the weight-4 plane ripples into the counter
                uint32_t s1, c1, c2, f;
                pb_csa(s1, c1, ck[0], xP[0], xP[1]);
                pb_csa(ck[0], c2, s1, xP[2], xP[3]);
                pb_csa(ck[1], f, ck[1], c1, c2);
#pragma unroll
                for (int b = 2; b < 6; ++b) { const uint32_t tt = ck[b] & f; ck[b] ^= f; f = tt; }
                kover |= f;
                pb_csa(s1, c1, ch[0], xH[0], xH[1]);
                pb_csa(ch[0], c2, s1, xH[2], xH[3]);
                pb_csa(ch[1], f, ch[1], c1, c2);
#pragma unroll
                for (int b = 2; b < 4; ++b) { const uint32_t tt = ch[b] & f; ch[b] ^= f; f = tt; }
                hover |= f;
            }
        }
    }
    // add the partial counters of the PB_FAST_G threads of a strip (bit-sliced ripple adders)
#pragma unroll
    for (int o = 1; o < PB_FAST_G; o <<= 1) {
        uint32_t carry = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, ck[b], o), x = ck[b] ^ y;
            const uint32_t c2 = (ck[b] & y) | (x & carry);
            ck[b] = x ^ carry; carry = c2;
        }
        kover |= carry | __shfl_xor_sync(0xffffffffu, kover, o);
        carry = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, ch[b], o), x = ch[b] ^ y;
            const uint32_t c2 = (ch[b] & y) | (x & carry);
            ch[b] = x ^ carry; carry = c2;
        }
        hover |= carry | __shfl_xor_sync(0xffffffffu, hover, o);
        {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, mism, o);
            mism2 |= (mism & y) | __shfl_xor_sync(0xffffffffu, mism2, o);
            mism |= y;
            mismH |= __shfl_xor_sync(0xffffffffu, mismH, o);
        }
        lowq |= __shfl_xor_sync(0xffffffffu, lowq, o);
    }
    // After the butterfly all PB_FAST_G threads of a strip hold the totals, so the range tests are shared out: thread g
    // evaluates "depth in [lo_g, hi_g]" for one of the four ranges (bit-sliced compares against run-time constants)
    const int k0lo = a.fp->k0lo, k0hi = a.fp->k0hi, k2lo = a.fp->k2lo, k2hi = a.fp->k2hi, k3lo = a.fp->k3lo, k3hi = a.fp->k3hi, nseg = a.fp->nseg;
    static_assert(PB_FAST_G == 4, "the range tests are shared out over four threads of a strip");
    const int lo_g = g == 0 ? k0lo : g == 1 ? k3lo : g == 2 ? k2lo : max(a.min_depth, 0);
    const int hi_g = g == 0 ? k0hi : g == 1 ? k3hi : g == 2 ? k2hi : 63;
    uint32_t in_range = 0;
    if (hi_g >= lo_g && lo_g <= 63) in_range = pb_bs_le6(ck, min(hi_g, 63)) & (lo_g > 0 ? ~pb_bs_le6(ck, lo_g - 1) : 0xffffffffu);
    // the H-plane rule: thread g takes the segments g, g + 4
    uint32_t r_k1 = 0;
    for (int sg = g; sg < nseg; sg += PB_FAST_G) {
        const int slo = __ldg(&a.fp->seg_lo[sg]), shi = __ldg(&a.fp->seg_hi[sg]), C = __ldg(&a.fp->seg_h[sg]) - 1;
        uint32_t in_k = pb_bs_le6(ck, min(shi, 63)) & (slo > 0 ? ~pb_bs_le6(ck, slo - 1) : 0xffffffffu);
        // khi >= h  <=>  not (khi <= h - 1), or the 4-plane counter overflowed (>= 16 > h)
        uint32_t lt = 0, eq = 0xffffffffu;
#pragma unroll
        for (int b = 3; b >= 0; --b) {
            const uint32_t cb = 0u - (uint32_t)((C >> b) & 1);
            lt |= eq & ~ch[b] & cb;
            eq &= ~(ch[b] ^ cb);
        }
        r_k1 |= in_k & (~(lt | eq) | hover);
    }
    r_k1 |= __shfl_xor_sync(0xffffffffu, r_k1, 1);
    r_k1 |= __shfl_xor_sync(0xffffffffu, r_k1, 2);
    const int lane0 = (tid & 31) & ~(PB_FAST_G - 1);
    const uint32_t r_k0 = __shfl_sync(0xffffffffu, in_range, lane0), r_k3 = __shfl_sync(0xffffffffu, in_range, lane0 + 1);
    const uint32_t r_k2 = __shfl_sync(0xffffffffu, in_range, lane0 + 2), dge = __shfl_sync(0xffffffffu, in_range, lane0 + 3);
    uint32_t hard = 0;
    if (live && g == 0) {
        const uint32_t valid = (strip_last - S + 1) >= 32 ? 0xffffffffu : (1u << (strip_last - S + 1)) - 1u;
        const uint32_t nonzero = (ck[0] | ck[1] | ck[2] | ck[3] | ck[4] | ck[5] | kover) & valid;
        // depth alone / count of high-quality bases proves the unanimous shortcut (pb_need_entry); one stray base at a
        // depth where it provably cannot change the homozygous-reference call (pb_one_stray_entry)
        const uint32_t easy = nonzero & ~lowq & ~kover & ((~mism & (r_k0 | r_k1)) | (mism & ~mism2 & (r_k2 | (~mismH & r_k3))));
        // qfilter for easy cells: rms >= min_rmsQ holds because every contributing read has mapQ >= min_rmsQ;
        // depth <= max_depth holds because the cap cannot bind; depth >= min_depth is the fourth range test
        a.cov32[(size_t)s * a.n_strips + strip] = easy & dge;
        hard = nonzero & ~easy;
    }
    // ---- the cells left over: directory entry + base codes, one warp per cell
    const uint32_t n_cand = my_hi - my_lo;                               // upper bound of a cell's depth: the strip's candidate records
    const uint32_t my_cells = (uint32_t)__popc(hard);
    __syncthreads();                                                      // the last walk is done: shared memory is reused below
    uint32_t tot_cells, tot_codes;
    uint32_t cell_at = pb_block_exscan(my_cells, &tot_cells);
    uint32_t code_at = pb_block_exscan(my_cells * n_cand, &tot_codes);
    if (tot_cells == 0) return;
    uint16_t *cellS = reinterpret_cast<uint16_t *>(smem_raw);            // [<= 32 * PB_FAST_STRIPS] strip << 5 | bit
    uint32_t *codeS = reinterpret_cast<uint32_t *>(smem_raw + 2 * 32 * PB_FAST_STRIPS);      // first code slot, relative to the CTA's reservation
    if (tid == 0) {
        unsigned long long cb = atomicAdd(&a.ctr->n_cells, (unsigned long long)tot_cells);
        const unsigned long long kb = atomicAdd(&a.ctr->n_codes, (unsigned long long)tot_codes);
        if (cb + tot_cells > a.cell_cap || kb + tot_codes > a.code_cap) { a.ctr->arena_overflow = 1; cb = ~0ULL; }
        s_base[0] = cb; s_base[1] = kb;
    }
    for (uint32_t hm = hard; hm; hm &= hm - 1) {
        cellS[cell_at] = (uint16_t)((uint32_t)(tid / PB_FAST_G) << 5 | (uint32_t)(__ffs(hm) - 1));
        codeS[cell_at] = code_at;
        ++cell_at; code_at += n_cand;
    }
    __syncthreads();
    if (s_base[0] == ~0ULL) return;
    const int lane = tid & 31;
    for (uint32_t c = (uint32_t)(tid >> 5); c < tot_cells; c += PB_FAST_THREADS / 32) {
        const uint32_t id = cellS[c];
        const int cstrip = t0 + (int)(id >> 5);
        const int pos = a.span_beg + cstrip * 32 + (int)(id & 31u);
        const uint32_t j0 = __ldg(Fs + cstrip), j1 = __ldg(Fs + cstrip + M + 1);
        uint16_t *out = a.codes + (s_base[1] + codeS[c]);
        uint32_t k = 0;
        int rmsq = 0;
        for (uint32_t jb = j0; jb < j1; jb += 32) {
            const uint32_t j = jb + (uint32_t)lane;
            bool ok = false;
            uint32_t code = 0;
            int mq = 0;
            if (j < j1) {
                const int4 r = __ldg(&a.srec[j]);
                const uint32_t z = (uint32_t)r.z;
                const uint32_t u = (uint32_t)(pos - r.y);
                if (u < (z & 0xffffu)) {
                    // call_base's filter and code for this base (popbam.cpp:268-284)
                    const int64_t off = (int64_t)(((uint64_t)(z >> 26) << 32) | (uint32_t)r.w) + u;
                    int bq = (int)__ldg(a.qual + off);
                    const uint32_t sbyte = __ldg(a.seq4 + (off >> 1));
                    const uint32_t nib = (off & 1) ? (sbyte & 15u) : (sbyte >> 4);
                    const uint32_t nt = (uint32_t)((PB_NT16_NT4_LUT >> (nib * 4)) & 0xf);
                    if (a.illumina) bq = bq > 31 ? bq - 31 : 0;
                    if (nt <= 3u && bq >= a.min_baseQ) {
                        mq = (int)((z >> 16) & 0xffu);
                        const int qq = max(4, min(63, min(bq, mq)));
                        code = (uint32_t)qq << 5 | ((z >> 24) & 1u) << 4 | nt;
                        ok = true;
                    }
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, ok);
            if (ok) { out[k + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint16_t)code; rmsq += mq * mq; }
            k += (uint32_t)__popc(bal);
        }
        rmsq = __reduce_add_sync(0xffffffffu, rmsq);
        if (lane == 0) a.cells[s_base[0] + c] = make_uint4((uint32_t)pos, (uint32_t)s | k << 8, (uint32_t)rmsq, (uint32_t)(s_base[1] + codeS[c]));
    }
}

struct PbHardArgs {
    const uint4 *cells;                      // directory written by k_pile_fast
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
    const uint32_t *cov32;                   // [n_samples][n_strips]
    uint64_t *acc_cov;                       // [span] coverage bits of the hard cells      (zeroed before k_hard_cells)
    uint32_t *acc_cnt4;                      // [span] derived-base counts of the hard cells (zeroed)
    uint64_t *site_type;                     // [span] derived-allele bits                   (zeroed; the hard cells are the only writers)
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

// The cells the bit-sliced pass could not settle, one per THREAD, straight from the directory (consecutive threads
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
    const unsigned long long total = a.ctr->arena_overflow ? 0ULL : a.ctr->n_cells;
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
    const int strip = (int)(o >> 5), bit = (int)(o & 31);
    uint64_t cov = a.acc_cov[o];
    for (int s = 0; s < a.n_samples; ++s) cov |= (uint64_t)((__ldg(a.cov32 + (size_t)s * a.n_strips + strip) >> bit) & 1u) << s;
    const int fq = pb_site_fq(a.acc_cnt4[o]);
    int lo = 0, hi = a.n_windows;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a.win_end + mid) > p) hi = mid; else lo = mid + 1; }
    const bool in_win = lo < a.n_windows && __ldg(a.win_beg + lo) <= p;
    const bool used = in_win && __popcll(cov) == a.n_samples;
    a.site_flag[o] = (uint8_t)((used ? 1 : 0) | ((used && fq > 0) ? 2 : 0));
}
