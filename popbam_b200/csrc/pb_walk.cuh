// pb_walk.cuh -- the error-model sums of errmod_cal (pop_utils.cpp:298-314) computed from a
// per-cell histogram of base codes instead of a sorted code list.
//
// The reference sorts the cell's 16-bit codes (qual<<5 | strand<<4 | base) ascending and consumes them
// from the top, i.e. by (qual desc, strand desc, base desc).  Two codes of equal value are
// indistinguishable, the counters c[] / sums bsum[] are per base and w[] is per (strand, base), so the
// result only depends, for every base separately, on the sequence of (qual, strand) in descending
// order.  A histogram over (quality level, strand) x base, walked from the highest level
// down, yields exactly that sequence with no sort.
//
// Histogram layout: one 32-bit word per level-word lw = level*2 + strand, holding four byte counters
// (base b in byte b).  Counts never exceed 255 because the raw depth cap (max_depth <= 255) precedes
// the filters (popbam.cpp:242-248).  `level` indexes the region's table of distinct quality values
// qval[] (ascending), built by the pre-pass kernels.
#pragma once
#include "pb_cell.cuh"

// Hist: callable `uint32_t take(int lw)` returning the word of level-word lw and clearing it.
// QV: anything indexable by level that yields the level's quality value (a `const uint8_t *` table of the region's
// levels, or a per-cell view for kernels that rank the levels of every cell separately).
// r4: rotation applied to the base index so that the lanes of a warp (different positions, different
// reference bases) run their heavy loop -- the reference base -- in the same unrolled slot.
template <class Hist, class QV>
PB_HD void pb_walk_hist(Hist &take, int n_lw, QV qval, int k, int r4, const double *fk,
                        const double *__restrict__ beta, double bsum[4], int c[4]) {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int cc0 = 0, cc1 = 0, cc2 = 0, cc3 = 0;
    // w[strand][rotated base] packed as bytes: counts <= 255
    uint32_t wf = 0, wr = 0;   // forward (strand 0) / reverse (strand 1)
    for (int lw = n_lw - 1; lw >= 0; --lw) {
        const uint32_t word = take(lw);
        if (word == 0) continue;
        const int q = qval[lw >> 1];
        const int st = lw & 1;
        const double *row = beta + ((size_t)q << 16 | (size_t)k << 8);
        const uint32_t wsel = st ? wr : wf;
        uint32_t wadd = 0;
#define PB_WALK_SLOT(J, ACC, CC)                                                     \
        {                                                                            \
            const int b = (r4 + J) & 3;                                              \
            const int m = (int)((word >> (8 * b)) & 255u);                           \
            const int w0 = (int)((wsel >> (8 * J)) & 255u);                          \
            for (int t = 0; t < m; ++t) ACC = pb_errmod_step(ACC, fk[w0 + t], PB_LDG(row + CC + t)); \
            CC += m;                                                                 \
            wadd |= (uint32_t)m << (8 * J);                                          \
        }
        PB_WALK_SLOT(0, acc0, cc0)
        if (word & ~(0xffu << (8 * (r4 & 3)))) {      // most words only hold the cell's dominant base (slot 0)
            PB_WALK_SLOT(1, acc1, cc1)
            PB_WALK_SLOT(2, acc2, cc2)
            PB_WALK_SLOT(3, acc3, cc3)
        }
#undef PB_WALK_SLOT
        // per-byte add cannot carry: each (strand, base) count is <= 255 in total
        if (st) wr += wadd; else wf += wadd;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int j = (b - r4) & 3;
        bsum[b] = j == 0 ? acc0 : j == 1 ? acc1 : j == 2 ? acc2 : acc3;
        c[b] = j == 0 ? cc0 : j == 1 ? cc1 : j == 2 ? cc2 : cc3;
    }
}

// ------------------------------------------------------------------------------------------------
// Cells whose bases all agree (one base b, count k).  Then errmod_cal's sums are bsum[b] = B > 0 and
// 0 for the other bases, and its ten likelihoods (pop_utils.cpp:316-362) are
//     (b,b)                      : 0                       (no other base present -> q stays 0)
//     (j,b) j<b   /  (b,m) m>b   : float(-4.343*lhet[k<<8|k])  /  float(-4.343*lhet[k<<8|0])   ("het" values)
//     every other (i,j)          : float(B)                (float accumulation of B and zeros; lhet[0] == 0)
// so gl2cns (pop_utils.cpp:66-100) returns genotype (b,b) and snpQ = (u64)(second smallest + 0.499) with
// second smallest = min(het values, float(B)).  B is a sum of non-negative terms accumulated in
// descending code order, so once float(partial B) >= the smallest het value the remaining terms
// cannot change the result and the walk stops.  pb_unanimous_het returns that smallest het value,
// or 0 when the shortcut must not be used.
PB_HD float pb_unanimous_het(const double *__restrict__ lhet, int k, int b) {
    float h = 3.402823466e+38f;
    if (b < 3) { const float h0 = PB_D2F(PB_DMUL(-4.343, PB_LDG(lhet + (k << 8)))); h = h0 < h ? h0 : h; }
    if (b > 0) { const float h1 = PB_D2F(PB_DMUL(-4.343, PB_LDG(lhet + (k << 8 | k)))); h = h1 < h ? h1 : h; }
    return h > 0.0f ? h : 0.0f;
}

// Walk of a unanimous cell: only byte b of every histogram word is populated.  Returns true when it
// stopped early (result fully determined by hmin); otherwise *bsum_b is the complete sum.
// `peek(lw)` returns a word; the caller clears the histogram afterwards.
template <class Hist, class QV>
PB_HD bool pb_walk_unanimous(Hist &peek, int n_lw, QV qval, int k, int b, const double *fk,
                             const double *__restrict__ beta, float hmin, double *bsum_b) {
    double acc = 0.0;
    int c = 0, wf = 0, wr = 0;
    for (int lw = n_lw - 1; lw >= 0; --lw) {
        const uint32_t word = peek(lw);
        if (word == 0) continue;
        const int m = (int)((word >> (8 * b)) & 255u);
        const int st = lw & 1;
        const double *row = beta + ((size_t)qval[lw >> 1] << 16 | (size_t)k << 8);
        const int w0 = st ? wr : wf;
        for (int t = 0; t < m; ++t) {
            acc = pb_errmod_step(acc, fk[w0 + t], PB_LDG(row + c + t));
            if (hmin > 0.0f && PB_D2F(acc) >= hmin) return true;
        }
        c += m;
        if (st) wr += m; else wf += m;
    }
    *bsum_b = acc;
    return false;
}

// tot4: the cell's per-base counts, one byte per base (== c[] of errmod_cal).
PB_HD int pb_tot4_k(uint32_t tot4) { return (int)((tot4 & 255u) + ((tot4 >> 8) & 255u) + ((tot4 >> 16) & 255u) + (tot4 >> 24)); }
PB_HD bool pb_tot4_unanimous(uint32_t tot4) {
    return ((tot4 & 255u) != 0) + ((tot4 >> 8 & 255u) != 0) + ((tot4 >> 16 & 255u) != 0) + ((tot4 >> 24) != 0) == 1;
}

// call_base for a cell whose bases all agree (tot4 has one populated byte): errmod_cal + gl2cns + rms
// packing (popbam.cpp:288-298) with the early exit described above.  The histogram is only read.
template <class Hist, class QV>
PB_HD uint64_t pb_call_unanimous(Hist &peek, int n_lw, QV qval, uint32_t tot4, int rmsq, const double *fk,
                                 const double *__restrict__ beta, const double *__restrict__ lhet) {
    const int b = (tot4 >> 8 & 255u) ? 1 : (tot4 >> 16 & 255u) ? 2 : (tot4 >> 24) ? 3 : 0;
    const int k = pb_tot4_k(tot4);
    const float hmin = pb_unanimous_het(lhet, k, b);
    double B = 0.0;
    if (pb_walk_unanimous(peek, n_lw, qval, k, b, fk, beta, hmin, &B)) {
        const uint64_t snpq = (uint64_t)PB_DADD((double)PB_FSUB(hmin, 0.0f), 0.499);
        const uint64_t cb = (snpq << 32) + ((uint64_t)(unsigned)k << 16) + ((uint64_t)(unsigned)(b << 2 | b) << 8);
        const uint64_t rms = (uint64_t)PB_DADD((double)PB_FSQRT(PB_FDIV((float)rmsq, (float)k)), 0.499);
        return cb | rms << 48;
    }
    double bsum[4];
    int c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bsum[i] = i == b ? B : 0.0; c[i] = i == b ? k : 0; }
    return pb_finish_cell(bsum, c, k, rmsq, lhet);
}

// call_base for any cell with k > 0 bases (the general path).  Leaves the histogram cleared.
template <class Hist, class QV>
PB_HD uint64_t pb_call_general(Hist &take, int n_lw, QV qval, uint32_t tot4, int rmsq, const double *fk,
                               const double *__restrict__ beta, const double *__restrict__ lhet) {
    double bsum[4];
    int c[4];
    const int k = pb_tot4_k(tot4);
    // rotate by the most frequent base so the lanes of a warp run their long loop in the same slot
    const uint32_t c0 = tot4 & 255u, c1 = tot4 >> 8 & 255u, c2 = tot4 >> 16 & 255u, c3 = tot4 >> 24;
    const uint32_t m01 = c0 >= c1 ? c0 : c1, m23 = c2 >= c3 ? c2 : c3;
    const int r4 = m01 >= m23 ? (c0 >= c1 ? 0 : 1) : (c2 >= c3 ? 2 : 3);
    pb_walk_hist(take, n_lw, qval, k, r4, fk, beta, bsum, c);
    return pb_finish_cell(bsum, c, k, rmsq, lhet);
}

// ------------------------------------------------------------------------------------------------
// Walk-free test for unanimous cells.  Let the cell hold `cum` bases of quality >= q_L (level L).  In
// errmod_cal's descending walk these are terms c = 0 .. cum-1, each fk[w]*beta[q',k,c] with w <= c and
// q' >= q_L, hence each term (and, rounding being monotone, every partial sum) is at least the
// corresponding one of
//     LB(L,k,m) = sum_{c<m} min_{w<=c} fk[w] * min_{L'>=L} beta[q_L',k,c]      (same order, same rounding)
// need[L][k] is the smallest m with float(LB(L,k,m)) >= max(het values of k); a cell with cum >= need
// therefore has float(bsum) >= its het value and takes the shortcut result without touching beta.
// 0 means "never".  One table per region (the levels are per region), built by k_need_table.
template <class QV>
PB_HD uint8_t pb_need_entry(int L, int nl, QV qval, int k, const double *fk, const double *__restrict__ beta,
                            const double *__restrict__ lhet) {
    if (k < 1 || k > 255) return 0;
    const float h0 = pb_unanimous_het(lhet, k, 0), h3 = pb_unanimous_het(lhet, k, 3);    // the two het values
    if (!(h0 > 0.0f) || !(h3 > 0.0f)) return 0;
    const float hcap = h0 > h3 ? h0 : h3;
    double acc = 0.0, fkmin = fk[0];
    for (int c = 0; c < k; ++c) {
        if (fk[c] < fkmin) fkmin = fk[c];
        double bmin = PB_LDG(beta + ((size_t)qval[L] << 16 | (size_t)k << 8 | (size_t)c));
        for (int l2 = L + 1; l2 < nl; ++l2) {
            const double v = PB_LDG(beta + ((size_t)qval[l2] << 16 | (size_t)k << 8 | (size_t)c));
            if (v < bmin) bmin = v;
        }
        if (!(bmin >= 0.0)) return 0;
        acc = pb_errmod_step(acc, fkmin, bmin);
        if (PB_D2F(acc) >= hcap) return (uint8_t)(c + 1);
    }
    return 0;
}

// One stray base.  A cell with k-1 bases of one letter b and ONE base of another letter e (any quality level of
// the region, any strand) is still called homozygous (b,b) once k is large enough: errmod_cal gives
// bsum[e] = fk[0]*beta[q_e,k,0] exactly (the stray base is the first and only one of its letter), every
// likelihood other than (b,b)'s either is a table value of the counts alone (the two (b,e) heterozygotes) or
// contains bsum[b] or adds a non-negative term to bsum[e].  bsum[b] is bounded below as in pb_need_entry (same
// order, same rounding, every factor at its minimum), and every step of pb_finish_cell is monotone in it, so if
// (b,b) wins STRICTLY with that bound for all twelve (b,e) letter pairs and every level of e, it wins for the
// real cell.  Returns 1 when that holds for depth k: the cell is then homozygous b whatever its snpQ, which is
// all the per-site logic needs when b is the reference base (pb_site_sample leaves such a word alone).
// The stray base's level is taken from [le_lo, le_hi): all levels, or only those below the H plane's level when the
// caller knows the stray base is not a high-quality one (a low-quality stray base is harmless at smaller depths).
PB_HD uint8_t pb_one_stray_entry(int nl, const uint8_t *qval, int k, const double *fk, const double *__restrict__ beta,
                                 const double *__restrict__ lhet, int le_lo, int le_hi) {
    if (k < 2 || k > 255 || nl < 1) return 0;
    double acc = 0.0, fkmin = fk[0];
    for (int c = 0; c < k - 1; ++c) {
        if (fk[c] < fkmin) fkmin = fk[c];
        double bmin = PB_LDG(beta + ((size_t)qval[0] << 16 | (size_t)k << 8 | (size_t)c));
        for (int l2 = 1; l2 < nl; ++l2) {
            const double v = PB_LDG(beta + ((size_t)qval[l2] << 16 | (size_t)k << 8 | (size_t)c));
            if (v < bmin) bmin = v;
        }
        if (!(bmin >= 0.0)) return 0;
        acc = pb_errmod_step(acc, fkmin, bmin);
    }
    if (le_lo < 0 || le_hi > nl || le_lo >= le_hi) return 0;
    for (int le = le_lo; le < le_hi; ++le) {
        const double be = pb_errmod_step(0.0, fk[0], PB_LDG(beta + ((size_t)qval[le] << 16 | (size_t)k << 8)));
        if (!(be >= 0.0)) return 0;
        for (int b = 0; b < 4; ++b)
            for (int e = 0; e < 4; ++e) {
                if (e == b) continue;
                double bs[4] = {0.0, 0.0, 0.0, 0.0};
                int c[4] = {0, 0, 0, 0};
                bs[b] = acc; c[b] = k - 1; bs[e] = be; c[e] = 1;
                const uint64_t cb = pb_finish_cell(bs, c, k, 0, lhet);
                // genotype (b,b) with a margin of at least one unit of snpQ between the best and the second best
                if ((int)((cb >> 8) & 0xff) != (b << 2 | b) || (int)((cb >> 32) & 0xffff) < 1) return 0;
            }
    }
    return 1;
}

// The shortcut result of a unanimous cell (see pb_unanimous_het): genotype (b,b), snpQ from the het value.
PB_HD uint64_t pb_unanimous_result(const double *__restrict__ lhet, int k, int b, int rmsq) {
    const float hmin = pb_unanimous_het(lhet, k, b);
    const uint64_t snpq = (uint64_t)PB_DADD((double)PB_FSUB(hmin, 0.0f), 0.499);
    const uint64_t cb = (snpq << 32) + ((uint64_t)(unsigned)k << 16) + ((uint64_t)(unsigned)(b << 2 | b) << 8);
    const uint64_t rms = (uint64_t)PB_DADD((double)PB_FSQRT(PB_FDIV((float)rmsq, (float)k)), 0.499);
    return cb | rms << 48;
}

// The same test for a kernel that ranks the quality levels of every cell separately (qval[L]: the cell's own
// ascending quality values): need_raw[q][k] is pb_need_entry for the level set {q, q+1, ..., 63}, a lower bound
// that holds whatever higher levels the cell contains.
template <class Hist, class QV>
PB_HD bool pb_unanimous_by_count_raw(Hist &peek, int nl, QV qval, const uint8_t *__restrict__ need_raw, int k, int b) {
    int cum = 0;
    for (int L = nl - 1; L >= 0; --L) {
        cum += (int)((peek(2 * L) >> (8 * b)) & 255u) + (int)((peek(2 * L + 1) >> (8 * b)) & 255u);
        const int nd = PB_LDG(need_raw + (int)qval[L] * 256 + k);
        if (nd && cum >= nd) return true;
        if (cum == k) break;
    }
    return false;
}
struct PbIota { int base; PB_HD int operator[](int i) const { return base + i; } };   // level set {base, base + 1, ...}

// need: [nl][256].  Returns true when the cell provably takes the shortcut.
template <class Hist>
PB_HD bool pb_unanimous_by_count(Hist &peek, int nl, const uint8_t *need, int k, int b) {
    int cum = 0;
    for (int L = nl - 1; L >= 0; --L) {
        cum += (int)((peek(2 * L) >> (8 * b)) & 255u) + (int)((peek(2 * L + 1) >> (8 * b)) & 255u);
        const int nd = need[L * 256 + k];
        if (nd && cum >= nd) return true;
        if (cum == k) break;
    }
    return false;
}
