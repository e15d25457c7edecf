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
// r4: rotation applied to the base index so that the lanes of a warp (different positions, different
// reference bases) run their heavy loop -- the reference base -- in the same unrolled slot.
template <class Hist>
PB_HD void pb_walk_hist(Hist &take, int n_lw, const uint8_t *qval, int k, int r4, const double *fk,
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
        PB_WALK_SLOT(1, acc1, cc1)
        PB_WALK_SLOT(2, acc2, cc2)
        PB_WALK_SLOT(3, acc3, cc3)
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
