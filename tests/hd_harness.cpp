// hd_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the product's inlinable device functions
// (popbam_b200/csrc/pb_cell.cuh, pb_walk.cuh) for the host, so that their arithmetic can be checked
// against the reference's known-answer vectors on a machine without a GPU.  Built by tests only
// (g++ -ffp-contract=off); never linked into libpopbam_b200.so.
#include <cstdint>
#include <cstring>
#include "../popbam_b200/csrc/pb_cell.cuh"
#include "../popbam_b200/csrc/pb_walk.cuh"

static long g_by_count = 0;
extern "C" long hd_by_count() { return g_by_count; }

extern "C" uint64_t hd_call_cell(const double *fk, const double *beta, const double *lhet, const uint16_t *codes, int k, int rmsq,
                                 int r4) {
    // level table of the codes present, ascending (what k_level_table builds for a region)
    uint8_t qrank[64], qval[64];
    uint64_t present = 0;
    for (int i = 0; i < k; ++i) {
        int q = codes[i] >> 5; q = q < 4 ? 4 : q > 63 ? 63 : q;
        present |= 1ULL << q;
    }
    int nl = 0;
    for (int q = 0; q < 64; ++q) { qrank[q] = (uint8_t)nl; if (present >> q & 1) qval[nl++] = (uint8_t)q; }
    uint32_t hist[128];
    memset(hist, 0, sizeof hist);
    for (int i = 0; i < k; ++i) {
        int q = codes[i] >> 5; q = q < 4 ? 4 : q > 63 ? 63 : q;
        hist[qrank[q] * 2 + ((codes[i] >> 4) & 1)] += 1u << (8 * (codes[i] & 3));
    }
    uint32_t tot4 = 0;
    for (int i = 0; i < k; ++i) tot4 += 1u << (8 * (codes[i] & 3));
    (void)r4;
    auto take = [&](int lw) -> uint32_t { uint32_t w = hist[lw]; hist[lw] = 0; return w; };
    auto peek = [&](int lw) -> uint32_t { return hist[lw]; };
    uint64_t cb;
    if (k == 0) { double bs[4] = {0, 0, 0, 0}; int c[4] = {0, 0, 0, 0}; cb = pb_finish_cell(bs, c, 0, rmsq, lhet); }
    else if (pb_tot4_unanimous(tot4)) {
        // the kernel's order: count test against the need table first, exact early-exit walk otherwise
        static thread_local uint8_t need[64 * 256];
        for (int L = 0; L < nl; ++L) need[L * 256 + k] = pb_need_entry(L, nl, qval, k, fk, beta, lhet);
        const int b = (tot4 >> 8 & 255u) ? 1 : (tot4 >> 16 & 255u) ? 2 : (tot4 >> 24) ? 3 : 0;
        if (pb_unanimous_by_count(peek, nl, need, k, b)) { cb = pb_unanimous_result(lhet, k, b, rmsq); ++g_by_count; }
        else cb = pb_call_unanimous(peek, 2 * nl, qval, tot4, rmsq, fk, beta, lhet);
        memset(hist, 0, sizeof hist);
    }
    else cb = pb_call_general(take, 2 * nl, qval, tot4, rmsq, fk, beta, lhet);
    for (int lw = 0; lw < 2 * nl; ++lw) if (hist[lw]) return ~0ULL;     // the histogram must come back cleared
    return cb;
}

extern "C" int hd_site_logic(uint64_t *cb, int n, int ref, int het_mode, int min_snpQ, int min_rmsQ, int min_depth, int max_depth,
                             uint64_t *cov, uint64_t *type) {
    return pb_site_logic(cb, 1, n, ref, het_mode, min_snpQ, min_rmsQ, min_depth, max_depth, cov, type);
}

// pb_one_stray_entry for a level set given as quality values (ascending)
extern "C" int hd_one_stray(const double *fk, const double *beta, const double *lhet, const uint8_t *qval, int nl, int k, int le_lo, int le_hi) {
    return pb_one_stray_entry(nl, qval, k, fk, beta, lhet, le_lo, le_hi);
}
