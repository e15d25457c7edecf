// pb_tables.cpp -- host-side construction of the error-model tables, i.e. what the reference's
// errmod_init(1.0-0.83) -> cal_coef(depcorr, 0.03) builds (pop_utils.cpp:203-266) with its LogGamma
// (gamma.cpp:126-166).  The tables depend on x87 long-double expl/logl, which has no device
// equivalent, so they are always built here with the host libm and uploaded.
// Compile with -ffp-contract=off.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/popbam_b200.h"

namespace {

// log Gamma(x) for the integer arguments 1..256 that cal_coef uses.  Below 12 the reference takes
// log|Gamma(x)| where Gamma reduces x to y = 1 (its rational approximation is exactly 1 there) and
// multiplies y, y+1, ... back up (gamma.cpp:44-111); from 12 on it is the Stirling series of
// gamma.cpp:140-165, evaluated in the same order.
double lgamma_int(int xi) {
    const double x = (double)xi;
    if (x < 12.0) {
        double y = 1.0, g = 1.0;
        for (int i = 0; i < xi - 1; ++i) { g *= y; y += 1.0; }
        return std::log(std::fabs(g));
    }
    static const double c[8] = {1.0 / 12.0,   -1.0 / 360.0,      1.0 / 1260.0, -1.0 / 1680.0,
                                1.0 / 1188.0, -691.0 / 360360.0, 1.0 / 156.0,  -3617.0 / 122400.0};
    const double z = 1.0 / (x * x);
    double sum = c[7];
    for (int i = 6; i >= 0; --i) { sum *= z; sum += c[i]; }
    const double series = sum / x;
    const double half_log_two_pi = 0.91893853320467274178032973640562;
    return (x - 0.5) * std::log(x) - x + half_log_two_pi + series;
}

const double kLn2 = 0.69314718055994530942;    // the reference's own constants (pop_utils.cpp:32-33)
const double kLn10 = 2.30258509299404568402;

}  // namespace

extern "C" int pb_build_errmod_tables(double *fk, double *beta, double *lhet) {
    if (!fk || !beta || !lhet) return PB_ERR_ARG;
    // errmod_init takes depcorr as a float: 1.0-0.83 rounds to float first
    const double depcorr = (double)(float)(1.0 - 0.83), eta = 0.03;
    std::vector<double> lC(256 * 256, 0.0);
    std::memset(beta, 0, sizeof(double) * 64 * 256 * 256);
    fk[0] = 1.0;
    for (int n = 1; n != 256; ++n) fk[n] = std::pow(1.0 - depcorr, n) * (1.0 - eta) + eta;
    for (int n = 1; n != 256; ++n) {
        const double lgn = lgamma_int(n + 1);
        for (int k = 1; k <= n; ++k) lC[n << 8 | k] = lgn - lgamma_int(k + 1) - lgamma_int(n - k + 1);
    }
    // the 63 quality rows are independent: a few host threads (same arithmetic per entry, so the table is identical)
    std::atomic<int> next_q{63};
    auto rows = [&]() {
        for (int q = next_q.fetch_sub(1); q >= 1; q = next_q.fetch_sub(1)) {   // rows differ a lot in cost: hand them out one by one
            const double e = std::pow(10.0, -q / 10.0);
            const double le = std::log(e), le1 = std::log(1.0 - e);
            for (int n = 1; n <= 255; ++n) {
                double *row = beta + (q << 16 | n << 8);
                long double sum = 0.0, sum1 = 0.0;
                for (int k = n; k >= 0; --k, sum1 = sum) {
                    sum = sum1 + expl(lC[n << 8 | k] + k * le + (n - k) * le1);
                    row[k] = -10.0 / kLn10 * logl(sum1 / sum);
                }
            }
        }
    };
    {
        const int nt = 8;
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(rows);
        for (auto &x : th) x.join();
    }
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 256; ++k) lhet[n << 8 | k] = lC[n << 8 | k] - kLn2 * n;
    return PB_OK;
}

extern "C" int64_t pb_window_grid(int32_t beg, int32_t end, int32_t win_size, int64_t cap, int32_t *win_beg, int32_t *win_end) {
    // main_nucdiv (pop_nucdiv.cpp:48-78): num_windows = ((end-beg)-1)/win_size; window cw is the region
    // string "chr:beg+cw*W+1-(cw+1)*W+(beg-1)" parsed by bam_parse_region (pop_utils.cpp:386-461), whose
    // 1-based inclusive end is then used as an exclusive bound.
    if (win_size <= 0) {
        if (cap > 0 && win_beg && win_end) { win_beg[0] = beg; win_end[0] = end; }
        return 1;
    }
    const int64_t nw = ((int64_t)(end - beg) - 1) / win_size;
    for (int64_t cw = 0; cw < nw && cw < cap; ++cw) {
        const int64_t first = (int64_t)beg + cw * win_size + 1;       // 1-based start in the region string
        const int64_t last = (cw + 1) * win_size + ((int64_t)beg - 1);
        win_beg[cw] = (int32_t)(first > 0 ? first - 1 : first);
        win_end[cw] = (int32_t)last;
    }
    return nw;
}
