// pb_tables.cpp -- host-side construction of the error-model tables, i.e. what the reference's
// errmod_init(1.0-0.83) -> cal_coef(depcorr, 0.03) builds (pop_utils.cpp:203-266) with its LogGamma
// (gamma.cpp:126-166).  The tables depend on x87 long-double expl/logl, which has no device
// equivalent, so they are always built here with the host libm and uploaded.
// Compile with -ffp-contract=off.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
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

namespace {

// fingerprint of the host arithmetic the tables depend on: the libm calls of cal_coef at fixed arguments
uint64_t libm_fingerprint() {
    volatile double a = 0.17, b = -3.7, c = 1.0e-7;
    volatile long double la = -12.345678901234567L, lb = 1.2345678901234567e-7L;
    const double v[4] = {std::pow(1.0 - (double)a, 37), std::pow(10.0, (double)b), std::log((double)c), std::log(1.0 - (double)c)};
    const long double w[2] = {expl(la), logl(lb)};
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&](const void *p, size_t n) { const unsigned char *q = (const unsigned char *)p; for (size_t i = 0; i < n; ++i) { h ^= q[i]; h *= 0x100000001b3ULL; } };
    mix(v, sizeof v);
    unsigned char lw[2][10];
    memcpy(lw[0], &w[0], 10); memcpy(lw[1], &w[1], 10);           // the 80 significant bits of an x87 long double
    mix(lw, sizeof lw);
    return h;
}
uint64_t word_hash(const double *p, size_t n, uint64_t h) {
    for (size_t i = 0; i < n; ++i) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 0x9e3779b97f4a7c15ULL; h ^= h >> 29; }
    return h;
}
const uint64_t kCacheMagic = 0x31764d5245425050ULL;      // "PPBERMv1"
const size_t kNFk = 256, kNBeta = (size_t)64 * 256 * 256, kNLhet = 65536;

std::string cache_dir_default() {
    const char *e = getenv("POPBAM_B200_CACHE_DIR");
    if (e && *e) return e;
    e = getenv("XDG_CACHE_HOME");
    if (e && *e) return std::string(e) + "/popbam_b200";
    e = getenv("HOME");
    if (e && *e) return std::string(e) + "/.cache/popbam_b200";
    return "/tmp/popbam_b200-" + std::to_string((long)getuid());
}
void mkdirs(const std::string &d) {
    for (size_t i = 1; i <= d.size(); ++i)
        if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0755);
}

}  // namespace

extern "C" int pb_errmod_tables_cached_ex(double *fk, double *beta, double *lhet, const char *cache_dir, int *from_cache) {
    if (!fk || !beta || !lhet) return PB_ERR_ARG;
    if (from_cache) *from_cache = 0;
    const char *off = getenv("POPBAM_B200_NO_TABLE_CACHE");
    if (off && *off == '1') return pb_build_errmod_tables(fk, beta, lhet);
    const uint64_t fp = libm_fingerprint();
    const std::string dir = cache_dir && *cache_dir ? cache_dir : cache_dir_default();
    char name[64];
    snprintf(name, sizeof name, "/errmod-v1-%016llx.bin", (unsigned long long)fp);
    const std::string path = dir + name;
    if (FILE *f = fopen(path.c_str(), "rb")) {
        uint64_t hdr[3] = {0, 0, 0};
        bool ok = fread(hdr, 8, 3, f) == 3 && hdr[0] == kCacheMagic && hdr[1] == fp;
        ok = ok && fread(fk, 8, kNFk, f) == kNFk && fread(beta, 8, kNBeta, f) == kNBeta && fread(lhet, 8, kNLhet, f) == kNLhet;
        char extra;
        ok = ok && fread(&extra, 1, 1, f) == 0;
        fclose(f);
        if (ok && word_hash(lhet, kNLhet, word_hash(beta, kNBeta, word_hash(fk, kNFk, fp))) == hdr[2]) {
            if (from_cache) *from_cache = 1;
            return PB_OK;
        }
    }
    const int rc = pb_build_errmod_tables(fk, beta, lhet);
    if (rc != PB_OK) return rc;
    // write next to the final name and rename: concurrent starts never see a half-written file
    mkdirs(dir);
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    if (FILE *f = fopen(tmp.c_str(), "wb")) {
        const uint64_t hdr[3] = {kCacheMagic, fp, word_hash(lhet, kNLhet, word_hash(beta, kNBeta, word_hash(fk, kNFk, fp)))};
        const bool ok = fwrite(hdr, 8, 3, f) == 3 && fwrite(fk, 8, kNFk, f) == kNFk && fwrite(beta, 8, kNBeta, f) == kNBeta && fwrite(lhet, 8, kNLhet, f) == kNLhet;
        if (fclose(f) == 0 && ok) { if (rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str()); }
        else remove(tmp.c_str());
    }
    return PB_OK;
}
extern "C" int pb_errmod_tables_cached(double *fk, double *beta, double *lhet, const char *cache_dir) {
    return pb_errmod_tables_cached_ex(fk, beta, lhet, cache_dir, nullptr);
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
