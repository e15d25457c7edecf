/*
 * pb_oracle.c -- plain-C CPU restatement of POPBAM 0.3's per-window statistics path.
 *
 * TEST INFRASTRUCTURE ONLY (see pb_oracle.h).  Every function cites the reference code it
 * restates (paths relative to /root/reference).  Written from SURVEY.md Appendix A/D and the
 * cited lines; pinned against the compiled reference (oracle/_ref/popbam, oracle/_ref/refdump)
 * by tests/test_oracle_pin.py and the committed goldens in tests/golden/.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (x86-64, no FMA contraction: the reference is
 * built -O2 for baseline x86-64, so every double multiply and add rounds separately).
 */
#define _GNU_SOURCE
#include "pb_oracle.h"
#include <float.h>
#include <stdarg.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ lookup tables
 * popbam.cpp:9-51.  nt16 -> nt4, genotype nibble pair -> IUPAC letter, letter -> 0..3.  */
static const int kNt16ToNt4[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};
static const char kIupac[16] = {'A', 'M', 'R', 'W', 'N', 'C', 'S', 'Y', 'N', 'N', 'G', 'K', 'N', 'N', 'N', 'T'};
static int iupac_rev(int c) {           /* popbam.cpp:33-51 */
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 14;
    }
}
static int nt16_of_char(int c) {        /* popbam.cpp:13-31 bam_nt16_table */
    switch (c) {
    case '=': return 0;
    case 'A': case 'a': return 1;  case 'C': case 'c': return 2;  case 'M': case 'm': return 3;
    case 'G': case 'g': return 4;  case 'R': case 'r': return 5;  case 'S': case 's': return 6;
    case 'V': case 'v': return 7;  case 'T': case 't': return 8;  case 'W': case 'w': return 9;
    case 'Y': case 'y': return 10; case 'H': case 'h': return 11; case 'K': case 'k': return 12;
    case 'D': case 'd': return 13; case 'B': case 'b': return 14;
    case '0': return 1; case '1': return 2; case '2': return 4; case '3': return 8;
    default: return 15;
    }
}
static const char kNt16Rev[17] = "=ACMGRSVTWYHKDBN";   /* bam.h bam_nt16_rev_table */

static int popc64(uint64_t x) { return __builtin_popcountll(x); }

/* ------------------------------------------------------------------ error-model tables
 * gamma.cpp:126-166 LogGamma.  Only integer arguments 1..256 occur (pop_utils.cpp:219-224).
 * For x < 12 the reference evaluates log|Gamma(x)| where Gamma reduces the argument to
 * y = 1 (z = y-1 = 0: the rational approximation is exactly 0/q7 + 1 = 1) and multiplies
 * back 1*2*...*(x-1) in double (gamma.cpp:44-111).  For x >= 12 it is the Stirling series
 * of gamma.cpp:140-165 evaluated in exactly this order.                                  */
static double log_gamma_int(int xi) {
    double x = (double)xi;
    if (x < 12.0) {
        double y = 1.0, result = 1.0;
        int n = xi - 1;
        for (int i = 0; i < n; ++i) { result *= y; y += 1.0; }
        return log(fabs(result));
    }
    static const double c[8] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0,
                                1.0 / 1188.0, -691.0 / 360360.0, 1.0 / 156.0, -3617.0 / 122400.0};
    double z = 1.0 / (x * x);
    double sum = c[7];
    for (int i = 6; i >= 0; --i) { sum *= z; sum += c[i]; }
    double series = sum / x;
    static const double halfLogTwoPi = 0.91893853320467274178032973640562;
    return (x - 0.5) * log(x) - x + halfLogTwoPi + series;
}

#define PBO_LN2  0.69314718055994530942     /* pop_utils.cpp:32-33 (the file's own M_LN2/M_LN10) */
#define PBO_LN10 2.30258509299404568402

/* pop_utils.cpp:203-266: errmod_init(float depcorr = 1.0-0.83) -> cal_coef(depcorr, 0.03) */
int pbo_build_tables(double *fk, double *beta, double *lhet) {
    float depcorr_f = (float)(1.0 - 0.83);      /* the float parameter of errmod_init */
    double depcorr = depcorr_f, eta = 0.03;
    double *lC = (double *)calloc(256 * 256, sizeof(double));
    if (!lC) return -1;
    memset(beta, 0, sizeof(double) * 64 * 256 * 256);
    fk[0] = 1.0;
    for (int n = 1; n != 256; ++n) fk[n] = pow(1.0 - depcorr, n) * (1.0 - eta) + eta;
    for (int n = 1; n != 256; ++n) {
        double lgn = log_gamma_int(n + 1);
        for (int k = 1; k <= n; ++k) lC[n << 8 | k] = lgn - log_gamma_int(k + 1) - log_gamma_int(n - k + 1);
    }
    for (int q = 1; q != 64; ++q) {
        double e = pow(10.0, -q / 10.0);
        double le = log(e);
        double le1 = log(1.0 - e);
        for (int n = 1; n <= 255; ++n) {
            double *b = beta + (q << 16 | n << 8);
            long double sum, sum1;
            sum1 = sum = 0.0;
            for (int k = n; k >= 0; --k, sum1 = sum) {
                sum = sum1 + expl(lC[n << 8 | k] + k * le + (n - k) * le1);
                b[k] = -10.0 / PBO_LN10 * logl(sum1 / sum);
            }
        }
    }
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 256; ++k) lhet[n << 8 | k] = lC[n << 8 | k] - PBO_LN2 * n;
    free(lC);
    return 0;
}

/* ------------------------------------------------------------------ per-cell call */
static int cmp_u16(const void *a, const void *b) {
    uint16_t x = *(const uint16_t *)a, y = *(const uint16_t *)b;
    return (x > y) - (x < y);
}

/* errmod_cal (pop_utils.cpp:280-365), m = 4.  q[16] out.  n <= 255 (the >255 shuffle path,
 * :293-297, is unreachable with max_depth <= 255).                                       */
static void errmod_cal4(const double *fk, const double *beta, const double *lhet, int n, uint16_t *bases, float *q) {
    double bsum[4] = {0, 0, 0, 0};
    unsigned c[4] = {0, 0, 0, 0};
    int w[32];
    memset(q, 0, 16 * sizeof(float));
    if (n == 0) return;
    qsort(bases, (size_t)n, sizeof(uint16_t), cmp_u16);     /* ks_introsort: ascending */
    memset(w, 0, sizeof w);
    for (int j = n - 1; j >= 0; --j) {                      /* consumed from the top (:303) */
        unsigned b = bases[j];
        int qq = (int)(b >> 5) < 4 ? 4 : (int)(b >> 5);
        if (qq > 63) qq = 63;
        int k = b & 0x1f;
        /* fsum only feeds the dead bar_e computation (:333-337) -- omitted */
        bsum[k & 0xf] += fk[w[k]] * beta[qq << 16 | n << 8 | c[k & 0xf]];
        ++c[k & 0xf];
        ++w[k];
    }
    for (int j = 0; j != 4; ++j) {
        float tmp1; int tmp2;
        tmp1 = 0.0f; tmp2 = 0;
        for (int k = 0; k != 4; ++k) {
            if (k == j) continue;
            tmp1 += bsum[k];            /* float += double: (float)((double)tmp1 + bsum[k]) */
            tmp2 += (int)c[k];
        }
        if (tmp2) q[j * 4 + j] = tmp1;
        for (int k = j + 1; k < 4; ++k) {
            int cjk = (int)(c[j] + c[k]);
            tmp1 = 0.0f; tmp2 = 0;
            for (int i = 0; i < 4; ++i) {
                if (i == j || i == k) continue;
                tmp1 += bsum[i];
                tmp2 += (int)c[i];
            }
            if (tmp2) q[j * 4 + k] = q[k * 4 + j] = -4.343 * lhet[cjk << 8 | c[k]] + tmp1;
            else      q[j * 4 + k] = q[k * 4 + j] = -4.343 * lhet[cjk << 8 | c[k]];
        }
        for (int k = 0; k != 4; ++k) if (q[j * 4 + k] < 0.0) q[j * 4 + k] = 0.0;
    }
}

/* gl2cns (pop_utils.cpp:66-100) */
static uint64_t gl2cns(const float q[16], unsigned k) {
    unsigned min_ij = 0;
    float mn = FLT_MAX, mn_next = FLT_MAX;
    for (unsigned i = 0; i < 4; ++i)
        for (unsigned j = i; j < 4; ++j) {
            float l = q[i << 2 | j];
            if (l < mn) { min_ij = i << 2 | j; mn_next = mn; mn = l; }
            else if (l < mn_next) mn_next = l;
        }
    uint64_t snpq = (uint64_t)((mn_next - mn) + 0.499) << 32;
    uint64_t nreads = (uint64_t)k << 16;
    uint64_t gt = (uint64_t)min_ij << 8;
    return snpq + nreads + gt;
}

uint64_t pbo_call_cell(const double *fk, const double *beta, const double *lhet,
                       uint16_t *codes, int k, int rmsq, float *q16) {
    float q[16];
    errmod_cal4(fk, beta, lhet, k, codes, q);
    if (q16) memcpy(q16, q, sizeof q);
    uint64_t cb = gl2cns(q, (unsigned)k);
    /* popbam.cpp:292: rms = (u64)(sqrt((float)rmsq/k)+0.499); the float overload of sqrt */
    if (k > 0) {
        uint64_t rms = (uint64_t)((double)sqrtf((float)rmsq / (float)k) + 0.499);
        cb |= rms << 48;
    }   /* k==0: NaN -> conversion leaves the low 16 bits zero after <<48 */
    return cb;
}

/* ------------------------------------------------------------------ per-site logic */
int pbo_site_logic(const pbo_params *p, uint64_t *cb, char ref, uint64_t *cov_out, uint64_t *type_out) {
    const int n = p->n_samples;
    const int r = iupac_rev((unsigned char)ref);
    /* clean_heterozygotes (pop_utils.cpp:170-201) unless -z */
    if (!(p->flags & PBO_FLAG_HETEROZYGOTE)) {
        for (int i = 0; i < n; ++i) {
            int g = (int)((cb[i] >> 8) & 0xff), a1 = (g >> 2) & 3, a2 = g & 3;
            int sq = (int)((cb[i] >> 32) & 0xffff);
            if (a1 != a2 && sq >= p->min_snpQ) {
                if (a1 == r) cb[i] += (uint64_t)(int64_t)((a2 - a1) * 1024);
                if (a2 == r) cb[i] -= (uint64_t)(int64_t)((a2 - a1) * 256);
            }
            if (a1 != a2 && sq < p->min_snpQ) {
                if (a1 != r) cb[i] += (uint64_t)(int64_t)((a2 - a1) * 1024);
                if (a2 != r) cb[i] -= (uint64_t)(int64_t)((a2 - a1) * 256);
            }
        }
    }
    /* segbase (pop_utils.cpp:122-168) */
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        int g = (int)((cb[i] >> 8) & 0xff), a1 = (g >> 2) & 3, a2 = g & 3;
        int sq = (int)((cb[i] >> 32) & 0xffff);
        if (a1 == a2 && kIupac[g & 15] != ref) {       /* g < 16 whenever a1==a2 is evaluated on a fresh call */
            if (sq >= p->min_snpQ) { cb[i] |= 2; ++cnt[a1]; }
            else {
                cb[i] -= (uint64_t)(int64_t)((g - r) * 256);
                cb[i] -= (uint64_t)(int64_t)((g - r) * 1024);
            }
        }
    }
    int nder = 0, last = 0;
    for (int b = 0; b < 4; ++b) if (cnt[b] > 0) { ++nder; last = b; }
    int fq = nder > 1 ? -1 : cnt[last];
    /* qfilter (pop_utils.cpp:102-120) */
    uint64_t cov = 0;
    for (int i = 0; i < n; ++i) {
        int rms = (int)((cb[i] >> 48) & 0xffff), nr = (int)((cb[i] >> 16) & 0xffff);
        if (rms >= p->min_rmsQ && nr >= p->min_depth && nr <= p->max_depth) { cb[i] |= 1; cov |= 1ULL << i; }
    }
    /* cal_site_type (popbam.cpp:173-184) */
    uint64_t t = 0;
    for (int i = 0; i < n; ++i) if ((cb[i] & 3) == 3) t |= 1ULL << i;
    *cov_out = cov; *type_out = t;
    return fq;
}

/* ------------------------------------------------------------------ pileup restatement
 * bam_plp_push filter (bam_pileup.c:371-374), bam_calend (bam.c:20-70), resolve_cigar2
 * (bam_pileup.c:90-221) as a from-scratch CIGAR walk per (read, position).               */
static int cig_refend(const pbo_batch *b, int64_t r, int *aligned) {
    int x = b->pos[r], al = 0;
    for (uint32_t k = b->cig_off[r]; k < b->cig_off[r + 1]; ++k) {
        int op = (int)(b->cigar[k] & 15), len = (int)(b->cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) { x += len; al += len; }
        else if (op == 2 || op == 3) x += len;
    }
    if (aligned) *aligned = al;
    return x;
}
/* returns 1 and *qpos for a base, 0 for deletion / ref-skip at p (is_del), -1 if not covered */
static int cig_resolve(const pbo_batch *b, int64_t r, int p, int *qpos) {
    int x = b->pos[r], y = 0;
    for (uint32_t k = b->cig_off[r]; k < b->cig_off[r + 1]; ++k) {
        int op = (int)(b->cigar[k] & 15), len = (int)(b->cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (p >= x && p < x + len) { *qpos = y + (p - x); return 1; }
            x += len; y += len;
        } else if (op == 2 || op == 3) {
            if (p >= x && p < x + len) return 0;
            x += len;
        } else if (op == 1 || op == 4) y += len;
    }
    return -1;
}

typedef struct { int64_t *v; int64_t n, cap; } ivec;
static void ivec_push(ivec *a, int64_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = (int64_t *)realloc(a->v, sizeof(int64_t) * (size_t)a->cap); }
    a->v[a->n++] = x;
}

/* ------------------------------------------------------------------ window statistics */
static void calc_diff(const pbo_params *p, const uint64_t *T, int S, uint16_t *diff /* n*n */) {
    /* calc_diff_matrix (pop_nucdiv.cpp:242-256) + hamming_distance (pop_utils.cpp:51-64):
     * popcount of XORed haplotype words == number of segsites where the two samples differ;
     * unsigned short accumulation wraps mod 65536.                                        */
    int n = p->n_samples;
    memset(diff, 0, sizeof(uint16_t) * (size_t)n * n);
    for (int i = 0; i < n - 1; ++i)
        for (int j = i + 1; j < n; ++j) {
            unsigned d = 0;
            for (int s = 0; s < S; ++s) d += (unsigned)(((T[s] >> i) ^ (T[s] >> j)) & 1);
            diff[j * n + i] = diff[i * n + j] = (uint16_t)d;
        }
}

static void calc_tree_diff(const pbo_params *p, const uint64_t *T, int S, uint16_t *d /* (n+1)*(n+1) */) {
    /* calc_diff_matrix of the tree subcommand (pop_tree.cpp:472-494): taxon 0 is the reference sequence, whose
     * difference to sample i is the number of segregating sites where i carries the derived allele; samples among
     * themselves as in calc_diff.  unsigned short, wraps.                                                       */
    int n = p->n_samples, m = n + 1;
    memset(d, 0, sizeof(uint16_t) * (size_t)m * m);
    for (int i = 0; i < n; ++i) {
        unsigned c = 0;
        for (int s = 0; s < S; ++s) c += (unsigned)((T[s] >> i) & 1);
        d[(i + 1) * m] = d[i + 1] = (uint16_t)c;
    }
    for (int i = 0; i < n - 1; ++i)
        for (int j = i + 1; j < n; ++j) {
            unsigned c = 0;
            for (int s = 0; s < S; ++s) c += (unsigned)(((T[s] >> i) ^ (T[s] >> j)) & 1);
            d[(j + 1) * m + i + 1] = d[(i + 1) * m + j + 1] = (uint16_t)c;
        }
}

static void calc_nucdiv(const pbo_params *p, const uint16_t *diff, double *piw, double *pib, uint16_t *mind) {
    /* calc_nucdiv (pop_nucdiv.cpp:206-239) and calc_minDxy (pop_haplo.cpp:325-363) */
    int P = p->n_pops, n = p->n_samples;
    for (int i = 0; i < P; ++i) piw[i] = 0.0;
    for (int i = 0; i < P * P; ++i) { pib[i] = 0.0; if (mind) mind[i] = 0; }
    for (int i = 0; i < P; ++i)
        for (int j = i; j < P; ++j) {
            if (i != j && mind) mind[i * P + (j - (i + 1))] = 65535;
            for (int v = 0; v < n - 1; ++v)
                for (int w = v + 1; w < n; ++w)
                    if ((p->pop_mask[i] >> v & 1) && (p->pop_mask[j] >> w & 1)) {
                        if (i == j) piw[i] += (double)diff[v * n + w];
                        else {
                            int x = i * P + (j - (i + 1));
                            pib[x] += (double)diff[v * n + w];
                            if (mind && diff[v * n + w] < mind[x]) mind[x] = diff[v * n + w];
                        }
                    }
            if (i != j) pib[i * P + (j - (i + 1))] *= 1.0 / (double)(p->pop_nsmpl[i] * p->pop_nsmpl[j]);
            else {
                piw[i] *= 2.0 / (double)(p->pop_nsmpl[i] * (p->pop_nsmpl[i] - 1));
                if (isnan(piw[i])) piw[i] = 0.0;
            }
        }
}

static void calc_sfs(const pbo_params *p, const uint64_t *T, int S, int32_t *num_snps, double *td, double *fwh) {
    /* calc_sfs + calc_a1/a2/e1/e2 (pop_sfs.cpp:227-291, :511-571) */
    int P = p->n_pops, N = p->n_samples;
    double *a1 = (double *)calloc((size_t)N + 2, sizeof(double)), *a2 = (double *)calloc((size_t)N + 3, sizeof(double));
    double *e1 = (double *)calloc((size_t)N + 2, sizeof(double)), *e2 = (double *)calloc((size_t)N + 2, sizeof(double));
    a1[0] = a1[1] = 1.0; a2[0] = a2[1] = 1.0; e1[0] = e1[1] = 1.0; e2[0] = e2[1] = 1.0;
    for (int i = 2; i <= N; ++i) { a1[i] = 0; for (int j = 1; j < i; ++j) a1[i] += 1.0 / (double)j; }
    for (int i = 2; i <= N + 1; ++i) { a2[i] = 0; for (int j = 1; j < i; ++j) a2[i] += 1.0 / (double)(j * j); }
    for (int i = 2; i <= N; ++i) { double b1 = (i + 1.0) / (3.0 * (i - 1)); e1[i] = (b1 - (1.0 / a1[i])) / a1[i]; }
    for (int i = 2; i <= N; ++i) {
        double b2 = (2.0 * ((i * i) + i + 3.0)) / (9.0 * i * (i - 1));
        e2[i] = (b2 - ((i + 2.0) / (a1[i] * i)) + (a2[i] / (a1[i] * a1[i]))) / ((a1[i] * a1[i]) + a2[i]);
    }
    for (int i = 0; i < P; ++i) {
        int n = p->pop_nsmpl[i];
        int *sfs = (int *)calloc((size_t)n + 1, sizeof(int));
        num_snps[i] = 0; td[i] = 0.0; fwh[i] = 0.0;
        for (int s = 0; s < S; ++s) {
            uint64_t pt = T[s] & p->pop_mask[i];
            unsigned f;
            if ((p->flags & PBO_FLAG_OUTGROUP) && (T[s] >> p->outidx & 1)) f = (unsigned)(n - popc64(pt));
            else f = (unsigned)popc64(pt);
            f &= 0xffff;
            if (f <= (unsigned)n) ++sfs[f];
            if (f > 0 && f < (unsigned)n) ++num_snps[i];
        }
        int ns = num_snps[i];
        if (ns > 0 && n > 1) {
            for (int j = 1; j < n; ++j) {
                td[i] += sfs[j] * (((2.0 * j * (n - j)) / (n * (n - 1))) - (1.0 / a1[n]));
                fwh[i] += sfs[j] * ((1.0 / a1[n]) - ((double)j / (n - 1)));
            }
            td[i] /= sqrt(e1[n] * ns + e2[n] * ns * (ns - 1));
            fwh[i] /= sqrt(((n - 2) * (ns / a1[n]) / (6.0 * (n - 1))) +
                           ((ns * (ns - 1) / ((a1[n] * a1[n]) + a2[n])) *
                            (18.0 * (n * n) * (3.0 * n + 2.0) * a2[n + 1] - (88.0 * n * n * n + 9.0 * (n * n) - 13.0 * n + 6.0)) /
                            (9.0 * n * ((n - 1) * (n - 1)))));
        } else { td[i] = NAN; fwh[i] = NAN; }
        free(sfs);
    }
    free(a1); free(a2); free(e1); free(e2);
}

static double r2_pair(uint64_t t1, uint64_t t2, int m1, int m2, int np) {
    /* pop_ld.cpp:238-242 */
    double x0 = (double)m1 / np, x1 = (double)m2 / np;
    double x11 = (double)popc64(t1 & t2) / np;
    double d = x11 - x0 * x1;
    return (d * d) / (x0 * (1. - x0) * x1 * (1. - x1));
}

static void calc_zns(const pbo_params *p, const uint64_t *T, int S, int32_t *num_snps, double *zns) {
    /* calc_zns (pop_ld.cpp:201-252) */
    int P = p->n_pops;
    for (int i = 0; i < P; ++i) { num_snps[i] = 0; zns[i] = 0.0; }
    if (S < 1) return;
    for (int i = 0; i < P; ++i) {
        int np = p->pop_nsmpl[i], mf = p->min_freq;
        for (int j = 0; j < S - 1; ++j) {
            uint64_t t1 = T[j] & p->pop_mask[i]; int m1 = popc64(t1);
            if (m1 >= mf && m1 <= np - mf) {
                ++num_snps[i];
                for (int k = j + 1; k < S; ++k) {
                    uint64_t t2 = T[k] & p->pop_mask[i]; int m2 = popc64(t2);
                    if (m2 >= mf && m2 <= np - mf) zns[i] += r2_pair(t1, t2, m1, m2, np);
                }
            }
        }
        ++num_snps[i];
        zns[i] *= 2.0 / (num_snps[i] * (num_snps[i] - 1));
    }
}

static void calc_omegamax(const pbo_params *p, const uint64_t *T, int S, int32_t *num_snps, double *omax) {
    /* calc_omegamax (pop_ld.cpp:254-373): literal O(S^3) scan with non-reset running sums */
    int P = p->n_pops;
    for (int j = 0; j < P; ++j) { num_snps[j] = 0; omax[j] = 0.0; }
    if (S < 1) return;
    double *r2 = (double *)malloc(sizeof(double) * (size_t)S * (size_t)S);
    for (int j = 0; j < P; ++j) {
        memset(r2, 0, sizeof(double) * (size_t)S * (size_t)S);
        int np = p->pop_nsmpl[j], mf = p->min_freq, count1 = 0, count2 = 0;
        for (int i = 0; i < S - 1; ++i) {
            uint64_t t1 = T[i] & p->pop_mask[j]; int m1 = popc64(t1);
            if (m1 >= mf && m1 <= np - mf) {
                ++num_snps[j]; count2 = count1;
                for (int k = i + 1; k < S; ++k) {
                    uint64_t t2 = T[k] & p->pop_mask[j]; int m2 = popc64(t2);
                    if (m2 >= mf && m2 <= np - mf) {
                        ++count2;
                        r2[(size_t)count1 * S + count2] = r2_pair(t1, t2, m1, m2, np);
                        r2[(size_t)count2 * S + count1] = r2[(size_t)count1 * S + count2];
                    }
                }
                ++count1;
            }
        }
        ++num_snps[j];
        int ns = num_snps[j];
        double sl = 0, sr = 0, sb = 0;
        omax[j] = 0;
        for (int i = 1; i < ns - 1; ++i) {
            for (int k = 0; k < i; ++k) for (int m = k + 1; m <= i; ++m) sl += r2[(size_t)k * S + m];
            for (int k = i + 1; k < ns; ++k) for (int m = 0; m <= i; ++m) sb += r2[(size_t)k * S + m];
            for (int k = i + 1; k < ns - 1; ++k) for (int m = k + 1; m < ns; ++m) sr += r2[(size_t)k * S + m];
            int left = i + 1, right = ns - left;
            double omega = (sl + sr) / (((left * (left - 1)) / 2.0) + ((right * (right - 1)) / 2.0));
            omega *= left * right / sb;
            omax[j] = omega > omax[j] ? omega : omax[j];
        }
    }
    free(r2);
}

static void calc_wall(const pbo_params *p, const uint64_t *T, int S, int32_t *num_snps, double *wb, double *wq) {
    /* calc_wall (pop_ld.cpp:375-458): one last_type shared by all populations */
    int P = p->n_pops, n = p->n_samples;
    for (int j = 0; j < P; ++j) { num_snps[j] = 0; wb[j] = 0.0; wq[j] = 0.0; }
    if (S < 1) return;
    int *ncong = (int *)calloc((size_t)P, sizeof(int)), *npart = (int *)calloc((size_t)P, sizeof(int));
    uint64_t **uniq = (uint64_t **)calloc((size_t)P, sizeof(uint64_t *));
    int *nu = (int *)calloc((size_t)P, sizeof(int));
    for (int j = 0; j < P; ++j) uniq[j] = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(S + 1));
    uint64_t last = 0;
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < P; ++j) {
            uint64_t type = 0, comp = 0;
            for (int k = 0; k < n; ++k) {
                uint64_t tb = T[i] & (1ULL << k), mb = p->pop_mask[j] & (1ULL << k);
                if (tb & mb) type |= 1ULL << k;
                else if (~tb & mb) comp |= 1ULL << k;
            }
            if (type > 0 && type < p->pop_mask[j]) {
                if (num_snps[j] == 0) { uniq[j][nu[j]++] = type; last = type; num_snps[j]++; }
                else {
                    if (type == last || comp == last) {
                        ncong[j]++;
                        int x = 0, y = 0;
                        for (int u = 0; u < nu[j]; ++u) { x += uniq[j][u] == type; y += uniq[j][u] == comp; }
                        if (x == 0 && y == 0) { uniq[j][nu[j]++] = type; npart[j]++; }
                    }
                    num_snps[j]++;
                    last = type;
                }
            }
        }
    for (int i = 0; i < P; ++i) {
        wb[i] = (double)ncong[i] / (double)(num_snps[i] - 1);
        wq[i] = (double)(ncong[i] + npart[i]) / num_snps[i];
    }
    for (int j = 0; j < P; ++j) free(uniq[j]);
    free(uniq); free(nu); free(ncong); free(npart);
}

static void calc_diverge(const pbo_params *p, const uint64_t *T, int S, uint16_t *ind, uint16_t *pdiv, int32_t *num_snps) {
    /* calc_diverge (pop_diverge.cpp:220-257) */
    int P = p->n_pops, n = p->n_samples;
    if (ind) for (int i = 0; i < n; ++i) {
        unsigned d = 0;
        for (int s = 0; s < S; ++s) d += (unsigned)(T[s] >> i & 1);
        ind[i] = (uint16_t)d;
    }
    if (pdiv) for (int i = 0; i < P; ++i) {
        int np = p->pop_nsmpl[i]; unsigned dv = 0; num_snps[i] = 0;
        for (int s = 0; s < S; ++s) {
            uint64_t pt = T[s] & p->pop_mask[i]; unsigned f;
            if ((p->flags & PBO_FLAG_OUTGROUP) && (T[s] >> p->outidx & 1)) f = (unsigned)(np - popc64(pt)) & 0xffff;
            else f = (unsigned)popc64(pt);
            if (f > 0 && f < (unsigned)np) ++num_snps[i];
            else if (f == (unsigned)np) ++dv;
        }
        pdiv[i] = (uint16_t)dv;
    }
}

static void calc_nhaps(const pbo_params *p, const uint16_t *diff, int32_t *nhaps, double *hdiv) {
    /* calc_nhaps (pop_haplo.cpp:208-254): diff_matrix indexed by within-population RANKS */
    int P = p->n_pops, n = p->n_samples;
    for (int i = 0; i < P; ++i) {
        int nelem = p->pop_nsmpl[i];
        nhaps[i] = 0; hdiv[i] = 0.0;
        if (nelem > 1) {
            int b[PBO_MAX_SAMPLES], m = 0;
            for (int j = 0; j < n; ++j) if (p->pop_mask[i] >> j & 1) b[m++] = j;
            for (int j = 0; j < nelem - 1; ++j)
                for (int k = j + 1; k < nelem; ++k)
                    if (diff[j * n + k] == 0 && b[k] > b[j]) b[k] = j;
            int ff = 0;
            for (int j = 0; j < m; ++j) {
                int f = 0;
                for (int k = 0; k < m; ++k) f += b[k] == j;
                if (f > 0) ++nhaps[i];
                ff += f * f;
            }
            double sh = (double)ff / (double)(nelem * nelem);
            hdiv[i] = 1.0 - ((1.0 - sh) * (double)(nelem / (nelem - 1)));
        } else { nhaps[i] = 1; hdiv[i] = 1.0; }
    }
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
static void calc_ehhs(const pbo_params *p, const uint64_t *T, int S, const double *hdiv, double *ehhs) {
    /* calc_ehhs (pop_haplo.cpp:256-323): list removal of each unique partition and of the
     * (never reset, so == whole population) "complement"; first maximum in ascending order wins */
    int P = p->n_pops;
    uint64_t *L = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(S + 1));
    uint64_t *U = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(S + 1));
    for (int i = 0; i < P; ++i) {
        int np = p->pop_nsmpl[i];
        if (np < 4) { ehhs[i] = NAN; continue; }
        int nl = 0;
        for (int s = 0; s < S; ++s) {
            uint64_t pt = T[s] & p->pop_mask[i]; int f = popc64(pt);
            if (f > 1 && f < np - 1) L[nl++] = pt;
        }
        int nu = nl; memcpy(U, L, sizeof(uint64_t) * (size_t)nl);
        qsort(U, (size_t)nu, sizeof(uint64_t), cmp_u64);
        int k = 0; for (int s = 0; s < nu; ++s) if (s == 0 || U[s] != U[s - 1]) U[k++] = U[s];
        nu = k;
        uint64_t comp = 0, max_site = 0; int best = 0;
        for (int u = 0; u < nu; ++u) {
            uint64_t t = U[u];
            comp |= p->pop_mask[i];       /* ~CHECK_BIT(..) && CHECK_BIT(pop_mask,j) is true for every member */
            int before = nl, w = 0;
            for (int s = 0; s < nl; ++s) if (L[s] != t) L[w++] = L[s];
            nl = w; w = 0;
            for (int s = 0; s < nl; ++s) if (L[s] != comp) L[w++] = L[s];
            nl = w;
            int cnt = (before - nl) + 1;
            if (cnt > best) { best = cnt; max_site = t; }
        }
        int f = popc64(max_site);
        double sh = (1.0 - ((double)((f * f) + ((np - f) * (np - f))) / (np * np))) * (double)(np / (np - 1));
        ehhs[i] = hdiv[i] / (1.0 - sh);
    }
    free(L); free(U);
}

/* ------------------------------------------------------------------ region driver */
int64_t pbo_window_grid(int32_t beg, int32_t end, int32_t win_size, int64_t cap, int32_t *wb, int32_t *we) {
    /* pop_nucdiv.cpp:48-78 + bam_parse_region (pop_utils.cpp:386-461), SURVEY Q14 */
    if (win_size <= 0) { if (cap > 0) { wb[0] = beg; we[0] = end; } return 1; }
    int64_t nw = ((int64_t)(end - beg) - 1) / win_size;
    for (int64_t cw = 0; cw < nw && cw < cap; ++cw) {
        int64_t first = beg + cw * win_size + 1, lastc = (cw + 1) * win_size + (beg - 1);
        wb[cw] = (int32_t)(first > 0 ? first - 1 : first);
        we[cw] = (int32_t)lastc;
    }
    return nw;
}

#define ALLOC(ptr, type, count) do { (ptr) = (type *)calloc((size_t)((count) > 0 ? (count) : 1), sizeof(type)); if (!(ptr)) return -4; } while (0)

int pbo_run_region(const pbo_params *p, const double *fk, const double *beta, const double *lhet,
                   const pbo_batch *b, const char *ref, int64_t ref_len, uint32_t analyses,
                   int32_t NW, const int32_t *win_beg, const int32_t *win_end, pbo_result *out) {
    const int n = p->n_samples, P = p->n_pops;
    if (n < 1 || n > PBO_MAX_SAMPLES || P < 1 || NW < 1) return -1;
    if (p->max_depth > 255 || p->max_depth < 1) return -6;
    memset(out, 0, sizeof *out);
    out->n_windows = NW; out->n_pops = P; out->n_samples = n; out->analyses = analyses;
    out->span_beg = win_beg[0]; out->span_end = win_end[NW - 1];
    const int64_t span = (int64_t)out->span_end - out->span_beg;
    ALLOC(out->win_beg, int32_t, NW); ALLOC(out->win_end, int32_t, NW);
    ALLOC(out->num_sites, int32_t, NW); ALLOC(out->segsites, int32_t, NW); ALLOC(out->seg_off, int64_t, NW + 1);
    ALLOC(out->piw, double, (int64_t)NW * P); ALLOC(out->pib, double, (int64_t)NW * P * P); ALLOC(out->min_dxy, uint16_t, (int64_t)NW * P * P);
    ALLOC(out->sfs_num_snps, int32_t, (int64_t)NW * P); ALLOC(out->td, double, (int64_t)NW * P); ALLOC(out->fwh, double, (int64_t)NW * P);
    ALLOC(out->ld_num_snps, int32_t, (int64_t)NW * P); ALLOC(out->zns, double, (int64_t)NW * P); ALLOC(out->omegamax, double, (int64_t)NW * P);
    ALLOC(out->wall_num_snps, int32_t, (int64_t)NW * P); ALLOC(out->wallb, double, (int64_t)NW * P); ALLOC(out->wallq, double, (int64_t)NW * P);
    ALLOC(out->tree_diff, uint16_t, (int64_t)NW * (n + 1) * (n + 1));
    ALLOC(out->ind_div, uint16_t, (int64_t)NW * n); ALLOC(out->pop_div, uint16_t, (int64_t)NW * P); ALLOC(out->div_num_snps, int32_t, (int64_t)NW * P);
    ALLOC(out->nhaps, int32_t, (int64_t)NW * P); ALLOC(out->hdiv, double, (int64_t)NW * P); ALLOC(out->ehhs, double, (int64_t)NW * P);
    ALLOC(out->site_type, uint64_t, span); ALLOC(out->site_flag, uint8_t, span);
    if (p->flags & PBO_FLAG_EMIT_CB) ALLOC(out->cb, uint64_t, span * n);

    /* reads that bam_plp_push keeps (bam_pileup.c:371-374), with their reference end */
    const int64_t N = b->n_reads;
    int32_t *rend = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
    uint8_t *keep = (uint8_t *)malloc((size_t)N + 1);
    out->reads_pushed = N;
    for (int64_t r = 0; r < N; ++r) {
        int al = 0;
        rend[r] = cig_refend(b, r, &al);
        unsigned flag = b->meta[r] >> 16;
        keep[r] = !(flag & 0x704) && rend[r] > b->pos[r];
        if (r > 0 && b->pos[r] < b->pos[r - 1]) { free(rend); free(keep); return -5; }
        if (keep[r]) { out->reads_used++; out->aligned_bases += al; }
    }

    /* growable per-region segregating-site storage */
    int64_t seg_cap = 1024, seg_n = 0;
    uint32_t *seg_pos = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)seg_cap), *seg_idx = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)seg_cap);
    uint64_t *seg_type = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)seg_cap);
    uint8_t *seg_ref = (uint8_t *)malloc((size_t)seg_cap);
    uint64_t *seg_cb = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)seg_cap * (size_t)n);

    ivec act = {0, 0, 0};
    int64_t next = 0;
    uint64_t cbv[PBO_MAX_SAMPLES];
    uint16_t *codes = (uint16_t *)malloc(sizeof(uint16_t) * 256 * (size_t)n);
    int depth[PBO_MAX_SAMPLES], kk[PBO_MAX_SAMPLES], rmsq[PBO_MAX_SAMPLES];
    uint16_t *diff = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n * n);

    for (int w = 0; w < NW; ++w) {
        const int beg = win_beg[w], end = win_end[w];
        out->win_beg[w] = beg; out->win_end[w] = end; out->seg_off[w] = seg_n;
        int num_sites = 0, S = 0;
        for (int pos = beg; pos < end; ++pos) {
            /* live reads at pos, file order (bam_plp_next, bam_pileup.c:283-363) */
            while (next < N && b->pos[next] <= pos) { if (keep[next] && rend[next] > pos) ivec_push(&act, next); ++next; }
            int64_t wj = 0;
            for (int64_t a = 0; a < act.n; ++a) if (rend[act.v[a]] > pos) act.v[wj++] = act.v[a];
            act.n = wj;
            if (act.n == 0) continue;                          /* no callback without live reads */
            /* call_base (popbam.cpp:186-313) */
            for (int s = 0; s < n; ++s) { depth[s] = 0; kk[s] = 0; rmsq[s] = 0; cbv[s] = 0; }
            for (int64_t a = 0; a < act.n; ++a) {
                int64_t r = act.v[a];
                int qpos;
                if (cig_resolve(b, r, pos, &qpos) != 1) continue;       /* is_del / is_refskip (:222) */
                unsigned smp = b->meta[r] & 0xff;
                if (smp >= (unsigned)n) continue;                       /* no RG tag (:227-228) */
                if (depth[smp] >= p->max_depth) continue;               /* cap before filters (:242-248) */
                depth[smp]++;
                int mapq = (int)((b->meta[r] >> 8) & 0xff), strand = (int)((b->meta[r] >> 16) & 16) ? 1 : 0;
                int bq = b->qual[(size_t)b->base_off[r] + (size_t)qpos];
                if (p->flags & PBO_FLAG_ILLUMINA) bq = bq > 31 ? bq - 31 : 0;
                if (bq < p->min_baseQ || mapq < p->min_mapQ) continue;
                int nib = (b->seq4[((size_t)b->base_off[r] >> 1) + ((size_t)qpos >> 1)] >> ((~qpos & 1) << 2)) & 0xf;
                int b4 = kNt16ToNt4[nib];
                if (b4 > 3) continue;
                int qq = bq < mapq ? bq : mapq;
                if (qq < 4) qq = 4;
                if (qq > 63) qq = 63;
                codes[smp * 256 + kk[smp]++] = (uint16_t)(qq << 5 | strand << 4 | b4);
                rmsq[smp] += mapq * mapq;
            }
            for (int s = 0; s < n; ++s)
                if (depth[s] > 0) cbv[s] = pbo_call_cell(fk, beta, lhet, codes + s * 256, kk[s], rmsq[s], 0);
            char rc = (pos >= 0 && pos < ref_len) ? ref[pos] : 'N';
            uint64_t cov, type;
            int fq = pbo_site_logic(p, cbv, rc, &cov, &type);
            int64_t so = (int64_t)pos - out->span_beg;
            if (out->cb) memcpy(out->cb + so * n, cbv, sizeof(uint64_t) * (size_t)n);
            out->site_type[so] = type;
            int used = popc64(cov) == n;
            out->site_flag[so] = (uint8_t)(4 | (used ? 1 : 0) | ((used && fq > 0) ? 2 : 0));
            if (used) {
                if (fq > 0) {
                    if (seg_n == seg_cap) {
                        seg_cap *= 2;
                        seg_pos = (uint32_t *)realloc(seg_pos, sizeof(uint32_t) * (size_t)seg_cap); seg_idx = (uint32_t *)realloc(seg_idx, sizeof(uint32_t) * (size_t)seg_cap);
                        seg_type = (uint64_t *)realloc(seg_type, sizeof(uint64_t) * (size_t)seg_cap); seg_ref = (uint8_t *)realloc(seg_ref, (size_t)seg_cap);
                        seg_cb = (uint64_t *)realloc(seg_cb, sizeof(uint64_t) * (size_t)seg_cap * (size_t)n);
                    }
                    seg_pos[seg_n] = (uint32_t)pos; seg_idx[seg_n] = (uint32_t)num_sites; seg_type[seg_n] = type; seg_ref[seg_n] = (uint8_t)rc;
                    memcpy(seg_cb + seg_n * n, cbv, sizeof(uint64_t) * (size_t)n);
                    ++seg_n; ++S;
                }
                ++num_sites;
            }
        }
        out->num_sites[w] = num_sites; out->segsites[w] = S;
        const uint64_t *T = seg_type + out->seg_off[w];
        if (analyses & (PBO_AN_NUCDIV | PBO_AN_HAPLO_K | PBO_AN_HAPLO_EHHS | PBO_AN_HAPLO_DXY)) calc_diff(p, T, S, diff);
        if (analyses & (PBO_AN_NUCDIV | PBO_AN_HAPLO_DXY)) calc_nucdiv(p, diff, out->piw + (int64_t)w * P, out->pib + (int64_t)w * P * P, out->min_dxy + (int64_t)w * P * P);
        if (analyses & PBO_AN_TREE) calc_tree_diff(p, T, S, out->tree_diff + (int64_t)w * (n + 1) * (n + 1));
        if (analyses & PBO_AN_SFS) calc_sfs(p, T, S, out->sfs_num_snps + (int64_t)w * P, out->td + (int64_t)w * P, out->fwh + (int64_t)w * P);
        if (analyses & PBO_AN_LD_ZNS) calc_zns(p, T, S, out->ld_num_snps + (int64_t)w * P, out->zns + (int64_t)w * P);
        if (analyses & PBO_AN_LD_OMEGA) calc_omegamax(p, T, S, out->ld_num_snps + (int64_t)w * P, out->omegamax + (int64_t)w * P);
        if (analyses & PBO_AN_LD_WALL) calc_wall(p, T, S, out->wall_num_snps + (int64_t)w * P, out->wallb + (int64_t)w * P, out->wallq + (int64_t)w * P);
        if (analyses & (PBO_AN_DIVERGE_IND | PBO_AN_DIVERGE_POP))
            calc_diverge(p, T, S, (analyses & PBO_AN_DIVERGE_IND) ? out->ind_div + (int64_t)w * n : 0,
                         (analyses & PBO_AN_DIVERGE_POP) ? out->pop_div + (int64_t)w * P : 0, out->div_num_snps + (int64_t)w * P);
        if (analyses & (PBO_AN_HAPLO_K | PBO_AN_HAPLO_EHHS)) calc_nhaps(p, diff, out->nhaps + (int64_t)w * P, out->hdiv + (int64_t)w * P);
        if (analyses & PBO_AN_HAPLO_EHHS) calc_ehhs(p, T, S, out->hdiv + (int64_t)w * P, out->ehhs + (int64_t)w * P);
    }
    out->seg_off[NW] = seg_n;
    out->seg_pos = seg_pos; out->seg_idx = seg_idx; out->seg_type = seg_type; out->seg_ref = seg_ref; out->seg_cb = seg_cb;
    free(act.v); free(codes); free(diff); free(rend); free(keep);
    return 0;
}

void pbo_free_result(pbo_result *r) {
    free(r->win_beg); free(r->win_end); free(r->num_sites); free(r->segsites); free(r->seg_off);
    free(r->seg_pos); free(r->seg_idx); free(r->seg_type); free(r->seg_ref); free(r->seg_cb);
    free(r->piw); free(r->pib); free(r->min_dxy); free(r->sfs_num_snps); free(r->td); free(r->fwh);
    free(r->ld_num_snps); free(r->zns); free(r->omegamax); free(r->wall_num_snps); free(r->wallb); free(r->wallq);
    free(r->tree_diff);
    free(r->ind_div); free(r->pop_div); free(r->div_num_snps); free(r->nhaps); free(r->hdiv); free(r->ehhs);
    free(r->cb); free(r->site_type); free(r->site_flag);
    memset(r, 0, sizeof *r);
}

/* ------------------------------------------------------------------ printers
 * print_nucdiv (pop_nucdiv.cpp:258-289), print_sfs (pop_sfs.cpp:293-317), print_ld
 * (pop_ld.cpp:650-712), print_diverge (pop_diverge.cpp:496-574), print_haplo
 * (pop_haplo.cpp:365-442), print_popbam_snp/print_sweep/print_ms (pop_snp.cpp:224-303).
 * iostream "fixed << setprecision(5)" == "%.5f"; setw(7) << "NA" == "%7s".                */
typedef struct { char *buf; int64_t cap, len; } sbuf;
static void sb_printf(sbuf *s, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

static void sb_printf(sbuf *s, const char *fmt, ...) {
    char tmp[512];
    va_list ap; va_start(ap, fmt);
    int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    if (k < 0) return;
    if (k >= (int)sizeof tmp) k = (int)sizeof tmp - 1;
    for (int i = 0; i < k; ++i) { if (s->len + 1 < s->cap) s->buf[s->len] = tmp[i]; s->len++; }
    if (s->cap > 0) s->buf[s->len < s->cap ? s->len : s->cap - 1] = 0;
}
static void sb_stat(sbuf *s, const char *name, const char *pop, int ok, double v) {
    if (ok) sb_printf(s, "\t%s[%s]:\t%.5f", name, pop, v);
    else sb_printf(s, "\t%s[%s]:\t%7s", name, pop, "NA");
}

/* ---- neighbour joining of the tree subcommand (pop_tree.cpp:208-470), kept in the reference's own terms: a tip is a
 * single node, an interior node a ring of three; a branch is a pair of nodes pointing at each other.            */
typedef struct njnode { struct njnode *next, *back; int tip, index; double v; } njnode;

static void nj_print(sbuf *s, const njnode *q, const njnode *start, const pbo_print_opts *o) {
    /* print_tree (pop_tree.cpp:439-470) */
    if (q->tip) sb_printf(s, "%s", q->index == 1 ? o->ref_name : o->sample_names[q->index - 2]);
    else {
        sb_printf(s, "(");
        nj_print(s, q->next->back, start, o);
        sb_printf(s, ",");
        nj_print(s, q->next->next->back, start, o);
        if (q == start) { sb_printf(s, ","); nj_print(s, q->back, start, o); }
        sb_printf(s, ")");
    }
    if (q == start) sb_printf(s, ";\n");
    else if (q->v < 0) sb_printf(s, ":0.00000");
    else sb_printf(s, ":%.5f", q->v);
}

static void nj_link(njnode *a, njnode *b) { a->back = b; b->back = a; }

static void nj_tree(sbuf *s, const uint16_t *diff, int ntaxa, int num_sites, const pbo_print_opts *o) {
    /* calc_dist_matrix (pop_tree.cpp:496-515) */
    double *x = (double *)calloc((size_t)ntaxa * ntaxa, sizeof(double));
    for (int i = 0; i < ntaxa - 1; ++i)
        for (int j = i + 1; j < ntaxa; ++j) {
            double d = (double)diff[i * ntaxa + j] / num_sites;
            if (o->jc) d = -0.75 * log(1.0 - (4.0 * d / 3.0));
            x[i * ntaxa + j] = x[j * ntaxa + i] = d;
        }
    /* tree_init + setup_tree (pop_tree.cpp:517-566): tips 1..ntaxa, rings ntaxa+1..2*ntaxa-1 */
    int nnodes = 2 * ntaxa - 1;
    njnode *pool = (njnode *)calloc((size_t)ntaxa + 3 * (size_t)(nnodes - ntaxa), sizeof(njnode));
    njnode **nodep = (njnode **)calloc((size_t)nnodes, sizeof(njnode *));
    for (int i = 0; i < ntaxa; ++i) { nodep[i] = pool + i; nodep[i]->tip = 1; nodep[i]->index = i + 1; }
    for (int i = ntaxa; i < nnodes; ++i) {
        njnode *a = pool + ntaxa + 3 * (i - ntaxa);
        a[0].next = a + 1; a[1].next = a + 2; a[2].next = a;
        a[0].index = a[1].index = a[2].index = i + 1;
        nodep[i] = a;
    }
    /* join_tree (pop_tree.cpp:254-429) */
    njnode **cluster = (njnode **)calloc((size_t)ntaxa, sizeof(njnode *));
    double *av = (double *)calloc((size_t)ntaxa, sizeof(double)), *R = (double *)calloc((size_t)ntaxa, sizeof(double));
    for (int i = 0; i < ntaxa; ++i) cluster[i] = nodep[i];
#define X(a, b) x[(size_t)(a) * ntaxa + (b)]
    for (int i = 0; i < ntaxa - 1; ++i)
        for (int j = i + 1; j < ntaxa; ++j) { double da = (X(i, j) + X(j, i)) / 2.0; X(i, j) = da; X(j, i) = da; }
    double fotu2 = ntaxa - 2.0, total = 0, tmin, dmin;
    int nextnode = ntaxa + 1, mini = 0, minj = 0;
    for (int nc = 1; nc <= ntaxa - 3; ++nc) {
        for (int j = 2; j <= ntaxa; ++j)
            for (int i = 0; i <= j - 2; ++i) X(j - 1, i) = X(i, j - 1);
        tmin = DBL_MAX;
        for (int i = 0; i < ntaxa; ++i) R[i] = 0.0;
        /* the enter order is the identity (make_nj, pop_tree.cpp:231-232) */
        for (int jj = 2; jj <= ntaxa; ++jj) {
            if (!cluster[jj - 1]) continue;
            for (int ii = 1; ii <= jj - 1; ++ii)
                if (cluster[ii - 1]) { R[ii - 1] += X(ii - 1, jj - 1); R[jj - 1] += X(ii - 1, jj - 1); }
        }
        for (int jj = 2; jj <= ntaxa; ++jj) {
            if (!cluster[jj - 1]) continue;
            for (int ii = 1; ii <= jj - 1; ++ii) {
                if (cluster[ii - 1]) total = fotu2 * X(ii - 1, jj - 1) - R[ii - 1] - R[jj - 1];
                if (total < tmin) { tmin = total; mini = ii; minj = jj; }       /* a stale total is compared again (:332-340) */
            }
        }
        double dio = 0.0, djo = 0.0;
        for (int i = 0; i < ntaxa; ++i) { dio += X(i, mini - 1); djo += X(i, minj - 1); }
        dmin = X(mini - 1, minj - 1);
        dio = (dio - dmin) / fotu2; djo = (djo - dmin) / fotu2;
        double bi = (dmin + dio - djo) * 0.5, bj = dmin - bi;
        bi -= av[mini - 1]; bj -= av[minj - 1];
        nj_link(nodep[nextnode - 1]->next, cluster[mini - 1]);
        nj_link(nodep[nextnode - 1]->next->next, cluster[minj - 1]);
        cluster[mini - 1]->v = bi; cluster[minj - 1]->v = bj;
        cluster[mini - 1]->back->v = bi; cluster[minj - 1]->back->v = bj;
        cluster[mini - 1] = nodep[nextnode - 1]; cluster[minj - 1] = 0;
        ++nextnode;
        av[mini - 1] = dmin * 0.5;
        fotu2 -= 1.0;
        for (int j = 0; j < ntaxa; ++j)
            if (cluster[j]) {
                double da = (X(mini - 1, j) + X(minj - 1, j)) * 0.5;
                if (mini - j - 1 < 0) X(mini - 1, j) = da;
                if (mini - j - 1 > 0) X(j, mini - 1) = da;
            }
        for (int j = 0; j < ntaxa; ++j) { X(minj - 1, j) = 0.0; X(j, minj - 1) = 0.0; }
    }
    int el[3], nude = 0;
    for (int i = 1; i <= ntaxa && nude < 3; ++i) if (cluster[i - 1]) el[nude++] = i;
    double bi = (X(el[0] - 1, el[1] - 1) + X(el[0] - 1, el[2] - 1) - X(el[1] - 1, el[2] - 1)) * 0.5;
    double bj = X(el[0] - 1, el[1] - 1) - bi, bk = X(el[0] - 1, el[2] - 1) - bi;
    bi -= av[el[0] - 1]; bj -= av[el[1] - 1]; bk -= av[el[2] - 1];
    nj_link(nodep[nextnode - 1], cluster[el[0] - 1]);
    nj_link(nodep[nextnode - 1]->next, cluster[el[1] - 1]);
    nj_link(nodep[nextnode - 1]->next->next, cluster[el[2] - 1]);
    cluster[el[0] - 1]->v = bi; cluster[el[1] - 1]->v = bj; cluster[el[2] - 1]->v = bk;
    cluster[el[0] - 1]->back->v = bi; cluster[el[1] - 1]->back->v = bj; cluster[el[2] - 1]->back->v = bk;
#undef X
    nj_print(s, nodep[0]->back, nodep[0]->back, o);        /* make_nj: start = nodep[0]->back (pop_tree.cpp:246-248) */
    free(x); free(pool); free(nodep); free(cluster); free(av); free(R);
}

int64_t pbo_format_window(const pbo_params *p, const pbo_result *r, int32_t w, uint32_t an,
                          const pbo_print_opts *o, char *buf, int64_t cap) {
    sbuf s = {buf, cap, 0};
    const int P = r->n_pops, n = r->n_samples;
    const int ns = r->num_sites[w];
    const int64_t so = r->seg_off[w]; const int S = r->segsites[w];
    if (an == PBO_AN_SNP) {
        if (o->snp_output == 0) {
            for (int i = 0; i < S; ++i) {
                sb_printf(&s, "%s\t%u\t%c", o->chrom, r->seg_pos[so + i] + 1, kNt16Rev[nt16_of_char(r->seg_ref[so + i])]);
                for (int j = 0; j < n; ++j) {
                    uint64_t cb = r->seg_cb[(so + i) * n + j];
                    unsigned g = (unsigned)((cb >> 8) & 0xff);
                    char base = g < 16 ? kNt16Rev[nt16_of_char(kIupac[g])] : '?';   /* g>=16: iupac[] read out of bounds in the reference */
                    sb_printf(&s, "\t%c\t%u\t%u\t%u", base, (unsigned)((cb >> 32) & 0xffff), (unsigned)((cb >> 48) & 0xffff), (unsigned)((cb >> 16) & 0xffff));
                }
                sb_printf(&s, "\n");
            }
        } else if (o->snp_output == 1) {
            for (int i = 0; i < S; ++i) {
                sb_printf(&s, "%s\t%u", o->chrom, r->seg_pos[so + i] + 1);
                uint64_t T = r->seg_type[so + i];
                for (int j = 0; j < P; ++j) {
                    int pn = popc64(p->pop_mask[j]); unsigned f;
                    if ((p->flags & PBO_FLAG_OUTGROUP) && (T >> p->outidx & 1)) f = (unsigned)(pn - popc64(T & p->pop_mask[j])) & 0xffff;
                    else f = (unsigned)popc64(T & p->pop_mask[j]);
                    sb_printf(&s, "\t%u\t%d", f, pn);
                }
                sb_printf(&s, "\n");
            }
        } else {
            sb_printf(&s, "//\nsegsites: %d\npositions: ", S);
            for (int i = 0; i < S; ++i)
                sb_printf(&s, "%.8g ", (double)(r->seg_pos[so + i] - (unsigned)r->win_beg[w]) / (r->win_end[w] - r->win_beg[w]));
            sb_printf(&s, "\n");
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < S; ++j) {
                    uint64_t T = r->seg_type[so + j]; int bit = (int)(T >> i & 1);
                    if ((p->flags & PBO_FLAG_OUTGROUP) && (T >> p->outidx & 1)) bit = !bit;
                    sb_printf(&s, "%c", bit ? '1' : '0');
                }
                sb_printf(&s, "\n");
            }
            sb_printf(&s, "\n");
        }
        return s.len;
    }
    sb_printf(&s, "%s\t%d\t%d\t%d", o->chrom, r->win_beg[w] + 1, r->win_end[w] + 1, ns);
    if (an == PBO_AN_TREE) {     /* make_nj (pop_tree.cpp:208-252) */
        if (ns < o->min_sites || S < 1) sb_printf(&s, "\tNA\n");
        else { sb_printf(&s, "\t"); nj_tree(&s, r->tree_diff + (int64_t)w * (n + 1) * (n + 1), n + 1, ns, o); }
        return s.len;
    }
    const int ok = ns >= o->min_sites;
    if (an == PBO_AN_NUCDIV) {
        for (int i = 0; i < P; ++i) sb_stat(&s, "pi", o->pop_names[i], ok, r->piw[(int64_t)w * P + i] / ns);
        for (int i = 0; i < P - 1; ++i) for (int j = i + 1; j < P; ++j) {
            char nm[256]; snprintf(nm, sizeof nm, "%s-%s", o->pop_names[i], o->pop_names[j]);
            sb_stat(&s, "dxy", nm, ok, r->pib[(int64_t)w * P * P + i * P + (j - (i + 1))] / ns);
        }
    } else if (an == PBO_AN_SFS) {
        for (int i = 0; i < P; ++i) {
            double d = r->td[(int64_t)w * P + i], h = r->fwh[(int64_t)w * P + i];
            sb_stat(&s, "D", o->pop_names[i], !isnan(d), d);
            sb_stat(&s, "H", o->pop_names[i], !isnan(h), h);
        }
    } else if (an == PBO_AN_LD_ZNS || an == PBO_AN_LD_OMEGA || an == PBO_AN_LD_WALL) {
        for (int i = 0; i < P; ++i) {
            int nsnp = an == PBO_AN_LD_WALL ? r->wall_num_snps[(int64_t)w * P + i] : r->ld_num_snps[(int64_t)w * P + i];
            sb_printf(&s, "\tS[%s]:\t%d", o->pop_names[i], nsnp);
            int good = nsnp >= o->min_snps;
            if (an == PBO_AN_LD_ZNS) sb_stat(&s, "Zns", o->pop_names[i], good, r->zns[(int64_t)w * P + i]);
            else if (an == PBO_AN_LD_OMEGA) sb_stat(&s, "omax", o->pop_names[i], good, r->omegamax[(int64_t)w * P + i]);
            else { sb_stat(&s, "B", o->pop_names[i], good, r->wallb[(int64_t)w * P + i]); sb_stat(&s, "Q", o->pop_names[i], good, r->wallq[(int64_t)w * P + i]); }
        }
    } else if (an == PBO_AN_DIVERGE_IND) {
        for (int i = 0; i < n; ++i) {
            double pd = (double)r->ind_div[(int64_t)w * n + i] / ns;
            if (o->jc) pd = -0.75 * log(1.0 - pd * (4.0 / 3.0));
            sb_stat(&s, "d", o->sample_names[i], ok, pd);
        }
    } else if (an == PBO_AN_DIVERGE_POP) {
        for (int i = 0; i < P; ++i) {
            if (ok) {
                int fx = r->pop_div[(int64_t)w * P + i], sg = r->div_num_snps[(int64_t)w * P + i];
                double pd = (p->flags & PBO_FLAG_SUBSTITUTE) ? (double)fx / ns : (double)(fx + sg) / ns;
                if (o->jc) pd = -0.75 * log(1.0 - pd * (4.0 / 3.0));
                sb_printf(&s, "\tFixed[%s]:\t%d\tSeg[%s]:\t%d\td[%s]:\t%.5f", o->pop_names[i], fx, o->pop_names[i], sg, o->pop_names[i], pd);
            } else
                sb_printf(&s, "\tFixed[%s]:\t%7s\tSeg[%s]:\t%7s\td[%s]:\t%7s", o->pop_names[i], "NA", o->pop_names[i], "NA", o->pop_names[i], "NA");
        }
    } else if (an == PBO_AN_HAPLO_K) {
        for (int i = 0; i < P; ++i) {
            if (ok) sb_printf(&s, "\tK[%s]:\t%d\tKdiv[%s]:\t%.5f", o->pop_names[i], r->nhaps[(int64_t)w * P + i], o->pop_names[i], 1.0 - r->hdiv[(int64_t)w * P + i]);
            else sb_printf(&s, "\tK[%s]:\t%7s\tKdiv[%s]:\t%7s", o->pop_names[i], "NA", o->pop_names[i], "NA");
        }
    } else if (an == PBO_AN_HAPLO_EHHS) {
        for (int i = 0; i < P; ++i) { double e = r->ehhs[(int64_t)w * P + i]; sb_stat(&s, "EHHS", o->pop_names[i], ok && !isnan(e), e); }
    } else if (an == PBO_AN_HAPLO_DXY) {
        for (int i = 0; i < P; ++i) sb_stat(&s, "pi", o->pop_names[i], ok, r->piw[(int64_t)w * P + i]);
        for (int i = 0; i < P - 1; ++i) for (int j = i + 1; j < P; ++j) {
            char nm[256]; snprintf(nm, sizeof nm, "%s-%s", o->pop_names[i], o->pop_names[j]);
            int x = i * P + (j - (i + 1));
            if (ok) sb_printf(&s, "\tdxy[%s]:\t%.5f\tmin[%s]:\t%u", nm, r->pib[(int64_t)w * P * P + x], nm, (unsigned)r->min_dxy[(int64_t)w * P * P + x]);
            else sb_printf(&s, "\tdxy[%s]:\t%7s\tmin[%s]:\t%7s", nm, "NA", nm, "NA");
        }
    } else return -1;
    sb_printf(&s, "\n");
    return s.len;
}
