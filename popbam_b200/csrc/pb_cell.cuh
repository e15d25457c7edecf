// pb_cell.cuh -- per-(site,sample) consensus call and per-site logic, as inlinable device functions.
//
// These restate, operation for operation, the arithmetic of the reference's
//   errmod_cal   (pop_utils.cpp:280-365)    likelihoods from the sorted multiset of base codes
//   gl2cns       (pop_utils.cpp:66-100)     best / runner-up genotype -> snpQ
//   call_base    (popbam.cpp:288-298)       rms mapping quality, cb word packing
//   clean_heterozygotes / segbase / qfilter / cal_site_type
//                (pop_utils.cpp:170-201, :122-168, :102-120; popbam.cpp:173-184)
// Bit-exactness rules: every double multiply and add rounds separately (the
// reference is x86-64 -O2 without FMA contraction), float accumulators of doubles round after each
// add, tables come from the host.  Hence the explicit _rn intrinsics below.
//
// The functions are also compiled for the host (PB_HOST_EMU, g++ -ffp-contract=off) by the unit
// test harness tests/hd_harness.cpp, which checks them against the reference's known-answer vectors
// without a GPU.  That harness is test infrastructure; the library has no host compute path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PB_DMUL(a, b) __dmul_rn((a), (b))
#define PB_DADD(a, b) __dadd_rn((a), (b))
#define PB_FSUB(a, b) __fsub_rn((a), (b))
#define PB_FDIV(a, b) __fdiv_rn((a), (b))
#define PB_FSQRT(a) __fsqrt_rn((a))
#define PB_D2F(a) __double2float_rn((a))
#define PB_LDG(p) __ldg((p))
#else
#include <math.h>
#define PB_DMUL(a, b) ((double)(a) * (double)(b))
#define PB_DADD(a, b) ((double)(a) + (double)(b))
#define PB_FSUB(a, b) ((float)(a) - (float)(b))
#define PB_FDIV(a, b) ((float)(a) / (float)(b))
#define PB_FSQRT(a) sqrtf((a))
#define PB_D2F(a) ((float)(a))
#define PB_LDG(p) (*(p))
#endif

// nt16 -> nt4 (popbam.cpp:9): A=1->0, C=2->1, G=4->2, T=8->3, everything else 4; one nibble per entry
#define PB_NT16_NT4_LUT 0x4444444344424104ULL

// letter -> 0..3 / 14 (popbam.cpp:33-51 iupac_rev)
PB_HD int pb_iupac_rev(int c) {
    c &= ~0x20;   // the table maps both cases
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 14;
}

// One term of the error-model sums (pop_utils.cpp:311): bsum += fk[w] * beta[q<<16|n<<8|c]
PB_HD double pb_errmod_step(double bsum, double fkw, double betav) { return PB_DADD(bsum, PB_DMUL(fkw, betav)); }

// Likelihoods of the 10 genotypes i<=j from the per-base sums (pop_utils.cpp:316-362), followed by
// gl2cns (pop_utils.cpp:66-100) and the rms / cb packing of call_base (popbam.cpp:292-298).
// bsum[b], c[b]: error-model sum and count of base b; k = number of codes (== sum of c), rmsq = sum mapq^2.
PB_HD uint64_t pb_finish_cell(const double bsum[4], const int c[4], int k, int rmsq, const double *__restrict__ lhet) {
    float mn = 3.402823466e+38f, mn_next = 3.402823466e+38f;
    unsigned min_ij = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // homozygous j: float accumulation of the other bases' sums
        float t1 = 0.0f;
        int t2 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i == j) continue;
            t1 = PB_D2F(PB_DADD((double)t1, bsum[i]));
            t2 += c[i];
        }
        float l = (k > 0 && t2) ? t1 : 0.0f;
        if (l < 0.0f) l = 0.0f;
        if (l < mn) { min_ij = (unsigned)(j << 2 | j); mn_next = mn; mn = l; }
        else if (l < mn_next) mn_next = l;
#pragma unroll
        for (int m = j + 1; m < 4; ++m) {
            t1 = 0.0f; t2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i == j || i == m) continue;
                t1 = PB_D2F(PB_DADD((double)t1, bsum[i]));
                t2 += c[i];
            }
            float h = 0.0f;
            if (k > 0) {
                double het = PB_DMUL(-4.343, PB_LDG(lhet + ((c[j] + c[m]) << 8 | c[m])));
                h = t2 ? PB_D2F(PB_DADD(het, (double)t1)) : PB_D2F(het);
                if (h < 0.0f) h = 0.0f;
            }
            if (h < mn) { min_ij = (unsigned)(j << 2 | m); mn_next = mn; mn = h; }
            else if (h < mn_next) mn_next = h;
        }
    }
    uint64_t snpq = (uint64_t)PB_DADD((double)PB_FSUB(mn_next, mn), 0.499);
    uint64_t cb = (snpq << 32) + ((uint64_t)(unsigned)k << 16) + ((uint64_t)min_ij << 8);
    if (k > 0) {
        uint64_t rms = (uint64_t)PB_DADD((double)PB_FSQRT(PB_FDIV((float)rmsq, (float)k)), 0.499);
        cb |= rms << 48;
    }
    return cb;
}

// clean_heterozygotes + segbase + qfilter + cal_site_type for ONE sample of a site; the three reference
// loops are independent per sample except for segbase's derived-allele counts, which accumulate in
// *cnt4 (one byte per base; n <= 64).  Returns the final cb word; *covered = passes qfilter,
// *derived = (cb & 3) == 3 (cal_site_type).  r = iupac_rev[ref].
PB_HD uint64_t pb_site_sample(uint64_t w, int ref, int r, int het_mode, int min_snpQ, int min_rmsQ, int min_depth, int max_depth,
                              uint32_t *cnt4, bool *covered, bool *derived) {
    if (!het_mode) {   // clean_heterozygotes (pop_utils.cpp:170-201): both tests use the original alleles
        const int g = (int)((w >> 8) & 0xff), a1 = (g >> 2) & 3, a2 = g & 3, sq = (int)((w >> 32) & 0xffff);
        if (a1 != a2) {
            if (sq >= min_snpQ) {
                if (a1 == r) w += (uint64_t)(int64_t)((a2 - a1) * 1024);
                if (a2 == r) w -= (uint64_t)(int64_t)((a2 - a1) * 256);
            } else {
                if (a1 != r) w += (uint64_t)(int64_t)((a2 - a1) * 1024);
                if (a2 != r) w -= (uint64_t)(int64_t)((a2 - a1) * 256);
            }
        }
    }
    {   // segbase (pop_utils.cpp:122-168)
        const int g = (int)((w >> 8) & 0xff), a1 = (g >> 2) & 3, a2 = g & 3, sq = (int)((w >> 32) & 0xffff);
        // iupac[g] for a homozygous genotype byte (a1 == a2) is "ACGT"[a1]; the byte is < 16 here because
        // the steps before only produce two-bit alleles
        const int letter = (g < 16) ? (int)"ACGT"[a1] : 'N';
        if (a1 == a2 && letter != ref) {
            if (sq >= min_snpQ) { w |= 2; *cnt4 += 1u << (8 * a1); }
            else {
                w -= (uint64_t)(int64_t)((g - r) * 256);
                w -= (uint64_t)(int64_t)((g - r) * 1024);
            }
        }
    }
    {   // qfilter (pop_utils.cpp:102-120)
        const int rms = (int)((w >> 48) & 0xffff), nr = (int)((w >> 16) & 0xffff);
        const bool ok = rms >= min_rmsQ && nr >= min_depth && nr <= max_depth;
        if (ok) w |= 1;
        *covered = ok;
    }
    *derived = (w & 3) == 3;
    return w;
}

// segbase's return value from the derived-allele counts: -1 if more than one derived base, else the count
PB_HD int pb_site_fq(uint32_t cnt4) {
    const int c0 = cnt4 & 255, c1 = (cnt4 >> 8) & 255, c2 = (cnt4 >> 16) & 255, c3 = cnt4 >> 24;
    const int nder = (c0 > 0) + (c1 > 0) + (c2 > 0) + (c3 > 0);
    return nder > 1 ? -1 : c0 + c1 + c2 + c3;
}

// All n samples of one site (in place); kept for the unit-test harness and small callers.
PB_HD int pb_site_logic(uint64_t *cb, int stride, int n, int ref, int het_mode, int min_snpQ, int min_rmsQ, int min_depth,
                        int max_depth, uint64_t *cov_out, uint64_t *type_out) {
    const int r = pb_iupac_rev(ref);
    uint32_t cnt4 = 0;
    uint64_t cov = 0, type = 0;
    for (int i = 0; i < n; ++i) {
        bool c, d;
        cb[i * stride] = pb_site_sample(cb[i * stride], ref, r, het_mode, min_snpQ, min_rmsQ, min_depth, max_depth, &cnt4, &c, &d);
        if (c) cov |= 1ULL << i;
        if (d) type |= 1ULL << i;
    }
    *cov_out = cov; *type_out = type;
    return pb_site_fq(cnt4);
}
