// refdump -- known-answer vector generator linked against the UNMODIFIED reference objects
// (oracle/_ref/obj/*.o, built by oracle/Makefile from /root/reference).  Test infrastructure.
//
//   refdump tables OUT          fk[256] | beta[64*256*256] | lhet[256*256] (doubles) from
//                               errmod_init(1.0-0.83)              (pop_utils.cpp:257)
//   refdump cells SEED N OUT    N random (site,sample) cells: the reference's errmod_cal +
//                               gl2cns + the rms packing of popbam.cpp:288-298
//   refdump sites SEED N OUT    N random sites of n samples through clean_heterozygotes,
//                               segbase, qfilter (pop_nucdiv.cpp:164-176 order)
//
// Record layouts are documented in tests/golden/README.md and parsed by tests/refkat.py.
#include "popbam.h"
#include <cstdio>
#include <cstdint>
#include <vector>

static uint64_t sm64(uint64_t &x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: refdump tables OUT | cells SEED N OUT | sites SEED N OUT\n"); return 2; }
    std::string mode = argv[1];
    errmod_t *em = errmod_init(1.0 - 0.83);
    if (mode == "tables") {
        FILE *f = fopen(argv[2], "wb");
        if (!f) return 1;
        fwrite(em->coef->fk, sizeof(double), 256, f);
        fwrite(em->coef->beta, sizeof(double), 64 * 256 * 256, f);
        fwrite(em->coef->lhet, sizeof(double), 256 * 256, f);
        fclose(f);
        return 0;
    }
    if (argc < 5) return 2;
    uint64_t seed = strtoull(argv[2], 0, 10);
    int N = atoi(argv[3]);
    FILE *f = fopen(argv[4], "wb");
    if (!f) return 1;
    if (mode == "cells") {
        // record: u16 k | u16 pad | i32 rmsq | u16 codes[256] (unsorted input) | float q[16] | u64 cb
        static const int quals[] = {4, 10, 13, 20, 29, 30, 35, 40, 41, 60, 63};
        for (int r = 0; r < N; ++r) {
            uint64_t w = sm64(seed);
            int style = w & 7;
            int k = style == 0 ? (int)((w >> 8) % 256) : style == 1 ? (int)((w >> 8) % 4) : 1 + (int)((w >> 8) % 60);
            int major = (w >> 20) & 3, minor = (w >> 22) & 3;
            double pminor = style == 2 ? 0.5 : style == 3 ? 0.02 : ((w >> 24) & 0xff) / 512.0;
            uint16_t codes[256] = {0}, work[256];
            int rmsq = 0;
            for (int i = 0; i < k; ++i) {
                uint64_t u = sm64(seed);
                int q = (style == 4) ? 4 + (int)(u % 60) : quals[u % 11];
                int strand = (u >> 16) & 1;
                int b = ((u >> 20) & 0xffff) / 65536.0 < pminor ? minor : major;
                if (((u >> 40) & 0xff) < 3) b = (u >> 48) & 3;
                codes[i] = (uint16_t)(q << 5 | strand << 4 | b);
                int mq = (int)((u >> 52) & 0xff);
                rmsq += mq * mq;
            }
            memcpy(work, codes, sizeof work);
            float q[16];
            errmod_cal(em, (unsigned short)k, NBASES, work, q);
            unsigned long long rms = (unsigned long long)(sqrt((float)(rmsq) / k) + 0.499);
            unsigned long long cb = gl2cns(q, (unsigned short)k);
            cb |= rms << (CHAR_BIT * 6);
            uint16_t hk[2] = {(uint16_t)k, 0};
            int32_t rq = rmsq;
            uint64_t cb64 = cb;
            fwrite(hk, 2, 2, f); fwrite(&rq, 4, 1, f); fwrite(codes, 2, 256, f); fwrite(q, 4, 16, f); fwrite(&cb64, 8, 1, f);
        }
    } else if (mode == "sites") {
        // record: i32 n | i32 min_snpq | i32 min_rmsq | i32 min_depth | i32 max_depth | i32 het | u8 ref | pad[7]
        //         | u64 cb_in[64] | u64 cb_out[64] | i32 fq | i32 pad | u64 cov
        static const char refs[] = {'A', 'C', 'G', 'T', 'a', 'c', 'g', 't', 'N', 'n'};
        for (int r = 0; r < N; ++r) {
            uint64_t w = sm64(seed);
            int n = 1 + (int)(w % 64);
            int min_snpq = (int)((w >> 8) % 3 == 0 ? (w >> 10) % 60 : 25);
            int min_rmsq = (int)((w >> 20) % 3 == 0 ? (w >> 22) % 60 : 25);
            int min_depth = (int)((w >> 30) % 4 == 0 ? (w >> 32) % 10 : 3);
            int max_depth = (int)((w >> 40) % 4 == 0 ? 10 + (w >> 42) % 200 : 255);
            int het = (int)((w >> 52) & 1);
            char ref = refs[(w >> 54) % ((w >> 60) ? 4 : 10)];
            unsigned long long in[64] = {0}, out[64] = {0};
            for (int i = 0; i < n; ++i) {
                uint64_t u = sm64(seed);
                if ((u & 15) == 0) { in[i] = 0; continue; }          // sample without reads
                unsigned a1 = (u >> 4) & 3, a2 = (u >> 6) & 3;
                if ((u >> 8) & 3) a2 = a1;                            // mostly homozygous
                if (a1 > a2) { unsigned t = a1; a1 = a2; a2 = t; }    // gl2cns emits i<=j
                unsigned long long snpq = ((u >> 10) & 3) == 0 ? (u >> 12) % 300 : (u >> 12) % 60;
                unsigned long long k = ((u >> 24) & 7) == 0 ? (u >> 27) % 256 : (u >> 27) % 40;
                unsigned long long rms = (u >> 40) % 64;
                in[i] = (snpq << 32) + (k << 16) + ((unsigned long long)(a1 << 2 | a2) << 8);
                in[i] |= rms << 48;
            }
            memcpy(out, in, sizeof out);
            if (!het) clean_heterozygotes(n, out, (int)ref, min_snpq);
            int fq = segbase(n, out, ref, min_snpq);
            unsigned long long cov = qfilter(n, out, min_rmsq, min_depth, max_depth);
            int32_t hdr[6] = {n, min_snpq, min_rmsq, min_depth, max_depth, het};
            uint8_t rb[8] = {(uint8_t)ref, 0, 0, 0, 0, 0, 0, 0};
            int32_t fqv[2] = {fq, 0};
            uint64_t cov64 = cov;
            fwrite(hdr, 4, 6, f); fwrite(rb, 1, 8, f); fwrite(in, 8, 64, f); fwrite(out, 8, 64, f); fwrite(fqv, 4, 2, f); fwrite(&cov64, 8, 1, f);
        }
    } else return 2;
    fclose(f);
    return 0;
}
