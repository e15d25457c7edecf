// pb_fast.cuh -- the bit-sliced pileup: the default formulation of the pileup / call / site stage.
// Used when the raw-depth cap cannot bind (k_depth_bound), the caller did not ask for the
// per-(site,sample) words, and min_depth / min_snpQ are positive; k_pileup_call otherwise.
//
// Why.  k_pileup_call spends ~29 warp instructions per (record, 32 positions): one code load, one
// histogram update, one total per LANE per record.  But ~97 % of the cells are "easy": every passing
// base equals the reference base (or all but one do) and the depth alone proves that call_base would
// call the cell homozygous reference (pb_need_entry, pb_one_stray_entry in pb_walk.cuh), so all the
// site needs from the cell is qfilter's coverage bit (pop_utils.cpp:102-120).  For such cells nothing
// per base has to be looked at one by one: with the bases' properties stored as BIT-PLANES (one bit per
// base), a single THREAD handles 32 positions of one sample with 32-bit logic operations:
//     planes (k_planes, from qual[] / seq4[]):  P passing base, B0/B1 its two base bits, H quality level >= hi
//     per record:  window = funnel-shift of the planes to the strip, masked to the segment
//                  stray    = P & ((B0 ^ R0) | (B1 ^ R1))          (R: the reference strip's planes), counted 0 / 1 / 2+
//                  lowq    |= P when the read's mapQ < min_rmsQ    (else rms >= min_rmsQ is guaranteed)
//                  k  += P,  khi += H    as bit-sliced counters (half-adder chains over 6 / 4 planes)
//     per strip:   easy = no lowq, and (no stray base, k or khi in a proven range) or (one stray base, k in a proven range)
// i.e. ~70 thread instructions per record for 32 cells instead of ~29 warp instructions (928 thread
// slots).  The remaining cells (two stray bases, a variant, a low-mapQ read, an odd depth) are listed
// in bit masks and called one cell per thread by k_hard_cells with the exact machinery of
// pb_cell.cuh / pb_walk.cuh; k_fast_sites puts the two together into the per-site result.
// Kernels: k_ref_planes (per contig), k_fast_params (per level set), k_planes, k_strip_index,
// k_pile_fast, k_hard_cells, k_fast_sites.
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>
#include "pb_kernels.cuh"

#define PB_H_QUALITY 30            // quality value of the H plane

struct PbFastParams {        // derived from the need table (k_fast_params), cached with it
    int k0lo, k0hi;          // for k0lo <= k <= k0hi: need[0][k] != 0 and need[0][k] <= k  (the depth alone suffices)
    int k1lo, k1hi, hmin;    // for k1lo <= k <= k1hi: need[hi][k] != 0 and <= hmin <= 15   (khi >= hmin suffices)
    int hi_level;            // level of the H plane (n_levels: none)
    int k2lo, k2hi;          // for k2lo <= k <= k2hi: one stray base cannot change a homozygous call (pb_one_stray_entry)
    int k3lo, k3hi;          // the same for a stray base below the H plane's level (a wider range of depths)
};

// ---- bit-planes of the bases, interleaved: planes[1 + w] = {P, B0, B1, H} of the 32 bases at byte offsets 32w .. 32w+31
// of qual[] (one pad element in front, PB_PLANE_PAD behind):
//     P   the base passes call_base's filters (popbam.cpp:268-284): read kept and mapQ >= min_mapQ, baseQ' >= min_baseQ, A/C/G/T
//     B0, B1   its two base bits (A 0, C 1, G 2, T 3), zero where P is zero
//     H   P and clamp(min(baseQ', mapQ), 4, 63) >= the quality value of level hi  (<=> baseQ' >= qv and mapQ >= qv)
// straight from qual[] / seq4[] with packed arithmetic -- no per-base table lookups and no codes[] in between:
// "byte >= T" for four quality bytes is three operations (T <= 128: add 128 - T to the low seven bits, or in bit 7),
// the four flags of a word are gathered by one multiply (0x00204081 moves byte i's bit 7 to bit 28 + i; the sixteen
// partial products land on distinct bits, so nothing carries), a byte of seq4 (two bases) is one shared-memory lookup
// giving "is A/C/G/T" and the two base bits of both bases, and per-read conditions (alive, mapQ >= qv) are bit masks
// below / above the one read boundary a 64-byte chunk can have.  Chunks with several boundaries (reads shorter than 64 bases)
// and the tail of the array go byte by byte.
#define PB_PLANE_PAD 16
#define PB_PL_CHUNK 64
__device__ __forceinline__ uint32_t pb_ge4(uint32_t w, uint32_t add) { return ((((w & 0x7f7f7f7fu) + add) | w) & 0x80808080u) * 0x00204081u >> 28; }
__global__ void __launch_bounds__(256) k_planes(int64_t n, const uint32_t *__restrict__ meta, int n_samples,
                                                const uint64_t *__restrict__ base, const uint8_t *__restrict__ seq4,
                                                const uint8_t *__restrict__ qual, int64_t n_bytes, double reads_per_byte,
                                                int min_mapQ, int min_baseQ, int illumina, PbCounters *__restrict__ ctr,
                                                uint4 *__restrict__ planes) {
    // seq4 byte -> valid bits 0-1, B0 bits 8-9, B1 bits 16-17 of its two bases (high nibble = first base = lower bit)
    __shared__ uint32_t seq_s[256];
    {
        const uint32_t n0 = (uint32_t)((PB_NT16_NT4_LUT >> ((threadIdx.x >> 4) * 4)) & 0xf), n1 = (uint32_t)((PB_NT16_NT4_LUT >> ((threadIdx.x & 15) * 4)) & 0xf);
        const uint32_t v0 = n0 < 4u, v1 = n1 < 4u;
        seq_s[threadIdx.x] = v0 | v1 << 1 | (v0 & n0 & 1u) << 8 | (v1 & n1 & 1u) << 9 | (v0 & (n0 >> 1) & 1u) << 16 | (v1 & (n1 >> 1) & 1u) << 17;
    }
    // which raw quality values occur (k_level_table builds the region's quality levels from them): one flag per value
    __shared__ uint8_t seen[256];
    seen[threadIdx.x] = 0;
    __syncthreads();
    const int qv = PB_H_QUALITY;                                                        // quality value of the H plane
    const int tp = min_baseQ <= 0 ? 0 : min_baseQ + (illumina ? 31 : 0);                // raw quality byte thresholds (host: tp <= 128)
    const int th = min(128, qv + (illumina ? 31 : 0));
    const uint32_t addP = (uint32_t)(128 - tp) * 0x01010101u, addH = (uint32_t)(128 - th) * 0x01010101u;
    // per-read conditions from meta alone (so this pass does not wait for k_read_prep): bit 0 alive -- not flagged
    // 0x704 (bam_pileup.c:371-374), a listed sample, mapQ >= min_mapQ -- bit 1 mapQ >= qv.  A read without reference span
    // is dropped by k_read_prep as well, but no segment record points at its bases, so its bits are never looked at.
    auto flags_of = [&](int64_t rr) -> uint32_t {
        if (rr >= n) return 0u;
        const uint32_t m = __ldg(meta + rr);
        const int mq = (int)((m >> 8) & 0xffu);
        if (((m >> 16) & 0x704u) || (m & 0xffu) >= (uint32_t)n_samples || mq < min_mapQ) return 0u;
        return 1u | (min(mq, 63) >= qv ? 2u : 0u);
    };
    const int64_t n_chunks = (n_bytes + PB_PL_CHUNK - 1) / PB_PL_CHUNK;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_chunks; t += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t o = (uint64_t)t * PB_PL_CHUNK;
        const int nb = (int)min((uint64_t)PB_PL_CHUNK, (uint64_t)n_bytes - o);
        // owner of byte o: last r with base[r] <= o (reads lie back to back: start from a proportional guess)
        int64_t lo = (int64_t)((double)o * reads_per_byte);
        if (lo >= n) lo = n - 1;
        int64_t hi2 = lo + 1, step = 1;
        while (lo > 0 && __ldg(base + lo) > o) { hi2 = lo; lo = max((int64_t)0, lo - step); step <<= 1; }
        step = 1;
        while (hi2 < n && __ldg(base + hi2) <= o) { lo = hi2; hi2 = min(n, hi2 + step); step <<= 1; }
        while (hi2 - lo > 1) { const int64_t mid = (lo + hi2) >> 1; if (__ldg(base + mid) <= o) lo = mid; else hi2 = mid; }
        const int64_t r = lo;
        const uint64_t next = r + 1 < n ? __ldg(base + r + 1) : ~0ULL;
        const uint64_t next2 = r + 2 < n ? __ldg(base + r + 2) : ~0ULL;
        uint32_t P[2] = {0, 0}, B0[2] = {0, 0}, B1[2] = {0, 0}, H[2] = {0, 0};
        if (nb == PB_PL_CHUNK && next2 - o >= (uint64_t)PB_PL_CHUNK) {
            uint4 qv4[4], sv[2];
#pragma unroll
            for (int v = 0; v < 4; ++v) qv4[v] = __ldg(reinterpret_cast<const uint4 *>(qual + o) + v);
#pragma unroll
            for (int v = 0; v < 2; ++v) sv[v] = __ldg(reinterpret_cast<const uint4 *>(seq4 + (o >> 1)) + v);
            const uint32_t qw[16] = {qv4[0].x, qv4[0].y, qv4[0].z, qv4[0].w, qv4[1].x, qv4[1].y, qv4[1].z, qv4[1].w,
                                     qv4[2].x, qv4[2].y, qv4[2].z, qv4[2].w, qv4[3].x, qv4[3].y, qv4[3].z, qv4[3].w};
            const uint32_t sw[8] = {sv[0].x, sv[0].y, sv[0].z, sv[0].w, sv[1].x, sv[1].y, sv[1].z, sv[1].w};
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                P[g >> 3] |= pb_ge4(qw[g], addP) << (4 * (g & 7));
                H[g >> 3] |= pb_ge4(qw[g], addH) << (4 * (g & 7));
#pragma unroll
                for (int b = 0; b < 4; ++b) seen[(qw[g] >> (8 * b)) & 0xffu] = 1;
            }
            uint32_t V[2] = {0, 0};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // four lookups per word: 8 bases -> V in bits 0-7, B0 in 8-15, B1 in 16-23
                const uint32_t acc = seq_s[sw[j] & 0xffu] | seq_s[(sw[j] >> 8) & 0xffu] << 2 | seq_s[(sw[j] >> 16) & 0xffu] << 4 | seq_s[sw[j] >> 24] << 6;
                V[j >> 2] |= (acc & 0xffu) << (8 * (j & 3)); B0[j >> 2] |= ((acc >> 8) & 0xffu) << (8 * (j & 3)); B1[j >> 2] |= ((acc >> 16) & 0xffu) << (8 * (j & 3));
            }
            // per-read conditions below / above the boundary
            const uint32_t fA = flags_of(r);
            const uint32_t cut = next - o < (uint64_t)PB_PL_CHUNK ? (uint32_t)(next - o) : (uint32_t)PB_PL_CHUNK;
            const uint32_t fB = cut < PB_PL_CHUNK ? flags_of(r + 1) : fA;
            const uint64_t below = cut >= 64 ? ~0ULL : (1ULL << cut) - 1ULL;
            const uint64_t alive = ((fA & 1u) ? below : 0ULL) | ((fB & 1u) ? ~below : 0ULL);
            const uint64_t hiok = ((fA & 2u) ? below : 0ULL) | ((fB & 2u) ? ~below : 0ULL);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                P[h] &= V[h] & (uint32_t)(alive >> (32 * h));
                H[h] &= P[h] & (uint32_t)(hiok >> (32 * h));
                B0[h] &= P[h]; B1[h] &= P[h];
            }
        } else {
            int64_t rr = r;
            uint64_t nx = next;
            uint32_t f = flags_of(rr);
            for (int i = 0; i < nb; ++i) {
                while (o + i >= nx) {
                    ++rr;
                    nx = rr + 1 < n ? __ldg(base + rr + 1) : ~0ULL;
                    f = flags_of(rr);
                }
                const uint32_t q = qual[o + i];
                seen[q] = 1;
                const uint32_t sbyte = seq4[(o + i) >> 1];
                const uint32_t nib = ((o + i) & 1) ? (sbyte & 15u) : (sbyte >> 4);
                const uint32_t nt = (uint32_t)((PB_NT16_NT4_LUT >> (nib * 4)) & 0xf);
                if (!(f & 1u) || nt > 3u || (int)q < tp) continue;
                const uint32_t bit = 1u << (i & 31);
                P[i >> 5] |= bit;
                if (nt & 1u) B0[i >> 5] |= bit;
                if (nt & 2u) B1[i >> 5] |= bit;
                if ((f & 2u) && (int)q >= th) H[i >> 5] |= bit;
            }
        }
        planes[2 * t + 1] = make_uint4(P[0], B0[0], B1[0], H[0]);
        if (nb > 32) planes[2 * t + 2] = make_uint4(P[1], B0[1], B1[1], H[1]);
    }
    if (blockIdx.x == 0 && threadIdx.x <= PB_PLANE_PAD) {                           // pad: one in front, PB_PLANE_PAD behind
        const int64_t n_words = (n_bytes + 31) >> 5;
        const int64_t i = threadIdx.x ? n_words + threadIdx.x : 0;
        planes[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    // one thread per raw quality value: transform, filter, clamp (as k_qual_mask)
    unsigned long long mask = 0;
    if (seen[threadIdx.x]) {
        int q = (int)threadIdx.x;
        if (illumina) q = q > 31 ? q - 31 : 0;
        if (q >= min_baseQ) mask = 1ULL << (q > 63 ? 63 : q);
    }
    for (int o = 16; o > 0; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
    if ((threadIdx.x & 31) == 0 && mask) atomicOr(&ctr->qual_mask, mask);
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
    int M, NI, RC;                           // RC: records staged per pass
    const uint4 *planes;                     // {P, B0, B1, H} per 32 code bytes (one pad element in front)
    const uint32_t *r0, *r1, *rv;            // reference planes, bit = absolute position
    int64_t ref_len;
    int span_beg, span_end;
    int n_samples, n_strips, n_sblocks;      // n_sblocks = strip blocks of PB_FAST_STRIPS strips
    int min_depth, min_rmsQ;
    int W;                                   // plane elements staged per record: covers 31 + the longest segment
    int max_span;                            // longest reference span of a kept read
    const PbCounters *ctr;
    const PbFastParams *fp;
    uint32_t *cov32, *hard32;                // [n_samples][n_strips]
    uint32_t *hcount;                        // [n_samples * n_strips + 1] popc(hard32), scanned into cell offsets afterwards
};

#define PB_FAST_STRIPS 64          // strips of 32 positions per CTA (one sample)
#define PB_FAST_G 4                // threads sharing a strip (each takes every G-th record; partial counters are added)
static inline int pb_fast_words(int max_span) { return ((31 + (max_span > 0 ? max_span - 1 : 0)) >> 5) + 1; }
static inline int pb_fast_rc(int W) { const int rc = (48 * 1024) / (16 * (1 + W)); return rc > 512 ? 512 : rc; }   // records staged per pass
static inline size_t pb_fast_smem(int W) { return (size_t)pb_fast_rc(W) * 16 * (1 + (size_t)W); }

// One CTA = one sample x 64 strips of 32 positions; PB_FAST_G threads share a strip.
// The CTA stages its records (read start in (first position - max_span, last position]) and the plane
// elements their bases live in into shared memory with wide coalesced-per-record loads -- every record is
// needed by four or five neighbouring strips, and staging makes that reuse explicit instead of leaving it
// to L2 (the unstaged version of this kernel moved 18 GB through L2 for 0.2 GB of distinct data).
// Then every thread walks the records that can cover its strip.
__global__ void __launch_bounds__(PB_FAST_STRIPS * PB_FAST_G) k_pile_fast(const PbFastArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4 *recS = reinterpret_cast<int4 *>(smem_raw);                      // [RC] {seg start, z, bit offset in staged words, read start}
    uint4 *plS = reinterpret_cast<uint4 *>(recS + a.RC);                  // [RC][W]
    const int tid = threadIdx.x;
    const int sb = (int)(blockIdx.x / a.n_samples), s = (int)(blockIdx.x % a.n_samples);     // the samples of a strip block run together: one pass through L2
    const int strip = sb * PB_FAST_STRIPS + tid / PB_FAST_G, g = tid % PB_FAST_G;
    const int W = a.W;
    const uint32_t *Fs = a.F + (size_t)s * a.NI;
    const int t0 = sb * PB_FAST_STRIPS;
    const uint32_t clo = __ldg(Fs + t0), chi = __ldg(Fs + min(t0 + PB_FAST_STRIPS, a.n_strips) + a.M);
    const int S = a.span_beg + strip * 32;
    const bool live = strip < a.n_strips;
    const int strip_last = min(S + 31, a.span_end - 1);
    const uint32_t my_lo = live ? __ldg(Fs + strip) : 0u, my_hi = live ? __ldg(Fs + strip + a.M + 1) : 0u;
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
    for (uint32_t c0 = clo; c0 < chi; c0 += (uint32_t)a.RC) {
        const int cnt = (int)min((uint32_t)a.RC, chi - c0);
        if (c0 != clo) __syncthreads();
        for (int r = tid; r < cnt; r += PB_FAST_STRIPS * PB_FAST_G) {
            const int4 rc = __ldg(&a.srec[c0 + r]);
            const uint32_t z = (uint32_t)rc.z;
            const int64_t gb = (int64_t)(((uint64_t)(z >> 26) << 32) | (uint32_t)rc.w) + 32;      // bit index of the segment's first base
            const uint4 *src = a.planes + (gb >> 5);
            recS[r] = make_int4(rc.y, rc.z, (int)(gb & 31), rc.x);
            // asynchronous 16-byte copies global -> shared (no registers, no wait until the whole pass is issued)
            uint4 *dst = plS + (size_t)r * W;
            for (int k = 0; k < W; ++k) __pipeline_memcpy_async(dst + k, src + k, 16);
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncthreads();
        if (live) {
            int lo = (int)(max(my_lo, c0) - c0);
            const int j1 = (int)min((uint32_t)cnt, max(my_hi, c0) - c0);
            // the strip index is 32 positions coarse: step over the leading records whose read ends before the strip
            while (lo < j1 && recS[lo].w + a.max_span <= S) ++lo;
            for (int j = lo + g; j < j1; j += PB_FAST_G) {
                const int4 r = recS[j];
                const uint32_t z = (uint32_t)r.y;
                const int len = (int)(z & 0xffffu);
                const int u = S - r.x;                                   // segment-relative index of the strip's first position
                const int i0 = max(0, -u), i1 = min(32, len - u);
                if (i1 <= i0) continue;
                const uint32_t m = (i1 == 32 ? 0xffffffffu : (1u << i1) - 1u) & ~((1u << i0) - 1u);
                const int t = r.z + u;                                   // bit index in the staged elements (>= -31)
                const int wi = t >> 5, sh = t & 31;
                const uint4 *pw = plS + (size_t)j * W;
                const uint4 zero = make_uint4(0, 0, 0, 0);
                const uint4 lo4 = wi >= 0 ? pw[wi] : zero;
                const uint4 hi4 = pw[min(wi + 1, W - 1)];               // past the staged elements only masked bits are read
                const uint32_t P = __funnelshift_r(lo4.x, hi4.x, sh) & m;
                const uint32_t B0 = __funnelshift_r(lo4.y, hi4.y, sh);
                const uint32_t B1 = __funnelshift_r(lo4.z, hi4.z, sh);
                const uint32_t H = __funnelshift_r(lo4.w, hi4.w, sh) & m;
                const uint32_t mm = P & (((B0 ^ R0) | (B1 ^ R1)) | ~RV);
                mism2 |= mism & mm; mism |= mm; mismH |= mm & H;
                if ((int)((z >> 16) & 0xffu) < a.min_rmsQ) lowq |= P;
                uint32_t carry = P;
#pragma unroll
                for (int b = 0; b < 6; ++b) { const uint32_t tt = ck[b] & carry; ck[b] ^= carry; carry = tt; }
                kover |= carry;
                carry = H;
#pragma unroll
                for (int b = 0; b < 4; ++b) { const uint32_t tt = ch[b] & carry; ch[b] ^= carry; carry = tt; }
                hover |= carry;
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
    const PbFastParams fp = *a.fp;
    static_assert(PB_FAST_G == 4, "four range tests, one per thread of a strip");
    const int lo_g = g == 0 ? fp.k0lo : g == 1 ? fp.k1lo : g == 2 ? fp.k2lo : max(a.min_depth, 0);
    const int hi_g = g == 0 ? fp.k0hi : g == 1 ? ((fp.hmin > 0) ? fp.k1hi : 0) : g == 2 ? fp.k2hi : 63;
    uint32_t in_range = 0;
    if (hi_g >= lo_g && lo_g <= 63) in_range = pb_bs_le6(ck, min(hi_g, 63)) & (lo_g > 0 ? ~pb_bs_le6(ck, lo_g - 1) : 0xffffffffu);
    if (g == 1 && in_range) {
        // khi >= hmin  <=>  not (khi <= hmin - 1), or the 4-plane counter overflowed (>= 16 > hmin)
        uint32_t lt = 0, eq = 0xffffffffu;
        const int C = fp.hmin - 1;
#pragma unroll
        for (int b = 3; b >= 0; --b) {
            const uint32_t cb = 0u - (uint32_t)((C >> b) & 1);
            lt |= eq & ~ch[b] & cb;
            eq &= ~(ch[b] ^ cb);
        }
        in_range &= ~(lt | eq) | hover;
    }
    const int lane0 = (tid & 31) & ~(PB_FAST_G - 1);
    const uint32_t r_k0 = __shfl_sync(0xffffffffu, in_range, lane0), r_k1 = __shfl_sync(0xffffffffu, in_range, lane0 + 1);
    const uint32_t r_k2 = __shfl_sync(0xffffffffu, in_range, lane0 + 2), dge = __shfl_sync(0xffffffffu, in_range, lane0 + 3);
    // fifth range (a low-quality stray base), evaluated by the thread that had the cheapest test
    uint32_t r_k3 = 0;
    if (g == 3 && fp.k3hi >= fp.k3lo) r_k3 = pb_bs_le6(ck, min(fp.k3hi, 63)) & (fp.k3lo > 0 ? ~pb_bs_le6(ck, fp.k3lo - 1) : 0xffffffffu);
    r_k3 = __shfl_sync(0xffffffffu, r_k3, lane0 + 3);
    if (!live || g != 0) return;
    const uint32_t valid = (strip_last - S + 1) >= 32 ? 0xffffffffu : (1u << (strip_last - S + 1)) - 1u;
    const uint32_t nonzero = (ck[0] | ck[1] | ck[2] | ck[3] | ck[4] | ck[5] | kover) & valid;
    // depth alone / count of high-quality bases proves the unanimous shortcut (pb_need_entry); one stray base at a
    // depth where it provably cannot change the homozygous-reference call (pb_one_stray_entry)
    const uint32_t easy = nonzero & ~lowq & ~kover & ((~mism & (r_k0 | r_k1)) | (mism & ~mism2 & (r_k2 | (~mismH & r_k3))));
    // qfilter for easy cells: rms >= min_rmsQ holds because every contributing read has mapQ >= min_rmsQ;
    // depth <= max_depth holds because the cap cannot bind; depth >= min_depth is the fourth range test
    a.cov32[(size_t)s * a.n_strips + strip] = easy & dge;
    a.hard32[(size_t)s * a.n_strips + strip] = nonzero & ~easy;
    a.hcount[(size_t)s * a.n_strips + strip] = (uint32_t)__popc(nonzero & ~easy);
    if (s == 0 && strip == 0) a.hcount[(size_t)a.n_samples * a.n_strips] = 0;
}

struct PbHardArgs {
    const int4 *srec;
    const uint32_t *F;                       // strip index (k_strip_index)
    int M, NI;
    const uint8_t *qual, *seq4;              // the read batch's bases: a cell's codes are formed on the fly (no codes[] on this path)
    const uint8_t *qtab;                     // [64][256] k_qual_table
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
    const uint8_t *need;
    const uint32_t *cov32, *hard32;          // [n_samples][n_strips]
    const uint32_t *hoff;                    // [n_samples * n_strips + 1] exclusive scan of popc(hard32); last = number of hard cells
    uint64_t *acc_cov;                       // [span] coverage bits of the hard cells      (zeroed before k_hard_cells)
    uint32_t *acc_cnt4;                      // [span] derived-base counts of the hard cells (zeroed)
    uint64_t *site_type;                     // [span] derived-allele bits                   (zeroed; the hard cells are the only writers)
    uint8_t *site_flag;
};

#define PB_HARD_THREADS 128
#define PB_HARD_SLICE 2048         // cell offsets of mask words a CTA keeps in shared memory
static inline size_t pb_hard_smem(int nl) {
    return (size_t)2 * nl * PB_HARD_THREADS * 4 + 256 * 8 + 64 + (size_t)nl * 256 + 64 + (PB_HARD_SLICE + 1) * 4;
}

// The cells the bit-sliced pass could not settle, one per THREAD.  The hard masks are numbered by an exclusive
// scan of their sizes (hoff); every CTA takes an equal, contiguous share of the cell numbers, so all warps stay
// busy however unevenly the cells are spread, and consecutive lanes hold neighbouring positions of one sample
// (they read the same segment records: L1 hits).  A thread finds its cell's mask word by a binary search over the
// CTA's slice of hoff in shared memory, walks the sample's records that can cover the strip (strip index, no
// search) in file order -- exactly the bases call_base sees -- into a private shared-memory histogram and
// calls the cell with the exact machinery of pb_cell.cuh / pb_walk.cuh.  What the site needs from the cell
// (pb_site_sample: coverage bit, derived-allele bit, derived-base counts) goes into per-position accumulators
// with integer atomics, so the result does not depend on the order of the cells.
__global__ void __launch_bounds__(PB_HARD_THREADS) k_hard_cells(const PbHardArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_w[2];
    const int nl = a.ctr->n_levels;
    const int n_lw = 2 * nl;
    const int tid = threadIdx.x;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                             // [n_lw][PB_HARD_THREADS]
    double *fk_s = reinterpret_cast<double *>(hist + (size_t)n_lw * PB_HARD_THREADS);
    uint8_t *qval_s = reinterpret_cast<uint8_t *>(fk_s + 256);                           // [64]
    uint8_t *need_s = qval_s + 64;                                                       // [nl][256]
    uint32_t *hoff_s = reinterpret_cast<uint32_t *>(need_s + (size_t)nl * 256 + 64);     // [PB_HARD_SLICE + 1]
    const uint32_t n_words = (uint32_t)a.n_samples * (uint32_t)a.n_strips;
    const uint32_t total = a.hoff[n_words];
    // the CTA's contiguous share of the cells and the mask words they live in
    const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
    const uint32_t g_beg = min(total, blockIdx.x * per), g_end = min(total, g_beg + per);
    if (g_beg >= g_end) return;
    for (int i = tid; i < 256; i += PB_HARD_THREADS) fk_s[i] = a.fk[i];
    if (tid < 64) qval_s[tid] = a.ctr->qval[tid];
    for (int i = tid; i < nl * 64; i += PB_HARD_THREADS) reinterpret_cast<uint32_t *>(need_s)[i] = reinterpret_cast<const uint32_t *>(a.need)[i];
    uint32_t *const my_hist = hist + tid;
    for (int lw = 0; lw < n_lw; ++lw) my_hist[lw * PB_HARD_THREADS] = 0;
    if (tid < 2) {
        const uint32_t g = tid ? g_end - 1 : g_beg;
        uint32_t lo = 0, hi = n_words;                         // last idx with hoff[idx] <= g
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.hoff + mid) > g) hi = mid; else lo = mid + 1; }
        s_w[tid] = lo - 1;
    }
    __syncthreads();
    const uint32_t w_beg = s_w[0], w_cnt = s_w[1] - s_w[0] + 1;
    const bool sliced = w_cnt <= PB_HARD_SLICE;
    if (sliced) for (uint32_t i = tid; i <= w_cnt; i += PB_HARD_THREADS) hoff_s[i] = __ldg(a.hoff + w_beg + i);
    __syncthreads();
    for (uint32_t g = g_beg + tid; g < g_end; g += PB_HARD_THREADS) {
        uint32_t idx;
        if (sliced) {
            uint32_t lo = 0, hi = w_cnt;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (hoff_s[mid] > g) hi = mid; else lo = mid + 1; }
            idx = w_beg + lo - 1;
        } else {
            uint32_t lo = w_beg, hi = w_beg + w_cnt;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.hoff + mid) > g) hi = mid; else lo = mid + 1; }
            idx = lo - 1;
        }
        const int bit = (int)__fns(__ldg(a.hard32 + idx), 0, (int)(g - __ldg(a.hoff + idx)) + 1);
        const int smp = (int)(idx / (uint32_t)a.n_strips), strip = (int)(idx % (uint32_t)a.n_strips);
        const int pos = a.span_beg + strip * 32 + bit;
        // ---- call_base for (pos, smp): the records that can cover the strip, in file order
        const uint32_t *Fs = a.F + (size_t)smp * a.NI + strip;
        const uint32_t j0 = __ldg(Fs), j1 = __ldg(Fs + a.M + 1);
        uint32_t tot4 = 0;
        int rmsq = 0;
        for (uint32_t j = j0; j < j1; j += 8) {               // eight records per step: their code loads (DRAM latency) overlap
            uint32_t code[8], zz[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                code[q] = PB_CODE_NONE; zz[q] = 0;
                if (j + q < j1) {
                    const int4 r = __ldg(&a.srec[j + q]);
                    const uint32_t z = (uint32_t)r.z;
                    const uint32_t u = (uint32_t)(pos - r.y);
                    zz[q] = z;
                    if (u < (z & 0xffffu)) {
                        // call_base's filter and code for this base (popbam.cpp:268-284), as k_encode forms them
                        const int64_t off = (int64_t)(((uint64_t)(z >> 26) << 32) | (uint32_t)r.w) + u;
                        const uint32_t qb = __ldg(a.qual + off), sb = __ldg(a.seq4 + (off >> 1));
                        const uint32_t nib = (off & 1) ? (sb & 15u) : (sb >> 4);
                        const uint32_t nt = (uint32_t)((PB_NT16_NT4_LUT >> (nib * 4)) & 0xf);
                        const uint32_t lv = __ldg(a.qtab + (min((z >> 16) & 0xffu, 63u) << 8) + qb);
                        code[q] = nt > 3u ? PB_CODE_NONE : (lv | nt);      // 0xff | nt stays PB_CODE_NONE
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (code[q] == PB_CODE_NONE) continue;
                const uint32_t inc = 1u << ((code[q] & 3u) << 3);
                my_hist[((code[q] >> 2) * 2 + ((zz[q] >> 24) & 1u)) * PB_HARD_THREADS] += inc;
                tot4 += inc;
                const int mq = (int)((zz[q] >> 16) & 0xffu);
                rmsq += mq * mq;
            }
        }
        if (tot4 == 0) continue;                               // cannot happen for a listed cell; nothing to fold anyway
        const int rc = (pos >= 0 && pos < a.ref_len) ? (int)(unsigned char)a.ref[pos] : 'N';
        uint64_t cb;
        if (pb_tot4_unanimous(tot4)) {
            auto peek = [&](int lw) -> uint32_t { return my_hist[lw * PB_HARD_THREADS]; };
            const int kk = pb_tot4_k(tot4);
            const int bb = (tot4 >> 8 & 255u) ? 1 : (tot4 >> 16 & 255u) ? 2 : (tot4 >> 24) ? 3 : 0;
            cb = pb_unanimous_by_count(peek, nl, need_s, kk, bb) ? pb_unanimous_result(a.lhet, kk, bb, rmsq)
                                                                  : pb_call_unanimous(peek, n_lw, qval_s, tot4, rmsq, fk_s, a.beta, a.lhet);
            for (int lw = 0; lw < n_lw; ++lw) my_hist[lw * PB_HARD_THREADS] = 0;
        } else {
            auto take = [&](int lw) -> uint32_t { const uint32_t w = my_hist[lw * PB_HARD_THREADS]; my_hist[lw * PB_HARD_THREADS] = 0; return w; };
            cb = pb_call_general(take, n_lw, qval_s, tot4, rmsq, fk_s, a.beta, a.lhet);
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
