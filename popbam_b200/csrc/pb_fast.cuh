// pb_fast.cuh -- the counting pileup: the default formulation of the pileup / call / site stage.
// Used when the raw-depth cap cannot bind (k_depth_bound), the caller did not ask for the
// per-(site,sample) words, and min_depth / min_snpQ are positive; k_pileup_call otherwise.
//
// Why.  k_pileup_call runs call_base's whole bookkeeping (histogram update, per-base totals, sum of mapq^2) for every
// base of every cell.  But ~99 % of the cells are "easy": every passing base equals the reference base (or all but
// one do) and a few COUNTS prove that call_base would call the cell homozygous reference (pb_need_entry,
// pb_one_stray_entry in pb_walk.cuh), so all the site needs from the cell is qfilter's coverage bit
// (pop_utils.cpp:102-120).  The counts per (position, sample) cell:
//     k     bases that pass call_base's filters (popbam.cpp:268-284)
//     khi   those with quality level >= 30 (baseQ' >= 30 and mapQ >= 30)
//     m     stray bases (passing, different from the reference base), and whether one of them is a high-quality one
//     lowq  a passing base of a read with mapQ < min_rmsQ is present (else rms >= min_rmsQ is guaranteed)
// and the rule, with per-depth tables built once per context (k_fast_tables):
//     easy = k > 0 and no lowq and ( m == 0 and (depth k alone suffices  or  khi >= hneed[k])
//                                 or m == 1 and (one stray base is harmless at depth k, or a low-quality one is) )
//
// k_pile_count is the north star's "CIGAR-expanding scatter of reads into a per-window [position x sample] integer
// count tensor, using shared-memory staging and atomics": a CTA owns one sample x a block of positions and keeps the
// counts as BYTE counters in shared memory, four positions per 32-bit word (the depth cap cannot bind, so no count
// exceeds 255 and no byte carries into its neighbour).  One THREAD takes one aligned segment of a read:
//   stage    its quality bytes and packed bases are brought into shared memory by asynchronous 16-byte copies
//            (LDGSTS: no registers; eight lanes copy the eight chunks of one record, so a record is one 128-byte request)
//   scatter  four bases per step with packed-byte arithmetic: PRMT aligns the quality bytes to the position grid,
//            "byte + (128 - T)" puts a threshold test into bit 7, a PRMT with the four base nibbles as its selector
//            is a 16-entry table lookup for four bases at once (valid A/C/G/T, base code), an XOR against the
//            reference's code bytes finds stray bases; the 0/1 bytes are added to the counters with one shared-memory
//            reduction (RED.ADD) per counter and four positions.
//   classify one thread per position reads the four counts and two table bytes: coverage bit (ballot -> cov32) or
//            "hard" (two stray bases, a variant, a low-mapQ read, an unproven depth: ~1 % of the cells).
//   emit     the hard cells' base codes, exactly as call_base forms them, go to a compact arena, one warp per cell
//            (the bytes were staged by this CTA moments ago: L2 hits); k_hard_cells calls them one cell per thread
//            with the exact machinery of pb_cell.cuh / pb_walk.cuh; k_fast_sites puts the two together.
// Round 1 built bit-planes in a separate streaming kernel and walked them per (record, 32 positions) with bit-sliced
// counters: 420 M warp instructions per 2.3 Mb shard; the byte counters need no per-strip walk at all.
// Kernels: k_set_levels / k_need_raw / k_fast_tables (per context), k_strip_index, k_pile_count, k_hard_cells,
// k_fast_sites.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_kernels.cuh"

#define PB_H_QUALITY 30            // quality value of the khi count

// ---- quality levels of this path.  Its proofs (pb_need_entry, pb_one_stray_entry) hold for any level SET that contains
// every level a passing base can have, so it does not look for the values present (round 1 paid two shared-memory
// stores per base for that): the set is the RANGE [qlo, qhi] with
//     qlo = clamp(min(min_baseQ, min_mapQ), 4, 63)   -- call_base keeps a base only if baseQ' >= min_baseQ and
//                                                       mapQ >= min_mapQ, and codes clamp(min(baseQ', mapQ), 4, 63)
//     qhi = the context's quality ceiling            -- an ASSUMPTION: k_pile_count checks every quality byte it
//                                                       touches against it (one add and one OR per four bytes) and
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

// Per-depth tables of the rule above, for the level range of k_set_levels (need: k_need_table for that range):
//   flags[k]  bit 0  need[0][k] <= k: k unanimous bases always take the shortcut (every base is at least level 0)
//             bit 1  one stray base of any level cannot change the homozygous call at depth k (pb_one_stray_entry)
//             bit 2  the same for a stray base below the khi level
//             bit 3  k >= min_depth (qfilter; k <= max_depth holds because the cap cannot bind)
//   hneed[k]  khi >= hneed[k] bases at or above the khi level prove the shortcut (0: never)
struct PbFastTables { uint8_t flags[256]; uint8_t hneed[256]; };
__global__ void __launch_bounds__(256) k_fast_tables(const PbCounters *__restrict__ ctr, const uint8_t *__restrict__ need,
                                                     const double *__restrict__ fk, const double *__restrict__ beta,
                                                     const double *__restrict__ lhet, int min_depth, PbFastTables *__restrict__ tab) {
    const int nl = ctr->n_levels, k = threadIdx.x;
    int hi = 0;
    while (hi < nl && (int)ctr->qval[hi] < PB_H_QUALITY) ++hi;
    uint32_t f = 0;
    const int nd = (k >= 1 && nl > 0) ? need[k] : 0;
    if (nd && nd <= k) f |= 1u;
    if (pb_one_stray_entry(nl, ctr->qval, k, fk, beta, lhet, 0, nl)) f |= 2u;
    if (hi > 0 && hi < nl && pb_one_stray_entry(nl, ctr->qval, k, fk, beta, lhet, 0, hi)) f |= 4u;
    if (k >= min_depth) f |= 8u;
    tab->flags[k] = (uint8_t)f;
    tab->hneed[k] = (k >= 1 && hi < nl) ? need[hi * 256 + k] : 0;
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

// Reference code bytes of a contig, in the code the PRMT lookup of k_pile_count yields for a read base (bits 1-4):
// A 0x02, C 0x04, G 0x08, T 0x1e; 0 for anything else (an upper-case A/C/G/T is the only thing a called base can equal,
// pop_utils.cpp:139 / SURVEY Q7) and for the padding behind the contig.
#define PB_REFCODE_PAD 4096
__global__ void __launch_bounds__(256) k_ref_codes(const char *__restrict__ ref, int64_t ref_len, uint8_t *__restrict__ code) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ref_len + PB_REFCODE_PAD; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = i < ref_len ? (int)(unsigned char)ref[i] : 'N';
        code[i] = (uint8_t)(c == 'A' ? 0x02 : c == 'C' ? 0x04 : c == 'G' ? 0x08 : c == 'T' ? 0x1e : 0x00);
    }
}

struct PbCountArgs {
    const int4 *srec;
    const uint32_t *F;                       // strip index (k_strip_index), [n_samples][NI]
    int NI;
    const uint8_t *qual, *seq4;              // the read batch's bases (16-byte aligned, padded by 64 zero bytes)
    const uint8_t *refcode;                  // reference code bytes of the contig (k_ref_codes)
    int span_beg, span_end;
    int n_samples, n_strips;
    int spc;                                 // strips of 32 positions per CTA
    int min_rmsQ, min_baseQ, illumina;
    int qual_ceiling;                        // assumed largest (adjusted) base quality, checked on every byte touched
    int qslot, sslot;                        // bytes of a record's staging slots (16 bytes of front padding included)
    int lq;                                  // log2 of the lanes that stage one record (>= the 16-byte chunks of either slot)
    PbCounters *ctr;
    const PbFastTables *tab;
    uint32_t *cov32;                         // [n_samples][n_strips] easy and covered
    uint4 *cells;                            // directory of the cells left for k_hard_cells: {pos, sample | k << 8, sum mapq^2, first code}
    uint16_t *codes;                         // their base codes  q << 5 | strand << 4 | base  (popbam.cpp:279-284)
    unsigned long long cell_cap, code_cap;
    uint32_t *carry;                         // [blocks][n_samples][4][halo / 4] counts a CTA's reads add behind its block
    uint32_t *carry_flag;                    // [blocks][n_samples] set once they are published (zeroed per region)
    int halo;                                // positions of that tail: max_span rounded up to 32 (<= 32 * spc)
};

#define PB_CNT_THREADS 256         // threads per CTA = records staged per pass
#define PB_CNT_SPC_MAX 64          // strips per CTA at most (cell ids are 16 bits, the emit scratch is sized for it)
#define PB_CNT_STRIDE (PB_CNT_SPC_MAX * 32)      // bytes between the four counter arrays (fixed: immediate offsets in the scatter)
// slots: 16 bytes of front padding (the first position word of a record can start up to three bytes before its first
// base), then the 16-byte chunks that hold the segment; an odd number of chunks, so that consecutive slots start in
// different bank groups
__host__ __device__ static inline int pb_cnt_qslot(int max_span) { int c = 1 + (15 + max_span + 15) / 16; return 16 * (c | 1); }
static inline int pb_cnt_sslot(int max_span) { int c = 1 + (15 + (max_span + 1) / 2 + 1 + 15) / 16; return 16 * (c | 1); }
static inline int pb_cnt_halo(int max_span) { return (max_span + 31) & ~31; }
static inline size_t pb_cnt_smem(int spc, int max_span) {
    const size_t pb = (size_t)spc * 32, halo = (size_t)pb_cnt_halo(max_span);
    const size_t slots = (size_t)PB_CNT_THREADS * (pb_cnt_qslot(max_span) + pb_cnt_sslot(max_span));
    const size_t emit = (size_t)pb * 6;                       // cell ids + code offsets, in the slots' place
    return 4 * PB_CNT_STRIDE + pb + halo + 16 + 512 + 64 * 4 + (slots > emit ? slots : emit) + 16;
}

// ---- 16-byte asynchronous copy global -> shared (LDGSTS, L2 only).  A 1-D bulk copy per record (cp.async.bulk, UBLKCP)
// was tried first: its operands live in uniform registers, so the compiler serialises the 32 lanes of a warp in an
// ELECT / R2UR loop of ten instructions per copy -- as many issue slots as the scatter itself.  Eight lanes copying the
// eight chunks of one record with one LDGSTS each cost a fifth of that and the 128 bytes of a record are one request.
__device__ __forceinline__ uint32_t pb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_cp16(uint32_t dst_shared, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void pb_cp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// prmt.b32 in its default mode: selector nibble n picks byte n & 7 of {a, b}; with bit 3 of the nibble set the SIGN of that
// byte is replicated over the result byte (__byte_perm masks bit 3 away, hence the inline PTX)
__device__ __forceinline__ uint32_t pb_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// swap the two nibbles of every byte: seq4 keeps the first base of a byte in the HIGH nibble (bam.h:245-258); after the
// swap base j of a little-endian word sits in bits 4j .. 4j+3, and any nibble shift keeps the order
__device__ __forceinline__ uint32_t pb_nibble_order(uint32_t w) { return ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu); }

// One CTA = one sample x `spc` strips of 32 positions.  ROBUST: quality bytes >= 128 were seen in this context (a BAM
// without qualities stores 0xff), so the packed threshold tests use the form that is right for any byte value.
template <bool ROBUST>
__global__ void __launch_bounds__(PB_CNT_THREADS) k_pile_count(const PbCountArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_base[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int max_span = a.ctr->max_span;
    if (!a.ctr->nocap || pb_cnt_qslot(max_span) > a.qslot || ((max_span + 31) & ~31) > a.halo) {      // launched on an assumption that does not hold: say so, do nothing
        if (tid == 0) a.ctr->spec_fail = 1;
        return;
    }
    const int PB = a.spc * 32;
    uint32_t *cK = reinterpret_cast<uint32_t *>(smem_raw);               // [PB / 4] passing bases, one byte per position
    uint32_t *cH = cK + PB_CNT_STRIDE / 4;                                // those at or above the khi level
    uint32_t *cM = cH + PB_CNT_STRIDE / 4;                                // stray bases
    uint32_t *cF = cM + PB_CNT_STRIDE / 4;                                // bit 0: a stray base at or above the khi level, bit 1: lowq
    uint8_t *rcode = reinterpret_cast<uint8_t *>(cF + PB_CNT_STRIDE / 4); // [PB + 16] reference code bytes (0: not A/C/G/T)
    uint8_t *tabS = rcode + PB + a.halo + 16;                                      // flags[256], hneed[256]
    uint32_t *hardS = reinterpret_cast<uint32_t *>(tabS + 512);           // [64] hard masks per strip
    unsigned char *slots = reinterpret_cast<unsigned char *>(hardS + 64); // [THREADS] quality slots, then [THREADS] sequence slots
    const int M = (max_span + 31) >> 5;
    const int sb = (int)(blockIdx.x / a.n_samples), s = (int)(blockIdx.x % a.n_samples);     // the samples of a block run together: one pass through L2
    const int t0 = sb * a.spc;
    const int p0 = a.span_beg + t0 * 32, p1 = min(p0 + PB, a.span_end);
    const uint32_t *Fs = a.F + (size_t)s * a.NI;
    // Every record belongs to ONE CTA, the one whose block holds its read's start (the first block also takes the reads
    // that start before the span), and is scattered in full: the counters reach `halo` positions behind the block, and
    // what lands there is handed to the next block's CTA through global memory (below).
    const uint32_t clo = __ldg(Fs + (sb == 0 ? 0 : t0 + M)), chi = __ldg(Fs + min(t0 + a.spc, a.n_strips) + M);
    const int pend = p0 + PB + a.halo;                                    // end of the counters
    for (int i = tid; i < (PB + a.halo) / 4 + 1; i += PB_CNT_THREADS) { cK[i] = 0; cH[i] = 0; cM[i] = 0; cF[i] = 0; }
    for (int i = tid; i < PB + a.halo + 16; i += PB_CNT_THREADS) rcode[i] = __ldg(a.refcode + (int64_t)p0 + i);       // k_ref_codes; padded behind the contig
    for (int i = tid; i < 128; i += PB_CNT_THREADS) reinterpret_cast<uint32_t *>(tabS)[i] = __ldg(reinterpret_cast<const uint32_t *>(a.tab) + i);
    *reinterpret_cast<uint4 *>(slots + (size_t)tid * a.qslot) = make_uint4(0, 0, 0, 0);        // front padding of the quality slot: read (never counted) by a segment's first word
    __syncthreads();
    // raw quality byte thresholds (host: all <= 128): passing, khi level, above the assumed ceiling
    const int qoff = a.illumina ? 31 : 0;
    const int tp = a.min_baseQ <= 0 ? 0 : a.min_baseQ + qoff;
    const int th = min(128, PB_H_QUALITY + qoff);
    // ceiling 63: levels are clamped there (popbam.cpp:281), nothing can exceed it; only bytes >= 128 matter then
    const int tc = a.qual_ceiling >= 63 ? 128 : min(128, a.qual_ceiling + 1 + qoff);
    const bool check = !ROBUST || a.qual_ceiling < 63;
    const uint32_t addP = (uint32_t)(128 - tp) * 0x01010101u, addC = (uint32_t)(128 - tc) * 0x01010101u;
    unsigned char *qslot = slots + (size_t)tid * a.qslot;
    unsigned char *sslot = slots + (size_t)PB_CNT_THREADS * a.qslot + (size_t)tid * a.sslot;
    uint32_t over = 0;
    for (uint32_t c0 = clo; c0 < chi; c0 += PB_CNT_THREADS) {
        // ---- stage this thread's record
        int x = 0, len = 0, pa = 0, pb = 0;
        uint32_t z = 0;
        uint64_t o = 0;
        if (c0 + tid < chi) {
            const int4 rc = __ldg(&a.srec[c0 + tid]);
            z = (uint32_t)rc.z; x = rc.y; len = (int)(z & 0xffffu);
            o = ((uint64_t)(z >> 26) << 32) | (uint32_t)rc.w;                                // byte offset of the segment's first base
            pa = max(x, p0); pb = min(x + len, pend);
        }
        const bool active = pb > pa;
        if (c0 + (uint32_t)(tid & ~31) >= chi) break;                             // none of this warp's threads has a record (later passes neither)
        {
            // 2^lq lanes stage one record: lane c of the group copies chunk c of its quality bytes and of its packed bases.
            // Every warp stages the 32 records of its own threads, so a warp-level wait is all the scatter needs.
            const uint32_t qc = (uint32_t)(o >> 4);                                        // first 16-byte chunk of qual[] (seq4[]: qc >> 1)
            const uint32_t nq = (uint32_t)((o + (uint64_t)len + 15) >> 4) - qc, ns = (uint32_t)((((o + (uint64_t)len + 1) >> 1) + 15) >> 4) - (qc >> 1);
            const uint32_t d_cnt = active ? (nq | ns << 8) : 0u;
            const int per = 32 >> a.lq, cl = lane & ((1 << a.lq) - 1);
            uint32_t qd = pb_smem_addr(slots) + (uint32_t)(((tid & ~31) + (lane >> a.lq)) * a.qslot + 16 + 16 * cl);
            uint32_t sd = pb_smem_addr(slots) + (uint32_t)(PB_CNT_THREADS * a.qslot + ((tid & ~31) + (lane >> a.lq)) * a.sslot + 16 + 16 * cl);
            const uint8_t *qsrc = a.qual + 16 * (size_t)cl, *ssrc = a.seq4 + 16 * (size_t)cl;
            const uint32_t qstep = (uint32_t)(per * a.qslot), sstep = (uint32_t)(per * a.sslot);
            for (int rl = lane >> a.lq; rl < 32; rl += per) {                               // record (lane of its owner) staged by this lane now
                const uint32_t r_qc = __shfl_sync(0xffffffffu, qc, rl), r_cnt = __shfl_sync(0xffffffffu, d_cnt, rl);
                if ((uint32_t)cl < (r_cnt & 0xffu)) pb_cp16(qd, qsrc + ((size_t)r_qc << 4));
                if ((uint32_t)cl < (r_cnt >> 8)) pb_cp16(sd, ssrc + ((size_t)(r_qc >> 1) << 4));
                qd += qstep; sd += sstep;
            }
            pb_cp_wait_all();
            __syncwarp();
        }
        // ---- scatter: position words j0 .. j1 of the block (four positions each), two per step
        if (active) {
            const int mq = (int)((z >> 16) & 0xffu);
            const uint32_t addH = mq >= PB_H_QUALITY ? (uint32_t)(128 - th) * 0x01010101u : 0u;      // mapQ below the khi level: no byte reaches bit 7
            const uint32_t hmask = mq >= PB_H_QUALITY ? 0xffffffffu : 0u;
            const bool lowq = mq < a.min_rmsQ;
            const int j0 = (pa - p0) >> 2, j1 = (pb - 1 - p0) >> 2;
            const int i0 = p0 + 4 * j0 - x;                                                  // base index of word j0's first byte (>= -3)
            const int bq = 16 + (int)(o & 15u) + i0;                                         // its byte in the quality slot (>= 13)
            const uint32_t *qw = reinterpret_cast<const uint32_t *>(qslot) + (bq >> 2);
            const uint32_t selq = 0x3210u + 0x1111u * (uint32_t)(bq & 3);
            const int nb = 32 + (int)(o & 31u) + i0;                                         // its nibble in the sequence slot (>= 29)
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(sslot) + (nb >> 3);
            const int sh = 4 * (nb & 7);
            // Segment ends.  The first and the last position word of a segment also hold bytes of its neighbours in qual[] /
            // seq4[] (or padding).  The slot is this thread's private copy, so those bases are simply made invalid there:
            // a zero nibble is no base (popbam.cpp:276-278), the lookup below gives it "not valid", and it passes no test.
            {
                const int lim = pb - x;                                                      // end of the segment's part inside the counters
                uint8_t *sb8 = sslot;
                for (int i = i0; i < 0; ++i) { const int n = 32 + (int)(o & 31u) + i; sb8[n >> 1] &= (n & 1) ? 0xf0 : 0x0f; }
                const int iend = i0 + 8 * ((j1 - j0 + 2) >> 1);                              // one past the last base the pairs below look at
                for (int i = lim; i < iend; ++i) { const int n = 32 + (int)(o & 31u) + i; sb8[n >> 1] &= (n & 1) ? 0xf0 : 0x0f; }
            }
            // One position word: quality bytes qv, base nibbles in the low 16 bits of sx; cp = the word's four counters
            // (arrays PB_CNT_STRIDE bytes apart: immediate offsets), rv = the reference's code bytes.
#define PB_CNT_WORD(cp, rv, qv_, sx_)                                                                                      \
            {                                                                                                              \
                /* 16-entry lookup for four bases: nibble 1 (A) -> 0x03, 2 (C) -> 0x05, 4 (G) -> 0x09, 8 (T) -> 0xff       \
                   (selector bit 3 replicates the sign of entry 0, 0x80), everything else -> bit 0 clear */               \
                const uint32_t sq = pb_prmt(0x00050380u, 0x00000009u, (sx_));                                              \
                const uint32_t qm = (qv_);                                                                                 \
                uint32_t fa, fh;                                                                                           \
                if (ROBUST) {                                                                                              \
                    const uint32_t lo7 = qm & 0x7f7f7f7fu;                                                                 \
                    fa = (lo7 + addP) | qm; fh = ((lo7 + addH) | qm) & hmask;                                              \
                    if (check) over |= (lo7 + addC) | qm;                                                                  \
                } else {                                                                                                   \
                    fa = qm + addP; fh = qm + addH;                                                                        \
                    over |= qm | (qm + addC);                                                                              \
                }                                                                                                          \
                const uint32_t P = (fa >> 7) & sq & 0x01010101u;                                                           \
                const uint32_t H = (fh >> 7) & P;                                                                          \
                const uint32_t y = (sq ^ (rv)) & 0x1e1e1e1eu;                                                              \
                const uint32_t mmw = ((y + 0x7f7f7f7fu) >> 7) & P;                                                         \
                atomicAdd((cp), P);                                                                                        \
                atomicAdd((cp) + PB_CNT_STRIDE / 4, H);                                                                    \
                if (mmw) {                                                                                                 \
                    atomicAdd((cp) + 2 * (PB_CNT_STRIDE / 4), mmw);                                                        \
                    if (mmw & H) atomicOr((cp) + 3 * (PB_CNT_STRIDE / 4), mmw & H);                                        \
                }                                                                                                          \
            }
            // pairs of position words share one 32-bit window of the nibble stream
            const int NP = (j1 - j0 + 2) >> 1;
            uint32_t *cp = cK + j0;
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(rcode) + j0;
            uint32_t wq0 = qw[0], sn0 = pb_nibble_order(sw[0]);
            for (int p = 0; p < NP; ++p) {
                const uint32_t wq1 = qw[1], wq2 = qw[2];
                const uint32_t sn1 = pb_nibble_order(sw[1]);
                const uint32_t sx = __funnelshift_r(sn0, sn1, sh);
                const uint32_t qa = __byte_perm(wq0, wq1, selq), qb = __byte_perm(wq1, wq2, selq);
                sn0 = sn1; wq0 = wq2;
                PB_CNT_WORD(cp, rp[0], qa, sx)
                PB_CNT_WORD(cp + 1, rp[1], qb, sx >> 16)
                qw += 2; sw += 1; cp += 2; rp += 2;
            }
#undef PB_CNT_WORD
            if (lowq) {
                // a read below min_rmsQ: every cell it covers leaves the easy path (its bases may or may not pass; k_hard_cells
                // computes the exact rms).  Rare, and outside the loop above.
                for (int j = j0; j <= j1; ++j) atomicOr(cF + j, 0x02020202u);
            }
            if (over & 0x80808080u) {
                // some byte this thread touched (its own or a neighbour's) is >= 128 or above the assumed ceiling: look at
                // the segment's own bytes one by one and tell the host, which raises the ceiling and runs the region again
                int mx = 0;
                const unsigned char *qb0 = qslot + 16 + (int)(o & 15u);
                for (int i = 0; i < len; ++i) mx = max(mx, (int)qb0[i]);
                const int adj = min(63, a.illumina ? (mx > 31 ? mx - 31 : 0) : mx);
                if (mx >= 128 && !ROBUST) { a.ctr->qual_high = 1; a.ctr->qual_over = 1; }
                if (adj > a.qual_ceiling) { atomicMax(&a.ctr->qual_max_seen, adj); a.ctr->qual_over = 1; }
                over = 0;
            }
        }
        __syncwarp();                                                         // the warp's slots are free again
    }
    __syncthreads();                                                          // this CTA's reads are counted
    {
        // publish what they added behind the block, take what the previous block's reads added to the front of this one.
        // The previous CTA of this sample has a lower block index, so it was scheduled no later than this one and does not
        // wait for anything itself: the wait below ends (decoupled look-back, as in a single-pass scan).
        const int hw = a.halo / 4;
        uint32_t *mine = a.carry + ((size_t)blockIdx.x * 4) * hw;
        for (int i = tid; i < 4 * hw; i += PB_CNT_THREADS) mine[i] = cK[(i / hw) * (PB_CNT_STRIDE / 4) + PB / 4 + (i % hw)];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicExch(a.carry_flag + blockIdx.x, 1u);
            if (sb > 0) while (atomicAdd(a.carry_flag + (blockIdx.x - a.n_samples), 0u) == 0u) __nanosleep(20);
            __threadfence();
        }
        __syncthreads();
        if (sb > 0) {
            const uint32_t *prev = a.carry + ((size_t)(blockIdx.x - a.n_samples) * 4) * hw;
            for (int i = tid; i < 4 * hw; i += PB_CNT_THREADS) {
                const uint32_t v = __ldcg(prev + i);
                uint32_t *d = cK + (i / hw) * (PB_CNT_STRIDE / 4) + (i % hw);
                if (i / hw == 3) *d |= v; else *d += v;                       // counts add byte-wise (no cell exceeds 255); flags or
            }
            __syncthreads();
        }
    }
    // ---- classify: one thread per position, 32 consecutive positions per warp
    const uint8_t *bK = reinterpret_cast<const uint8_t *>(cK), *bH = reinterpret_cast<const uint8_t *>(cH);
    const uint8_t *bM = reinterpret_cast<const uint8_t *>(cM), *bF = reinterpret_cast<const uint8_t *>(cF);
    for (int q = tid; q < PB; q += PB_CNT_THREADS) {
        const int k = bK[q], kh = bH[q], m = bM[q], f = bF[q];
        const uint32_t fl = tabS[k], hn = tabS[256 + k];
        const bool unan = (fl & 1u) || (hn && (uint32_t)kh >= hn);                       // the depth alone / the count of high-quality bases proves the shortcut
        const bool stray = (fl & 2u) || (!(f & 1) && (fl & 4u));                         // one stray base that provably cannot change the call
        const bool inside = p0 + q < p1;
        const bool easy = k > 0 && !(f & 2) && ((m == 0 && unan) || (m == 1 && stray));
        // qfilter for easy cells: rms >= min_rmsQ holds because every contributing read has mapQ >= min_rmsQ;
        // depth <= max_depth holds because the cap cannot bind; depth >= min_depth is bit 3 of the table
        const uint32_t covb = __ballot_sync(0xffffffffu, inside && easy && (fl & 8u));
        const uint32_t hardb = __ballot_sync(0xffffffffu, inside && k > 0 && !easy);
        if (lane == 0) {
            if (t0 + (q >> 5) < a.n_strips) a.cov32[(size_t)s * a.n_strips + t0 + (q >> 5)] = covb;
            hardS[q >> 5] = hardb;
        }
    }
    __syncthreads();
    // ---- the cells left over: directory entry + base codes, one warp per cell
    uint32_t hard = 0, n_cand = 0;
    if (tid < a.spc && t0 + tid < a.n_strips) {
        hard = hardS[tid];
        n_cand = __ldg(Fs + t0 + tid + M + 1) - __ldg(Fs + t0 + tid);                  // upper bound of a cell's depth: the strip's candidate records
    }
    const uint32_t my_cells = (uint32_t)__popc(hard);
    // offsets of the strips' cells and code slots: at most 64 strips, so the first two warps scan and the rest wait
    __shared__ uint32_t s_scan[4];
    uint32_t cell_at = 0, code_at = 0;
    if (tid < 64) {
        uint32_t xc = my_cells, xk = my_cells * n_cand;
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const uint32_t yc = __shfl_up_sync(0xffffffffu, xc, o2), yk = __shfl_up_sync(0xffffffffu, xk, o2);
            if (lane >= o2) { xc += yc; xk += yk; }
        }
        if (lane == 31) { s_scan[(tid >> 5) * 2] = xc; s_scan[(tid >> 5) * 2 + 1] = xk; }
        cell_at = xc - my_cells; code_at = xk - my_cells * n_cand;
    }
    __syncthreads();
    const uint32_t tot_cells = s_scan[0] + s_scan[2], tot_codes = s_scan[1] + s_scan[3];
    if (tid >= 32 && tid < 64) { cell_at += s_scan[0]; code_at += s_scan[1]; }
    if (tot_cells == 0) return;
    uint16_t *cellS = reinterpret_cast<uint16_t *>(slots);                              // [<= PB] strip << 5 | bit
    uint32_t *codeS = reinterpret_cast<uint32_t *>(slots + 2 * (size_t)PB);             // first code slot, relative to the CTA's reservation
    if (tid == 0) {
        unsigned long long cb = atomicAdd(&a.ctr->n_cells, (unsigned long long)tot_cells);
        const unsigned long long kb = atomicAdd(&a.ctr->n_codes, (unsigned long long)tot_codes);
        if (cb + tot_cells > a.cell_cap || kb + tot_codes > a.code_cap) { a.ctr->arena_overflow = 1; cb = ~0ULL; }
        s_base[0] = cb; s_base[1] = kb;
    }
    for (uint32_t hm = hard; hm; hm &= hm - 1) {
        cellS[cell_at] = (uint16_t)((uint32_t)tid << 5 | (uint32_t)(__ffs(hm) - 1));
        codeS[cell_at] = code_at;
        ++cell_at; code_at += n_cand;
    }
    __syncthreads();
    if (s_base[0] == ~0ULL) return;
    for (uint32_t c = (uint32_t)(tid >> 5); c < tot_cells; c += PB_CNT_THREADS / 32) {
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
