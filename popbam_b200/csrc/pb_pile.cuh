// pb_pile.cuh -- the counting pileup: reads in FILE order, all samples of a block of positions in one CTA.
//
// This is the north star's "CIGAR-expanding scatter of reads into a per-window [position x sample] integer count
// tensor, using shared-memory staging and atomics".  What the counts are for and why ~99 % of the cells need nothing
// else is explained at the top of pb_fast.cuh; this file holds the kernel that touches the reads:
//
//   k_pile_reads   one CTA = `spc` strips of 32 positions x ALL samples.  Its reads are the ones that START in the block
//                  (the reads are sorted, bam_pileup.c:384-395, so they are one contiguous run of the batch, found by two
//                  binary searches in pos[]): no partition by sample, no segment records, no strip index -- the per-read
//                  arrays of the C ABI are the kernel's input as they are.  A WARP takes 32 consecutive reads per step:
//                  their quality bytes and packed bases lie back to back in qual[] / seq4[], so the tile is brought into
//                  shared memory by two 1-D bulk copies (cp.async.bulk + mbarrier, issued by one lane) and every lane
//                  then walks its own read: CIGAR -> aligned segments (resolve_cigar2, bam_pileup.c:90-221), four bases
//                  per step with packed-byte arithmetic (below), counts added to the sample's byte counters with one
//                  shared-memory reduction per counter and four positions.  A read that starts in an earlier block and
//                  reaches into this one is walked by this CTA as well (clipped to the block): no CTA waits for another.
//                  Then one thread per position classifies the cells of all samples (easy / hard), writes the position's
//                  coverage mask, and the hard cells get a directory entry and room for their base codes, which the CTA
//                  collects itself while its reads are in L2 (pb_block_codes; k_cell_codes is the same function as a
//                  kernel of its own, for the blocks whose CTA lacked the shared memory for the read lists).
//
// Packed-byte arithmetic of one position word (four positions, one byte each):
//   PRMT aligns the quality bytes to the position grid; "byte + (128 - T)" puts a threshold test into bit 7; a PRMT with
//   the four base nibbles as its selector is a 16-entry table lookup for four bases at once (valid A/C/G/T, base code);
//   an XOR against the reference's code bytes finds stray bases.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pb_kernels.cuh"
#include "pb_fast.cuh"             // PB_H_QUALITY, PbFastTables

struct PbPileReadsArgs {
    // the read batch as pushed (file order)
    const int32_t *pos;
    const uint32_t *meta, *cigstart, *ncig, *cigar;
    const uint64_t *base;                    // byte offset of a read's first base in qual[] (seq4[]: half of it)
    int64_t n_reads;
    uint64_t n_bytes;
    const uint8_t *qual, *seq4;              // 16-byte aligned, at least 64 readable bytes behind the last base
    const uint32_t *refcode;                 // reference code nibbles of the contig (k_ref_codes), padded behind its end
    int span_beg, span_end;
    int n_samples, n_strips;
    int spc;                                 // strips of 32 positions per CTA
    int asw;                                 // words between the four counter arrays of a sample: 8 * spc + 1
    int tile_q;                              // bytes of a warp's quality tile (multiple of 32); its packed-base tile: tile_q / 2 + 16
    int tail_bytes;                          // shared memory from the tiles' start to the end (pb_pile_tail_bytes)
    int max_span;                            // > 0: the largest read span ASSUMED (the per-read chain runs beside this kernel; the host compares afterwards, and
                                             // likewise "the depth cap cannot bind"); 0: read both from the counters
    int qcap;                                // further aligned segments (reads with deletions ...) a CTA can queue for its last pass
    int min_mapQ, min_rmsQ, min_baseQ, illumina;
    int qual_ceiling;                        // largest (adjusted) quality of a stray base the one-stray-base rule of the tables covers
    PbCounters *ctr;
    const PbFastTables *tab;
    uint64_t *acc_cov;                       // [span] out: samples whose cell is settled here and covered (k_hard_cells adds its cells)
    uint32_t *acc_cnt4;                      // [span] out: derived-base counts of the cells settled here (k_hard_cells adds its cells')
    uint64_t *site_type;                     // [span] out: derived-allele bits of the cells settled here (k_hard_cells adds its cells')
    uint4 *blk;                              // out: per block {first directory entry, entries (bit 31: codes done), first read that can cover the block, end of the block's reads}
    uint4 *cells;                            // out: directory of the cells left for k_hard_cells {pos, sample | k << 8, -, first code}
    unsigned long long cell_cap, code_cap;
    uint16_t *codes;                         // the code arena: the block's CTA fills in its cells' codes itself (pb_block_codes); null: all left to k_cell_codes
    int codes_skip;                          // (tests) blocks with (index & codes_skip) != 0 leave their codes to k_cell_codes all the same
};

// in the tiles' place once the reads are counted: per-position masks, the classification queue, the hard-cell list (and its
// code offsets), some slack (the read lists of pb_block_codes take the counters' place, not this one)
__host__ __device__ static inline int pb_pile_hcap(int n_samples, int spc) { return n_samples * 32 * spc; }      // (every cell may end up in the list)
__host__ __device__ static inline size_t pb_pile_tail_smem(int n_samples, int spc, int warps) {
    return (size_t)20 * 32 * spc + ((((size_t)n_samples * 8 * spc) + 1) & ~(size_t)1) * 2 + (4 + (size_t)1 / 8) * (size_t)pb_pile_hcap(n_samples, spc) + (size_t)pb_pile_hcap(n_samples, spc) / 8 + 8 + (2 * (size_t)n_samples + 1) * 4 +
           256 * (size_t)warps + 64;
}
static inline int pb_pile_asw(int spc) { return 8 * spc + 1; }
// dynamic shared memory: counters, reference nibbles (two copies), tables, barriers, per-warp tiles (16 bytes of
// padding around each)
static inline size_t pb_pile_reads_smem(int n_samples, int spc, int tile_q, int warps, int qcap) {
    const size_t cnt = (size_t)n_samples * (5 * (size_t)pb_pile_asw(spc) + 1) * 4;
    const size_t rc = 2 * ((size_t)(32 * spc) / 8 + 2) * 4 + 32;
    const size_t tiles = (size_t)warps * ((size_t)tile_q + 32 + (size_t)tile_q / 2 + 16 + 32);
    const size_t scan = pb_pile_tail_smem(n_samples, spc, warps);
    return ((cnt + 15) & ~(size_t)15) + rc + ((sizeof(PbFastTables) + 15) & ~(size_t)15) + 16 * (size_t)warps + 16 * (size_t)qcap + (tiles > scan ? tiles : scan) + 64;
}

static inline size_t pb_pile_tail_bytes(int n_samples, int spc, int tile_q, int warps) {
    const size_t tiles = (size_t)warps * ((size_t)tile_q + 32 + (size_t)tile_q / 2 + 16 + 32), scan = pb_pile_tail_smem(n_samples, spc, warps);
    return tiles > scan ? tiles : scan;
}

__device__ __forceinline__ uint32_t pb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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

// ---- mbarrier + 1-D bulk copy (TMA without a tensor map): global -> shared, completion counted in bytes on the barrier
__device__ __forceinline__ void pb_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void pb_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PB_DONE_%=;\n"
        "bra PB_WAIT_%=;\n"
        "PB_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// call_base's filter and code for one base (popbam.cpp:268-284): code = q << 5 | strand << 4 | nt4, q = clamp(min(baseQ', mapQ), 4, 63)
__device__ __forceinline__ bool pb_base_code(const uint8_t *__restrict__ qual, const uint8_t *__restrict__ seq4, uint64_t off, int illumina, int min_baseQ,
                                             int mq, uint32_t strand, uint32_t *code) {
    int bq = (int)__ldg(qual + off);
    const uint32_t sbyte = __ldg(seq4 + (off >> 1));
    const uint32_t nib = (off & 1) ? (sbyte & 15u) : (sbyte >> 4);
    const uint32_t nt = (uint32_t)((PB_NT16_NT4_LUT >> (nib * 4)) & 0xf);
    if (illumina) bq = bq > 31 ? bq - 31 : 0;
    if (nt > 3u || bq < min_baseQ) return false;
    const int qq = max(4, min(63, min(bq, mq)));
    *code = (uint32_t)qq << 5 | strand << 4 | nt;
    return true;
}

// One position word (four positions, one byte each) of one read: quality bytes qv aligned to the positions, the four base
// nibbles in the low 16 bits of sx, xn = those nibbles XOR the reference's (non-zero: a stray base), vm = 0x01 in the
// bytes that belong to the segment.  Adds the passing bases (P) and the high-quality ones (H) to the counters at cp / cp +
// ASW and returns the stray passing bases; the caller handles those (rare).
template <bool ROBUST, bool MASKED, bool DQ>
__device__ __forceinline__ uint32_t pb_count_word(uint32_t *cp, int ASW, uint32_t qm, uint32_t sx, uint32_t xn, uint32_t vm, uint32_t addP, uint32_t addH,
                                                  uint32_t hmask, uint32_t dq, uint32_t *Hout) {
    // 16-entry lookup for four bases: nibble 1 (A), 2 (C), 4 (G), 8 (T) -> bit 0 set (T: selector bit 3 replicates the sign
    // of entry 0, 0x80, over the byte), everything else -> bit 0 clear
    const uint32_t sq = pb_prmt(0x00050380u, 0x00000009u, sx);
    uint32_t fa, fh;
    if (ROBUST) {
        const uint32_t lo7 = qm & 0x7f7f7f7fu;
        fa = (lo7 + addP) | qm; fh = ((lo7 + addH) | qm) & hmask;
    } else {
        fa = qm + addP; fh = qm + addH;
    }
    const uint32_t P = (fa >> 7) & sq & vm;
    const uint32_t H = (fh >> 7) & P;
    // nibble != 0 -> bit 0 (entries 1..7: 0x81; entry 0: 0x80, and its sign makes nibble 8 0xff)
    const uint32_t mm = pb_prmt(0x81818180u, 0x81818181u, xn) & P;
    if (!MASKED || P) {
        atomicAdd(cp, P);
        atomicAdd(cp + ASW, H);
        if (DQ && dq) atomicAdd(cp + 4 * ASW, P * dq);    // a read below the top mapping-quality class (PbFastTables::dq)
    }
    *Hout = H;
    return mm;
}
// The stray bases of a word (rare): count them and OR into the cells' flag bytes
//   bits 0-3  the stray letters seen (one-hot in seq4's code, bam.h:245-258): a cell whose bases are all the same stray
//             letter is a homozygous variant, settled by the classification
//   bit 4     a stray base at or above the khi level
//   bit 5     off the easy path: a stray base with a quality above the ceiling of the one-stray-base rule (or, set by the
//             caller, a read below min_rmsQ)
template <bool ROBUST>
__device__ __forceinline__ void pb_count_stray(uint32_t *cp, int ASW, uint32_t mm, uint32_t H, uint32_t qm, uint32_t sx, uint32_t addC) {
    atomicAdd(cp + 2 * ASW, mm);
    const uint32_t oc = ((ROBUST ? (((qm & 0x7f7f7f7fu) + addC) | qm) : (qm + addC)) >> 7) & mm;
    // the four nibbles of sx, one per byte
    const uint32_t letters = (sx & 0xfu) | (sx & 0xf0u) << 4 | (sx & 0xf00u) << 8 | (sx & 0xf000u) << 12;
    atomicOr(cp + 3 * ASW, (letters & mm * 15u) | (mm & H) << 4 | oc << 5);
}

// What the scatter of one aligned segment needs besides the segment itself.
struct PbScatter {
    uint32_t *cnt;                 // counters [n][K, H, M, F, D][ASW]
    const uint32_t *refA, *refB;   // reference nibbles of the block, and the same stream one position word further on
    int ASW, RW;
    int p0, pend;                  // the counters' positions
    uint32_t addP, addC, addH1;    // packed thresholds: passing, ceiling of the one-stray-base rule, khi level
    const uint8_t *dqtab;          // mapping quality -> deficit (PbFastTables::dq)
    PbCounters *ctr;
};
template <bool GLOBAL> __device__ __forceinline__ uint32_t pb_ld32(const uint32_t *p) {
    if (GLOBAL) return __ldg(p);
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(pb_smem_addr(p)));
    return v;
}
// One aligned segment of one read, by ONE thread: reference [sx0, sx0 + slen), its first base at byte qbase[qi] / nibble
// 2 * sbase + qi... given as (word pointers + offsets) by the caller: `qb` / `sb` point to 4-byte aligned memory that holds
// base i of the segment at byte qoff + i resp. nibble noff + i (both offsets may be negative by up to 3 for the masked bytes
// in front of the segment; GLOBAL: the caller guarantees that memory is readable from 4 bytes before base 0, or qoff/noff >= 3).
// DQ: some read of the warp is below the top mapping-quality class (the deficit counter is then updated, one more reduction
// per word for those reads; the common case -- no such read among the warp's 32 -- runs without it).
template <bool ROBUST, bool GLOBAL, bool DQ>
__device__ __forceinline__ void pb_scatter_segment(const PbScatter &c, int sx0, int slen, uint32_t smp, int mq, const unsigned char *qb, long long qoff,
                                                   const unsigned char *sb, long long noff) {
    const int pa = max(sx0, c.p0), pb = min(sx0 + slen, c.pend);
    if (pb <= pa) return;
    const int ASW = c.ASW;
    uint32_t *row = c.cnt + (size_t)smp * c.RW;
    uint32_t addH = mq >= PB_H_QUALITY ? c.addH1 : 0u;                    // mapQ below the khi level: no byte reaches bit 7
    uint32_t hmask = mq >= PB_H_QUALITY ? 0xffffffffu : 0u;
    const uint32_t addP = c.addP, addC = c.addC;
    uint32_t dq = DQ ? c.dqtab[min(mq, 255)] : 0u;
    asm volatile("" : "+r"(addH), "+r"(hmask), "+r"(dq));                 // keep them in registers (the compiler would recompute them per word)
    const int j0 = (pa - c.p0) >> 2, j1 = (pb - 1 - c.p0) >> 2;           // position words of the block (four positions each)
    const int i0b = c.p0 + 4 * j0 - sx0;                                  // base index of word j0's first byte (>= -3)
    const long long bq = qoff + i0b;                                      // its byte (>= -3 relative to base 0)
    const uint32_t *qw = reinterpret_cast<const uint32_t *>(qb + (bq & ~3LL));
    const uint32_t selq = 0x3210u + 0x1111u * (uint32_t)(bq & 3);
    const long long nb = noff + i0b;                                      // its nibble
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sb + ((nb >> 3) << 2));
    const int sh = 4 * (int)(nb & 7);
    // bytes of the first / last word that belong to the segment (and to the counters)
    const uint32_t mfirst = 0x01010101u << (8 * ((pa - c.p0) & 3));
    const uint32_t mlast = 0x01010101u >> (8 * (4 - (pb - c.p0 - 4 * j1)));
    uint32_t *cp = row + j0;
    const uint32_t *rp = ((j0 & 1) ? c.refB : c.refA) + (j0 >> 1);        // reference nibbles of a pair of position words
    uint32_t over = 0;
    // pairs of position words share one 32-bit window of the nibble stream
#define PB_CNT_PAIR(vmA, vmB, MASKED)                                                                                        \
    {                                                                                                                        \
        const uint32_t wq1 = pb_ld32<GLOBAL>(qw + 1), wq2 = pb_ld32<GLOBAL>(qw + 2);                                         \
        const uint32_t sn1 = pb_nibble_order(pb_ld32<GLOBAL>(sw + 1));                                                       \
        const uint32_t sxw = __funnelshift_r(sn0, sn1, sh);                                                                  \
        const uint32_t qa = __byte_perm(wq0, wq1, selq), qb_ = __byte_perm(wq1, wq2, selq);                                  \
        const uint32_t xn = sxw ^ rp[0];                                                                                     \
        sn0 = sn1; wq0 = wq2;                                                                                                \
        if (!ROBUST) over |= MASKED ? ((qa & (vmA) << 7) | (qb_ & (vmB) << 7)) : (qa | qb_);                                 \
        uint32_t hA, hB;                                                                                                     \
        const uint32_t mmA = pb_count_word<ROBUST, MASKED, DQ>(cp, ASW, qa, sxw, xn, (vmA), addP, addH, hmask, dq, &hA);             \
        const uint32_t mmB = pb_count_word<ROBUST, MASKED, DQ>(cp + 1, ASW, qb_, sxw >> 16, xn >> 16, (vmB), addP, addH, hmask, dq, &hB); \
        if (mmA | mmB) {                                   /* stray bases: rare, one branch per pair */                     \
            if (mmA) pb_count_stray<ROBUST>(cp, ASW, mmA, hA, qa, sxw, addC);                                                \
            if (mmB) pb_count_stray<ROBUST>(cp + 1, ASW, mmB, hB, qb_, sxw >> 16, addC);                                     \
        }                                                                                                                    \
        qw += 2; sw += 1; cp += 2; rp += 1;                                                                                  \
    }
    const int NP = (j1 - j0 + 2) >> 1;                                    // pairs; the last one may hold one word only
    const bool odd = ((j1 - j0) & 1) == 0;
    // GLOBAL: the word in front of base 0 may lie in front of the array
    uint32_t wq0 = (GLOBAL && bq < 0) ? 0u : pb_ld32<GLOBAL>(qw);
    uint32_t sn0 = (GLOBAL && nb < 0) ? 0u : pb_nibble_order(pb_ld32<GLOBAL>(sw));
    {
        // first pair (also the last one of a short segment)
        uint32_t vmA = mfirst, vmB = 0x01010101u;
        if (NP == 1) { if (odd) { vmA &= mlast; vmB = 0u; } else vmB = mlast; }
        PB_CNT_PAIR(vmA, vmB, true)
    }
    for (int p = 1; p < NP - 1; ++p) PB_CNT_PAIR(0x01010101u, 0x01010101u, false)
    if (NP > 1) {
        const uint32_t vmA = odd ? mlast : 0x01010101u, vmB = odd ? 0u : mlast;
        PB_CNT_PAIR(vmA, vmB, true)
    }
#undef PB_CNT_PAIR
    if (!ROBUST && (over & 0x80808080u)) { c.ctr->qual_high = 1; c.ctr->qual_over = 1; }     // a quality byte >= 128: the host runs the region again with the robust variant
}

// first index in [0, n) with pos[i] >= target (n if none), by a whole warp: 32 probes per step
__device__ __forceinline__ long long pb_warp_lower_bound(const int32_t *__restrict__ pos, long long n, int target, int lane) {
    long long lo = 0, hi = n;                                            // the answer is in [lo, hi]
    while (hi - lo > 32) {
        const long long stride = (hi - lo + 31) >> 5;
        const long long idx = min(hi - 1, lo + (long long)(lane + 1) * stride - 1);
        const uint32_t ge = __ballot_sync(0xffffffffu, __ldg(pos + idx) >= target);
        if (!ge) { lo = min(hi, lo + 32 * stride); if (lo >= hi) return hi; continue; }
        const int L = __ffs((int)ge) - 1;
        hi = min(hi - 1, lo + (long long)(L + 1) * stride - 1);         // that probe is >= target
        lo = lo + (long long)L * stride;
    }
    const long long idx = lo + lane;
    const uint32_t ge = __ballot_sync(0xffffffffu, idx < hi && __ldg(pos + idx) >= target);
    return ge ? lo + (__ffs((int)ge) - 1) : hi;
}


// ROBUST: quality bytes >= 128 were seen in this context (a BAM without qualities stores 0xff), so the packed threshold
// tests use the form that is right for any byte value.
// shared memory of pb_block_codes: list starts and cursors, a flag, every warp's gathered reads, the lists
__host__ __device__ static inline size_t pb_block_codes_smem(int n_samples, int warps, int lcap) { return (2 * (size_t)n_samples + 4) * 4 + (size_t)warps * 64 * 4 + 4 * (size_t)lcap; }

// The base codes of the cells a block of k_pile_reads leaves for k_hard_cells, exactly as call_base forms them
// (popbam.cpp:268-284).  The reads that can cover a position of the block -- a run of the sorted batch, fresh in L2 -- are
// listed per sample in shared memory with their start positions; a WARP takes a cell, scans its sample's list, gathers
// the reads that start in (position - max_span, position] and handles them 32 at a time (CIGAR walk to the position,
// base filter, code).  All threads of the CTA call it; false: more reads than the lists hold (nothing written).  FIXED: the
// lists have equal room (no counting pass over the reads first).
struct PbBlockCodes {
    const int32_t *pos;
    const uint32_t *meta, *cigstart, *ncig, *cigar;
    const uint64_t *base;
    const uint8_t *qual, *seq4;
    int n_samples, min_mapQ, min_baseQ, illumina, lcap, max_span;
    uint4 *cells;
    uint16_t *codes;
};
template <bool FIXED>
__device__ __forceinline__ bool pb_block_codes(const PbBlockCodes &a, unsigned long long first_cell, int nh, long long rback, long long rend, int p0, unsigned char *smem) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = (int)blockDim.x, NWARP = NT >> 5, n = a.n_samples;
    const int max_span = a.max_span;
    int *cntS = reinterpret_cast<int *>(smem);                                 // [n + 1] reads per sample; then: their lists' starts
    int *curS = cntS + n + 1;                                                  // [n + 1] fill cursors of the lists
    int *badS = curS + n + 1;                                                  // [2]
    uint32_t *wlist = reinterpret_cast<uint32_t *>(badS + 2);                  // [warps][64] a warp's gathered reads
    uint32_t *rlist = wlist + 64 * NWARP;                                      // [lcap] the lists: read (relative to rback) | (start - p0) << 16
    const int cap_s = a.lcap / max(n, 1);                                      // FIXED: every sample's list has this room
    if (FIXED) {
        for (int i = tid; i <= n; i += NT) { cntS[i] = i * cap_s; curS[i] = i * cap_s; }
        if (tid == 0) badS[0] = (rend - rback > 0xffff) ? 1 : 0;
    } else {
        for (int i = tid; i <= n; i += NT) cntS[i] = 0;
        if (tid == 0) badS[0] = (rend - rback > 0xffff) ? 1 : 0;
        __syncthreads();
        for (long long r = rback + tid; r < rend; r += NT) {
            const uint32_t meta = __ldg(a.meta + r);
            if (!((meta >> 16) & 0x704u) && (meta & 0xffu) < (uint32_t)n && (int)((meta >> 8) & 0xffu) >= a.min_mapQ) atomicAdd(&cntS[meta & 0xffu], 1);
        }
        __syncthreads();
        if (wid == 0) {
            // starts of the samples' lists (n <= 64: two per lane)
            const int c0 = lane < n ? cntS[lane] : 0, c1 = lane + 32 < n ? cntS[lane + 32] : 0;
            int x0 = c0, x1 = c1;
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const int y0 = __shfl_up_sync(0xffffffffu, x0, o2), y1 = __shfl_up_sync(0xffffffffu, x1, o2); if (lane >= o2) { x0 += y0; x1 += y1; } }
            const int t0 = __shfl_sync(0xffffffffu, x0, 31), t1 = __shfl_sync(0xffffffffu, x1, 31);
            __syncwarp();
            if (lane < n) { cntS[lane] = x0 - c0; curS[lane] = x0 - c0; }
            if (lane + 32 < n) { cntS[lane + 32] = t0 + x1 - c1; curS[lane + 32] = t0 + x1 - c1; }
            if (lane == 0) { cntS[n] = t0 + t1; if (t0 + t1 > a.lcap) badS[0] = 1; }
        }
    }
    __syncthreads();
    if (badS[0]) return false;
    for (long long r = rback + tid; r < rend; r += NT) {
        const uint32_t meta = __ldg(a.meta + r);
        if (!((meta >> 16) & 0x704u) && (meta & 0xffu) < (uint32_t)n && (int)((meta >> 8) & 0xffu) >= a.min_mapQ) {
            const int rel = max(-32768, __ldg(a.pos + r) - p0);                 // (a read that starts further back than that covers nothing here)
            const int at = atomicAdd(&curS[meta & 0xffu], 1);
            if (!FIXED || at < cntS[meta & 0xffu] + cap_s) rlist[at] = (uint32_t)(r - rback) | (uint32_t)(uint16_t)(int16_t)rel << 16;
            else badS[0] = 1;                                                  // (a sample with more reads than its share of the room)
        }
    }
    __syncthreads();
    if (FIXED && badS[0]) return false;
    uint32_t *gl = wlist + 64 * wid;
    for (int i = wid; i < nh; i += NWARP) {
        const unsigned long long c = first_cell + i;
        const uint4 cell = a.cells[c];
        const int pos = (int)cell.x, q = pos - p0, smp = (int)(cell.y & 0xffu);
        const uint32_t k = (cell.y >> 8) & 0xffu;
        uint16_t *out = a.codes + cell.w;
        uint32_t kk = 0, n_acc = 0;
        int rmsq = 0;
        // the first `cnt` gathered reads: code of the base at the cell's position, if the read has one there and it passes
        auto take = [&](uint32_t cnt) {
            bool ok = false;
            uint32_t code = 0;
            int mq = 0;
            if ((uint32_t)lane < cnt) {
                const long long r = rback + (long long)(gl[lane] & 0xffffu);
                const uint32_t meta = __ldg(a.meta + r);
                mq = (int)((meta >> 8) & 0xffu);
                int x = __ldg(a.pos + r);
                const uint32_t c0 = __ldg(a.cigstart + r), ncg = __ldg(a.ncig + r);
                uint64_t qo = __ldg(a.base + r);
                for (uint32_t ci = 0; ci < ncg && x <= pos; ++ci) {
                    const uint32_t cg = __ldg(a.cigar + c0 + ci);
                    const uint32_t op = cg & 15u;
                    const int len = (int)(cg >> 4);
                    if ((0x181u >> op) & 1u) {
                        if (pos < x + len) { ok = pb_base_code(a.qual, a.seq4, qo + (uint64_t)(pos - x), a.illumina, a.min_baseQ, mq, (meta >> 20) & 1u, &code); break; }
                        x += len; qo += (uint64_t)len;
                    } else if ((0x12u >> op) & 1u) qo += (uint64_t)len;
                    else if ((0x0cu >> op) & 1u) x += len;
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const uint32_t at = kk + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                if (at < k) out[at] = (uint16_t)code;
                rmsq += mq * mq;
            }
            kk += (uint32_t)__popc(bal);
        };
        const int jend = FIXED ? min(curS[smp], cntS[smp] + cap_s) : cntS[smp + 1];
        for (int jb = cntS[smp]; jb < jend; jb += 32) {
            const int j = jb + lane;
            uint32_t ent = 0;
            bool acc = false;
            if (j < jend) {
                ent = rlist[j];
                const int rel = (int)(int16_t)(uint16_t)(ent >> 16);
                acc = rel <= q && rel + max_span > q;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, acc);
            if (acc) gl[n_acc + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = ent;
            n_acc += (uint32_t)__popc(bal);
            __syncwarp();
            if (n_acc >= 32) {
                take(32);
                const uint32_t rest = n_acc - 32, moved = (uint32_t)lane < rest ? gl[32 + lane] : 0u;
                __syncwarp();
                if ((uint32_t)lane < rest) gl[lane] = moved;
                n_acc = rest;
                __syncwarp();
            }
        }
        if (n_acc) take(n_acc);
        rmsq = __reduce_add_sync(0xffffffffu, rmsq);
        if (lane == 0) { a.cells[c].y = (uint32_t)smp | min(kk, k) << 8; a.cells[c].z = (uint32_t)rmsq; }
        __syncwarp();
    }
    return true;
}

template <bool ROBUST>
__global__ void __launch_bounds__(512) k_pile_reads(const PbPileReadsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_range[2];
    __shared__ unsigned long long s_resv[2];
    __shared__ uint32_t s_wsum[2][16];
    __shared__ int s_next, s_qn, s_nh, s_nc;
    __shared__ long long s_rback;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = (int)blockDim.x, NWARP = NT >> 5;
    const int max_span = a.max_span > 0 ? a.max_span : a.ctr->max_span;
    if (a.max_span <= 0 && !a.ctr->nocap) {                           // launched on an assumption that does not hold: say so, do nothing
        if (tid == 0) a.ctr->spec_fail = 1;
        return;
    }
    const int n = a.n_samples;
    const int PB = a.spc * 32;
    const int ASW = a.asw, RW = 5 * ASW + 1;                                // words: array stride, sample stride (the samples' rows start in different banks)
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem_raw);               // [n][K, H, M, F, D][ASW]: passing bases, those at or above the khi level, stray bases, flags (pb_count_stray), mapping-quality deficits (PbFastTables)
    const size_t cnt_bytes = (((size_t)n * RW * 4) + 15) & ~(size_t)15;
    const int NRW = PB / 8 + 2;                                            // words of reference nibbles (eight positions each)
    uint32_t *refA = reinterpret_cast<uint32_t *>(smem_raw + cnt_bytes);  // [NRW] position 8 i of the block in bits 0-3 of word i
    uint32_t *refB = refA + NRW;                                          // [NRW] the same stream 16 bits (one position word) further on
    uint8_t *tabS = smem_raw + ((cnt_bytes + (size_t)2 * NRW * 4 + 15) & ~(size_t)15);     // PbFastTables
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tabS + ((sizeof(PbFastTables) + 15) & ~(size_t)15));     // one per warp
    int4 *queue = reinterpret_cast<int4 *>(reinterpret_cast<unsigned char *>(mbar) + 16 * (size_t)NWARP);       // [qcap] {sx0, offset lo, len | offset hi << 16 | sample << 24, mapq}
    unsigned char *tiles = reinterpret_cast<unsigned char *>(queue + a.qcap);
    const int tile_s = a.tile_q / 2 + 16;
    const size_t tile_bytes = (size_t)a.tile_q + 32 + (size_t)tile_s + 32;
    const int t0s = (int)blockIdx.x * a.spc;                                    // first strip of the block
    const int p0 = a.span_beg + t0s * 32, p1 = min(p0 + PB, a.span_end);
    // the block's reads: every read that can reach a position of the block, i.e. starts in (p0 - max_span, p0 + PB) -- one
    // contiguous run of the sorted batch.  Two warps look the ends up while the others clear the counters.
    if (wid == 0) {
        const long long rlo = pb_warp_lower_bound(a.pos, a.n_reads, p0 - max_span + 1, lane);
        if (lane == 0) { s_range[0] = rlo; s_rback = rlo; s_next = 0; s_qn = 0; s_nh = 0; s_nc = 0; }
    } else if (wid == 1) {
        const long long rhi = pb_warp_lower_bound(a.pos, a.n_reads, p0 + PB, lane);
        if (lane == 0) s_range[1] = rhi;
    }
    for (int i = tid; i < (int)(cnt_bytes / 16); i += NT) reinterpret_cast<uint4 *>(cnt)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < NRW; i += NT) {
        // k_ref_codes packs by absolute position; the block starts at p0 (padded behind the contig)
        const uint32_t *g = a.refcode + ((int64_t)p0 >> 3) + i;
        const uint32_t w0 = __ldg(g), w1 = __ldg(g + 1), w2 = __ldg(g + 2);
        const int sh = 4 * (p0 & 7);
        const uint32_t c0 = __funnelshift_r(w0, w1, sh), c1 = __funnelshift_r(w1, w2, sh);
        refA[i] = c0; refB[i] = c0 >> 16 | c1 << 16;
    }
    for (int i = tid; i < (int)(sizeof(PbFastTables) / 4); i += NT) reinterpret_cast<uint32_t *>(tabS)[i] = __ldg(reinterpret_cast<const uint32_t *>(a.tab) + i);
    if (lane == 0) pb_mbar_init(pb_smem_addr(mbar + 2 * wid), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    PbScatter sc;
    sc.cnt = cnt; sc.refA = refA; sc.refB = refB; sc.ASW = ASW; sc.RW = RW; sc.p0 = p0; sc.pend = p0 + PB; sc.dqtab = reinterpret_cast<const PbFastTables *>(tabS)->dq; sc.ctr = a.ctr;
    {
        // raw quality byte thresholds (host: all <= 128): passing, khi level, above the ceiling of the one-stray-base rule
        const int qoff = a.illumina ? 31 : 0;
        const int tp = a.min_baseQ <= 0 ? 0 : a.min_baseQ + qoff;
        const int th = min(128, PB_H_QUALITY + qoff);
        const int tc = min(128, min(63, a.qual_ceiling) + 1 + qoff);
        uint32_t addP = (uint32_t)(128 - tp) * 0x01010101u, addC = (uint32_t)(128 - tc) * 0x01010101u, addH1 = (uint32_t)(128 - th) * 0x01010101u;
        asm volatile("" : "+r"(addP), "+r"(addC), "+r"(addH1));            // computed once (the compiler would recompute them inside the scatter loop)
        sc.addP = addP; sc.addC = addC; sc.addH1 = addH1;
    }
    const int64_t rlo = s_range[0];
    const int n_blk = (int)(s_range[1] - rlo);                             // reads of the block
    unsigned char *tq = tiles + (size_t)wid * tile_bytes + 16;             // quality tile (16 bytes of padding in front: a segment's first word may start 3 bytes early)
    unsigned char *ts = tq + a.tile_q + 16 + 16;                           // packed-base tile
    const uint32_t bar = pb_smem_addr(mbar + 2 * wid);
    uint32_t phase = 0;
    // ---- scatter.  A warp takes the block's reads 32 at a time (a shared cursor: the warps stay busy whatever the reads look
    // like), brings their bytes -- back to back in qual[] / seq4[] -- into its tile, and every lane walks the first aligned
    // segment of its own read.  Further segments (a read with a deletion has two) are queued for the CTA's last pass.
    for (;;) {
        int ch = 0;
        if (lane == 0) ch = atomicAdd(&s_next, 32);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= n_blk) break;
        const int cnt_l = min(32, n_blk - ch);
        const int64_t r = rlo + ch + lane;
        uint64_t b0 = 0, b1 = 0;
        uint32_t meta_r = 0x7040000u, c0_r = 0, ncg_r = 0;                     // (no read: dropped by the flag filter)
        int x_r = 0;
        if (lane < cnt_l) {
            // the whole header at once (the loads do not wait for one another; a dropped read's are wasted)
            b0 = __ldg(a.base + r); b1 = r + 1 < a.n_reads ? __ldg(a.base + r + 1) : a.n_bytes;
            meta_r = __ldg(a.meta + r); x_r = __ldg(a.pos + r); c0_r = __ldg(a.cigstart + r); ncg_r = __ldg(a.ncig + r);
        }
        bool give_up = false;
        for (int start = 0; start < cnt_l;) {
            const uint64_t tq0 = __shfl_sync(0xffffffffu, b0, start) & ~(uint64_t)15;                // first byte of the quality tile
            const bool fits = lane >= start && lane < cnt_l && b0 >= tq0 && b1 >= b0 && b1 <= a.n_bytes && b1 - tq0 <= (uint64_t)a.tile_q;
            const uint32_t fm = __ballot_sync(0xffffffffu, fits) >> start;
            const int nt = fm == 0xffffffffu ? 32 : __ffs((int)~fm) - 1;                             // reads of this tile (the offsets ascend)
            if (nt == 0) {                                                                           // one read larger than a tile, or offsets out of order
                if (lane == 0) a.ctr->spec_fail = 1;
                give_up = true;
                break;
            }
            const uint64_t tend = __shfl_sync(0xffffffffu, b1, start + nt - 1);
            const uint64_t ts0 = (tq0 >> 1) & ~(uint64_t)15;                                          // first byte of the packed-base tile
            if (lane == 0) {
                const uint32_t nbq = (uint32_t)((tend - tq0 + 15) & ~(uint64_t)15), nbs = (uint32_t)((((tend + 1) >> 1) - ts0 + 15) & ~(uint64_t)15);
                pb_mbar_expect_tx(bar, nbq + nbs);
                if (nbq) pb_bulk_g2s(pb_smem_addr(tq), a.qual + tq0, nbq, bar);
                if (nbs) pb_bulk_g2s(pb_smem_addr(ts), a.seq4 + ts0, nbs, bar);
            }
            // the read's header and aligned segments while the tile is on its way.  Dropped: flag filter of bam_plp_push
            // (bam_pileup.c:371-374), no sample, below min_mapQ (the raw-depth cap cannot bind, so such a read reaches no cell:
            // popbam.cpp:242-266)
            int sx0 = 0, slen = 0, mq = 0;
            uint64_t so = 0;
            uint32_t smp = 0;
            if (lane >= start && lane < start + nt) {
                const uint32_t meta = meta_r;
                mq = (int)((meta >> 8) & 0xffu); smp = meta & 0xffu;
                if (!((meta >> 16) & 0x704u) && smp < (uint32_t)n && mq >= a.min_mapQ) {
                    int x = x_r;
                    const uint32_t c0 = c0_r, ncg = ncg_r;
                    uint64_t qo = b0;
                    for (uint32_t ci = 0; ci < ncg; ++ci) {
                        const uint32_t cg = __ldg(a.cigar + c0 + ci);
                        const uint32_t op = cg & 15u, len = cg >> 4;
                        if ((0x181u >> op) & 1u) {                                                   // M = X: reference and query
                            if (qo + len > b1 || len > 0xffffu) { a.ctr->spec_fail = 1; break; }     // CIGAR longer than the read's bases (or a segment the queue entry cannot hold)
                            if (len) {
                                if (!slen) { sx0 = x; so = qo; slen = (int)len; }
                                else if (x < p0 + PB && x + (int)len > p0) {
                                    const int qi = atomicAdd(&s_qn, 1);
                                    if (qi < a.qcap) queue[qi] = make_int4(x, (int)(uint32_t)qo, (int)(len | (uint32_t)(qo >> 32) << 16 | smp << 24), mq);
                                    else a.ctr->spec_fail = 1;                                       // (more such segments than the queue holds -- it is sized for a fifth of the block's reads: the host takes the other path)
                                }
                            }
                            x += (int)len; qo += len;
                        } else if ((0x12u >> op) & 1u) qo += len;                                    // I, S: query only
                        else if ((0x0cu >> op) & 1u) x += (int)len;                                  // D, N: reference only
                    }
                }
            }
            pb_mbar_wait(bar, phase);
            __syncwarp();                                                     // the lanes leave the wait loop one by one: walk the segments together
            phase ^= 1u;
            if (__any_sync(0xffffffffu, slen && sc.dqtab[min(mq, 255)])) {
                if (slen) pb_scatter_segment<ROBUST, false, true>(sc, sx0, slen, smp, mq, tq, (long long)(so - tq0), ts, (long long)(so - 2 * ts0));
            } else if (slen) pb_scatter_segment<ROBUST, false, false>(sc, sx0, slen, smp, mq, tq, (long long)(so - tq0), ts, (long long)(so - 2 * ts0));
            __syncwarp();                                                     // the warp's tile is free again
            start += nt;
        }
        if (give_up) break;
    }
    __syncthreads();
    // ---- the queued segments, one per thread, bases straight from global memory (a few per cent of the reads)
    {
        const int nq = min(s_qn, a.qcap);
        for (int i = tid; i < nq; i += NT) {
            const int4 e = queue[i];
            const uint32_t z = (uint32_t)e.z;
            const long long so = (long long)(((uint64_t)((z >> 16) & 0xffu) << 32) | (uint32_t)e.y);
            pb_scatter_segment<ROBUST, true, true>(sc, e.x, (int)(z & 0xffffu), z >> 24, e.w, a.qual, so, a.seq4, so);
        }
    }
    __syncthreads();                                                          // this CTA's reads are counted
    // ---- classify.  Pass 1, one thread per position word (four positions), all samples one after the other: four covered
    // homozygous-reference cells at once by two packed compares; the (word, sample) pairs that fail the test are queued.
    // Pass 2, one thread per queued pair: the cells one by one.  Pass 3: the positions' masks go to global memory.
    // In the tiles' place (they are free now):
    uint32_t *covS = reinterpret_cast<uint32_t *>(tiles);                      // [PB][2] samples whose cell is settled and covered
    uint32_t *derS = covS + 2 * PB;                                            // [PB][2] ... and holds a derived allele
    uint32_t *dcntS = derS + 2 * PB;                                           // [PB] derived-base counts (one byte per base)
    uint16_t *cq = reinterpret_cast<uint16_t *>(dcntS + PB);                   // [n * PB / 4] queued pairs: word | sample << 9
    uint32_t *hlist = reinterpret_cast<uint32_t *>(cq + (((size_t)n * (PB / 4) + 1) & ~(size_t)1));     // [hcap] cells left for k_hard_cells: k | position of the block << 8 | sample << 20
    const int hcap = pb_pile_hcap(n, a.spc);
    uint32_t *hoff = hlist + hcap;                                             // [hcap / 32 + 1] first code of every 32nd cell, relative to the CTA's reservation

    const PbFastTables *T = reinterpret_cast<const PbFastTables *>(tabS);
    {
        // bit 7 of (x + (128 - t)) says x >= t for bytes below 128: three depth runs, the count of high-quality bases in two of them
        uint32_t aLo[3], aHi[3], aH[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            aLo[r] = (uint32_t)(128 - T->rlo[r]) * 0x01010101u; aHi[r] = (uint32_t)(127 - T->rhi[r]) * 0x01010101u;
            aH[r] = (uint32_t)(128 - T->rh[r]) * 0x01010101u;
        }
        for (int w = tid; w < PB / 4; w += NT) {
            const bool whole = p0 + 4 * w + 3 < p1;                            // the word's positions all inside the span
            unsigned long long cov = 0;
            for (int s = 0; s < n; ++s) {
                const uint32_t *rw = cnt + (size_t)s * RW + w;
                const uint32_t K4 = rw[0], H4 = rw[ASW], M4 = rw[2 * ASW], F4 = rw[3 * ASW] | rw[4 * ASW];      // (flags or deficits: not for the packed test)
                const uint32_t ok = ((K4 + aLo[0]) & ~(K4 + aHi[0])) | ((K4 + aLo[1]) & ~(K4 + aHi[1]) & (H4 + aH[1])) | ((K4 + aLo[2]) & ~(K4 + aHi[2]) & (H4 + aH[2]));
                if (whole && !(M4 | F4) && ((ok & ~K4 & 0x80808080u) == 0x80808080u)) cov |= 1ULL << s;
                else if (K4) cq[atomicAdd(&s_nc, 1)] = (uint16_t)(w | s << 9);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                covS[2 * (4 * w + j)] = (uint32_t)cov; covS[2 * (4 * w + j) + 1] = (uint32_t)(cov >> 32);
                derS[2 * (4 * w + j)] = 0; derS[2 * (4 * w + j) + 1] = 0; dcntS[4 * w + j] = 0;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < s_nc; i += NT) {
        const int w = cq[i] & 511, s = cq[i] >> 9, pw = p0 + 4 * w;
        const uint32_t *rw = cnt + (size_t)s * RW + w;
        const uint32_t K4 = rw[0], H4 = rw[ASW], M4 = rw[2 * ASW], F4 = rw[3 * ASW], D4 = rw[4 * ASW];
        const uint32_t ref16 = (refA[w >> 1] >> (16 * (w & 1))) & 0xffffu;
        const uint32_t sbit = 1u << (s & 31);
        const int hs = s >> 5;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k = (K4 >> (8 * j)) & 0xffu, kh = (H4 >> (8 * j)) & 0xffu, m = (M4 >> (8 * j)) & 0xffu, f = (F4 >> (8 * j)) & 0xffu;
            if (k == 0 || pw + j >= p1) continue;                                        // an empty cell is not covered
            const uint32_t fl = T->flags[k], hn = T->hneed[k];
            const bool unan = (fl & 1u) || (hn && kh >= hn);                             // the depth alone / the count of high-quality bases proves the shortcut
            const bool stray = (fl & 2u) || (!(f & 0x10u) && (fl & 4u));                 // one stray base that provably cannot change the call
            // qfilter for settled cells: rms >= min_rmsQ is proven from the mapping-quality deficits (PbFastTables); depth <=
            // max_depth because the cap cannot bind; depth >= min_depth is bit 3 of the table
            const int q = 4 * w + j;
            const uint32_t d = (D4 >> (8 * j)) & 0xffu;
            const bool rms_ok = d == 0 ? (fl & 0x40u) != 0
                                       : (k <= PB_RMS_KMAX && (int)k * T->m0sq + T->qstep * (3 * (int)k - (int)d) >= T->rthr[k]);
            bool settled = false;
            if (!(f & 0x20u) && rms_ok) {
                if ((m == 0 && unan) || (m == 1 && stray)) {                             // homozygous reference
                    settled = true;
                    if (fl & 8u) atomicOr(&covS[2 * q + hs], sbit);
                } else if (m == k && unan && ((ref16 >> (4 * j)) & 0xfu)) {              // every base a stray one, the reference base a proper one
                    const uint32_t al = f & 0xfu;
                    if (__popc(al) == 1) {                                               // ... and all the same letter: homozygous variant
                        const int b = __ffs((int)al) - 1;
                        if ((T->altok[k] >> b) & 1u) {                                   // segbase keeps it (pop_utils.cpp:139-150): derived allele
                            settled = true;
                            atomicAdd(&dcntS[q], 1u << (8 * b));
                            if (fl & 8u) { atomicOr(&covS[2 * q + hs], sbit); atomicOr(&derS[2 * q + hs], sbit); }
                        }
                    }
                }
            }
            if (!settled) {
                const int hi = atomicAdd(&s_nh, 1);
                if (hi < hcap) hlist[hi] = k | (uint32_t)q << 8 | (uint32_t)s << 20;
            }
        }
    }
    __syncthreads();
    for (int q = tid; q < PB; q += NT)
        if (p0 + q < p1) {
            const int64_t o = (int64_t)(p0 + q) - a.span_beg;
            a.acc_cov[o] = (unsigned long long)covS[2 * q + 1] << 32 | covS[2 * q];
            a.acc_cnt4[o] = dcntS[q];
            a.site_type[o] = (unsigned long long)derS[2 * q + 1] << 32 | derS[2 * q];
        }
    // ---- the cells left over: a directory entry each {position, sample | k << 8, -, first code} and a record of the block's
    // run of the directory; k_cell_codes collects their base codes
    const int nh = s_nh;
    if (tid == 0) a.blk[blockIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    if (nh == 0) return;
    if (wid == 0) {
        uint32_t run = 0;
        for (int i0 = 0; i0 < nh; i0 += 32) {
            const int i = i0 + lane;
            uint32_t x = i < nh ? (hlist[i] & 0xffu) : 0u;
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o2); if (lane >= o2) x += y; }
            if (lane == 0) hoff[i0 >> 5] = run;
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) {
            unsigned long long cb = atomicAdd(&a.ctr->n_cells, (unsigned long long)nh);
            const unsigned long long kb = atomicAdd(&a.ctr->n_codes, (unsigned long long)run);
            if (cb + nh > a.cell_cap || kb + run > a.code_cap) { a.ctr->arena_overflow = 1; cb = ~0ULL; }
            else a.blk[blockIdx.x] = make_uint4((uint32_t)cb, (uint32_t)nh, (uint32_t)s_rback, (uint32_t)s_range[1]);
            s_resv[0] = cb; s_resv[1] = kb;
        }
    }
    __syncthreads();
    if (s_resv[0] == ~0ULL) return;                                             // no room: the host runs the region again with a larger arena
    for (int g = wid; 32 * g < nh; g += NWARP) {
        const int i = 32 * g + lane;
        const uint32_t e = i < nh ? hlist[i] : 0u, k = e & 0xffu;
        uint32_t x = k;
        for (int o2 = 1; o2 < 32; o2 <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o2); if (lane >= o2) x += y; }
        if (i < nh) a.cells[s_resv[0] + i] = make_uint4((uint32_t)(p0 + (int)((e >> 8) & 0xfffu)), (e >> 20) | k << 8, 0u, (uint32_t)(s_resv[1] + hoff[g] + x - k));
    }
    // ---- their base codes, while the block's reads are fresh in L2: the lists of pb_block_codes take the counters' place
    if (a.codes && !((int)blockIdx.x & a.codes_skip)) {
        const long long lroom = ((long long)cnt_bytes - (long long)pb_block_codes_smem(n, NWARP, 0)) / 4;
        if (lroom >= (long long)n_blk) {                                       // (else: k_cell_codes)
            __syncthreads();                                                   // the directory entries are written, the counters read
            PbBlockCodes b;
            b.pos = a.pos; b.meta = a.meta; b.cigstart = a.cigstart; b.ncig = a.ncig; b.cigar = a.cigar; b.base = a.base; b.qual = a.qual; b.seq4 = a.seq4;
            b.n_samples = n; b.min_mapQ = a.min_mapQ; b.min_baseQ = a.min_baseQ; b.illumina = a.illumina; b.lcap = (int)min(lroom, (long long)0x7fffffff);
            b.max_span = max_span; b.cells = a.cells; b.codes = a.codes;
            if (pb_block_codes<true>(b, s_resv[0], nh, (long long)rlo, (long long)rlo + n_blk, p0, smem_raw) && tid == 0)
                a.blk[blockIdx.x].y = (uint32_t)nh | 0x80000000u;              // done: nothing left for k_cell_codes here
        }
    }
}

struct PbCellCodesArgs {
    const int32_t *pos;
    const uint32_t *meta, *cigstart, *ncig, *cigar;
    const uint64_t *base;
    const uint8_t *qual, *seq4;
    int span_beg, n_samples, spc;
    int min_mapQ, min_baseQ, illumina;
    int lcap;                                // reads the shared-memory lists hold
    int max_span;                            // as in PbPileReadsArgs
    PbCounters *ctr;
    const uint4 *blk;                        // per block of k_pile_reads: {first directory entry, entries (bit 31: codes already there), first read that can cover the block, end of the block's reads}
    uint4 *cells;                            // in: {position, sample | k << 8, -, first code}; out: .y = sample | codes found << 8, .z = sum of mapq^2
    uint16_t *codes;
};
static inline size_t pb_cell_codes_smem(int n_samples, int lcap) { return pb_block_codes_smem(n_samples, 8, lcap) + 16; }

// pb_block_codes for the blocks whose k_pile_reads CTA did not do it itself (its lists were too small), one CTA per block
__global__ void __launch_bounds__(256) k_cell_codes(const PbCellCodesArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.ctr->arena_overflow || a.ctr->spec_fail) return;
    const uint4 rec = a.blk[blockIdx.x];
    if (rec.y == 0 || (rec.y >> 31)) return;
    PbBlockCodes b;
    b.pos = a.pos; b.meta = a.meta; b.cigstart = a.cigstart; b.ncig = a.ncig; b.cigar = a.cigar; b.base = a.base; b.qual = a.qual; b.seq4 = a.seq4;
    b.n_samples = a.n_samples; b.min_mapQ = a.min_mapQ; b.min_baseQ = a.min_baseQ; b.illumina = a.illumina; b.lcap = a.lcap;
    b.max_span = a.max_span > 0 ? a.max_span : a.ctr->max_span;
    b.cells = a.cells; b.codes = a.codes;
    const int p0 = a.span_beg + (int)blockIdx.x * a.spc * 32;
    if (!pb_block_codes<false>(b, (unsigned long long)rec.x, (int)rec.y, (long long)rec.z, (long long)rec.w, p0, smem_raw) && threadIdx.x == 0)
        a.ctr->arena_overflow = 1;                                             // (more reads than the lists hold: the host gives up on this path for the region)
}
