// pbsynth -- seeded synthetic fixture generator (test / bench infrastructure, not product).
//
// Produces, from one parameter block and one seed:
//   * per-contig read batches in the C-ABI layout of include/popbam_b200.h (pb_read_batch),
//     in coordinate-sorted "file order";
//   * the same data as on-disk files the unmodified reference can read: FASTA (+.fai),
//     coordinate-sorted BAM (BGZF) and its BAI, with @RG ID/SM/PO header lines.
// The on-disk formats are written from their specifications as summarised in SURVEY.md
// Appendix B (bgzf.c:47-61, bam.c:119-170/283-331, bam_index.c:447-528, faidx.c:178-230).
//
// Data model: uniform ACGT reference; reads of length R at uniform starts,
// 50 % reverse strand; base qualities i.i.d. from {10,20,30,35,40}; mapq i.i.d. from
// {10,29,40,60}; 3 % 40M2D60M-style, 3 % 5S45M3I47M-style CIGARs; sequencing errors at rate
// 10^(-Q/10); a SNP panel with per-sample carrier masks; ingroup split into two populations
// plus an optional outgroup sample in population "out".
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

extern "C" {

typedef struct pbsynth_params {
    int32_t n_contigs;
    int32_t contig_len;
    int32_t n_ingroup;        // ingroup samples (split into 2 populations)
    int32_t has_outgroup;     // +1 sample "og" in population "out"
    int32_t rg_per_sample;    // read groups per sample (C5 uses 2)
    double  depth;            // coverage per read group
    int32_t read_len;
    double  snp_density;      // SNPs per bp
    double  het_frac;         // fraction of carrier genotypes that are heterozygous
    double  frac_del;         // fraction of reads with a deletion CIGAR
    double  frac_ins;         // fraction of reads with soft clip + insertion CIGAR
    int32_t edge_mode;        // 1: also emit flagged reads, N bases, =/X/H/P/N ops, low-depth holes
                              // 2: realistic quality spectrum instead -- base qualities 2..41 (skewed to the top), 20 mapping qualities
    uint64_t seed;
    int32_t n_threads;
} pbsynth_params;

typedef struct pbsynth_batch {   // mirrors pb_read_batch
    int64_t n_reads, n_cigar, n_bases;
    const int32_t *pos;
    const uint32_t *meta;
    const uint32_t *cig_off;
    const uint32_t *cigar;
    const uint32_t *base_off;
    const uint8_t *seq4;
    const uint8_t *qual;
} pbsynth_batch;

}  // extern "C"

namespace {

static inline uint64_t splitmix64(uint64_t &x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) { (void)splitmix64(s); }
    uint64_t next() { return splitmix64(s); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

struct Snp {
    int32_t pos;
    uint8_t derived;    // 0..3
    uint64_t carriers;  // bit per sample
    uint64_t hets;      // subset of carriers that are heterozygous
};

struct Contig {
    std::string name;
    std::vector<char> ref;       // contig_len bytes
    std::vector<Snp> snps;       // sorted by pos
    // batch arrays
    std::vector<int32_t> pos;
    std::vector<uint32_t> meta, cig_off, cigar, base_off;
    std::vector<uint8_t> seq4, qual;
    std::vector<uint64_t> rid;   // generation id per read (names, rng stream)
    std::vector<uint8_t> kind;   // cigar template id
    std::vector<uint16_t> rg;    // read group per read
    int64_t aligned_bases = 0;
};

struct Synth {
    pbsynth_params p;
    int n_samples = 0, n_rg = 0;
    std::vector<std::string> sample_names, pop_of_sample, rg_names;
    std::vector<int> rg_sample;
    std::vector<Contig> contigs;
};

static const uint8_t kQualSet[5] = {10, 20, 30, 35, 40};
static const uint8_t kMapqSet[4] = {10, 29, 40, 60};
static const char kBases[4] = {'A', 'C', 'G', 'T'};
static const uint8_t kNt16[4] = {1, 2, 4, 8};
// error probability scaled to 2^32 for the 5 quality values
static uint32_t kErrThresh[5];
static uint32_t kErrThreshQ[64];     // the same for every quality value (edge_mode 2)
static const uint8_t kMapqWide[20] = {0, 3, 7, 10, 13, 16, 19, 22, 25, 27, 29, 30, 33, 37, 40, 44, 48, 52, 57, 60};

// CIGAR templates.  ops: M0 I1 D2 N3 S4 H5 P6 =7 X8
struct CigT { int n; uint32_t op[8]; };
static inline uint32_t cg(int len, int op) { return (uint32_t)len << 4 | (uint32_t)op; }

static void make_cigar(int kind, int R, CigT &c) {
    // all templates consume exactly R query bases
    switch (kind) {
    default:
    case 0: c.n = 1; c.op[0] = cg(R, 0); break;
    case 1: {  // aM 2D bM          (40M2D60M at R=100)
        int a = (R * 2) / 5; c.n = 3; c.op[0] = cg(a, 0); c.op[1] = cg(2, 2); c.op[2] = cg(R - a, 0); break; }
    case 2: {  // 5S aM 3I bM       (5S45M3I47M at R=100)
        int a = (R * 9) / 20; c.n = 4; c.op[0] = cg(5, 4); c.op[1] = cg(a, 0); c.op[2] = cg(3, 1);
        c.op[3] = cg(R - 8 - a, 0); break; }
    case 3: {  // 3H 4S a= 1X b= 6S  (edge: = / X / hard clip / trailing soft clip)
        int a = (R - 11) / 2; c.n = 6; c.op[0] = cg(3, 5); c.op[1] = cg(4, 4); c.op[2] = cg(a, 7);
        c.op[3] = cg(1, 8); c.op[4] = cg(R - 11 - a, 7); c.op[5] = cg(6, 4); break; }
    case 4: {  // aM 7N bM 1P 2I cM  (edge: ref skip, padding)
        int a = R / 4, b = R / 4; c.n = 6; c.op[0] = cg(a, 0); c.op[1] = cg(7, 3); c.op[2] = cg(b, 0);
        c.op[3] = cg(1, 6); c.op[4] = cg(2, 1); c.op[5] = cg(R - a - b - 2, 0); break; }
    case 5: {  // 2I aM 1D 1I bM    (edge: leading insertion, adjacent D/I)
        int a = R / 3; c.n = 5; c.op[0] = cg(2, 1); c.op[1] = cg(a, 0); c.op[2] = cg(1, 2); c.op[3] = cg(1, 1);
        c.op[4] = cg(R - 3 - a, 0); break; }
    }
}
static int cigar_refspan(const CigT &c) {
    int s = 0;
    for (int i = 0; i < c.n; ++i) { int op = c.op[i] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) s += c.op[i] >> 4; }
    return s;
}

static void build_panel(Synth &S, Contig &C, uint64_t seed) {
    const pbsynth_params &p = S.p;
    Rng r(seed);
    int L = p.contig_len;
    C.ref.resize(L);
    for (int i = 0; i < L; ++i) C.ref[i] = kBases[r.next() >> 62];
    if (p.edge_mode == 1) {  // a few lower-case and N reference bytes
        for (int i = 0; i < L / 997 + 1; ++i) { int q = r.below(L); C.ref[q] = (char)(C.ref[q] | 0x20); }
        for (int i = 0; i < L / 1999 + 1; ++i) { int q = r.below(L); C.ref[q] = 'N'; }
    }
    int64_t ns = (int64_t)std::llround(p.snp_density * L);
    std::vector<int32_t> ps(ns);
    for (auto &x : ps) x = (int32_t)r.below(L);
    std::sort(ps.begin(), ps.end());
    ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
    int n_in = p.n_ingroup, half = (n_in + 1) / 2;
    uint64_t in_mask = n_in >= 64 ? ~0ULL : ((1ULL << n_in) - 1);
    uint64_t popA = half >= 64 ? ~0ULL : ((1ULL << half) - 1), popB = in_mask & ~popA;
    uint64_t og = p.has_outgroup ? (1ULL << n_in) : 0;
    C.snps.reserve(ps.size());
    for (int32_t q : ps) {
        Snp s; s.pos = q;
        int rb = 0; char rc = (char)(C.ref[q] & ~0x20);
        for (int b = 0; b < 4; ++b) if (kBases[b] == rc) rb = b;
        s.derived = (uint8_t)((rb + 1 + r.below(3)) & 3);
        double u = r.uni();
        uint64_t car = 0;
        if (u < 0.70) {  // segregating in the ingroup with independent per-population frequency
            double fa = r.uni(), fb = r.uni();
            for (int i = 0; i < n_in; ++i) if (r.uni() < ((popA >> i & 1) ? fa : fb)) car |= 1ULL << i;
            if (og && r.uni() < 0.15) car |= og;
        } else if (u < 0.80) { car = in_mask; }            // fixed in ingroup vs outgroup/reference
        else if (u < 0.88) { car = og ? og : popA; }       // outgroup-only difference
        else if (u < 0.94) { car = popA; }                 // fixed between populations
        else { car = popB | ((r.uni() < 0.5) ? og : 0); }
        uint64_t het = 0;
        for (int i = 0; i < S.n_samples; ++i) if ((car >> i & 1) && r.uni() < p.het_frac) het |= 1ULL << i;
        s.carriers = car; s.hets = het;
        C.snps.push_back(s);
    }
}

struct ReadKey { int32_t pos; uint32_t rg; uint64_t id; uint8_t kind; };

static void gen_contig(Synth &S, int ci) {
    const pbsynth_params &p = S.p;
    Contig &C = S.contigs[ci];
    uint64_t cseed = p.seed * 0x100000001b3ULL + 0x51ed270b7a5ULL * (uint64_t)(ci + 1);
    build_panel(S, C, cseed);
    const int L = p.contig_len, R = p.read_len;
    // ---- 1. read starts per read group
    std::vector<ReadKey> keys;
    int64_t per_rg = (int64_t)std::llround(p.depth * (double)L / R);
    keys.reserve((size_t)per_rg * S.n_rg);
    uint64_t idc = 0;
    for (int g = 0; g < S.n_rg; ++g) {
        Rng r(cseed ^ (0xabcdef12345ULL * (uint64_t)(g + 7)));
        for (int64_t i = 0; i < per_rg; ++i) {
            double u = r.uni();
            int kind = 0;
            if (u < p.frac_del) kind = 1; else if (u < p.frac_del + p.frac_ins) kind = 2;
            else if (p.edge_mode == 1 && u < p.frac_del + p.frac_ins + 0.02) kind = 3 + (int)r.below(3);
            CigT c; make_cigar(kind, R, c);
            int span = cigar_refspan(c);
            if (span >= L) { kind = 0; make_cigar(0, R, c); span = R; }
            ReadKey k; k.pos = (int32_t)r.below((uint32_t)(L - span + 1)); k.rg = (uint32_t)g; k.id = idc++; k.kind = (uint8_t)kind;
            if (p.edge_mode == 1) {  // carve a low-coverage hole for one read group
                int hb = L / 3, he = hb + L / 50;
                if (g == 1 && k.pos + span > hb && k.pos < he) continue;
            }
            keys.push_back(k);
        }
    }
    std::stable_sort(keys.begin(), keys.end(), [](const ReadKey &a, const ReadKey &b) { return a.pos < b.pos; });
    const int64_t N = (int64_t)keys.size();
    // ---- 2. offsets
    C.pos.resize(N); C.meta.resize(N); C.cig_off.resize(N + 1); C.base_off.resize(N + 1);
    C.rid.resize(N); C.kind.resize(N); C.rg.resize(N);
    uint64_t co = 0, bo = 0;
    const int Rpad = (R + 1) & ~1;
    for (int64_t i = 0; i < N; ++i) {
        CigT c; make_cigar(keys[i].kind, R, c);
        C.cig_off[i] = (uint32_t)co; C.base_off[i] = (uint32_t)bo;
        co += c.n; bo += Rpad;
    }
    C.cig_off[N] = (uint32_t)co; C.base_off[N] = (uint32_t)bo;
    if (bo > 0xfffffff0ULL) { fprintf(stderr, "pbsynth: contig too deep for 32-bit base offsets\n"); abort(); }
    C.cigar.resize(co); C.qual.assign(bo, 0); C.seq4.assign(bo / 2, 0);
    // ---- 3. contents, in parallel (each read owns an RNG stream keyed by its generation id)
    int nt = std::max(1, p.n_threads);
    std::vector<std::thread> th;
    std::vector<int64_t> aligned(nt, 0);
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            int64_t lo = N * t / nt, hi = N * (t + 1) / nt, al = 0;
            std::vector<uint8_t> qb(R), bb(R);
            for (int64_t i = lo; i < hi; ++i) {
                const ReadKey &k = keys[i];
                Rng r(cseed + 0x9e3779b97f4a7c15ULL * (k.id + 1));
                int smp = S.rg_sample[k.rg];
                CigT c; make_cigar(k.kind, R, c);
                for (int j = 0; j < c.n; ++j) C.cigar[C.cig_off[i] + j] = c.op[j];
                uint64_t w = r.next();
                uint32_t flag = (w & 1) ? 16u : 0u;             // reverse strand
                uint32_t mapq = p.edge_mode == 2 ? kMapqWide[(w >> 1) % 20] : kMapqSet[(w >> 1) & 3];
                uint32_t smeta = (uint32_t)smp;
                if (p.edge_mode == 1) {
                    uint32_t e = (uint32_t)((w >> 8) & 0x3ff);
                    if (e < 6) flag |= 0x400;                   // duplicate
                    else if (e < 10) flag |= 0x200;             // QC fail
                    else if (e < 14) flag |= 0x100;             // secondary
                    else if (e < 16) flag |= 0x4;               // unmapped (placed)
                    else if (e < 20) mapq = 0;
                    else if (e < 24) mapq = 255;
                }
                // haplotype choice for heterozygous genotypes: one per read
                bool hap_alt = (w >> 20) & 1;
                // walk the CIGAR laying down query bases
                int x = k.pos, y = 0;
                size_t sidx = std::lower_bound(C.snps.begin(), C.snps.end(), k.pos,
                                               [](const Snp &s, int32_t v) { return s.pos < v; }) - C.snps.begin();
                for (int j = 0; j < c.n; ++j) {
                    int op = c.op[j] & 15, len = (int)(c.op[j] >> 4);
                    if (op == 0 || op == 7 || op == 8) {
                        for (int z = 0; z < len; ++z, ++x, ++y) {
                            char rc = (char)(C.ref[x] & ~0x20);
                            int b = rc == 'A' ? 0 : rc == 'C' ? 1 : rc == 'G' ? 2 : rc == 'T' ? 3 : (int)(r.next() >> 62);
                            while (sidx < C.snps.size() && C.snps[sidx].pos < x) ++sidx;
                            if (sidx < C.snps.size() && C.snps[sidx].pos == x) {
                                const Snp &s = C.snps[sidx];
                                if (s.carriers >> smp & 1) { if (!(s.hets >> smp & 1) || hap_alt) b = s.derived; }
                            }
                            bb[y] = (uint8_t)b;
                        }
                        al += len;
                    } else if (op == 1 || op == 4) {
                        for (int z = 0; z < len; ++z, ++y) bb[y] = (uint8_t)(r.next() >> 62);
                    } else if (op == 2 || op == 3) {
                        x += len;
                    }
                }
                // qualities + errors
                uint8_t *q = &C.qual[C.base_off[i]];
                uint8_t *s4 = &C.seq4[C.base_off[i] / 2];
                for (int y2 = 0; y2 < R; ++y2) {
                    uint64_t u = r.next();
                    int qi = (int)((u >> 60) % 5);
                    uint8_t code = kNt16[bb[y2]];
                    if (p.edge_mode == 2) {
                        // 2..41, the larger of two draws: most bases in the thirties, a tail down to 2 (as a real run's)
                        const int qa = (int)((u >> 52) % 40), qb2 = (int)((u >> 46) % 40), qv = 2 + (qa > qb2 ? qa : qb2);
                        if ((uint32_t)u < kErrThreshQ[qv]) code = kNt16[(bb[y2] + 1 + ((u >> 40) % 3)) & 3];
                        q[y2] = (uint8_t)qv;
                        if (y2 & 1) s4[y2 >> 1] |= code; else s4[y2 >> 1] = (uint8_t)(code << 4);
                        continue;
                    }
                    if ((uint32_t)u < kErrThresh[qi]) code = kNt16[(bb[y2] + 1 + ((u >> 40) % 3)) & 3];
                    if (p.edge_mode) {
                        uint32_t e = (uint32_t)((u >> 44) & 0xfff);
                        if (e < 8) code = 15;                    // N base
                        else if (e < 10) code = 3;               // ambiguity code M
                        if (e >= 10 && e < 14) q[y2] = (uint8_t)(e - 10);   // very low quality
                        else if (e == 14) q[y2] = 93;            // above the 63 clamp
                        else q[y2] = kQualSet[qi];
                    } else q[y2] = kQualSet[qi];
                    if (y2 & 1) s4[y2 >> 1] |= code; else s4[y2 >> 1] = (uint8_t)(code << 4);
                }
                C.pos[i] = k.pos;
                C.meta[i] = flag << 16 | mapq << 8 | smeta;
                C.rid[i] = k.id; C.kind[i] = k.kind; C.rg[i] = (uint16_t)k.rg;
            }
            aligned[t] = al;
        });
    }
    for (auto &x : th) x.join();
    for (int t = 0; t < nt; ++t) C.aligned_bases += aligned[t];
}

// ------------------------------------------------------------------ file writers
struct Bgzf {
    FILE *f; std::vector<uint8_t> buf; uint64_t coff = 0; int level;
    std::vector<uint8_t> out;
    explicit Bgzf(FILE *fp, int lvl) : f(fp), level(lvl) { buf.reserve(0xff00); out.resize(0x10000 + 64); }
    uint64_t tell() const { return coff << 16 | (uint64_t)buf.size(); }
    void flush_block() {
        // one gzip member with the BC extra field (bgzf.c:47-61); payload is raw deflate
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = buf.data(); zs.avail_in = (uInt)buf.size();
        zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 18 - 8);
        int rc = deflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END) { fprintf(stderr, "pbsynth: deflate failed\n"); abort(); }
        uint32_t clen = (uint32_t)zs.total_out; deflateEnd(&zs);
        uint32_t bsize = clen + 18 + 8;
        static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(out.data(), hdr, 16);
        out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf.data(), (uInt)buf.size());
        uint32_t isz = (uint32_t)buf.size();
        uint8_t *t = out.data() + 18 + clen;
        for (int i = 0; i < 4; ++i) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isz >> (8 * i)); }
        fwrite(out.data(), 1, bsize, f);
        coff += bsize; buf.clear();
    }
    void write(const void *p, size_t n) {
        const uint8_t *s = (const uint8_t *)p;
        while (n) {
            size_t room = 0xff00 - buf.size(), k = std::min(room, n);
            buf.insert(buf.end(), s, s + k); s += k; n -= k;
            if (buf.size() == 0xff00) flush_block();
        }
    }
    void finish() {
        if (!buf.empty()) flush_block();
        flush_block();   // empty block == EOF marker (bgzf.c:700-713)
    }
};

static inline int reg2bin(uint32_t beg, uint32_t end) {   // UCSC binning scheme, 5 levels, 16 kb leaves
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
}
static void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
static void put64(std::vector<uint8_t> &v, uint64_t x) { for (int i = 0; i < 8; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

struct BinChunks { std::vector<std::pair<uint64_t, uint64_t>> ch; };

static int write_files(Synth &S, const char *prefix, int level) {
    std::string fa = std::string(prefix) + ".fa", bam = std::string(prefix) + ".bam";
    // ---- FASTA + .fai (faidx.c:178-230: name, len, offset, line_blen, line_len)
    {
        FILE *f = fopen(fa.c_str(), "wb"), *fi = fopen((fa + ".fai").c_str(), "wb");
        if (!f || !fi) return -1;
        uint64_t off = 0; const int LW = 60;
        for (auto &C : S.contigs) {
            off += (uint64_t)fprintf(f, ">%s\n", C.name.c_str());
            fprintf(fi, "%s\t%d\t%llu\t%d\t%d\n", C.name.c_str(), (int)C.ref.size(), (unsigned long long)off, LW, LW + 1);
            for (size_t i = 0; i < C.ref.size(); i += LW) {
                size_t k = std::min((size_t)LW, C.ref.size() - i);
                fwrite(&C.ref[i], 1, k, f); fputc('\n', f); off += k + 1;
            }
        }
        fclose(f); fclose(fi);
    }
    // ---- BAM
    FILE *fb = fopen(bam.c_str(), "wb"); if (!fb) return -1;
    Bgzf bz(fb, level);
    std::string text = "@HD\tVN:1.0\tSO:coordinate\n";
    for (auto &C : S.contigs) text += "@SQ\tSN:" + C.name + "\tLN:" + std::to_string(C.ref.size()) + "\tAS:synth\n";
    for (int g = 0; g < S.n_rg; ++g) {
        int s = S.rg_sample[g];
        text += "@RG\tID:" + S.rg_names[g] + "\tSM:" + S.sample_names[s] + "\tPO:" + S.pop_of_sample[s] + "\n";
    }
    std::vector<uint8_t> h;
    h.push_back('B'); h.push_back('A'); h.push_back('M'); h.push_back(1);
    put32(h, (uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end());
    put32(h, (uint32_t)S.contigs.size());
    for (auto &C : S.contigs) { put32(h, (uint32_t)C.name.size() + 1); h.insert(h.end(), C.name.begin(), C.name.end()); h.push_back(0); put32(h, (uint32_t)C.ref.size()); }
    bz.write(h.data(), h.size());
    bz.flush_block();   // records start on a block boundary (not required, keeps offsets tidy)
    // ---- records + index (bam_index.c:447-528)
    std::vector<uint8_t> bai; bai.push_back('B'); bai.push_back('A'); bai.push_back('I'); bai.push_back(1);
    put32(bai, (uint32_t)S.contigs.size());
    std::vector<uint8_t> rec;
    for (size_t ci = 0; ci < S.contigs.size(); ++ci) {
        Contig &C = S.contigs[ci];
        std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> binrecs;  // bin -> (beg,end) voffsets, in file order
        size_t nlin = (C.ref.size() >> 14) + 1;
        std::vector<uint64_t> lin(nlin, 0);
        int64_t N = (int64_t)C.pos.size();
        for (int64_t i = 0; i < N; ++i) {
            uint32_t flag = C.meta[i] >> 16, mapq = (C.meta[i] >> 8) & 0xff, smp = C.meta[i] & 0xff;
            (void)smp;
            uint32_t ncig = C.cig_off[i + 1] - C.cig_off[i];
            int R = S.p.read_len;
            // ref end
            uint32_t end = (uint32_t)C.pos[i];
            for (uint32_t j = 0; j < ncig; ++j) { uint32_t c = C.cigar[C.cig_off[i] + j]; int op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += c >> 4; }
            if (end == (uint32_t)C.pos[i]) end += 1;
            char qname[32]; int lq = snprintf(qname, sizeof qname, "r%llu", (unsigned long long)C.rid[i]) + 1;
            const std::string &rgn = S.rg_names[C.rg[i]];
            int bin = reg2bin((uint32_t)C.pos[i], end);
            rec.clear();
            uint32_t l_data = (uint32_t)lq + 4 * ncig + (uint32_t)(R + 1) / 2 + (uint32_t)R + 3 + (uint32_t)rgn.size() + 1;
            put32(rec, 32 + l_data);
            put32(rec, (uint32_t)ci); put32(rec, (uint32_t)C.pos[i]);
            put32(rec, (uint32_t)bin << 16 | mapq << 8 | (uint32_t)lq);
            put32(rec, flag << 16 | ncig);
            put32(rec, (uint32_t)R); put32(rec, 0xffffffffu); put32(rec, 0xffffffffu); put32(rec, 0);
            rec.insert(rec.end(), qname, qname + lq);
            for (uint32_t j = 0; j < ncig; ++j) put32(rec, C.cigar[C.cig_off[i] + j]);
            const uint8_t *s4 = &C.seq4[C.base_off[i] / 2]; rec.insert(rec.end(), s4, s4 + (R + 1) / 2);
            const uint8_t *q = &C.qual[C.base_off[i]]; rec.insert(rec.end(), q, q + R);
            rec.push_back('R'); rec.push_back('G'); rec.push_back('Z'); rec.insert(rec.end(), rgn.begin(), rgn.end()); rec.push_back(0);
            uint64_t v0 = bz.tell();
            bz.write(rec.data(), rec.size());
            uint64_t v1 = bz.tell();
            if (!(flag & 4)) {
                binrecs.push_back({(uint32_t)bin, {v0, v1}});
                for (uint32_t w = (uint32_t)C.pos[i] >> 14; w <= (end - 1) >> 14 && w < nlin; ++w) if (lin[w] == 0) lin[w] = v0;
            } else {
                binrecs.push_back({(uint32_t)bin, {v0, v1}});
            }
        }
        // group by bin, merge adjacent chunks
        std::stable_sort(binrecs.begin(), binrecs.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        std::vector<std::pair<uint32_t, BinChunks>> bins;
        for (auto &br : binrecs) {
            if (bins.empty() || bins.back().first != br.first) bins.push_back({br.first, BinChunks()});
            auto &ch = bins.back().second.ch;
            if (!ch.empty() && (ch.back().second >> 16) == (br.second.first >> 16)) ch.back().second = br.second.second;
            else ch.push_back(br.second);
        }
        put32(bai, (uint32_t)bins.size());
        for (auto &b : bins) { put32(bai, b.first); put32(bai, (uint32_t)b.second.ch.size()); for (auto &c : b.second.ch) { put64(bai, c.first); put64(bai, c.second); } }
        // linear index: fill gaps with the next known offset going backwards is NOT what samtools
        // does; it fills forward (empty slot inherits the previous one)
        for (size_t w = 1; w < nlin; ++w) if (lin[w] == 0) lin[w] = lin[w - 1];
        put32(bai, (uint32_t)nlin); for (auto v : lin) put64(bai, v);
    }
    bz.finish(); fclose(fb);
    FILE *fi = fopen((bam + ".bai").c_str(), "wb"); if (!fi) return -1;
    fwrite(bai.data(), 1, bai.size(), fi); fclose(fi);
    return 0;
}

}  // namespace

extern "C" {

void pbsynth_default_params(pbsynth_params *p) {
    memset(p, 0, sizeof *p);
    p->n_contigs = 1; p->contig_len = 100000; p->n_ingroup = 10; p->has_outgroup = 1; p->rg_per_sample = 1;
    p->depth = 20.0; p->read_len = 100; p->snp_density = 0.01; p->het_frac = 0.05; p->frac_del = 0.03; p->frac_ins = 0.03;
    p->edge_mode = 0; p->seed = 1; p->n_threads = 4;
}

void *pbsynth_create(const pbsynth_params *pp) {
    for (int i = 0; i < 5; ++i) kErrThresh[i] = (uint32_t)(std::pow(10.0, -kQualSet[i] / 10.0) * 4294967296.0);
    for (int i = 0; i < 64; ++i) kErrThreshQ[i] = (uint32_t)(std::min(0.75, std::pow(10.0, -i / 10.0)) * 4294967296.0);
    Synth *S = new Synth();
    S->p = *pp;
    const pbsynth_params &p = S->p;
    S->n_samples = p.n_ingroup + (p.has_outgroup ? 1 : 0);
    if (S->n_samples > 64 || S->n_samples < 1 || p.read_len < 24 || p.contig_len <= p.read_len + 16) { delete S; return nullptr; }
    int half = (p.n_ingroup + 1) / 2;
    for (int i = 0; i < p.n_ingroup; ++i) { S->sample_names.push_back("s" + std::to_string(i)); S->pop_of_sample.push_back(i < half ? "popA" : "popB"); }
    if (p.has_outgroup) { S->sample_names.push_back("og"); S->pop_of_sample.push_back("out"); }
    for (int s = 0; s < S->n_samples; ++s)
        for (int k = 0; k < p.rg_per_sample; ++k) { S->rg_names.push_back("rg" + std::to_string(s) + "_" + std::to_string(k)); S->rg_sample.push_back(s); }
    S->n_rg = (int)S->rg_names.size();
    S->contigs.resize(p.n_contigs);
    for (int c = 0; c < p.n_contigs; ++c) { S->contigs[c].name = "chr" + std::to_string(c + 1); gen_contig(*S, c); }
    return S;
}
void pbsynth_destroy(void *h) { delete (Synth *)h; }
int pbsynth_n_samples(void *h) { return ((Synth *)h)->n_samples; }
int pbsynth_n_pops(void *h) { Synth *S = (Synth *)h; return (S->p.n_ingroup > 1 ? 2 : 1) + (S->p.has_outgroup ? 1 : 0); }
// population index of a sample in order of first appearance (popA, popB, out)
int pbsynth_sample_pop(void *h, int s) {
    Synth *S = (Synth *)h; const std::string &p = S->pop_of_sample[s];
    if (p == "popA") return 0; if (p == "popB") return 1; return S->p.n_ingroup > 1 ? 2 : 1;
}
const char *pbsynth_sample_name(void *h, int s) { return ((Synth *)h)->sample_names[s].c_str(); }
const char *pbsynth_pop_name(void *h, int p) {
    Synth *S = (Synth *)h; static const char *n[3] = {"popA", "popB", "out"};
    if (S->p.n_ingroup <= 1 && p == 1) return n[2];
    return n[p];
}
const char *pbsynth_contig_name(void *h, int c) { return ((Synth *)h)->contigs[c].name.c_str(); }
const char *pbsynth_ref(void *h, int c) { return ((Synth *)h)->contigs[c].ref.data(); }
int64_t pbsynth_aligned_bases(void *h, int c) { return ((Synth *)h)->contigs[c].aligned_bases; }
int pbsynth_batch_get(void *h, int c, pbsynth_batch *b) {
    Synth *S = (Synth *)h; if (c < 0 || c >= (int)S->contigs.size()) return -1;
    Contig &C = S->contigs[c];
    b->n_reads = (int64_t)C.pos.size(); b->n_cigar = (int64_t)C.cigar.size(); b->n_bases = (int64_t)C.qual.size();
    b->pos = C.pos.data(); b->meta = C.meta.data(); b->cig_off = C.cig_off.data(); b->cigar = C.cigar.data();
    b->base_off = C.base_off.data(); b->seq4 = C.seq4.data(); b->qual = C.qual.data();
    return 0;
}
int64_t pbsynth_n_snps(void *h, int c) { return (int64_t)((Synth *)h)->contigs[c].snps.size(); }
int pbsynth_write_files(void *h, const char *prefix, int level) { return write_files(*(Synth *)h, prefix, level); }

}  // extern "C"

#ifdef PBSYNTH_MAIN
static void usage() {
    fprintf(stderr,
            "usage: pbsynth -o PREFIX [-c contigs] [-l contig_len] [-n ingroup] [-g 0|1 outgroup] [-r rg_per_sample]\n"
            "               [-d depth] [-R read_len] [-s snp_density] [-H het_frac] [-e edge] [-S seed] [-t threads] [-z level]\n");
}
int main(int argc, char **argv) {
    pbsynth_params p; pbsynth_default_params(&p);
    const char *prefix = nullptr; int level = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto nxt = [&]() -> const char * { if (i + 1 >= argc) { usage(); exit(1); } return argv[++i]; };
        if (a == "-o") prefix = nxt(); else if (a == "-c") p.n_contigs = atoi(nxt()); else if (a == "-l") p.contig_len = atoi(nxt());
        else if (a == "-n") p.n_ingroup = atoi(nxt()); else if (a == "-g") p.has_outgroup = atoi(nxt()); else if (a == "-r") p.rg_per_sample = atoi(nxt());
        else if (a == "-d") p.depth = atof(nxt()); else if (a == "-R") p.read_len = atoi(nxt()); else if (a == "-s") p.snp_density = atof(nxt());
        else if (a == "-H") p.het_frac = atof(nxt()); else if (a == "-e") p.edge_mode = atoi(nxt()); else if (a == "-S") p.seed = strtoull(nxt(), 0, 10);
        else if (a == "-t") p.n_threads = atoi(nxt()); else if (a == "-z") level = atoi(nxt());
        else { usage(); return 1; }
    }
    if (!prefix) { usage(); return 1; }
    void *h = pbsynth_create(&p);
    if (!h) { fprintf(stderr, "pbsynth: bad parameters\n"); return 1; }
    if (pbsynth_write_files(h, prefix, level)) { fprintf(stderr, "pbsynth: cannot write %s.*\n", prefix); return 1; }
    int64_t nr = 0, nb = 0;
    for (int c = 0; c < p.n_contigs; ++c) { pbsynth_batch b; pbsynth_batch_get(h, c, &b); nr += b.n_reads; nb += pbsynth_aligned_bases(h, c); }
    fprintf(stderr, "pbsynth: %lld reads, %lld aligned bases -> %s.{fa,bam,bam.bai}\n", (long long)nr, (long long)nb, prefix);
    pbsynth_destroy(h);
    return 0;
}
#endif
