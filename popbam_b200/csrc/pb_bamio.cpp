// pb_bamio.cpp -- see pb_bamio.h.  Host feeder of the popbam command line.
#include "pb_bamio.h"
#include "pb_inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

namespace pbio {

namespace {
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint64_t le64(const uint8_t *p) { return (uint64_t)le32(p) | (uint64_t)le32(p + 4) << 32; }
inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
[[noreturn]] void fail(const std::string &m) { throw Error{m}; }
}  // namespace

// ------------------------------------------------------------------------------------------------ BGZF
BgzfFile::~BgzfFile() {
    if (map_) munmap(const_cast<uint8_t *>(map_), size_);
    if (fd_ >= 0) close(fd_);
}

void BgzfFile::open(const std::string &path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail("Cannot read BAM file " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 28) fail("Cannot read BAM file " + path);
    size_ = (uint64_t)st.st_size;
    void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) fail("Cannot map BAM file " + path);
    map_ = static_cast<const uint8_t *>(m);
    madvise(m, size_, MADV_SEQUENTIAL);
}

uint32_t BgzfFile::inflate_block(uint64_t coff, std::vector<uint8_t> &out) const {
    out.clear();
    if (coff + 18 > size_) return 0;
    const uint8_t *h = map_ + coff;
    // gzip member with a 'BC' extra subfield holding BSIZE (member size - 1)
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail("invalid BGZF block header");
    const uint32_t xlen = le16(h + 10);
    if (coff + 12 + xlen > size_) fail("truncated BGZF block");
    uint32_t bsize = 0;
    for (uint32_t o = 0; o + 4 <= xlen;) {
        const uint8_t *x = h + 12 + o;
        const uint32_t slen = le16(x + 2);
        if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = (uint32_t)le16(x + 4) + 1;
        o += 4 + slen;
    }
    if (!bsize || coff + bsize > size_) fail("truncated BGZF block");
    const uint32_t hdr = 12 + xlen;
    if ((uint64_t)hdr + 8 > bsize) fail("corrupted BGZF block (extra field longer than the block)");
    const uint32_t isize = le32(h + bsize - 4);
    if (isize > 65536) fail("corrupted BGZF block (uncompressed size above 64 KiB, bgzf.c:47-61)");
    out.resize(isize);
    if (isize && !inflate_raw(h + hdr, bsize - hdr - 8, out.data(), isize)) fail("BGZF inflate failed");
    return bsize;
}

void BgzfReader::seek(uint64_t voffset) {
    const uint64_t c = voffset >> 16;
    if (!loaded_ || c != coff_) { coff_ = c; loaded_ = false; }
    uoff_ = (uint32_t)(voffset & 0xffff);
}

bool BgzfReader::load() {
    csize_ = f_.inflate_block(coff_, buf_);
    loaded_ = csize_ != 0;
    return loaded_;
}

bool BgzfReader::read(void *dst, size_t n) {
    uint8_t *d = static_cast<uint8_t *>(dst);
    while (n) {
        if (!loaded_ && !load()) return false;
        if (uoff_ >= buf_.size()) {      // next block (empty blocks are skipped the same way)
            coff_ += csize_; uoff_ = 0; loaded_ = false;
            continue;
        }
        const size_t k = std::min(n, buf_.size() - uoff_);
        memcpy(d, buf_.data() + uoff_, k);
        d += k; n -= k; uoff_ += (uint32_t)k;
    }
    // normalise the position so tell() points at the next byte's block
    if (loaded_ && uoff_ >= buf_.size()) { coff_ += csize_; uoff_ = 0; loaded_ = false; }
    return true;
}

bool BgzfReader::eof() {
    for (;;) {
        if (!loaded_ && !load()) return true;
        if (uoff_ < buf_.size()) return false;
        coff_ += csize_; uoff_ = 0; loaded_ = false;
    }
}

// ------------------------------------------------------------------------------------------------ header
int BamHeader::tid_of(const std::string &name) const {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return (int)i;
    return -1;
}

BamHeader read_header(const BgzfFile &f) {
    BgzfReader r(f);
    r.seek(0);
    uint8_t b4[4];
    if (!r.read(b4, 4) || memcmp(b4, "BAM\1", 4) != 0) fail("Cannot read BAM header: bad magic");
    BamHeader h;
    if (!r.read(b4, 4)) fail("Cannot read BAM header");
    h.text.resize(le32(b4));
    if (!h.text.empty() && !r.read(&h.text[0], h.text.size())) fail("Cannot read BAM header text");
    while (!h.text.empty() && h.text.back() == '\0') h.text.pop_back();
    if (!r.read(b4, 4)) fail("Cannot read BAM header");
    const uint32_t nref = le32(b4);
    for (uint32_t i = 0; i < nref; ++i) {
        if (!r.read(b4, 4)) fail("Cannot read BAM header");
        std::string nm(le32(b4), '\0');
        if (!nm.empty() && !r.read(&nm[0], nm.size())) fail("Cannot read BAM header");
        while (!nm.empty() && nm.back() == '\0') nm.pop_back();
        if (!r.read(b4, 4)) fail("Cannot read BAM header");
        h.names.push_back(nm);
        h.lens.push_back((int32_t)le32(b4));
    }
    h.first_record_voffset = r.tell();
    return h;
}

// ------------------------------------------------------------------------------------------------ samples
namespace {
std::string tag_value(const char *p) {
    const char *e = p;
    while (*e && *e != '\t' && *e != '\n') ++e;
    return std::string(p, e);
}
}  // namespace

SampleTable build_samples(const std::string &text, const std::string &bam_path) {
    // bam_smpl_add (pop_sample.cpp:15-107): for every "@RG", the next "\tID:", "\tSM:" and "\tPO:" found from
    // there on (not bounded by the line); first-seen order of SM defines sample ids, of PO population ids
    SampleTable st;
    std::unordered_map<std::string, int> sm2id, pop2id, sm2pop;
    const char *p = text.c_str();
    int n = 0;
    while (const char *q0 = strstr(p, "@RG")) {
        p = q0 + 3;
        const char *q = strstr(p, "\tID:"), *r = strstr(p, "\tSM:"), *s = strstr(p, "\tPO:");
        if (!q || !r) break;
        const std::string id = tag_value(q + 4), smn = tag_value(r + 4);
        if (!st.rg2sample.count(id)) {       // duplicated @RG-ID keeps its first sample
            auto it = sm2id.find(smn);
            if (it == sm2id.end()) { it = sm2id.emplace(smn, (int)st.samples.size()).first; st.samples.push_back(smn); }
            st.rg2sample[id] = it->second;
        }
        if (s) {
            const std::string pon = tag_value(s + 4);
            if (!sm2pop.count(smn)) {        // a sample keeps the population of its first read group
                auto it = pop2id.find(pon);
                if (it == pop2id.end()) { it = pop2id.emplace(pon, (int)st.pops.size()).first; st.pops.push_back(pon); }
                sm2pop[smn] = it->second;
            }
        }
        const char *mx = q + 4;
        if (r + 4 > mx) mx = r + 4;
        if (s && s + 4 > mx) mx = s + 4;
        p = mx;
        ++n;
    }
    if (n == 0) {   // no read groups: the file is one sample in one population (pop_sample.cpp:97-101)
        st.samples.push_back(bam_path);
        st.pops.push_back(bam_path);
        sm2pop[bam_path] = 0;
    }
    if (st.samples.size() > 64) fail("popbam supports at most 64 samples");
    st.sample_pop.assign(st.samples.size(), -1);
    for (size_t i = 0; i < st.samples.size(); ++i) {      // assign_pops (popbam.cpp:145-171)
        auto it = sm2pop.find(st.samples[i]);
        if (it == sm2pop.end())
            fail("Sample " + st.samples[i] + " not assigned to a population.\nPlease check BAM header file definitions");
        st.sample_pop[i] = it->second;
        st.pop_mask[it->second] |= 1ULL << i;
        st.pop_nsmpl[it->second]++;
    }
    return st;
}

// ------------------------------------------------------------------------------------------------ BAI
void BamIndex::load(const std::string &path, size_t n_ref_expected) {
    std::ifstream in(path, std::ios::binary);
    if (!in) fail("Index file not available for BAM file " + path.substr(0, path.size() > 4 ? path.size() - 4 : 0));
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) fail("wrong magic number in the BAM index " + path);
    size_t o = 4;
    auto need = [&](size_t k) { if (o + k > d.size()) fail("truncated BAM index " + path); };
    need(4);
    const uint32_t nref = le32(&d[o]); o += 4;
    (void)n_ref_expected;
    refs_.resize(nref);
    for (uint32_t i = 0; i < nref; ++i) {
        need(4);
        const uint32_t nbin = le32(&d[o]); o += 4;
        for (uint32_t b = 0; b < nbin; ++b) {
            need(8);
            const uint32_t bin = le32(&d[o]), nch = le32(&d[o + 4]); o += 8;
            need((size_t)nch * 16);
            std::vector<Chunk> &v = refs_[i].bins[bin];
            v.resize(nch);
            for (uint32_t c = 0; c < nch; ++c) { v[c].beg = le64(&d[o]); v[c].end = le64(&d[o + 8]); o += 16; }
        }
        need(4);
        const uint32_t nint = le32(&d[o]); o += 4;
        need((size_t)nint * 8);
        refs_[i].linear.resize(nint);
        for (uint32_t k = 0; k < nint; ++k) { refs_[i].linear[k] = le64(&d[o]); o += 8; }
    }
}

std::vector<Chunk> BamIndex::query(int tid, int32_t beg, int32_t end) const {
    std::vector<Chunk> out;
    if (tid < 0 || (size_t)tid >= refs_.size()) return out;
    const Ref &R = refs_[tid];
    if (beg < 0) beg = 0;
    if (end <= beg) return out;
    // smallest file offset of any record overlapping the 16 kb tile containing `beg` (linear index)
    uint64_t min_off = 0;
    if (!R.linear.empty()) {
        const size_t w = (size_t)beg >> 14;
        min_off = w < R.linear.size() ? R.linear[w] : R.linear.back();
        if (min_off == 0) {     // unset slot: fall back to the nearest earlier one
            for (size_t k = std::min(w, R.linear.size() - 1); k-- > 0;) if (R.linear[k]) { min_off = R.linear[k]; break; }
        }
    }
    // bins of the UCSC scheme overlapping [beg, end)
    const uint32_t b = (uint32_t)beg, e = (uint32_t)end - 1;
    auto add_bin = [&](uint32_t bin) {
        auto it = R.bins.find(bin);
        if (it == R.bins.end()) return;
        for (const Chunk &c : it->second) if (c.end > min_off) out.push_back(c);
    };
    add_bin(0);
    for (uint32_t k = 1 + (b >> 26); k <= 1 + (e >> 26); ++k) add_bin(k);
    for (uint32_t k = 9 + (b >> 23); k <= 9 + (e >> 23); ++k) add_bin(k);
    for (uint32_t k = 73 + (b >> 20); k <= 73 + (e >> 20); ++k) add_bin(k);
    for (uint32_t k = 585 + (b >> 17); k <= 585 + (e >> 17); ++k) add_bin(k);
    for (uint32_t k = 4681 + (b >> 14); k <= 4681 + (e >> 14); ++k) add_bin(k);
    std::sort(out.begin(), out.end(), [](const Chunk &x, const Chunk &y) { return x.beg < y.beg; });
    // merge overlapping / adjacent chunks
    std::vector<Chunk> m;
    for (const Chunk &c : out) {
        if (!m.empty() && c.beg <= m.back().end) m.back().end = std::max(m.back().end, c.end);
        else m.push_back(c);
    }
    return m;
}

// ------------------------------------------------------------------------------------------------ records
void Batch::clear() {
    pos.clear(); meta.clear(); cig_off.clear(); cigar.clear(); base_off.clear(); seq4.clear(); qual.clear();
}

namespace {
AllocFn g_alloc = nullptr;
FreeFn g_free = nullptr;
}  // namespace
void set_batch_allocator(AllocFn alloc, FreeFn release) { g_alloc = alloc; g_free = release; }
void *batch_alloc(size_t bytes) {
    void *p = nullptr;
    if (g_alloc) p = g_alloc(bytes);
    else if (bytes >= (size_t)4 << 20) {
        // a decoded piece is tens to hundreds of megabytes written once: with 4 KB pages a quarter of the decode time is
        // page faults, so ask for huge pages (transparent_hugepage=madvise is the usual setting)
        const size_t huge = (size_t)2 << 20;
        if (posix_memalign(&p, huge, (bytes + huge - 1) & ~(huge - 1)) != 0) p = nullptr;
        else madvise(p, (bytes + huge - 1) & ~(huge - 1), MADV_HUGEPAGE);
    } else p = malloc(bytes);
    if (!p) fail("out of host memory for read batches");
    return p;
}
void batch_free(void *p) { if (g_free) g_free(p); else free(p); }

namespace {
// reg2bin (bam.h:697-708): the smallest bin of the UCSC scheme containing [beg, end)
inline uint32_t region_bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (uint32_t)(4681 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint32_t)(585 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint32_t)(73 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint32_t)(9 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint32_t)(1 + (beg >> 26));
    return 0;
}
}  // namespace

int64_t build_bai(const BgzfFile &f, const std::string &bai_path) {
    const BamHeader h = read_header(f);
    const size_t nref = h.names.size();
    const uint32_t kMetaBin = 37450;
    struct RefIdx { std::map<uint32_t, std::vector<Chunk>> bins; std::vector<uint64_t> linear; };
    std::vector<RefIdx> refs(nref);
    BgzfReader rd(f);
    rd.seek(h.first_record_voffset);
    std::vector<uint8_t> rec;
    int64_t n_rec = 0;
    uint64_t n_no_coor = 0, n_mapped = 0, n_unmapped = 0;
    int32_t last_tid = -1, last_pos = -1;
    bool have_run = false;                 // a run = consecutive records of one (reference, bin)
    int32_t run_tid = -1; uint32_t run_bin = 0; uint64_t run_beg = 0;
    uint64_t ref_beg = rd.tell();          // first virtual offset of the current reference's records
    uint64_t last_off = rd.tell();         // virtual offset of the record being read
    auto close_ref = [&](int32_t tid, uint64_t off_end) {
        if (tid < 0) return;
        std::vector<Chunk> &m = refs[(size_t)tid].bins[kMetaBin];
        m.push_back(Chunk{ref_beg, off_end});
        m.push_back(Chunk{n_mapped, n_unmapped});
        n_mapped = n_unmapped = 0;
        ref_beg = off_end;
    };
    for (;;) {
        last_off = rd.tell();
        uint8_t b4[4];
        if (!rd.read(b4, 4)) break;
        const uint32_t bs = le32(b4);
        if (bs < 32) fail("corrupted BAM record");
        rec.resize(bs);
        if (!rd.read(rec.data(), bs)) fail("truncated BAM record");
        ++n_rec;
        const int32_t tid = (int32_t)le32(&rec[0]), pos = (int32_t)le32(&rec[4]);
        const uint32_t bmq = le32(&rec[8]), fnc = le32(&rec[12]);
        const uint32_t l_qname = bmq & 0xff, flag = fnc >> 16, n_cig = fnc & 0xffff;
        if (tid < 0) { ++n_no_coor; if (have_run) { refs[(size_t)run_tid].bins[run_bin].push_back(Chunk{run_beg, last_off}); close_ref(run_tid, last_off); have_run = false; } last_tid = tid; continue; }
        if ((size_t)tid >= nref) fail("BAM record on an unknown reference");
        if (n_no_coor) fail("the alignment is not sorted: reads without coordinates prior to reads with coordinates");
        if (tid < last_tid || (tid == last_tid && pos < last_pos)) fail("the alignment is not sorted");
        // reference end (bam_calend) for the bin and the linear index
        int64_t end = pos;
        const size_t o_cig = 32 + l_qname;
        if (o_cig + 4 * (size_t)n_cig > bs) fail("corrupted BAM record");
        for (uint32_t i = 0; i < n_cig; ++i) {
            const uint32_t c = le32(&rec[o_cig + 4 * i]);
            const uint32_t op = c & 15;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += c >> 4;
        }
        if (end <= pos) end = (int64_t)pos + 1;
        const bool unmapped = (flag & 0x4) != 0;
        if (!unmapped) {                   // insert_offset2 (bam_index.c:110-142): smallest offset per 16 kb tile
            std::vector<uint64_t> &lin = refs[(size_t)tid].linear;
            const size_t w0 = (size_t)pos >> 14, w1 = (size_t)(end - 1) >> 14;
            if (lin.size() < w1 + 1) lin.resize(w1 + 1, 0);
            for (size_t w = w0; w <= w1; ++w) if (lin[w] == 0) lin[w] = last_off;
        }
        const uint32_t bin = region_bin(pos, end);
        if (!have_run || tid != run_tid || bin != run_bin) {
            if (have_run) {
                refs[(size_t)run_tid].bins[run_bin].push_back(Chunk{run_beg, last_off});
                if (tid != run_tid) close_ref(run_tid, last_off);
            }
            have_run = true; run_tid = tid; run_bin = bin; run_beg = last_off;
        }
        if (unmapped) ++n_unmapped; else ++n_mapped;
        last_tid = tid; last_pos = pos;
    }
    if (have_run) { refs[(size_t)run_tid].bins[run_bin].push_back(Chunk{run_beg, last_off}); close_ref(run_tid, last_off); }
    // merge_chunks (bam_index.c:141-177): neighbours of one bin that meet inside a BGZF block; fill_missing (:179-191)
    for (RefIdx &R : refs) {
        for (auto &kv : R.bins) {
            if (kv.first == kMetaBin) continue;
            std::vector<Chunk> &v = kv.second;
            size_t m = 0;
            for (size_t l = 1; l < v.size(); ++l) {
                if (v[m].end >> 16 == v[l].beg >> 16) v[m].end = v[l].end;
                else v[++m] = v[l];
            }
            v.resize(v.empty() ? 0 : m + 1);
        }
        for (size_t j = 1; j < R.linear.size(); ++j) if (R.linear[j] == 0) R.linear[j] = R.linear[j - 1];
    }
    // bam_index_save (bam_index.c:322-373)
    std::string out("BAI\1", 4);
    auto put32 = [&](uint32_t v) { char b[4]; for (int i = 0; i < 4; ++i) b[i] = (char)(v >> (8 * i)); out.append(b, 4); };
    auto put64 = [&](uint64_t v) { char b[8]; for (int i = 0; i < 8; ++i) b[i] = (char)(v >> (8 * i)); out.append(b, 8); };
    put32((uint32_t)nref);
    for (const RefIdx &R : refs) {
        put32((uint32_t)R.bins.size());
        for (const auto &kv : R.bins) {
            put32(kv.first); put32((uint32_t)kv.second.size());
            for (const Chunk &c : kv.second) { put64(c.beg); put64(c.end); }
        }
        put32((uint32_t)R.linear.size());
        for (uint64_t v : R.linear) put64(v);
    }
    put64(n_no_coor);
    std::ofstream of(bai_path, std::ios::binary);
    if (!of) fail("cannot write " + bai_path);
    of.write(out.data(), (std::streamsize)out.size());
    if (!of) fail("cannot write " + bai_path);
    return n_rec;
}

int64_t fetch_region(const BgzfFile &f, const BamIndex &idx, const SampleTable &st, int tid, int32_t beg, int32_t end, Batch &out) {
    return fetch_piece(f, idx, st, tid, beg, end, beg, end, out);
}

int64_t fetch_piece(const BgzfFile &f, const BamIndex &idx, const SampleTable &st, int tid, int32_t beg, int32_t end, int32_t lo, int32_t hi,
                    Batch &out) {
    const bool first = lo <= beg;
    if (hi > end) hi = end;
    if (lo >= hi) { if (out.cig_off.empty()) { out.cig_off.push_back(0); out.base_off.push_back(0); } return 0; }
    const std::vector<Chunk> chunks = idx.query(tid, first ? beg : lo, hi);
    if (out.cig_off.empty()) { out.cig_off.push_back(0); out.base_off.push_back(0); }
    BgzfReader rd(f);
    std::vector<uint8_t> rec;
    int64_t delivered = 0;
    std::string last_rg;
    int last_sample = -1;
    const bool file_sample = st.rg2sample.empty();       // header without @RG: a read WITH an RG tag falls back to the file's sample (popbam.cpp:231-232)
    for (const Chunk &ch : chunks) {
        rd.seek(ch.beg);
        while (rd.tell() < ch.end) {
            uint8_t b4[4];
            if (!rd.read(b4, 4)) return delivered;
            const uint32_t bs = le32(b4);
            if (bs < 32) fail("corrupted BAM record");
            rec.resize(bs);
            if (!rd.read(rec.data(), bs)) fail("truncated BAM record");
            const int32_t rtid = (int32_t)le32(&rec[0]), pos = (int32_t)le32(&rec[4]);
            if (rtid != tid || pos >= hi) return delivered;            // bam_iter_read: no need to proceed
            if (!first && pos < lo) continue;                          // an earlier piece's record
            const uint32_t bmq = le32(&rec[8]), fnc = le32(&rec[12]);
            const uint32_t l_qname = bmq & 0xff, mapq = (bmq >> 8) & 0xff, flag = fnc >> 16, n_cig = fnc & 0xffff;
            const int32_t l_seq = (int32_t)le32(&rec[16]);
            const size_t o_cig = 32 + l_qname, o_seq = o_cig + 4 * (size_t)n_cig, o_qual = o_seq + (size_t)(l_seq + 1) / 2,
                         o_aux = o_qual + (size_t)l_seq;
            if (l_seq < 0 || o_aux > bs) fail("corrupted BAM record");
            // is_overlap (bam_index.c:729-735)
            int64_t rend = pos;
            for (uint32_t i = 0; i < n_cig; ++i) {
                const uint32_t c = le32(&rec[o_cig + 4 * i]);
                const uint32_t op = c & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rend += c >> 4;
            }
            if (n_cig == 0) rend = (int64_t)pos + 1;
            if (!(rend > beg && pos < end)) continue;
            // RG:Z tag -> sample (call_base, popbam.cpp:224-240)
            int sample = 0xff;                                          // no RG tag: call_base skips the read (popbam.cpp:226-228)
            {
                size_t a = o_aux;
                const char *rg = nullptr;
                while (a + 3 <= bs) {
                    const uint8_t t0 = rec[a], t1 = rec[a + 1], ty = rec[a + 2];
                    a += 3;
                    if (ty == 'Z' || ty == 'H') {
                        const char *z = reinterpret_cast<const char *>(&rec[a]);
                        const size_t len = strnlen(z, bs - a);
                        if (t0 == 'R' && t1 == 'G' && ty == 'Z') { rg = z; break; }
                        a += len + 1;
                    } else if (ty == 'A' || ty == 'c' || ty == 'C') a += 1;
                    else if (ty == 's' || ty == 'S') a += 2;
                    else if (ty == 'i' || ty == 'I' || ty == 'f') a += 4;
                    else if (ty == 'd') a += 8;
                    else if (ty == 'B') {
                        if (a + 5 > bs) break;
                        const uint8_t sub = rec[a];
                        const uint32_t cnt = le32(&rec[a + 1]);
                        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                        a += 5 + es * cnt;
                    } else break;
                }
                if (rg) {
                    if (last_sample >= 0 && last_rg == rg) sample = last_sample;
                    else {
                        auto it = st.rg2sample.find(rg);
                        if (it == st.rg2sample.end() && !file_sample)
                            fail(std::string("Problem assigning read group ") + rg +
                                 " to a sample.\nPlease check BAM header for correct SM and PO tags");
                        sample = file_sample ? 0 : it->second; last_rg = rg; last_sample = sample;
                    }
                }
            }
            out.pos.push_back(pos);
            out.meta.push_back(flag << 16 | mapq << 8 | (uint32_t)sample);
            { uint32_t *cg = out.cigar.grow(n_cig); for (uint32_t i = 0; i < n_cig; ++i) cg[i] = le32(&rec[o_cig + 4 * i]); }
            out.cig_off.push_back((uint32_t)out.cigar.size());
            const size_t pad = ((size_t)l_seq + 3) & ~(size_t)3;      // reads start 4-byte aligned
            if (out.qual.size() + pad > 0xfffffff0u) fail("a single fetch holds more than 4 GiB of bases: use smaller pieces");
            uint8_t *qd = out.qual.grow(pad);
            memcpy(qd, &rec[o_qual], (size_t)l_seq);
            memset(qd + l_seq, 0, pad - (size_t)l_seq);
            uint8_t *sd = out.seq4.grow(pad / 2);
            memcpy(sd, &rec[o_seq], (size_t)(l_seq + 1) / 2);
            memset(sd + (l_seq + 1) / 2, 0, pad / 2 - (size_t)(l_seq + 1) / 2);
            out.base_off.push_back((uint32_t)out.qual.size());
            ++delivered;
        }
    }
    return delivered;
}

// ------------------------------------------------------------------------------------------------ FASTA
namespace {
struct FaiEntry { int64_t len, offset; int line_blen, line_len; };

bool load_fai(const std::string &path, std::unordered_map<std::string, FaiEntry> &m) {
    std::ifstream in(path);
    if (!in) return false;
    std::string name;
    FaiEntry e;
    while (in >> name >> e.len >> e.offset >> e.line_blen >> e.line_len) m[name] = e;
    return true;
}

void build_fai(const std::string &fa) {
    // one pass over the FASTA (faidx.c:68-176): name, length, offset of the first base, bases per line, bytes per line
    std::ifstream in(fa, std::ios::binary);
    if (!in) fail("Failed to load index for fastA reference file: " + fa);
    std::ofstream out(fa + ".fai");
    std::string line, name;
    int64_t off = 0, len = 0, seq_off = 0;
    int blen = 0, llen = 0;
    bool have = false;
    auto flush = [&]() { if (have) out << name << '\t' << len << '\t' << seq_off << '\t' << blen << '\t' << llen << '\n'; };
    while (std::getline(in, line)) {
        const int64_t raw = (int64_t)line.size() + 1;
        if (!line.empty() && line[0] == '>') {
            flush();
            size_t e = 1;
            while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
            name = line.substr(1, e - 1);
            have = true; len = 0; blen = 0; llen = 0; seq_off = off + raw;
        } else if (have) {
            std::string t = line;
            while (!t.empty() && (t.back() == '\r')) t.pop_back();
            if (blen == 0) { blen = (int)t.size(); llen = (int)raw; }
            len += (int64_t)t.size();
        }
        off += raw;
    }
    flush();
}
}  // namespace

std::string fetch_contig(const std::string &fa, const std::string &name) {
    std::unordered_map<std::string, FaiEntry> idx;
    if (!load_fai(fa + ".fai", idx)) {
        build_fai(fa);
        if (!load_fai(fa + ".fai", idx)) fail("Failed to load index for fastA reference file: " + fa);
    }
    auto it = idx.find(name);
    if (it == idx.end()) fail("Failed to find reference sequence " + name + " in " + fa);
    const FaiEntry &e = it->second;
    std::ifstream in(fa, std::ios::binary);
    if (!in) fail("Failed to open fastA reference file: " + fa);
    std::string seq;
    seq.reserve((size_t)e.len);
    const int64_t n_lines = e.line_blen > 0 ? (e.len + e.line_blen - 1) / e.line_blen : 0;
    std::string buf((size_t)(n_lines * e.line_len + 16), '\0');
    in.seekg(e.offset);
    in.read(&buf[0], (std::streamsize)buf.size());
    const size_t got = (size_t)in.gcount();
    for (size_t i = 0; i < got && (int64_t)seq.size() < e.len; ++i)
        if (isgraph((unsigned char)buf[i])) seq.push_back(buf[i]);
    return seq;
}

bool parse_region(const BamHeader &h, const std::string &region_in, int *tid, int32_t *beg, int32_t *end) {
    std::string region;
    for (char c : region_in) if (c != ' ' && c != ',') region.push_back(c);
    *tid = *beg = *end = -1;
    const size_t l = region.size();
    size_t name_end = region.find(':');
    if (name_end == std::string::npos) name_end = l;
    int id = -1;
    if (name_end < l) {
        const std::string coords = region.substr(name_end + 1);
        const size_t n_hyphen = (size_t)std::count(coords.begin(), coords.end(), '-');
        if (coords.find_first_not_of("0123456789,-") != std::string::npos || n_hyphen > 1) name_end = l;
        id = h.tid_of(region.substr(0, name_end));
        if (id < 0) {
            id = h.tid_of(region);
            if (id < 0) { fprintf(stderr, "Cannot find sequence name %s in header\n", region.c_str()); return false; }
        }
    } else id = h.tid_of(region);
    if (id < 0) return false;
    *tid = id;
    if (name_end < l) {
        const std::string coords = region.substr(name_end + 1);
        const size_t dash = coords.find('-');
        *beg = atoi(coords.substr(0, dash).c_str());
        if (*beg > 0) --*beg;
        // no '-' : substr(npos + 1) == substr(0) == the whole string, as in the reference
        *end = atoi((dash == std::string::npos ? coords : coords.substr(dash + 1)).c_str());
    } else { *beg = 0; *end = h.lens[id]; }
    return *beg <= *end;
}

}  // namespace pbio
