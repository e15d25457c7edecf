// pb_bamio.h -- host-side feeder: BGZF inflate, BAM header / record decode, BAI region slicing,
// FASTA (+.fai) contig fetch, @RG -> sample / population tables.
//
// Written from the on-disk formats as the reference reads them:
//   BGZF   bgzf.c:47-61, :366-399, :479-536, :715-747   (gzip members <= 64 KiB, virtual offsets)
//   BAM    bam.c:119-170 (header), :283-331 (bam_read1), bam.h:178-267 (record layout)
//   BAI    bam_index.c:447-528 (load), :704-727 (reg2bins), :751-861 (query), :729-735 (is_overlap)
//   FASTA  faidx.c:178-230 (.fai), :433-468 (fetch)
//   @RG    pop_sample.cpp:15-107 (bam_smpl_add), :152-226, popbam.cpp:145-171 (assign_pops)
// Everything here runs on host threads and only produces pb_read_batch arrays for the C ABI.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace pbio {

struct Error {
    std::string msg;
};

// ---- BGZF ---------------------------------------------------------------------------------------
class BgzfFile {
public:
    ~BgzfFile();
    void open(const std::string &path);                         // mmap; throws Error
    const uint8_t *data() const { return map_; }
    uint64_t size() const { return size_; }
    // Inflate the block at compressed offset `coff` into out (resized to ISIZE); returns the block's
    // compressed size (0 at end of file).  Thread-safe (no shared state).
    uint32_t inflate_block(uint64_t coff, std::vector<uint8_t> &out) const;

private:
    int fd_ = -1;
    const uint8_t *map_ = nullptr;
    uint64_t size_ = 0;
};

// Sequential reader over the uncompressed stream starting at a virtual offset.
class BgzfReader {
public:
    explicit BgzfReader(const BgzfFile &f) : f_(f) {}
    void seek(uint64_t voffset);
    uint64_t tell() const { return (coff_ << 16) | uoff_; }     // virtual offset of the next byte
    bool read(void *dst, size_t n);                             // false at end of file
    bool eof();

private:
    bool load();
    const BgzfFile &f_;
    std::vector<uint8_t> buf_;
    uint64_t coff_ = 0;      // compressed offset of the block in buf_
    uint32_t csize_ = 0;     // its compressed size
    uint32_t uoff_ = 0;
    bool loaded_ = false;
};

// ---- BAM header / samples ---------------------------------------------------------------------------
struct BamHeader {
    std::string text;
    std::vector<std::string> names;
    std::vector<int32_t> lens;
    uint64_t first_record_voffset = 0;
    int tid_of(const std::string &name) const;
};
BamHeader read_header(const BgzfFile &f);

struct SampleTable {
    std::vector<std::string> samples;       // order of first SM
    std::vector<std::string> pops;          // order of first PO
    std::vector<int> sample_pop;            // population of each sample
    std::unordered_map<std::string, int> rg2sample;
    uint64_t pop_mask[64] = {0};
    uint8_t pop_nsmpl[64] = {0};
};
// bam_smpl_add + assign_pops.  Throws Error with the reference's message when a sample has no population.
SampleTable build_samples(const std::string &header_text, const std::string &bam_path);

// ---- BAI --------------------------------------------------------------------------------------------
struct Chunk {
    uint64_t beg, end;      // virtual offsets
};
class BamIndex {
public:
    void load(const std::string &bai_path, size_t n_ref_expected);
    // merged chunk list whose records may overlap [beg, end) on reference tid
    std::vector<Chunk> query(int tid, int32_t beg, int32_t end) const;

private:
    struct Ref {
        std::unordered_map<uint32_t, std::vector<Chunk>> bins;
        std::vector<uint64_t> linear;
    };
    std::vector<Ref> refs_;
};

// bam_index_build / bam_index_core / bam_index_save (bam_index.c:193-320, :641-700, :322-373): writes <bam>.bai for a
// coordinate-sorted BAM -- UCSC bins with chunks of consecutive records (chunks meeting inside one BGZF block merged),
// the 16 kb linear index with gaps filled from the left, the per-reference pseudo-bin 37450 (file range, mapped /
// unmapped counts) and the count of records without coordinates.  Throws Error on unsorted input.  Returns the number
// of records seen.
int64_t build_bai(const BgzfFile &f, const std::string &bai_path);

// ---- records -> batch -------------------------------------------------------------------------------
// Where batch arrays live.  The command line points these at pb_host_alloc / pb_host_free of the C ABI (page-locked
// host memory: the device reads it with asynchronous copies); the default is malloc / free (tests without a GPU).
typedef void *(*AllocFn)(size_t);
typedef void (*FreeFn)(void *);
void set_batch_allocator(AllocFn alloc, FreeFn release);

// Minimal growable array on the batch allocator (grow-only capacity: a Batch is reused for many chunks).
template <class T>
class Vec {
public:
    Vec() = default;
    Vec(const Vec &) = delete;
    Vec &operator=(const Vec &) = delete;
    Vec(Vec &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
    Vec &operator=(Vec &&o) noexcept { if (this != &o) { release(); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = nullptr; o.n_ = o.cap_ = 0; } return *this; }
    ~Vec() { release(); }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T &back() { return p_[n_ - 1]; }
    void clear() { n_ = 0; }
    void reserve(size_t c);
    void push_back(const T &v) { if (n_ == cap_) reserve(cap_ ? cap_ * 2 : 1024); p_[n_++] = v; }
    T *grow(size_t k) { if (n_ + k > cap_) reserve(std::max(n_ + k, cap_ * 2)); T *r = p_ + n_; n_ += k; return r; }   // k uninitialised elements
private:
    void release();
    T *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

// Structure-of-arrays batch in the pb_read_batch layout (offsets 4-byte aligned per read).
struct Batch {
    Vec<int32_t> pos;
    Vec<uint32_t> meta, cig_off, cigar, base_off;
    Vec<uint8_t> seq4, qual;
    void clear();
    int64_t n_reads() const { return (int64_t)pos.size(); }
};

// bam_fetch over [beg, end) of tid (bam_index.c:943-957): every record overlapping the interval, in file
// order, appended to `out`.  Reads without an RG tag get sample PB_NO_SAMPLE (call_base skips them, popbam.cpp:227);
// an RG that is not in the header raises the reference's "Problem assigning read group" error.  Returns records delivered.
int64_t fetch_region(const BgzfFile &f, const BamIndex &idx, const SampleTable &st, int tid, int32_t beg, int32_t end, Batch &out);
// One PIECE of such a fetch, so that several threads can decode one region: of the records bam_fetch would deliver for
// [beg, end) exactly those whose start lies in [lo, hi) -- and, for the first piece (lo == beg), also the ones that start
// before beg.  The pieces of a partition of [beg, end), concatenated in order, equal fetch_region's batch.
int64_t fetch_piece(const BgzfFile &f, const BamIndex &idx, const SampleTable &st, int tid, int32_t beg, int32_t end, int32_t lo, int32_t hi,
                    Batch &out);

// ---- FASTA ------------------------------------------------------------------------------------------
// Whole contig, bytes verbatim (case preserved, faidx.c:433-468); builds <fa>.fai next to the FASTA if it
// is missing, as fai_load does (faidx.c:289-293).
std::string fetch_contig(const std::string &fasta_path, const std::string &name);

// "chr[:beg[-end]]" -> tid, 0-based beg, end (bam_parse_region, pop_utils.cpp:386-461)
bool parse_region(const BamHeader &h, const std::string &region, int *tid, int32_t *beg, int32_t *end);

void *batch_alloc(size_t bytes);
void batch_free(void *p);
template <class T> void Vec<T>::reserve(size_t c) {
    if (c <= cap_) return;
    T *np = static_cast<T *>(batch_alloc(c * sizeof(T)));
    if (n_) memcpy(np, p_, n_ * sizeof(T));
    if (p_) batch_free(p_);
    p_ = np; cap_ = c;
}
template <class T> void Vec<T>::release() { if (p_) batch_free(p_); p_ = nullptr; n_ = cap_ = 0; }

}  // namespace pbio
