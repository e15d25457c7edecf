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
#include <cstdint>
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
// Growable structure-of-arrays batch in the pb_read_batch layout (offsets 4-byte aligned per read).
struct Batch {
    std::vector<int32_t> pos;
    std::vector<uint32_t> meta, cig_off, cigar, base_off;
    std::vector<uint8_t> seq4, qual;
    void clear();
    int64_t n_reads() const { return (int64_t)pos.size(); }
};

// bam_fetch over [beg, end) of tid (bam_index.c:943-957): every record overlapping the interval, in file
// order, appended to `out`.  Reads without an RG tag get sample PB_NO_SAMPLE; an RG that is not in the
// header raises the reference's "Problem assigning read group" error.  Returns records delivered.
int64_t fetch_region(const BgzfFile &f, const BamIndex &idx, const SampleTable &st, int tid, int32_t beg, int32_t end, Batch &out);

// ---- FASTA ------------------------------------------------------------------------------------------
// Whole contig, bytes verbatim (case preserved, faidx.c:433-468); builds <fa>.fai next to the FASTA if it
// is missing, as fai_load does (faidx.c:289-293).
std::string fetch_contig(const std::string &fasta_path, const std::string &name);

// "chr[:beg[-end]]" -> tid, 0-based beg, end (bam_parse_region, pop_utils.cpp:386-461)
bool parse_region(const BamHeader &h, const std::string &region, int *tid, int32_t *beg, int32_t *end);

}  // namespace pbio
