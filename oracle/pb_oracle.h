/*
 * pb_oracle.h -- CPU restatement of POPBAM 0.3's per-window statistics path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library never does.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/popbam (the
 * unmodified reference compiled by oracle/Makefile) is run on seeded synthetic BAMs and its
 * stdout must equal pbo_format_window()'s text byte for byte for every subcommand/option
 * set (tests/test_oracle_pin.py; committed goldens under tests/golden/), and
 * oracle/_ref/refdump dumps errmod tables / errmod_cal / gl2cns / segbase known-answer
 * vectors from the reference's own object files (tests/golden/kat_*.bin).
 */
#ifndef PB_ORACLE_H
#define PB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PBO_MAX_SAMPLES 64

/* analyses / flags: numerically identical to include/popbam_b200.h */
#define PBO_FLAG_ILLUMINA     0x02u
#define PBO_FLAG_SUBSTITUTE   0x10u
#define PBO_FLAG_HETEROZYGOTE 0x20u
#define PBO_FLAG_OUTGROUP     0x40u
#define PBO_FLAG_EMIT_CB      0x10000u

#define PBO_AN_NUCDIV        0x001u
#define PBO_AN_SFS           0x002u
#define PBO_AN_LD_ZNS        0x004u
#define PBO_AN_LD_OMEGA      0x008u
#define PBO_AN_LD_WALL       0x010u
#define PBO_AN_DIVERGE_IND   0x020u
#define PBO_AN_DIVERGE_POP   0x040u
#define PBO_AN_HAPLO_K       0x080u
#define PBO_AN_HAPLO_EHHS    0x100u
#define PBO_AN_HAPLO_DXY     0x200u
#define PBO_AN_SNP           0x400u
#define PBO_AN_TREE          0x800u  /* pop_tree.cpp: difference matrix incl. the reference taxon */

typedef struct pbo_params {
    int32_t  n_samples, n_pops;
    uint64_t pop_mask[PBO_MAX_SAMPLES];
    uint8_t  pop_nsmpl[PBO_MAX_SAMPLES];
    int32_t  min_depth, max_depth, min_rmsQ, min_snpQ, min_mapQ, min_baseQ;
    uint32_t flags;
    int32_t  outidx;
    int32_t  min_freq;
    int32_t  device;      /* unused; keeps the layout equal to pb_params */
} pbo_params;

typedef struct pbo_batch {
    int64_t n_reads, n_cigar, n_bases;
    const int32_t  *pos;
    const uint32_t *meta;
    const uint32_t *cig_off;
    const uint32_t *cigar;
    const uint32_t *base_off;
    const uint8_t  *seq4;
    const uint8_t  *qual;
} pbo_batch;

typedef struct pbo_result {
    int32_t n_windows, n_pops, n_samples;
    uint32_t analyses;
    int32_t *win_beg, *win_end, *num_sites, *segsites;
    int64_t *seg_off;
    uint32_t *seg_pos, *seg_idx;
    uint64_t *seg_type;
    uint8_t  *seg_ref;
    uint64_t *seg_cb;
    double *piw, *pib; uint16_t *min_dxy;
    int32_t *sfs_num_snps; double *td, *fwh;
    int32_t *ld_num_snps; double *zns, *omegamax;
    int32_t *wall_num_snps; double *wallb, *wallq;
    uint16_t *ind_div, *pop_div; int32_t *div_num_snps;
    int32_t *nhaps; double *hdiv, *ehhs;
    int32_t span_beg, span_end;
    uint64_t *cb, *site_type; uint8_t *site_flag;
    int64_t reads_pushed, reads_used, aligned_bases;
    uint16_t *tree_diff;          /* [NW][(n+1)*(n+1)] treeData::diff_matrix, taxon 0 = the reference */
} pbo_result;

typedef struct pbo_print_opts {
    const char *chrom;
    const char *const *pop_names;
    const char *const *sample_names;
    int32_t min_sites, min_snps, jc, snp_output;
    const char *ref_name;         /* treeData::refid: the AS tag of the sequence dictionary */
} pbo_print_opts;

/* errmod_init(1.0-0.83) -> cal_coef (pop_utils.cpp:203-266).  Buffers: 256, 64*256*256, 256*256 */
int  pbo_build_tables(double *fk, double *beta, double *lhet);

/* errmod_cal + gl2cns + rms packing for one (site,sample) cell (pop_utils.cpp:280-365, :66-100;
 * popbam.cpp:288-298).  codes[] is modified (sorted).  rmsq = sum of mapq^2.  Returns the cb word;
 * q16 (optional) receives the 16 likelihoods.                                                  */
uint64_t pbo_call_cell(const double *fk, const double *beta, const double *lhet,
                       uint16_t *codes, int k, int rmsq, float *q16);

/* clean_heterozygotes + segbase + qfilter + cal_site_type on the n cb words of one site
 * (pop_utils.cpp:170-201, :122-168, :102-120; popbam.cpp:173-184).  Returns fq; writes coverage
 * and site type.                                                                              */
int pbo_site_logic(const pbo_params *p, uint64_t *cb, char ref, uint64_t *cov, uint64_t *type);

/* Whole region: every window's pileup -> calls -> stats.  Returns 0 or a negative code.       */
int  pbo_run_region(const pbo_params *p, const double *fk, const double *beta, const double *lhet,
                    const pbo_batch *b, const char *ref, int64_t ref_len, uint32_t analyses,
                    int32_t n_windows, const int32_t *win_beg, const int32_t *win_end,
                    pbo_result *out);
void pbo_free_result(pbo_result *r);

int64_t pbo_window_grid(int32_t beg, int32_t end, int32_t win_size, int64_t cap,
                        int32_t *win_beg, int32_t *win_end);

int64_t pbo_format_window(const pbo_params *p, const pbo_result *r, int32_t w, uint32_t analysis,
                          const pbo_print_opts *o, char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
