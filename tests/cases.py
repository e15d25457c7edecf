"""Parity case table shared by the golden generator, the oracle pin tests and the GPU parity tests.

Each case names a seeded synthetic fixture, the reference command line that produced the committed golden text
(tests/golden/<name>.txt, made by tests/golden/make_goldens.py with the UNMODIFIED reference binary), and the
equivalent analysis / parameter / printer settings of the C ABI (include/popbam_b200.h).
"""
from pbtest import AN, FLAG

# fixtures (pbsynth parameters); kept small so that the reference needs ~1-2 s per command
FIXTURES = {
    # SURVEY §8(d) C1-shaped, cut down: 10 ingroup + outgroup, 20x, 3 windows of 10 kb
    "c1": dict(contig_len=30500, n_ingroup=10, has_outgroup=1, depth=20.0, snp_density=0.01, seed=11),
    # edge cases: flagged reads, N bases, =/X/H/P/N cigar ops, lower-case / N reference, low-coverage hole
    "edge": dict(contig_len=20500, n_ingroup=6, has_outgroup=1, depth=14.0, snp_density=0.02, het_frac=0.3,
                 edge_mode=1, seed=12),
    # two read groups per sample, no outgroup, odd read length (seq4 padding)
    "rg2": dict(contig_len=20500, n_ingroup=5, has_outgroup=0, rg_per_sample=2, depth=9.0, read_len=75,
                snp_density=0.015, seed=13),
    # dense SNPs, more samples: LD stress (C3-shaped, cut down)
    "ld": dict(contig_len=20500, n_ingroup=23, has_outgroup=1, depth=12.0, snp_density=0.04, seed=14),
    # the 64-sample limit (popbam.1:508): sample masks use bit 63, the sample partition uses both register banks
    "n64": dict(contig_len=10500, n_ingroup=63, has_outgroup=1, depth=18.0, snp_density=0.03, het_frac=0.2, seed=15),
    # C3-shaped at a size the reference's O(S^3) omega_max still finishes in seconds: 64 samples, 20 kb windows with ~1000
    # segregating sites -- several 256-SNP row blocks and partner tiles per population in k_ld_rows
    "ldbig": dict(contig_len=40500, n_ingroup=63, has_outgroup=1, depth=24.0, snp_density=0.12, seed=16),
    # a realistic quality spectrum: base qualities 2..41 (40 values), 20 mapping qualities (0..60), some below min_rmsQ
    "wideq": dict(contig_len=20500, n_ingroup=9, has_outgroup=1, depth=30.0, snp_density=0.02, het_frac=0.3, edge_mode=2, seed=17),
}


def _b(v):
    """-a/-b are parsed into unsigned char: pass the byte whose value is the threshold."""
    return chr(v)


# name, fixture, reference argv (before "-f ref.fa ... bam region"), analysis, params kw, print kw
CASES = [
    ("nucdiv_c1", "c1", ["nucdiv", "-w", "10"], "NUCDIV", {}, {}),
    ("sfs_c1_og", "c1", ["sfs", "-w", "10", "-p", "og"], "SFS", dict(flags=FLAG["OUTGROUP"], outidx=10), {}),
    ("sfs_c1", "c1", ["sfs", "-w", "10"], "SFS", {}, {}),
    ("ld0_c1", "c1", ["ld", "-w", "10", "-o", "0"], "LD_ZNS", {}, {}),
    ("ld1_c1", "c1", ["ld", "-w", "10", "-o", "1"], "LD_OMEGA", {}, {}),
    ("ld2_c1", "c1", ["ld", "-w", "10", "-o", "2"], "LD_WALL", {}, {}),
    ("ld0e_c1", "c1", ["ld", "-w", "10", "-o", "0", "-e"], "LD_ZNS", dict(min_freq=2), {}),
    ("div0_c1", "c1", ["diverge", "-w", "10", "-o", "0"], "DIVERGE_IND", {}, {}),
    ("div0jc_c1", "c1", ["diverge", "-w", "10", "-o", "0", "-d", "jc"], "DIVERGE_IND", {}, dict(jc=1)),
    ("div1_c1_og", "c1", ["diverge", "-w", "10", "-o", "1", "-p", "og"], "DIVERGE_POP",
     dict(flags=FLAG["OUTGROUP"], outidx=10), {}),
    ("div1t_c1", "c1", ["diverge", "-w", "10", "-o", "1", "-t"], "DIVERGE_POP", dict(flags=FLAG["SUBSTITUTE"]), {}),
    ("hap0_c1", "c1", ["haplo", "-w", "10", "-o", "0"], "HAPLO_K", {}, {}),
    ("hap1_c1", "c1", ["haplo", "-w", "10", "-o", "1"], "HAPLO_EHHS", {}, {}),
    ("hap2_c1", "c1", ["haplo", "-w", "10", "-o", "2"], "HAPLO_DXY", {}, {}),
    ("snp0_c1", "c1", ["snp", "-o", "0"], "SNP", {}, dict(snp_output=0)),
    ("snp1_c1_og", "c1", ["snp", "-o", "1", "-p", "og"], "SNP", dict(flags=FLAG["OUTGROUP"], outidx=10),
     dict(snp_output=1)),
    ("snp2_c1", "c1", ["snp", "-o", "2", "-w", "10"], "SNP", {}, dict(snp_output=2)),
    # thresholds / option semantics (SURVEY Appendix D "option semantics")
    ("snp0_c1_x12", "c1", ["snp", "-o", "0", "-x", "12"], "SNP", dict(max_depth=12), dict(snp_output=0)),
    ("snp0_c1_ab", "c1", ["snp", "-o", "0", "-a", _b(30), "-b", _b(21)], "SNP", dict(min_mapQ=30, min_baseQ=21),
     dict(snp_output=0)),
    ("snp0_c1_qs", "c1", ["snp", "-o", "0", "-q", "40", "-s", "15", "-m", "5"], "SNP",
     dict(min_rmsQ=40, min_snpQ=15, min_depth=5), dict(snp_output=0)),
    ("snp0_c1_z", "c1", ["snp", "-o", "0", "-z"], "SNP", dict(flags=FLAG["HETEROZYGOTE"]), dict(snp_output=0)),
    ("snp0_c1_i", "c1", ["snp", "-o", "0", "-i", "-b", _b(4)], "SNP", dict(flags=FLAG["ILLUMINA"], min_baseQ=4),
     dict(snp_output=0)),
    ("nucdiv_c1_k", "c1", ["nucdiv", "-w", "10", "-k", "9900"], "NUCDIV", {}, dict(min_sites=9900)),
    ("nucdiv_c1_now", "c1", ["nucdiv"], "NUCDIV", {}, {}),
    # edge fixture
    ("snp0_edge", "edge", ["snp", "-o", "0"], "SNP", {}, dict(snp_output=0)),
    ("nucdiv_edge", "edge", ["nucdiv", "-w", "10"], "NUCDIV", {}, {}),
    ("sfs_edge_og", "edge", ["sfs", "-w", "10", "-p", "og"], "SFS", dict(flags=FLAG["OUTGROUP"], outidx=6), {}),
    ("hap0_edge", "edge", ["haplo", "-w", "10", "-o", "0"], "HAPLO_K", {}, {}),
    ("hap1_edge", "edge", ["haplo", "-w", "10", "-o", "1"], "HAPLO_EHHS", {}, {}),
    ("snp0_edge_x9", "edge", ["snp", "-o", "0", "-x", "9"], "SNP", dict(max_depth=9), dict(snp_output=0)),
    # two read groups per sample
    ("snp0_rg2", "rg2", ["snp", "-o", "0"], "SNP", {}, dict(snp_output=0)),
    ("nucdiv_rg2", "rg2", ["nucdiv", "-w", "10"], "NUCDIV", {}, {}),
    ("ld2_rg2", "rg2", ["ld", "-w", "10", "-o", "2"], "LD_WALL", {}, {}),
    # LD fixture
    ("ld0_ld", "ld", ["ld", "-w", "10", "-o", "0"], "LD_ZNS", {}, {}),
    ("ld1_ld", "ld", ["ld", "-w", "10", "-o", "1"], "LD_OMEGA", {}, {}),
    ("ld2_ld", "ld", ["ld", "-w", "10", "-o", "2"], "LD_WALL", {}, {}),
    ("hap1_ld", "ld", ["haplo", "-w", "10", "-o", "1"], "HAPLO_EHHS", {}, {}),
    ("sfs_ld_og", "ld", ["sfs", "-w", "10", "-p", "og"], "SFS", dict(flags=FLAG["OUTGROUP"], outidx=23), {}),
    # tree (neighbour joining on the difference matrix incl. the reference taxon)
    ("tree_c1", "c1", ["tree", "-w", "10"], "TREE", {}, {}),
    ("tree_c1_jc", "c1", ["tree", "-w", "10", "-d", "jc"], "TREE", {}, dict(jc=1)),
    ("tree_c1_now", "c1", ["tree"], "TREE", {}, {}),
    ("tree_c1_k", "c1", ["tree", "-w", "10", "-k", "9900"], "TREE", {}, dict(min_sites=9900)),
    ("tree_edge", "edge", ["tree", "-w", "10"], "TREE", {}, {}),
    ("tree_rg2", "rg2", ["tree", "-w", "10", "-d", "jc"], "TREE", {}, dict(jc=1)),
    ("tree_ld", "ld", ["tree", "-w", "10"], "TREE", {}, {}),
    ("tree_n64", "n64", ["tree", "-w", "5"], "TREE", {}, {}),
    # 64 samples
    ("nucdiv_n64", "n64", ["nucdiv", "-w", "5"], "NUCDIV", {}, {}),
    ("ld0_n64", "n64", ["ld", "-w", "5", "-o", "0"], "LD_ZNS", {}, {}),
    ("ld1_n64", "n64", ["ld", "-w", "5", "-o", "1"], "LD_OMEGA", {}, {}),
    ("sfs_n64_og", "n64", ["sfs", "-w", "5", "-p", "og"], "SFS", dict(flags=FLAG["OUTGROUP"], outidx=63), {}),
    ("hap0_n64", "n64", ["haplo", "-w", "5", "-o", "0"], "HAPLO_K", {}, {}),
    ("div1_n64_og", "n64", ["diverge", "-w", "5", "-o", "1", "-p", "og"], "DIVERGE_POP", dict(flags=FLAG["OUTGROUP"], outidx=63), {}),
    ("snp1_n64", "n64", ["snp", "-o", "1"], "SNP", {}, dict(snp_output=1)),
    # pairwise LD with ~1000 segregating sites per window (pop_ld.cpp:201-373)
    ("ld0_ldbig", "ldbig", ["ld", "-w", "20", "-o", "0"], "LD_ZNS", {}, {}),
    ("ld1_ldbig", "ldbig", ["ld", "-w", "20", "-o", "1"], "LD_OMEGA", {}, {}),
    ("ld2_ldbig", "ldbig", ["ld", "-w", "20", "-o", "2"], "LD_WALL", {}, {}),
    # wide quality spectrum
    ("snp0_wideq", "wideq", ["snp", "-o", "0"], "SNP", {}, dict(snp_output=0)),
    ("nucdiv_wideq", "wideq", ["nucdiv", "-w", "10"], "NUCDIV", {}, {}),
    ("sfs_wideq_og", "wideq", ["sfs", "-w", "10", "-p", "og"], "SFS", dict(flags=FLAG["OUTGROUP"], outidx=9), {}),
    ("ld0_wideq", "wideq", ["ld", "-w", "10", "-o", "0"], "LD_ZNS", {}, {}),
    ("snp0_wideq_q", "wideq", ["snp", "-o", "0", "-q", "12", "-a", _b(5), "-b", _b(3)], "SNP", dict(min_rmsQ=12, min_mapQ=5, min_baseQ=3), dict(snp_output=0)),
]


def win_kb(argv):
    """Window size in bp from a reference argv (-w is in kb, pop_nucdiv.cpp:321); 0 when absent."""
    return int(argv[argv.index("-w") + 1]) * 1000 if "-w" in argv else 0


def analysis_bit(name):
    return AN[name]
