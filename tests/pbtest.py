"""Shared test plumbing: builds and loads the TEST-ONLY libraries (oracle restatement, synthetic
fixture generator, reference binary) and exposes small numpy-friendly wrappers.

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
TOOLS_DIR = ROOT / "tools"
GOLDEN = ROOT / "tests" / "golden"
REF_BIN = ORACLE_DIR / "_ref" / "popbam"
REFDUMP = ORACLE_DIR / "_ref" / "refdump"

MAXS = 64

AN = dict(NUCDIV=0x001, SFS=0x002, LD_ZNS=0x004, LD_OMEGA=0x008, LD_WALL=0x010, DIVERGE_IND=0x020,
          DIVERGE_POP=0x040, HAPLO_K=0x080, HAPLO_EHHS=0x100, HAPLO_DXY=0x200, SNP=0x400, TREE=0x800)
FLAG = dict(ILLUMINA=0x02, SUBSTITUTE=0x10, HETEROZYGOTE=0x20, OUTGROUP=0x40, EMIT_CB=0x10000)


# result arrays that belong to each analysis bit: (integer arrays compared exactly, fp64 arrays compared to 1e-9)
BY_AN = {
    "NUCDIV": (["min_dxy"], ["piw", "pib"]), "HAPLO_DXY": (["min_dxy"], ["piw", "pib"]),
    "SFS": (["sfs_num_snps"], ["td", "fwh"]),
    "LD_ZNS": (["ld_num_snps"], ["zns"]), "LD_OMEGA": (["ld_num_snps"], ["omegamax"]),
    "LD_WALL": (["wall_num_snps"], ["wallb", "wallq"]),
    "DIVERGE_IND": (["ind_div"], []), "DIVERGE_POP": (["pop_div", "div_num_snps"], []),
    "HAPLO_K": (["nhaps"], ["hdiv"]), "HAPLO_EHHS": (["nhaps"], ["hdiv", "ehhs"]),
    "SNP": (["seg_cb"], []),
    "TREE": (["tree_diff"], []),
}

# The C-ABI structures are shared with the product binding (the oracle mirrors their layout on purpose:
# oracle/pb_oracle.h pbo_params / pbo_batch / pbo_result / pbo_print_opts).
sys.path.insert(0, str(ROOT))
from popbam_b200.capi import Batch, Params, PrintOpts, Result  # noqa: E402


def _p(t):
    return C.POINTER(t)


class SynthParams(C.Structure):
    _fields_ = [("n_contigs", C.c_int32), ("contig_len", C.c_int32), ("n_ingroup", C.c_int32),
                ("has_outgroup", C.c_int32), ("rg_per_sample", C.c_int32), ("depth", C.c_double),
                ("read_len", C.c_int32), ("snp_density", C.c_double), ("het_frac", C.c_double),
                ("frac_del", C.c_double), ("frac_ins", C.c_double), ("edge_mode", C.c_int32),
                ("seed", C.c_uint64), ("n_threads", C.c_int32)]


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))


def build_test_libs():
    """Compile the oracle restatement, the generator and (when /root/reference exists) oracle/_ref."""
    _run(["make", "-s", "oracle"], ORACLE_DIR)
    _run(["make", "-s", "ref"], ORACLE_DIR)
    _run(["make", "-s"], TOOLS_DIR)


_oracle = None
_synth = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        so = ORACLE_DIR / "_build" / "libpboracle.so"
        if not so.exists():
            _run(["make", "-s", "oracle"], ORACLE_DIR)
        L = C.CDLL(str(so))
        L.pbo_build_tables.argtypes = [_p(C.c_double)] * 3
        L.pbo_call_cell.restype = C.c_uint64
        L.pbo_call_cell.argtypes = [_p(C.c_double)] * 3 + [_p(C.c_uint16), C.c_int, C.c_int, _p(C.c_float)]
        L.pbo_site_logic.argtypes = [_p(Params), _p(C.c_uint64), C.c_char, _p(C.c_uint64), _p(C.c_uint64)]
        L.pbo_run_region.argtypes = [_p(Params)] + [_p(C.c_double)] * 3 + [_p(Batch), C.c_char_p, C.c_int64,
                                                                         C.c_uint32, C.c_int32, _p(C.c_int32),
                                                                         _p(C.c_int32), _p(Result)]
        L.pbo_free_result.argtypes = [_p(Result)]
        L.pbo_window_grid.restype = C.c_int64
        L.pbo_window_grid.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _p(C.c_int32), _p(C.c_int32)]
        L.pbo_format_window.restype = C.c_int64
        L.pbo_format_window.argtypes = [_p(Params), _p(Result), C.c_int32, C.c_uint32, _p(PrintOpts), C.c_char_p,
                                        C.c_int64]
        _oracle = L
    return _oracle


def synth_lib():
    global _synth
    if _synth is None:
        so = TOOLS_DIR / "_build" / "libpbsynth.so"
        if not so.exists():
            _run(["make", "-s"], TOOLS_DIR)
        L = C.CDLL(str(so))
        L.pbsynth_default_params.argtypes = [_p(SynthParams)]
        L.pbsynth_create.restype = C.c_void_p
        L.pbsynth_create.argtypes = [_p(SynthParams)]
        for fn in ("pbsynth_destroy",):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.pbsynth_n_samples.argtypes = [C.c_void_p]
        L.pbsynth_n_pops.argtypes = [C.c_void_p]
        L.pbsynth_sample_pop.argtypes = [C.c_void_p, C.c_int]
        for fn in ("pbsynth_sample_name", "pbsynth_pop_name", "pbsynth_contig_name"):
            getattr(L, fn).restype = C.c_char_p
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
        L.pbsynth_ref.restype = C.c_void_p
        L.pbsynth_ref.argtypes = [C.c_void_p, C.c_int]
        L.pbsynth_aligned_bases.restype = C.c_int64
        L.pbsynth_aligned_bases.argtypes = [C.c_void_p, C.c_int]
        L.pbsynth_batch_get.argtypes = [C.c_void_p, C.c_int, _p(Batch)]
        L.pbsynth_n_snps.restype = C.c_int64
        L.pbsynth_n_snps.argtypes = [C.c_void_p, C.c_int]
        L.pbsynth_write_files.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _synth = L
    return _synth


class Tables:
    """Error-model tables as numpy arrays (fk[256], beta[64*256*256], lhet[65536])."""

    def __init__(self, fk, beta, lhet):
        self.fk, self.beta, self.lhet = fk, beta, lhet

    def ptrs(self):
        return tuple(a.ctypes.data_as(_p(C.c_double)) for a in (self.fk, self.beta, self.lhet))


_tables = None


def oracle_tables():
    global _tables
    if _tables is None:
        fk = np.zeros(256); beta = np.zeros(64 * 256 * 256); lhet = np.zeros(65536)
        t = Tables(fk, beta, lhet)
        assert oracle_lib().pbo_build_tables(*t.ptrs()) == 0
        _tables = t
    return _tables


class Fixture:
    """A seeded synthetic data set: batches in the C-ABI layout (+ files on demand)."""

    def __init__(self, **kw):
        L = synth_lib()
        sp = SynthParams()
        L.pbsynth_default_params(C.byref(sp))
        for k, v in kw.items():
            if not hasattr(sp, k):
                raise KeyError(k)
            setattr(sp, k, v)
        self.sp = sp
        self.h = L.pbsynth_create(C.byref(sp))
        if not self.h:
            raise ValueError("bad synth parameters")
        self.L = L
        self.n_samples = L.pbsynth_n_samples(self.h)
        self.n_pops = L.pbsynth_n_pops(self.h)
        self.sample_names = [L.pbsynth_sample_name(self.h, i).decode() for i in range(self.n_samples)]
        self.pop_names = [L.pbsynth_pop_name(self.h, i).decode() for i in range(self.n_pops)]
        self.sample_pop = [L.pbsynth_sample_pop(self.h, i) for i in range(self.n_samples)]
        self.contig_len = sp.contig_len

    def close(self):
        if self.h:
            self.L.pbsynth_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def contig_name(self, c=0):
        return self.L.pbsynth_contig_name(self.h, c).decode()

    def batch(self, c=0):
        b = Batch()
        assert self.L.pbsynth_batch_get(self.h, c, C.byref(b)) == 0
        return b

    def ref(self, c=0):
        return C.string_at(self.L.pbsynth_ref(self.h, c), self.contig_len)

    def aligned_bases(self, c=0):
        return self.L.pbsynth_aligned_bases(self.h, c)

    def write_files(self, prefix, level=1):
        assert self.L.pbsynth_write_files(self.h, str(prefix).encode(), level) == 0
        return str(prefix) + ".bam", str(prefix) + ".fa"

    def params(self, **kw):
        """pb_params with the reference defaults (popbam.cpp:79-93) and this fixture's populations."""
        p = Params()
        p.n_samples, p.n_pops = self.n_samples, self.n_pops
        for s, pop in enumerate(self.sample_pop):
            p.pop_mask[pop] |= 1 << s
            p.pop_nsmpl[pop] += 1
        p.min_depth, p.max_depth, p.min_rmsQ, p.min_snpQ, p.min_mapQ, p.min_baseQ = 3, 255, 25, 25, 13, 13
        p.flags, p.outidx, p.min_freq, p.device = 0, 0, 1, 0
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        return p

    def print_opts(self, c=0, **kw):
        o = PrintOpts()
        self._keep = [(C.c_char_p * self.n_pops)(*[s.encode() for s in self.pop_names]),
                      (C.c_char_p * self.n_samples)(*[s.encode() for s in self.sample_names])]
        o.chrom = self.contig_name(c).encode()
        o.pop_names, o.sample_names = self._keep
        o.min_sites, o.min_snps, o.jc, o.snp_output = 10, 10, 0, 0
        o.ref_name = b"synth"          # the AS tag tools/pbsynth writes into @SQ (tree subcommand: treeData::refid)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


def window_grid(beg, end, win_size):
    L = oracle_lib()
    nw = L.pbo_window_grid(beg, end, win_size, 0, None, None)
    wb = (C.c_int32 * max(nw, 1))(); we = (C.c_int32 * max(nw, 1))()
    L.pbo_window_grid(beg, end, win_size, nw, wb, we)
    return np.array(wb[:nw], dtype=np.int32), np.array(we[:nw], dtype=np.int32)


def arr(ptr, n, dtype=None):
    """numpy view of a ctypes pointer (copy)."""
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).copy()


class OracleRun:
    """Runs the oracle restatement over a region and keeps the result alive."""

    def __init__(self, params, batch, ref, analyses, win_beg, win_end, tables=None):
        L = oracle_lib()
        t = tables or oracle_tables()
        self.L, self.p = L, params
        self.res = Result()
        wb = np.ascontiguousarray(win_beg, dtype=np.int32); we = np.ascontiguousarray(win_end, dtype=np.int32)
        rc = L.pbo_run_region(C.byref(params), *t.ptrs(), C.byref(batch), ref, len(ref), analyses, len(wb),
                              wb.ctypes.data_as(_p(C.c_int32)), we.ctypes.data_as(_p(C.c_int32)), C.byref(self.res))
        if rc != 0:
            raise RuntimeError("pbo_run_region failed: %d" % rc)

    def text(self, analysis, opts, windows=None):
        out = []
        buf = C.create_string_buffer(1 << 20)
        for w in (range(self.res.n_windows) if windows is None else windows):
            k = self.L.pbo_format_window(C.byref(self.p), C.byref(self.res), w, analysis, C.byref(opts), buf, len(buf))
            if k >= len(buf):
                buf = C.create_string_buffer(int(k) + 16)
                k = self.L.pbo_format_window(C.byref(self.p), C.byref(self.res), w, analysis, C.byref(opts), buf, len(buf))
            out.append(buf.raw[:k].decode())
        return "".join(out)

    def close(self):
        if self.res.win_beg:
            self.L.pbo_free_result(C.byref(self.res))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def result_arrays(res, want_cb=False):
    """Dict of numpy copies of every array in a pb_region_result / pbo_result."""
    NW, P, n = res.n_windows, res.n_pops, res.n_samples
    S = int(res.seg_off[NW]) if res.seg_off else 0
    span = res.span_end - res.span_beg
    d = dict(win_beg=arr(res.win_beg, NW), win_end=arr(res.win_end, NW), num_sites=arr(res.num_sites, NW),
             segsites=arr(res.segsites, NW), seg_off=arr(res.seg_off, NW + 1), seg_pos=arr(res.seg_pos, S),
             seg_idx=arr(res.seg_idx, S), seg_type=arr(res.seg_type, S), seg_ref=arr(res.seg_ref, S),
             seg_cb=arr(res.seg_cb, S * n), piw=arr(res.piw, NW * P), pib=arr(res.pib, NW * P * P),
             min_dxy=arr(res.min_dxy, NW * P * P), sfs_num_snps=arr(res.sfs_num_snps, NW * P),
             td=arr(res.td, NW * P), fwh=arr(res.fwh, NW * P), ld_num_snps=arr(res.ld_num_snps, NW * P),
             zns=arr(res.zns, NW * P), omegamax=arr(res.omegamax, NW * P),
             wall_num_snps=arr(res.wall_num_snps, NW * P), wallb=arr(res.wallb, NW * P), wallq=arr(res.wallq, NW * P),
             ind_div=arr(res.ind_div, NW * n), pop_div=arr(res.pop_div, NW * P),
             div_num_snps=arr(res.div_num_snps, NW * P), nhaps=arr(res.nhaps, NW * P), hdiv=arr(res.hdiv, NW * P),
             ehhs=arr(res.ehhs, NW * P), site_type=arr(res.site_type, span), site_flag=arr(res.site_flag, span),
             tree_diff=arr(res.tree_diff, NW * (n + 1) * (n + 1)))
    if want_cb:
        d["cb"] = arr(res.cb, span * n)
    return d


def have_ref():
    return REF_BIN.exists()


def run_ref(args, cwd=None):
    """Run the unmodified reference binary; returns stdout text."""
    r = subprocess.run([str(REF_BIN)] + list(args), cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if r.returncode != 0:
        raise RuntimeError("reference popbam failed (%d): %s" % (r.returncode, r.stderr.decode()[-2000:]))
    return r.stdout.decode()


# ---------------------------------------------------------------------------------------------
# parity-case helpers (tests/cases.py)
_fixture_cache = {}


def fixture(name):
    from cases import FIXTURES
    if name not in _fixture_cache:
        _fixture_cache[name] = Fixture(**FIXTURES[name])
    return _fixture_cache[name]


def ms_header(fx, n_windows):
    """print_ms_header (pop_snp.cpp:305-317)."""
    s = "ms %d %d -t 5.0 " % (fx.n_samples, n_windows)
    if fx.n_pops > 1:
        cnt = [fx.sample_pop.count(p) for p in range(fx.n_pops)]
        s += "-I %d %s " % (fx.n_pops, " ".join(str(c) for c in cnt))
    return s + "\n1350154902\n\n"


def case_setup(case):
    """(fixture, params, analysis bit, window arrays, print opts) of one tests/cases.py row."""
    from cases import win_kb
    name, fxname, argv, an, pkw, okw = case
    fx = fixture(fxname)
    p = fx.params(**pkw)
    wb, we = window_grid(0, fx.contig_len, win_kb(argv))
    o = fx.print_opts(**okw)
    return fx, p, AN[an], wb, we, o


def golden_text(case):
    return (GOLDEN / (case[0] + ".txt")).read_text()


def texts_equal(got, want, snp0=False):
    """Exact text comparison.  For `snp -o 0` rows the base letter of a sample whose genotype byte is out of range
    is undefined in the reference (iupac[] read out of bounds): the restatement prints '?' there and
    that one character is excluded; every other field must match."""
    if not snp0:
        return got == want, "text differs"
    gl, wl = got.splitlines(), want.splitlines()
    if len(gl) != len(wl):
        return False, "row count %d != %d" % (len(gl), len(wl))
    for i, (a, b) in enumerate(zip(gl, wl)):
        fa, fb = a.split("\t"), b.split("\t")
        if len(fa) != len(fb):
            return False, "row %d: field count" % i
        for j, (x, y) in enumerate(zip(fa, fb)):
            if x == y:
                continue
            if j >= 3 and (j - 3) % 4 == 0 and x == "?":
                continue
            return False, "row %d field %d: %r != %r" % (i, j, x, y)
    return True, ""
