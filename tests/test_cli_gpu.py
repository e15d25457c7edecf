"""End to end on the GPU box: the `popbam` command line (host feeder -> C ABI -> kernels -> printers) on BAM/BAI/FASTA
files must print what the unmodified reference printed for the same command line (tests/golden/*.txt)."""
import subprocess

import pytest

import pbtest
import popbam_b200
from cases import CASES, FIXTURES

pytestmark = pytest.mark.gpu
EXE = popbam_b200.capi.PKG / "_build" / "popbam"


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    out = {}
    for name in FIXTURES:
        out[name] = pbtest.fixture(name).write_files(d / name)
    return out


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cli_output_equals_reference(case, files):
    name, fxname, argv, an, _pk, okw = case
    bam, fa = files[fxname]
    r = subprocess.run([str(EXE)] + list(argv) + ["-f", fa, bam, "chr1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    ok, why = pbtest.texts_equal(r.stdout.decode(), pbtest.golden_text(case), snp0=(an == "SNP" and okw.get("snp_output", 0) == 0))
    assert ok, why


def test_cli_sharding_does_not_change_output(files):
    """Small shards (one window each) and sub-regions: rows are identical to the single-shard run."""
    bam, fa = files["c1"]
    base = subprocess.run([str(EXE), "sfs", "-w", "5", "-p", "og", "-f", fa, bam, "chr1"], stdout=subprocess.PIPE).stdout
    tiny = subprocess.run([str(EXE), "sfs", "-w", "5", "-p", "og", "--shard-mb", "0.004", "--threads", "3", "-f", fa, bam, "chr1"],
                          stdout=subprocess.PIPE).stdout
    assert base == tiny and base.count(b"\n") == 6
    sub = subprocess.run([str(EXE), "sfs", "-w", "5", "-p", "og", "-f", fa, bam, "chr1:10001-20001"], stdout=subprocess.PIPE).stdout
    assert sub == b"".join(base.splitlines(keepends=True)[2:4])


def test_cli_reports_errors(files):
    bam, fa = files["c1"]
    r = subprocess.run([str(EXE), "sfs", "-p", "nobody", "-f", fa, bam, "chr1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"Specified outgroup nobody not found" in r.stderr
    r = subprocess.run([str(EXE), "nucdiv", "-f", fa, bam, "chrZ"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"Bad genome coordinates" in r.stderr
