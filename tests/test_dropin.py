"""The compiled drop-in: the reference's OWN main() (src/ntSeqMatchCount.cpp, untouched) with class
FingerPrint replaced by include/FingerPrintB200.hpp over the C ABI of libntsm_b200.so
(`make -C oracle dropin`, built where /root/reference exists; the binary travels to the GPU box).
It proves the boundary from C++ -- the five-call seam of src/ntSeqMatchCount.cpp:177-181 -- and that the
reference's getopt front end drives the GPU path unchanged: every golden fixture made by the real
reference binary must come out byte-identical through it."""
import json
import os
import subprocess

import pytest

from conftest import ROOT, golden_cases
from test_host import _case_files
from test_oracle import _filter_err

DROPIN = os.path.join(ROOT, "oracle", "_ref", "ntsmCount_dropin")


def _have():
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/ntsmCount_dropin not built (make -C oracle dropin, where /root/reference exists)")


def test_dropin_links_the_library_and_has_no_cpu_path():
    _have()
    out = subprocess.run(["ldd", DROPIN], capture_output=True, text=True).stdout
    assert "libntsm_b200.so" in out and "not found" not in out.split("libntsm_b200.so")[1].splitlines()[0]
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    d, opts, files = _case_files("mini")
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    p = subprocess.run([DROPIN] + argv, cwd=d, capture_output=True)
    assert p.returncode == 1 and b"no CUDA device" in p.stderr and p.stdout == b""


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_cases())
def test_reference_main_over_the_c_abi_matches_reference_fixture(name):
    _have()
    d, opts, files = _case_files(name)
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    p = subprocess.run([DROPIN] + argv, cwd=d, capture_output=True)
    ref_rc = int(open(os.path.join(d, "rc.txt")).read())
    ref_out = open(os.path.join(d, "stdout.txt"), "rb").read()
    ref_err = open(os.path.join(d, "stderr.txt")).read()
    if ref_rc != 0:
        assert p.returncode in (134, -6)                 # uncaught std::out_of_range -> SIGABRT, as upstream
        assert b"out_of_range" in p.stderr
        assert [l for l in _filter_err(p.stderr.decode()) if "collision" in l] == [l for l in _filter_err(ref_err) if "collision" in l]
        return
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == ref_out
    keep = lambda t: [l for l in t.splitlines() if l.startswith(("Warning", "Reached", "Total ", "Distinct ", "Sites Covered"))]
    assert keep(p.stderr.decode()) == keep(ref_err)


# ---- the multi-sample matrix path: the reference's own ntsmVCF main over include/VCFConvertB200.hpp ----
VCF_DROPIN = os.path.join(ROOT, "oracle", "_ref", "ntsmVCF_dropin")


def _have_vcf():
    if not os.path.exists(VCF_DROPIN):
        pytest.skip("oracle/_ref/ntsmVCF_dropin not built (make -C oracle dropin_vcf, where /root/reference exists)")


def test_vcf_dropin_links_the_library_and_has_no_cpu_path():
    _have_vcf()
    out = subprocess.run(["ldd", VCF_DROPIN], capture_output=True, text=True).stdout
    assert "libntsm_b200.so" in out and "not found" not in out.split("libntsm_b200.so")[1].splitlines()[0]
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    from test_multi import _case
    d, a, _ = _case("basic")
    p = subprocess.run([VCF_DROPIN, "-s", "sites.fa", "-r", a["ref"], "-p", "/tmp/ntsm_dropin_should_not_exist", "in.vcf"], cwd=d, capture_output=True)
    assert p.returncode == 1 and b"no CUDA device" in p.stderr
    assert not os.path.exists("/tmp/ntsm_dropin_should_not_exist_matrix.tsv")


def _vcf_case_names():
    from test_multi import vcf_cases
    return vcf_cases()


@pytest.mark.gpu
@pytest.mark.parametrize("name", _vcf_case_names())
def test_reference_ntsmvcf_main_over_the_c_abi_matches_reference_classes_fixture(name, tmp_path):
    """src/ntSeqMatchVCF.cpp, untouched, compiled against VCFConvertB200.hpp: `ntsmVCF -p` produces the files the
    reference's own classes produce under tools/ref_vcf_harness.cpp (with them, that main crashes)."""
    _have_vcf()
    from test_multi import _case
    d, a, want_rc = _case(name)
    argv = [VCF_DROPIN, "-s", os.path.join(d, "sites.fa"), "-r", os.path.join(d, a["ref"]), "-k", str(a["k"]), "-m", str(a["multi"]),
            "-w", str(a["window"]), "-p", "out", "-t", "2"] + (["-d"] if a["dupes"] else []) + [os.path.join(d, "in.vcf")]
    p = subprocess.run(argv, cwd=str(tmp_path), capture_output=True)
    if want_rc:
        assert p.returncode in (134, -6) and b"out_of_range" in p.stderr     # uncaught exception -> SIGABRT, as upstream
        return
    assert p.returncode == 0, p.stderr.decode(errors="replace")
    for f in ("out_matrix.tsv", "out_center.txt"):
        assert open(os.path.join(d, f), "rb").read() == open(tmp_path / f, "rb").read(), f
    assert p.stderr[:p.stderr.rfind(b"Time: ")] == open(os.path.join(d, "stderr.bin"), "rb").read()
