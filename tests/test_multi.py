"""The multi-sample matrix path: MultiCount driven by VCFConvert
(src/MultiCount.hpp, src/VCFConvert.hpp, src/ntSeqMatchVCF.cpp).

Pins: tests/golden/vcf/<case>/ holds what the REFERENCE's two classes produced (tools/ref_vcf_harness.cpp
compiles them unmodified from /root/reference; tools/make_golden_vcf.py wrote the cases): the PCA matrix and
the centre file of printNormMatrix, printCountsMax per sample, the raw byte matrix, the warnings, and the
exit code where the reference dies.

CPU (`-m "not gpu"`): the oracle's restatement against every case and, where the harness binary is present,
against the live reference classes on fuzzed inputs; the ntsmVCF binary refuses to run without a GPU.
GPU (`-m gpu`): the library (C ABI through ctypes, and the ntsmVCF binary) against every case byte for byte,
and against the oracle on inputs the fixtures do not hold (many samples, batches, single inserts).
"""
import filecmp
import glob
import json
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

VCF_GOLDEN = os.path.join(GOLDEN, "vcf")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_vcf_harness")
NTSMVCF = os.path.join(ROOT, "ntsm_b200", "bin", "ntsmVCF")
SITES300 = os.path.join(GOLDEN, "shared", "sites300.fa")
GT = ["0|0", "0|1", "1|0", "1|1"]


def vcf_cases():
    return sorted(os.path.basename(d) for d in glob.glob(os.path.join(VCF_GOLDEN, "*")) if os.path.isdir(d))


def _case(name):
    d = os.path.join(VCF_GOLDEN, name)
    a = json.load(open(os.path.join(d, "args.json")))
    rc = int(open(os.path.join(d, "rc.txt")).read())
    return d, a, rc


def _same_outputs(golden_dir, got_prefix_dir, got_prefix="out"):
    """Every out_* file of the fixture equals the file of the same name produced under got_prefix_dir."""
    names = sorted(f for f in os.listdir(golden_dir) if f.startswith("out_"))
    assert names
    for f in names:
        g = os.path.join(got_prefix_dir, got_prefix + f[3:])
        assert os.path.exists(g), f
        assert filecmp.cmp(os.path.join(golden_dir, f), g, shallow=False), f


# --------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", vcf_cases())
def test_oracle_matches_reference_classes_fixture(oracle, name, tmp_path):
    d, a, want_rc = _case(name)
    rc = oracle.vcf_run(os.path.join(d, "sites.fa"), os.path.join(d, a["ref"]), os.path.join(d, "in.vcf"), str(tmp_path / "out"),
                        k=a["k"], dupes=a["dupes"], multi=a["multi"], window=a["window"])
    assert -rc == want_rc
    if want_rc:
        return
    _same_outputs(d, str(tmp_path))
    assert open(tmp_path / "out_stderr.txt", "rb").read() == open(os.path.join(d, "stderr.bin"), "rb").read()


def _fuzz_inputs(rng, d, n_sites, n_samples, k=19, window=31, repeats=0, overlap=False, odd_gt=0.03):
    """sites.fa + ref.fa + in.vcf under d for a random genome; returns nothing."""
    half = window // 2
    step = 7 if overlap else window + 30
    positions = [half + 60 + i * step + (0 if overlap else rng.randrange(0, 20)) for i in range(n_sites)]
    length = positions[-1] + 200
    g = [rng.choice("ACGT") for _ in range(length)]
    for p in positions:
        g[p - 1] = "A"
    if rng.random() < 0.5:
        for _ in range(length // 200):
            g[rng.randrange(length)] = rng.choice("Nnacgt")
    g = "".join(g)
    with open(os.path.join(d, "sites.fa"), "w") as fh:
        for i, p in enumerate(positions):
            w = g[p - half - 1:p + half].upper().replace("N", "A")
            v = w[:half] + "G" + w[half + 1:]
            n = window - k + 1
            fh.write(">rs%d\n%s\n>rs%d\n%s\n" % (i, "N".join(w[j:j + k] for j in range(n)), i, "N".join(v[j:j + k] for j in range(n))))
    open(os.path.join(d, "ref.fa"), "w").write(">chrA\n" + "\n".join(g[i:i + 61] for i in range(0, length, 61)) + "\n")
    samples = ["s%d" % i for i in range(n_samples)]
    lines = []
    for i, p in enumerate(positions):
        if rng.random() < 0.1:
            continue
        for _ in range(1 + (rng.randrange(repeats + 1) if repeats else 0)):
            gts = [rng.choice(GT) if rng.random() > odd_gt else rng.choice(["./.", "0/1", "1|2", ""]) for _ in samples]
            if gts[-1] == "":
                gts[-1] = "."                 # a tab at the end of the line is one column too few for the reference (fixture abort_two_trailing_tabs)
            alt = "G" if rng.random() > 0.05 else rng.choice(["GT", "<DEL>", "."])
            lines.append("chrA\t%d\trs%d\tA\t%s\t.\tPASS\t.\tGT%s\n" % (p, i, alt, "".join("\t" + x for x in gts)))
    with open(os.path.join(d, "in.vcf"), "w") as fh:
        fh.write("##fuzz\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" + "".join("\t" + s for s in samples) + "\n")
        fh.writelines(lines)


@pytest.mark.parametrize("seed", range(12))
def test_oracle_matches_live_reference_classes_on_fuzzed_vcfs(oracle, seed, tmp_path):
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/ref_vcf_harness not built (make -C oracle ref_vcf, where /root/reference exists)")
    rng = random.Random(1000 + seed)
    k, window = rng.choice([(19, 31), (19, 31), (21, 31), (17, 33), (25, 35)])
    overlap = seed % 3 == 0
    multi = rng.choice([20, 20, 1, 3, 100, 130])
    d = str(tmp_path)
    _fuzz_inputs(rng, d, rng.randrange(2, 40), rng.randrange(1, 9), k=k, window=window, repeats=seed % 2, overlap=overlap)
    dupes = 1 if overlap else 0
    p = subprocess.run([HARNESS, "sites.fa", "ref.fa", "in.vcf", "ref", str(k), str(multi), str(window), str(dupes)], cwd=d, capture_output=True)
    rc = oracle.vcf_run(d + "/sites.fa", d + "/ref.fa", d + "/in.vcf", d + "/orc", k=k, dupes=dupes, multi=multi, window=window)
    assert p.returncode == 0 and rc == 0, p.stderr[-300:]
    for f in sorted(glob.glob(d + "/ref_*")):
        assert filecmp.cmp(f, f.replace("/ref_", "/orc_"), shallow=False), os.path.basename(f)
    assert open(d + "/orc_stderr.txt", "rb").read() == p.stderr


def _oracle_matrix_from_parsed(oracle, sites, k, dupes, multi, n_samples, wins, geno):
    """MultiCount::insertCount (the oracle's) driven by the LIBRARY's parsed lines in VCFConvert::count's loop order."""
    omc = oracle.multicount(sites, n_samples, k=k, dupes=bool(dupes))
    for l, pair in enumerate(wins):
        for a in range(2):
            for (_, h, _, _) in oracle.iter(pair[a], k):
                for s in range(n_samples):
                    g = geno[l, s]
                    if g == (2 if a else 0):
                        omc.insert(s, h, multi * 2)
                    elif g == 1:
                        omc.insert(s, h, multi)
    return omc.matrix()


@pytest.mark.parametrize("name", vcf_cases())
def test_host_vcf_parser_feeds_the_reference_matrix(oracle, name):
    """The library's host half (genome, header, getSeqFromSite windows, genotype codes; ntsm_vcf_parse, no GPU) against the
    reference-made fixtures: the windows and genotypes it produces, pushed through the oracle's insertCount, give the
    reference's byte matrix; where the reference dies while reading the VCF, so does the parser; -t changes nothing."""
    from ntsm_b200.multicount import parse_vcf
    d, a, want_rc = _case(name)
    ref, vcf = os.path.join(d, a["ref"]), os.path.join(d, "in.vcf")
    rc, ids, wins, geno = parse_vcf(ref, vcf, window=a["window"], threads=1)
    rc7, ids7, wins7, geno7 = parse_vcf(ref, vcf, window=a["window"], threads=7)
    assert (rc, ids, wins) == (rc7, ids7, wins7) and np.array_equal(geno, geno7)
    dies_while_reading = name in ("abort_unknown_chrom", "abort_few_columns", "abort_empty_header_line", "abort_bad_pos", "abort_two_trailing_tabs")
    assert rc == (-134 if dies_while_reading else 0)
    if want_rc:
        return
    assert ids == open(os.path.join(d, "out_matrix.tsv")).readline().rstrip("\n").split("\t")[1:]   # the IDs the reference's header line carries
    want = np.fromfile(os.path.join(d, "out_mat.bin"), np.uint8)
    got = _oracle_matrix_from_parsed(oracle, os.path.join(d, "sites.fa"), a["k"], a["dupes"], a["multi"], len(ids), wins, geno)
    assert got.size == want.size and np.array_equal(got.ravel(), want)


@pytest.mark.parametrize("seed", range(6))
def test_host_vcf_parser_vs_oracle_on_fuzzed_vcfs(oracle, seed, tmp_path):
    from ntsm_b200.multicount import parse_vcf
    rng = random.Random(500 + seed)
    d = str(tmp_path)
    k, window = rng.choice([(19, 31), (21, 31), (17, 33)])
    _fuzz_inputs(rng, d, rng.randrange(5, 60), rng.randrange(1, 40), k=k, window=window, repeats=seed % 2, overlap=seed % 3 == 0)
    dupes = 1 if seed % 3 == 0 else 0
    assert oracle.vcf_run(d + "/sites.fa", d + "/ref.fa", d + "/in.vcf", d + "/orc", k=k, dupes=dupes, multi=20, window=window) == 0
    rc, ids, wins, geno = parse_vcf(d + "/ref.fa", d + "/in.vcf", window=window, threads=1 + seed)
    assert rc == 0
    got = _oracle_matrix_from_parsed(oracle, d + "/sites.fa", k, dupes, 20, len(ids), wins, geno)
    assert got.tobytes() == open(d + "/orc_mat.bin", "rb").read()


def _mutate_vcf(rng, lines):
    """A few line-level injuries of the kinds that decide how the reference's stringstream parsing behaves."""
    ls = list(lines)
    for _ in range(rng.randrange(1, 4)):
        i = rng.randrange(len(ls))
        f = ls[i].split(b"\t")
        op = rng.randrange(11)
        if op == 0 and len(f) > 1:
            f = f[:rng.randrange(1, len(f))]
        elif op == 1:
            f.append(rng.choice([b"0|1", b"", b"1|1"]))
        elif op == 2 and len(f) > 4:
            f[4] = rng.choice([b"", b"AC", b".", b"N", b"c", b"\0"])
        elif op == 3 and len(f) > 3:
            f[3] = rng.choice([b".", b"", b"ACG"])
        elif op == 4 and len(f) > 1:
            f[1] = rng.choice([b"", b"x", b"17", b" 300", b"300x", b"+400", b"99999999999", b"1e3", b"-5", b"0x20", b"\t"])
        elif op == 5:
            f = [b""]
        elif op == 6 and len(f) > 9:
            f[rng.randrange(9, len(f))] = rng.choice([b"0|0 ", b"1/1", b".", b"0|1\r", b"2|1", b""])
        elif op == 7:
            ls.insert(i, ls[i])
        elif op == 8:
            f[0] = rng.choice([b"chrZ", b"chr1 ", b""])
        elif op == 9:
            f = [b"#CHROM"] + f[1:]
        else:
            f = f + [b""]
        ls[i] = b"\t".join(f)
    if rng.random() < 0.2 and ls and ls[-1] == b"":
        ls.pop()
    return b"\n".join(ls)


@pytest.mark.parametrize("name", [n for n in vcf_cases() if not n.startswith("abort") and n != "overlap_no_dupes_aborts"])
def test_malformed_vcfs_three_ways(oracle, name, tmp_path):
    """Injured copies of every fixture's VCF through (a) the reference's classes, live, (b) the oracle, (c) the library's host
    parser + the oracle's insertCount: the same three files, or the same death.  (Inputs on which the reference has
    undefined behaviour -- a window that starts before its chromosome -- are recognised by (b) and (c) alike and skipped.)"""
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/ref_vcf_harness not built (make -C oracle ref_vcf, where /root/reference exists)")
    from ntsm_b200.multicount import parse_vcf
    d, a, _ = _case(name)
    rng = random.Random(name)
    lines = open(os.path.join(d, "in.vcf"), "rb").read().split(b"\n")
    sites, ref, t = os.path.join(d, "sites.fa"), os.path.join(d, a["ref"]), str(tmp_path)
    for _ in range(4):
        open(t + "/m.vcf", "wb").write(_mutate_vcf(rng, lines))
        p = subprocess.run([HARNESS, sites, ref, t + "/m.vcf", t + "/ref", str(a["k"]), str(a["multi"]), str(a["window"]), str(a["dupes"])], capture_output=True)
        orc_rc = oracle.vcf_run(sites, ref, t + "/m.vcf", t + "/orc", k=a["k"], dupes=a["dupes"], multi=a["multi"], window=a["window"])
        rc, ids, wins, geno = parse_vcf(ref, t + "/m.vcf", window=a["window"], threads=2)
        if orc_rc == -2 or rc == -1:
            assert (orc_rc, rc) == (-2, -1)
            continue
        if p.returncode != 0:
            assert (orc_rc, rc) == (-134, -134)
            continue
        assert orc_rc == 0 and rc == 0
        for x in ("_mat.bin", "_matrix.tsv", "_center.txt"):
            assert open(t + "/ref" + x, "rb").read() == open(t + "/orc" + x, "rb").read(), x
        got = _oracle_matrix_from_parsed(oracle, sites, a["k"], a["dupes"], a["multi"], len(ids), wins, geno)
        assert got.tobytes() == open(t + "/ref_mat.bin", "rb").read()


def test_gzipped_vcf_parses_like_plain(tmp_path):
    """An extension: a gzip / multi-member (bgzip-style) VCF is inflated by the library's own decoder and parsed like the plain
    file (upstream's ifstream would take the compressed bytes for text)."""
    import gzip
    from ntsm_b200.multicount import parse_vcf
    d, a, _ = _case("panel300_24samples")
    ref, raw = os.path.join(d, a["ref"]), open(os.path.join(d, "in.vcf"), "rb").read()
    want = parse_vcf(ref, os.path.join(d, "in.vcf"), threads=2)
    open(tmp_path / "one.vcf.gz", "wb").write(gzip.compress(raw))
    with open(tmp_path / "members.vcf.gz", "wb") as fh:
        for i in range(0, len(raw), 40000):
            fh.write(gzip.compress(raw[i:i + 40000]))
    import ntsm_b200
    L = ntsm_b200.lib()
    old = L.ntsm_vcf_stream_chunk(0)
    try:
        # such input is read in regions cut at line ends (64 MiB; here: smaller than a line, a few lines, the default)
        for chunk in (64, 3000, 100000, old):
            L.ntsm_vcf_stream_chunk(chunk)
            for f, threads in (("one.vcf.gz", 1), ("members.vcf.gz", 4)):
                got = parse_vcf(ref, str(tmp_path / f), threads=threads)
                assert got[:3] == want[:3] and np.array_equal(got[3], want[3]), (chunk, f)
        # every fixture, gzipped and read in small regions: the same lines, the same death
        L.ntsm_vcf_stream_chunk(700)
        for name in vcf_cases():
            d2, a2, _ = _case(name)
            open(tmp_path / "c.vcf.gz", "wb").write(gzip.compress(open(os.path.join(d2, "in.vcf"), "rb").read()))
            w = parse_vcf(os.path.join(d2, a2["ref"]), os.path.join(d2, "in.vcf"), window=a2["window"])
            g = parse_vcf(os.path.join(d2, a2["ref"]), str(tmp_path / "c.vcf.gz"), window=a2["window"], threads=3)
            assert g[:3] == w[:3] and np.array_equal(g[3], w[3]), name
    finally:
        L.ntsm_vcf_stream_chunk(old)
    assert parse_vcf(ref, str(tmp_path / "missing.vcf"))[0] == -5


def test_genotype_decoders_agree(tmp_path):
    """The VCF parser's column decoders -- byte-wise, AVX2 (8 columns a step), AVX-512 (16) -- give the same lines on every
    fixture and on wide fuzzed VCFs whose odd columns fall at every offset of a step."""
    import ntsm_b200
    from ntsm_b200.multicount import parse_vcf
    L = ntsm_b200.lib()
    have = L.ntsm_vcf_genotype_isa(-1)
    if have == 0:
        pytest.skip("no AVX2 on this CPU: only the byte-wise decoder exists")
    inputs = [(os.path.join(d, a["ref"]), os.path.join(d, "in.vcf"), a["window"]) for d, a, _ in map(_case, vcf_cases())]
    rng = random.Random(77)
    for i, n_samples in enumerate((7, 16, 33, 100, 257)):
        d = tmp_path / ("f%d" % i)
        os.makedirs(d)
        _fuzz_inputs(rng, str(d), 30, n_samples, odd_gt=0.02 * i)
        inputs.append((str(d / "ref.fa"), str(d / "in.vcf"), 31))
    try:
        for ref, vcf, window in inputs:
            L.ntsm_vcf_genotype_isa(0)
            want = parse_vcf(ref, vcf, window=window, threads=2)
            for isa in range(1, have + 1):
                assert L.ntsm_vcf_genotype_isa(isa) == isa
                got = parse_vcf(ref, vcf, window=window, threads=2)
                assert got[:3] == want[:3] and np.array_equal(got[3], want[3]), (vcf, isa)
    finally:
        L.ntsm_vcf_genotype_isa(-1)


def test_ntsmvcf_binary_has_no_cpu_path():
    assert os.path.exists(NTSMVCF), "make -C ntsm_b200/csrc"
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    d, a, _ = _case("basic")
    p = subprocess.run([NTSMVCF, "-s", "sites.fa", "-r", a["ref"], "-p", "/tmp/should_not_exist", "in.vcf"], cwd=d, capture_output=True)
    assert p.returncode == 1 and b"no CUDA device" in p.stderr
    assert not os.path.exists("/tmp/should_not_exist_matrix.tsv")
    # option handling comes before any device work, as upstream: missing reference / input -> exit 1 with the reference's words
    p = subprocess.run([NTSMVCF, "-s", "sites.fa", "in.vcf"], cwd=d, capture_output=True)
    assert p.returncode == 1 and b"Error: Unable to load reference file" in p.stderr and b"Try '--help'" in p.stderr
    p = subprocess.run([NTSMVCF, "-s", "sites.fa", "-r", a["ref"]], cwd=d, capture_output=True)
    assert p.returncode == 1 and b"Error: Need Input File" in p.stderr
    p = subprocess.run([NTSMVCF, "--help"], capture_output=True)
    assert p.returncode == 0 and b"Converts a multi vcf file to a set of counts files." in p.stderr


# --------------------------------------------------------------------------------------- GPU: the library
def _expected_stderr(d, sites_warnings):
    """The fixture's stderr = the site table's collision warnings, then insertCount's."""
    want = open(os.path.join(d, "stderr.bin"), "rb").read()
    head = "".join(w + "\n" for w in sites_warnings).encode()
    assert want.startswith(head)
    return want[len(head):]


@pytest.mark.gpu
@pytest.mark.parametrize("name", vcf_cases())
def test_abi_matches_reference_classes_fixture(name, tmp_path):
    import ntsm_b200
    d, a, want_rc = _case(name)
    vc = ntsm_b200.VCFConvert(os.path.join(d, "sites.fa"), os.path.join(d, a["ref"]), k=a["k"], dupes=bool(a["dupes"]), multi=a["multi"],
                              window=a["window"], threads=1 + len(name) % 3)
    if want_rc:
        with pytest.raises(KeyError):                       # the reference dies (uncaught exception / assert): rc 134
            vc.count(os.path.join(d, "in.vcf"))
            vc.outputMatrix(str(tmp_path / "out"))
        return
    vc.count(os.path.join(d, "in.vcf"))
    vc.outputMatrix(str(tmp_path / "out"))
    for j in range(len(vc.sample_ids)):
        open(tmp_path / ("out_counts_%d.txt" % j), "w").write(vc.counts.printCountsMax(j))
    vc.counts.matrix().tofile(str(tmp_path / "out_mat.bin"))
    _same_outputs(d, str(tmp_path))
    assert vc.counts.warnings_text == _expected_stderr(d, vc._fp.sites.warnings)
    assert vc.counts.launches > 0
    # outputCounts writes <sampleID>.counts.txt, one per sample, with the same bytes
    os.makedirs(tmp_path / "c")
    vc.outputCounts(str(tmp_path / "c"))
    for j, sid in enumerate(vc.sample_ids):
        assert open(tmp_path / "c" / (sid + ".counts.txt"), "rb").read() == open(os.path.join(d, "out_counts_%d.txt" % j), "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("name", vcf_cases())
def test_ntsmvcf_binary_matches_reference_classes_fixture(name, tmp_path):
    d, a, want_rc = _case(name)
    argv = [NTSMVCF, "-s", os.path.join(d, "sites.fa"), "-r", os.path.join(d, a["ref"]), "-k", str(a["k"]), "-m", str(a["multi"]),
            "-w", str(a["window"]), "-p", "out", "--counts", "-t", str(1 + len(name) % 5)] + (["-d"] if a["dupes"] else []) + [os.path.join(d, "in.vcf")]
    p = subprocess.run(argv, cwd=str(tmp_path), capture_output=True)
    if want_rc:
        assert p.returncode == 134 and b"terminate called" in p.stderr
        return
    assert p.returncode == 0, p.stderr.decode(errors="replace")
    for f in ("out_matrix.tsv", "out_center.txt"):
        assert filecmp.cmp(os.path.join(d, f), str(tmp_path / f), shallow=False), f
    ids = open(os.path.join(d, "out_matrix.tsv")).readline().rstrip("\n").split("\t")[1:]
    for j, sid in enumerate(ids):
        if ids.index(sid) == j:                              # sample IDs are file names here
            assert open(tmp_path / (sid + ".counts.txt"), "rb").read() == open(os.path.join(d, "out_counts_%d.txt" % j), "rb").read()
    want = open(os.path.join(d, "stderr.bin"), "rb").read()
    got = p.stderr[:p.stderr.rfind(b"Time: ")]
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_sites,n_samples,overlap", [(1, 60, 37, False), (2, 25, 5, True), (3, 120, 300, False), (4, 8, 1, True), (5, 200, 17, False)])
def test_abi_matches_oracle_on_fuzzed_vcfs(oracle, seed, n_sites, n_samples, overlap, tmp_path):
    import ntsm_b200
    rng = random.Random(seed)
    d = str(tmp_path)
    multi = rng.choice([20, 3, 130])
    _fuzz_inputs(rng, d, n_sites, n_samples, repeats=seed % 2, overlap=overlap)
    dupes = 1 if overlap else 0
    assert oracle.vcf_run(d + "/sites.fa", d + "/ref.fa", d + "/in.vcf", d + "/orc", dupes=dupes, multi=multi) == 0
    vc = ntsm_b200.VCFConvert(d + "/sites.fa", d + "/ref.fa", dupes=bool(dupes), multi=multi, threads=seed)
    vc.count(d + "/in.vcf")
    vc.outputMatrix(d + "/gpu")
    assert filecmp.cmp(d + "/orc_matrix.tsv", d + "/gpu_matrix.tsv", shallow=False)
    assert filecmp.cmp(d + "/orc_center.txt", d + "/gpu_center.txt", shallow=False)
    assert vc.counts.matrix().tobytes() == open(d + "/orc_mat.bin", "rb").read()
    for j in range(0, n_samples, max(1, n_samples // 7)):
        assert vc.counts.printCountsMax(j) == open(d + "/orc_counts_%d.txt" % j).read()
    site_warn = "".join(w + "\n" for w in vc._fp.sites.warnings).encode()
    assert site_warn + vc.counts.warnings_text == open(d + "/orc_stderr.txt", "rb").read()
    # the numbers behind the text: UNDEF where both maxima are zero, sums in sample order
    values, sums = vc.counts.normMatrix()
    undef = values == ntsm_b200.multicount.UNDEF
    rows = [l.rstrip("\n").split("\t") for l in open(d + "/orc_matrix.tsv")][1:]
    centers = [float(x) for x in open(d + "/orc_center.txt")]
    for i in (0, len(rows) // 2, len(rows) - 1):
        assert np.allclose(np.where(undef[i], centers[i], values[i]), [float(x) for x in rows[i][1:]], rtol=1e-5, atol=1e-12)
        assert abs(sums[i] / n_samples - centers[i]) < 1e-12


def _panel_windows():
    """[(ref window, alt window)] of the 300-site slice of the real panel (3-13 of a window's 13 k-mers are listed)."""
    import panel_util
    rng = random.Random(300)
    out = []
    for name, w, alt in panel_util.panel_windows(open(SITES300).read().splitlines(), rng):
        assert w is not None
        out.append((w.encode(), (w[:15] + alt + w[16:]).encode()))
    return out


@pytest.mark.gpu
def test_insert_windows_batches_and_single_inserts_vs_oracle(oracle):
    """ntsm_multi_insert_windows over more lines than one device batch takes (2^15), fed in uneven pieces, equals
    MultiCount::insertCount called k-mer by k-mer in the reference's loop order (the oracle's), including which
    cells the first writer keeps; ntsm_multi_insert_count (one insert per call) equals the same."""
    import ntsm_b200
    rng = random.Random(11)
    wins = _panel_windows()
    S, multi = 6, 20
    n_lines = 40000
    which = [rng.randrange(len(wins)) for _ in range(n_lines)]
    geno = np.array([[rng.randrange(3) for _ in range(S)] for _ in range(n_lines)], np.uint8)
    omc = oracle.multicount(SITES300, S)
    for l in range(n_lines):
        for a in range(2):
            for (pos, h, fw, rv) in oracle.iter(wins[which[l]][a], 19):
                for s in range(S):
                    g = geno[l, s]
                    if g == (2 if a else 0):
                        omc.insert(s, h, multi * 2)
                    elif g == 1:
                        omc.insert(s, h, multi)
    want = omc.matrix()
    mc = ntsm_b200.MultiCount(SITES300, ["s%d" % i for i in range(S)])
    at = 0
    for piece in (1, 4095, 4097, 33000, n_lines):
        end = min(n_lines, at + piece)
        mc.insertWindows([wins[w] for w in which[at:end]], geno[at:end], multi)
        at = end
    assert at == n_lines
    got = mc.matrix()
    assert got.shape == want.shape and np.array_equal(got, want)
    assert mc._fp.launches > 0
    # one insert per call
    omc2 = oracle.multicount(SITES300, 3)
    mc2 = ntsm_b200.MultiCount(SITES300, ["a", "b", "c"])
    hs = [h for w in wins[:40] for a in range(2) for (_, h, _, _) in oracle.iter(w[a], 19)]
    for t in range(600):
        s, h, m = rng.randrange(3), rng.choice(hs) if t % 10 else rng.getrandbits(38), rng.choice([1, 20, 40, 255, 256, 300])
        omc2.insert(s, h, m)
        mc2.insertCount(s, h, m)
    assert np.array_equal(mc2.matrix(), omc2.matrix())


@pytest.mark.gpu
def test_norm_matrix_many_samples_vs_oracle(oracle, tmp_path):
    """2 504 samples (the 1000 Genomes panel's width) over the 300-site slice: every digit of both files."""
    import ntsm_b200
    rng = random.Random(5)
    d = str(tmp_path)
    wins = _panel_windows()
    pieces, at, lines = [], 0, []
    samples = ["HG%05d" % i for i in range(2504)]
    for i, (w, v) in enumerate(wins):
        pad = "".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 50)))
        pieces.append(pad + w.decode())
        at += len(pad)
        if rng.random() > 0.05:
            p_alt = rng.random()
            gts = [GT[(rng.random() < p_alt) + 2 * (rng.random() < p_alt)] for _ in samples]
            lines.append("chr1\t%d\tx%d\t%s\t%s\t.\tPASS\t.\tGT\t%s\n" % (at + 16, i, chr(w[15]), chr(v[15]), "\t".join(gts)))
        at += 31
    open(d + "/ref.fa", "w").write(">chr1\n" + "".join(pieces) + "ACGT" * 30 + "\n")
    with open(d + "/in.vcf", "w") as fh:
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        fh.writelines(lines)
    assert oracle.vcf_run(SITES300, d + "/ref.fa", d + "/in.vcf", d + "/orc") == 0
    vc = ntsm_b200.VCFConvert(SITES300, d + "/ref.fa", threads=8)
    vc.count(d + "/in.vcf")
    assert vc.lines_counted == len(lines) and len(vc.sample_ids) == 2504
    vc.outputMatrix(d + "/gpu")
    assert filecmp.cmp(d + "/orc_matrix.tsv", d + "/gpu_matrix.tsv", shallow=False)
    assert filecmp.cmp(d + "/orc_center.txt", d + "/gpu_center.txt", shallow=False)
    assert vc.counts.matrix().tobytes() == open(d + "/orc_mat.bin", "rb").read()
