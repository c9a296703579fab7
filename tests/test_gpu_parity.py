"""GPU parity: the CUDA path, called through the C ABI (libntsm_b200.so) and through the
ntsmCount binary, against (a) the committed stdout/stderr of the real reference binary and
(b) the CPU oracle on seeded inputs.  Bit-exact: everything on this path is integer work."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases
from test_host import _case_files
from test_oracle import _filter_err

import ntsm_b200

pytestmark = pytest.mark.gpu

NTSMCOUNT = os.path.join(ROOT, "ntsm_b200", "bin", "ntsmCount")
PANEL = os.path.join(ROOT, "data", "human_sites_n10.fa.gz")
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp(b):
    return b.translate(COMP)[::-1]


# ---------------------------------------------------------------- reference fixtures
@pytest.mark.parametrize("name", golden_cases())
def test_cli_matches_reference_fixture(name):
    """Our ntsmCount binary vs what the reference binary printed for the same argv."""
    d, opts, files = _case_files(name)
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    p = subprocess.run([NTSMCOUNT] + argv, cwd=d, capture_output=True)
    ref_rc = int(open(os.path.join(d, "rc.txt")).read())
    ref_out = open(os.path.join(d, "stdout.txt"), "rb").read()
    ref_err = open(os.path.join(d, "stderr.txt")).read()
    if ref_rc != 0:
        assert p.returncode == 134
        assert [l for l in _filter_err(p.stderr.decode()) if "collision" in l] == [l for l in _filter_err(ref_err) if "collision" in l]
        return
    assert p.returncode == 0, p.stderr.decode()
    # -m with one parser thread stops after the very read the reference stops after (ntsm_trim_to_cap),
    # so these fixtures are compared byte for byte like all the others
    assert p.stdout == ref_out
    keep = lambda t: [l for l in t.splitlines() if l.startswith(("Warning", "Reached", "Total ", "Distinct ", "Sites Covered"))]
    assert keep(p.stderr.decode()) == keep(ref_err)


@pytest.mark.parametrize("name", [n for n in golden_cases() if not n.startswith(("dupes_abort", "odd_sites"))])
def test_abi_files_match_reference_fixture(name):
    """FingerPrint mirror over the C ABI (computeCounts on the same files)."""
    d, opts, files = _case_files(name)
    if opts["m"] > 0 and "Reached desired" in open(os.path.join(d, "stderr.txt")).read():
        pytest.skip("-m early stop: covered by the prefix-property test")
    fp = ntsm_b200.FingerPrint(opts["sites"], k=opts["k"], dupes=opts["dupes"], cov_thresh=opts["m"], batch_bases=1 << 14)
    fp.computeCounts(files, threads=opts["t"])
    assert fp.counts_text().encode() == open(os.path.join(d, "stdout.txt"), "rb").read()
    assert fp.printInfoSummary() in open(os.path.join(d, "stderr.txt")).read()
    fp.close()


# ---------------------------------------------------------------- oracle, seeded inputs
def _windows(sites_path, limit=None):
    import gzip
    op = gzip.open if sites_path.endswith(".gz") else open
    wins = []
    with op(sites_path, "rt") as fh:
        for i, line in enumerate(fh):
            if limit and i >= limit:
                break
            if not line.startswith(">"):
                wins.extend(line.strip().split("N"))
    return wins


def _reads(rng, wins, n, alphabet="ACGT", maxflank=80):
    out = []
    for i in range(n):
        r = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, maxflank))) + rng.choice(wins) + \
            "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, maxflank)))
        if rng.random() < 0.3:
            p = rng.randrange(len(r)); r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        r = r.encode()
        out.append(revcomp(r) if rng.random() < 0.5 else r)
    return out


def _check_against_oracle(oracle, sites, reads, k=19, dupes=False, **kw):
    fp = ntsm_b200.FingerPrint(sites, k=k, dupes=dupes, **kw)
    ofp = oracle.fingerprint(sites, k, dupes)
    for r in reads:
        fp.insertCount(r)
        ofp.insert(r)
    _, _, ocnt = ofp.lists()
    gcnt = fp.kmer_counts()
    live = ocnt != 0xFFFFFFFF
    assert np.array_equal(gcnt[live], ocnt[live])            # every k-mer's counter, not just the per-site rows
    mr, mv, sr, sv, t = fp.finalize()
    assert (int(t[0]), int(t[1]), int(t[2])) == (ofp.total_kmers, ofp.total_counts, ofp.total_bases)
    if fp.sites.printable():
        assert fp.counts_text() == ofp.counts_text()
        assert fp.printInfoSummary() == ofp.summary()
    fp.close()
    return ofp


def test_insert_count_vs_oracle_k19(oracle):
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(42)
    reads = _reads(rng, _windows(sites), 3000, "ACGTN")
    ofp = _check_against_oracle(oracle, sites, reads)
    assert ofp.total_counts > 3000


@pytest.mark.parametrize("k", [1, 2, 5, 11, 15, 16, 17, 18, 20, 21, 22, 24, 25, 27, 29, 30, 31])
def test_other_k_vs_oracle(oracle, k):
    """Every -k the reference takes (vendor/KseqHashIterator.hpp:28-33): k < 17 runs the generic kernel,
    17..31 the paired-seed kernel with k at run time (seed length min(14, k-5)); both against the oracle's
    per-k-mer counters, with and without the pair kernel for the k it applies to."""
    name = "k%d" % k if os.path.isdir(os.path.join(GOLDEN, "cases", "k%d" % k)) else "k11"
    sites = os.path.join(GOLDEN, "cases", name, "sites.fa")
    rng = random.Random(k)
    wins = [w for w in _windows(sites)]
    reads = _reads(rng, wins, 800, "ACGTN", maxflank=40)
    fp = ntsm_b200.FingerPrint(sites, k=k, dupes=True)
    assert fp.kernel_name == ("count_kernel_generic" if k < 17 else "count_kernel_pair<0,1024,2>")
    fp.close()
    _check_against_oracle(oracle, sites, reads, k=k, dupes=True)
    if k >= 17:
        _check_against_oracle(oracle, sites, reads, k=k, dupes=True, options={"kernel": 0})


@pytest.mark.parametrize("k", [17, 18, 21, 25, 28, 31])
def test_pair_kernel_other_k_on_a_large_panel(oracle, tmp_path, k):
    """A 30 000-site synthetic panel k-merized AT k (k - 6 k-mers per allele, like the human panel at 19),
    dense-hit reads made of its windows, N runs, both strands, odd batch sizes: count_kernel_pair<0> with
    k at run time against the oracle's per-k-mer counters."""
    from ntsm_b200 import synth_np
    sites = str(tmp_path / "sites.fa")
    win, n = synth_np.synthetic_panel(sites, 30000, 100 + k, k=k, flank=k - 4)
    assert n > 29000
    wins = ["".join("ACGT"[c] for c in win[i, a]) for i in range(0, 3000) for a in (0, 1)]
    rng = random.Random(500 + k)
    reads = []
    for j in range(1500):
        r = "".join(rng.choice(wins) for _ in range(rng.randrange(1, 6)))
        if j % 3 == 0:
            p = rng.randrange(len(r)); r = r[:p] + "N" * rng.randrange(1, 5) + r[p:]
        if j % 4 == 0:
            r = r[rng.randrange(0, 7):]
        r = r.encode()
        reads.append(revcomp(r) if rng.random() < 0.5 else r)
    reads += [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 300))) for _ in range(500)]
    for bb in (40001, 1 << 18):
        ofp = _check_against_oracle(oracle, sites, reads, k=k, batch_bases=bb)
    if k >= 19:                     # the wide table (16-mers 4 apart) takes any k >= 19 too
        _check_against_oracle(oracle, sites, reads, k=k, batch_bases=1 << 16, options={"kernel": 2})
    assert ofp.total_counts > 5 * len(reads)


def test_alphabet_and_edge_reads_vs_oracle(oracle):
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    w = _windows(sites)
    reads = [b"", b"A", w[0][:18].encode(), w[0][:19].encode(), w[0].lower().encode(), w[1].replace("T", "U").encode(),
             w[2].replace("T", "u").encode(), bytes("ACGT".index(c) for c in w[3]), w[4][:10].encode() + b"\xff" + w[4][10:].encode(),
             b"N" * 100, (w[5] + "N" + w[6] + "R" + w[7]).encode(), b"ACGT" * 5000 + w[8].encode() + b"TTTT" * 5000]
    _check_against_oracle(oracle, sites, reads)


def test_long_reads_split_across_batches(oracle):
    """Reads far longer than a batch are cut with a k-1 overlap: every k-mer still counted once."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(7)
    wins = _windows(sites)
    reads = []
    for _ in range(30):
        parts = []
        for _ in range(rng.randrange(5, 60)):
            parts.append("".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 2500))))
            parts.append(rng.choice(wins))
            if rng.random() < 0.2:
                parts.append("N" * rng.randrange(1, 60))
        r = "".join(parts).encode()
        reads.append(revcomp(r) if rng.random() < 0.5 else r)
    assert max(len(r) for r in reads) > 20000
    for bb in (4096, 5000, 1 << 14):
        _check_against_oracle(oracle, sites, reads, batch_bases=bb, n_buffers=2)


def test_bulk_insert_reads_vs_oracle(oracle):
    """ntsm_insert_reads / ntsm_insert_reads_fixed (multi-threaded pack into the pinned ring) ==
    one insertCount per read == the oracle, for ragged and fixed-length bulks."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(99)
    wins = _windows(sites)
    reads = _reads(rng, wins, 6000, "ACGTN") + [b"", b"A", b"N" * 40, wins[0].encode() * 300]
    ofp = oracle.fingerprint(sites, 19, False)
    for r in reads:
        ofp.insert(r)
    buf = np.frombuffer(b"".join(reads), np.uint8)
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    for threads, bb in ((1, 1 << 14), (4, 1 << 13), (7, 5000)):
        fp = ntsm_b200.FingerPrint(sites, batch_bases=bb, n_buffers=threads + 2)
        fp.insertReads(buf, off, threads=threads)
        assert fp.counts_text() == ofp.counts_text() and fp.printInfoSummary() == ofp.summary()
        fp.close()
    # dense matrix form: fixed-length reads with a row stride
    L_, stride, n = 151, 160, 4000
    mat = np.full((n, stride), ord("N"), np.uint8)
    ofp2 = oracle.fingerprint(sites, 19, False)
    for i in range(n):
        w = rng.choice(wins).encode()
        r = (bytes(rng.choice(b"ACGT") for _ in range(L_)) + w)[-L_:] if i % 3 else bytes(rng.choice(b"ACGTN") for _ in range(L_))
        mat[i, :L_] = np.frombuffer(r, np.uint8)
        ofp2.insert(r)
    fp = ntsm_b200.FingerPrint(sites, batch_bases=1 << 15, n_buffers=5)
    fp.insertReadsFixed(mat.ctypes.data, L_, stride, n, threads=3)
    assert fp.counts_text() == ofp2.counts_text() and fp.printInfoSummary() == ofp2.summary()
    fp.close()


def test_full_panel_vs_oracle(oracle):
    """The real 96 287-site panel, reads planted on panel windows (first 20k records) + random ones."""
    rng = random.Random(2026)
    wins = _windows(PANEL, limit=20000)
    reads = _reads(rng, wins, 20000) + [bytes(rng.choice(b"ACGT") for _ in range(150)) for _ in range(5000)]
    ofp = _check_against_oracle(oracle, PANEL, reads)
    assert ofp.n_sites == 96287 and ofp.table_size == 1270317


# ---------------------------------------------------------------- size-independent properties
def _device_pack(reads):
    import torch
    b2, mk, n_pos, _ = ntsm_b200.pack_reads(reads)
    return torch.from_numpy(b2.view(np.int32)).cuda(), torch.from_numpy(mk.view(np.int32)).cuda(), n_pos


def test_properties_linearity_strand_and_reset(oracle):
    """counts(A+B) = counts(A) + counts(B); reverse-complementing every read changes nothing;
    reset gives zeros; the device-resident entry point agrees with the pinned-batch path."""
    import torch
    rng = random.Random(5)
    wins = _windows(PANEL, limit=40000)
    A = _reads(rng, wins, 30000)
    B = _reads(rng, wins, 30000)
    fp = ntsm_b200.FingerPrint(PANEL)

    def run(reads):
        fp.reset()
        db, dm, n = _device_pack(reads)
        fp.count_packed_device(db.data_ptr(), dm.data_ptr(), n, sum(map(len, reads)), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        c = fp.kmer_counts().astype(np.uint64)
        t = fp.finalize()[4].copy()
        return c, t

    ca, ta = run(A)
    cb, tb = run(B)
    cab, tab = run(A + B)
    assert np.array_equal(ca + cb, cab) and np.array_equal(ta + tb, tab)
    crc, trc = run([revcomp(r) for r in A])
    assert np.array_equal(crc, ca) and np.array_equal(trc, ta)
    fp.reset()
    assert fp.kmer_counts().sum() == 0 and fp.finalize()[4].sum() == 0
    # pinned-batch path == device-resident path
    fp.reset()
    for r in A:
        fp.insertCount(r)
    assert np.array_equal(fp.kmer_counts().astype(np.uint64), ca)
    assert np.array_equal(fp.finalize()[4], ta)
    # and both equal the oracle
    ofp = oracle.fingerprint(PANEL, 19)
    buf = np.frombuffer(b"".join(A), np.uint8)
    off = np.zeros(len(A) + 1, np.uint64); np.cumsum([len(r) for r in A], out=off[1:])
    ofp.insert_many(buf, off, threads=8)
    _, _, oc = ofp.lists()
    assert np.array_equal(oc.astype(np.uint64), ca)
    assert (ofp.total_kmers, ofp.total_counts, ofp.total_bases) == tuple(int(x) for x in ta)
    fp.close()


@pytest.mark.parametrize("name", ["panel300_m1", "panel300_m0.5_two_files"])
@pytest.mark.parametrize("batch", ["4096", "30000", "1000003", None])
def test_m_cap_stops_after_the_same_read_as_the_reference(name, batch):
    """-m, one parser thread (the reference's deterministic -t 1): whatever the batch size -- the cap
    crossed in the first batch, the last one, one holding a single read or thousands -- the counts
    file and the totals are the reference's."""
    d, opts, files = _case_files(name)
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    if batch:
        argv = ["--batch-bases", batch] + argv
    p = subprocess.run([NTSMCOUNT] + argv, cwd=d, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == open(os.path.join(d, "stdout.txt"), "rb").read()
    keep = lambda t: [l for l in t.splitlines() if l.startswith(("Reached", "Total ", "Distinct ", "Sites Covered"))]
    assert keep(p.stderr.decode()) == keep(open(os.path.join(d, "stderr.txt")).read())


@pytest.mark.parametrize("flag", ["-v", "-vvv"])
def test_m_cap_verbose_line_matches_the_live_reference(ref_bin, flag):
    """-v with -m prints "max count reached at R reads, K k-mers, C total counts, and B total bases" the moment the
    cap is crossed (src/FingerPrint.hpp:477-484).  Compared with the reference binary run right here on the same
    fixture; R only moves under -vvv upstream (:70-72), so -v prints 0 reads there and here."""
    d, opts, files = _case_files("panel300_m1")
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    ref = subprocess.run([ref_bin, flag] + argv, cwd=d, capture_output=True)
    ours = subprocess.run([NTSMCOUNT, flag] + argv, cwd=d, capture_output=True)
    assert ref.returncode == 0 and ours.returncode == 0
    assert ours.stdout == ref.stdout
    line = lambda t: [l.strip() for l in t.decode().splitlines() if l.startswith("max count reached")]
    assert len(line(ref.stderr)) == 1
    want, got = line(ref.stderr)[0], line(ours.stderr)[0]
    assert got == want
    order = [l.split()[0] for l in ours.stderr.decode().splitlines() if l.startswith(("Opening", "max", "Reached"))]
    assert order == ["Opening", "Opening", "max", "Reached"]


def test_m_cap_exact_vs_oracle_many_caps(oracle, tmp_path):
    """computeCounts with a cap, one thread: for a spread of caps and batch sizes the counts equal the
    oracle's read-by-read stop (src/FingerPrint.hpp:473-488), per k-mer."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(19)
    reads = _reads(rng, _windows(sites), 3000) + [rng.choice(_windows(sites)).encode() * 40 for _ in range(30)]   # some long reads
    rng.shuffle(reads)
    f = tmp_path / "reads.fa"
    f.write_bytes(b"".join(b">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    n_early = 0
    for cov, bb in ((0.05, 8192), (0.3, 8192), (1.0, 50000), (2.0, 1 << 20), (3.5, 12345), (1000.0, 8192)):
        fp = ntsm_b200.FingerPrint(sites, cov_thresh=cov, batch_bases=bb)
        fp.computeCounts([str(f)], threads=1)
        o = oracle.fingerprint(sites, 19, False, cov)
        o.count_file(str(f))
        assert fp.counts_text() == o.counts_text(), (cov, bb)
        assert fp.printInfoSummary() == o.summary(), (cov, bb)
        assert np.array_equal(fp.kmer_counts(), o.lists()[2]), (cov, bb)
        assert bool(fp.early_term) == o.early_term, (cov, bb)
        n_early += o.early_term
        fp.close()
    assert n_early >= 3


def test_m_cap_crossed_inside_a_read_longer_than_a_batch(oracle, tmp_path):
    """-m with contig-sized reads: the cap is crossed by a piece of a read that spans several batches.  The
    reference counts a whole read before it looks at the cap (processSingleRead, src/FingerPrint.hpp:473-487),
    so the rest of that read is still counted -- and nothing after it."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(23)
    wins = _windows(sites)

    def contig(n_parts):
        parts = []
        for _ in range(n_parts):
            parts.append("".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 300))))
            parts.append(rng.choice(wins))
        return "".join(parts).encode()

    reads = [contig(3) for _ in range(20)] + [contig(400)] + [contig(3) for _ in range(20)] + [contig(700)] + [contig(5) for _ in range(50)]
    assert max(map(len, reads)) > 100000
    f = tmp_path / "contigs.fa"
    f.write_bytes(b"".join(b">c%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    n_early = 0
    for cov, bb in ((0.1, 8192), (0.25, 8192), (0.25, 20000), (0.6, 8192), (0.9, 16384), (50.0, 8192)):
        fp = ntsm_b200.FingerPrint(sites, cov_thresh=cov, batch_bases=bb)
        fp.computeCounts([str(f)], threads=1)
        o = oracle.fingerprint(sites, 19, False, cov)
        o.count_file(str(f))
        assert fp.printInfoSummary() == o.summary(), (cov, bb)
        assert fp.counts_text() == o.counts_text(), (cov, bb)
        assert np.array_equal(fp.kmer_counts(), o.lists()[2]), (cov, bb)
        assert bool(fp.early_term) == o.early_term, (cov, bb)
        n_early += o.early_term
        fp.close()
    assert n_early >= 4


def test_m_cap_prefix_property(oracle, tmp_path):
    """-m, the weaker property that also holds where the stop stays batch-granular (several parser
    threads, where the reference itself is racy): whatever prefix of the file was consumed, the
    counts must equal the oracle's counts on exactly that prefix, the cap must have been exceeded by
    it, and the prefix is at most a batch longer than the reference's."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(9)
    reads = _reads(rng, _windows(sites), 4000)
    f = tmp_path / "reads.fa"
    f.write_bytes(b"".join(b">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    fp = ntsm_b200.FingerPrint(sites, cov_thresh=1.0, batch_bases=8192)
    fp.computeCounts([str(f)], threads=1)
    assert fp.early_term
    mr, mv, sr, sv, t = fp.finalize()
    ofp = oracle.fingerprint(sites, 19)
    n = 0
    prev_hits = 0
    while ofp.total_bases < int(t[2]):
        prev_hits = ofp.total_counts
        ofp.insert(reads[n]); n += 1
    assert ofp.total_bases == int(t[2]) and n < len(reads)         # a strict prefix of whole reads
    assert fp.counts_text() == ofp.counts_text()
    assert ofp.total_counts > fp.max_counts                         # cap exceeded ...
    assert int(t[2]) - 0 <= (ofp.total_bases)                       # ... by this prefix
    # the reference (read-granular) would have stopped within the last batch we consumed
    o2 = oracle.fingerprint(sites, 19, False, 1.0)
    o2.count_file(str(f))
    assert o2.total_bases <= int(t[2]) <= o2.total_bases + 2 * 8192
    fp.close()


# ---------------------------------------------------------------- BASELINE configs 3 and 5, scaled
def _synthetic_panel(path, n_sites, seed):
    """cfg5-style panel: random 31-mers, centre base A/T vs C/G, all 13 k-mers per allele joined by N;
    sites whose k-mers collide with earlier ones are rejected (the reference aborts on duplicates)."""
    rng = np.random.default_rng(seed)
    seen = set()
    out = []
    wins = []
    n = 0
    comp = str.maketrans("ACGT", "TGCA")
    while n < n_sites:
        w = "".join("ACGT"[c] for c in rng.integers(0, 4, 31))
        at, cg = ("A" if rng.random() < 0.5 else "T"), ("C" if rng.random() < 0.5 else "G")
        alleles = [w[:15] + at + w[16:], w[:15] + cg + w[16:]]
        kms = [[a[j:j + 19] for j in range(13)] for a in alleles]
        canon = {min(k, k.translate(comp)[::-1]) for ks in kms for k in ks}
        if len(canon) != 26 or canon & seen:
            continue
        seen |= canon
        out.append(">s%d ref\n%s\n>s%d var\n%s\n" % (n, "N".join(kms[0]), n, "N".join(kms[1])))
        wins.extend(alleles)
        n += 1
    with open(path, "w") as fh:
        fh.write("".join(out))
    return wins


def test_cfg5_large_synthetic_panel_vs_oracle(oracle, tmp_path):
    """A 60 000-site synthetic panel (1.56 M k-mers, more than the human panel) built the cfg5 way."""
    sites = str(tmp_path / "sites.fa")
    wins = _synthetic_panel(sites, 60000, 5)
    rng = random.Random(55)
    reads = []
    for _ in range(30000):
        r = "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 60))) + rng.choice(wins) + \
            "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 60)))
        if rng.random() < 0.3:
            p = rng.randrange(len(r)); r = r[:p] + rng.choice("ACGTN") + r[p + 1:]
        reads.append(revcomp(r.encode()) if rng.random() < 0.5 else r.encode())
    ofp = _check_against_oracle(oracle, sites, reads)
    assert ofp.table_size == 60000 * 26 and ofp.total_counts > 300000


def test_cfg3_ont_like_long_reads_vs_oracle(oracle):
    """ONT-like: log-normal lengths (N50 ~ 20 kb), 5-10 % errors split sub/ins/del, N runs."""
    rng = np.random.default_rng(3)
    prng = random.Random(3)
    wins = _windows(PANEL, limit=4000)
    reads = []
    for _ in range(150):
        L = int(np.clip(rng.lognormal(9.6, 0.8), 200, 200000))
        s = []
        while sum(map(len, s)) < L:
            s.append("".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(20, 800)))))
            s.append(prng.choice(wins))
        seq = np.frombuffer("".join(s)[:L].encode(), np.uint8).copy()
        err = rng.uniform(0.05, 0.10)
        u = rng.random(len(seq))
        sub = u < err / 3
        seq[sub] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(sub.sum()))]
        keep = ~((u >= 2 * err / 3) & (u < err))                 # deletions
        ins = (u >= err / 3) & (u < 2 * err / 3)                # insertion after the base
        pieces = np.stack([seq, np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(seq))]], 1)
        mask = np.stack([keep, ins & keep], 1)
        r = pieces[mask].tobytes().decode()
        if rng.random() < 0.1:
            p = int(rng.integers(0, len(r))); r = r[:p] + "N" * int(rng.geometric(1 / 50)) + r[p:]
        reads.append(revcomp(r.encode()) if rng.random() < 0.5 else r.encode())
    assert max(map(len, reads)) > 50000
    _check_against_oracle(oracle, PANEL, reads, batch_bases=1 << 16)


# ---------------------------------------------------------------- kernel variants
@pytest.mark.parametrize("opts", [{"kernel": 0}, {"kernel": 1}, {"kernel": 1, "launch_shape": 0}, {"kernel": 1, "launch_shape": 2},
                                  {"kernel": 1, "launch_shape": 3}, {"kernel": 1, "l2_persist": 1}, {"kernel": 1, "filter_bits": 20},
                                  {"kernel": 2}, {"kernel": 2, "filter_bits": 22}])
def test_every_k19_kernel_variant_vs_oracle(oracle, opts):
    """ntsm_ctx_set_option picks the count kernel (generic / paired seeds), its launch shape, the L2
    access-policy window and the k-mer bitmap size: each must give the oracle's per-k-mer counters."""
    rng = random.Random(1900 + sum(opts.values()))
    wins = _windows(PANEL, limit=8000)
    reads = _reads(rng, wins, 6000, "ACGTN") + [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 400))) for _ in range(2000)]
    _check_against_oracle(oracle, PANEL, reads, batch_bases=1 << 18, options=opts)


@pytest.mark.parametrize("kernel", [1, 2])
def test_pair_kernel_dense_hits_and_every_phase(oracle, kernel):
    """Stress for count_kernel_pair and count_kernel_wide (pair.cuh): reads that are nothing but site windows back to back
    (nearly every seed marked, so the pooled tail runs several rounds per warp), read lengths that
    walk the read starts through every chunk phase and lane, N runs right before and after the
    seed positions, and batches of odd sizes so the last group is ragged."""
    rng = random.Random(77)
    wins = _windows(PANEL, limit=6000)
    reads = []
    for n in range(2500):
        parts = [rng.choice(wins) for _ in range(rng.randrange(1, 8))]
        r = "".join(parts)
        if n % 3 == 0:                      # an N every few bases around a window boundary
            p = rng.randrange(len(r))
            r = r[:p] + "N" * rng.randrange(1, 7) + r[p:]
        if n % 5 == 0:
            r = r[rng.randrange(0, 6):]
        r = r.encode()
        reads.append(revcomp(r) if rng.random() < 0.5 else r)
    reads += [wins[i % len(wins)][: 19 + i % 13].encode() for i in range(600)]        # 19..31-base reads: 1..13 windows each
    for bb in (1 << 12, 40001, 1 << 20):
        ofp = _check_against_oracle(oracle, PANEL, reads, batch_bases=bb, n_buffers=3, options={"kernel": kernel})
    assert ofp.total_counts > 3 * len(reads)


@pytest.mark.parametrize("fold", [0, 1, 2, 4])
def test_pair_table_fold_vs_oracle(oracle, fold):
    """The paired-seed table folded 2^fold : 1 (pair.cuh): a smaller level-1 table only lets more
    windows through to the exact path, so the counters must not change."""
    rng = random.Random(2100 + fold)
    wins = _windows(PANEL, limit=8000)
    reads = _reads(rng, wins, 5000, "ACGTN") + [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 400))) for _ in range(1500)]
    _check_against_oracle(oracle, PANEL, reads, batch_bases=1 << 18, options={"kernel": 1, "pair_fold": fold})


# ---------------------------------------------------------------- reads packed on the device
def _pinned_bytes(data: bytes):
    import torch
    t = torch.empty(max(1, len(data)), dtype=torch.uint8).pin_memory()
    t[:len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8) if data else t[:0]
    return t


@pytest.mark.parametrize("threads", [0, 2])
def test_device_packed_reads_vs_oracle(oracle, threads):
    """ntsm_insert_reads / ntsm_insert_reads_fixed from PAGE-LOCKED memory: a feeder thread DMAs the ASCII
    bytes as they are and pack_ascii_kernel (devpack.cuh) decodes + packs them on the GPU; threads = 0 is
    the device packer alone, threads = 2 races it against two host packers on the same queue of blocks.
    Same counts as the oracle, for ragged reads (incl. empty ones, N runs, lower case, U, raw 0-3 bytes,
    0xFF, a read longer than a batch) and for a strided matrix."""
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(4242)
    wins = _windows(sites)
    reads = _reads(rng, wins, 5000, "ACGTN") + [b"", b"A", b"N" * 40, b"", wins[0].encode() * 300, wins[1].lower().encode(),
                                                 wins[2].replace("T", "U").encode(), bytes("ACGT".index(c) for c in wins[3]),
                                                 wins[4][:10].encode() + b"\xff" + wins[4][10:].encode(), b""]
    ofp = oracle.fingerprint(sites, 19, False)
    for r in reads:
        ofp.insert(r)
    buf = _pinned_bytes(b"".join(reads))
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    both = {"device_pack": 1} if threads else {}          # threads = 2: force packers AND feeders (the default with 2 packers is feeders alone)
    for bb in (1 << 14, 5000, 1 << 20):
        fp = ntsm_b200.FingerPrint(sites, batch_bases=bb, n_buffers=threads + 3, options=both)
        l0 = fp.launches
        check_rc = ntsm_b200._lib.lib().ntsm_insert_reads((ntsm_b200._lib.C.c_void_p * 1)(fp._ctx), 1, buf.data_ptr(), off.ctypes.data,
                                                          len(reads), threads)
        assert check_rc == 0, ntsm_b200._lib.lib().ntsm_last_error(None)
        assert fp.counts_text() == ofp.counts_text() and fp.printInfoSummary() == ofp.summary()
        assert fp.launches > l0
        fp.close()
    # dense matrix with a row stride, in pinned memory
    L_, stride, n = 151, 160, 4000
    mat = np.full((n, stride), ord("N"), np.uint8)
    ofp2 = oracle.fingerprint(sites, 19, False)
    for i in range(n):
        w = rng.choice(wins).encode()
        r = (bytes(rng.choice(b"ACGT") for _ in range(L_)) + w)[-L_:] if i % 3 else bytes(rng.choice(b"ACGTN") for _ in range(L_))
        mat[i, :L_] = np.frombuffer(r, np.uint8)
        ofp2.insert(r)
    pm = _pinned_bytes(mat.tobytes())
    for bb, stride_, ptr in ((1 << 15, stride, pm.data_ptr()), (3000, stride, pm.data_ptr())):
        fp = ntsm_b200.FingerPrint(sites, batch_bases=bb, n_buffers=threads + 3, options=both)
        fp.insertReadsFixed(ptr, L_, stride_, n, threads=threads)
        assert fp.counts_text() == ofp2.counts_text() and fp.printInfoSummary() == ofp2.summary()
        fp.close()
    # the same bytes from pageable memory with device_pack off: host packers only, same answer
    fp = ntsm_b200.FingerPrint(sites, batch_bases=1 << 15, n_buffers=4, options={"device_pack": 0})
    fp.insertReadsFixed(pm.data_ptr(), L_, stride, n, threads=max(1, threads))
    assert fp.counts_text() == ofp2.counts_text()
    fp.close()


def test_host_register_makes_a_buffer_device_packable(oracle):
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(5)
    wins = _windows(sites)
    L_, n = 150, 3000
    mat = np.empty((n, L_), np.uint8)
    ofp = oracle.fingerprint(sites, 19, False)
    for i in range(n):
        r = (bytes(rng.choice(b"ACGT") for _ in range(L_)) + rng.choice(wins).encode())[-L_:]
        mat[i] = np.frombuffer(r, np.uint8)
        ofp.insert(r)
    Lb = ntsm_b200._lib.lib()
    assert Lb.ntsm_host_register(mat.ctypes.data, mat.nbytes) == 0
    try:
        fp = ntsm_b200.FingerPrint(sites, batch_bases=1 << 16)
        l0 = fp.launches
        h0 = fp.pcie_bytes[0]
        fp.insertReadsFixed(mat.ctypes.data, L_, L_, n, threads=0)      # threads = 0: only the device packer can do the work
        assert fp.counts_text() == ofp.counts_text()
        assert fp.launches - l0 >= 2
        assert fp.pcie_bytes[0] - h0 == mat.nbytes                       # every base crossed PCIe as one ASCII byte, nothing else did
        fp.close()
    finally:
        assert Lb.ntsm_host_unregister(mat.ctypes.data) == 0


# ---------------------------------------------------------------- parsers as worker processes
@pytest.mark.parametrize("name", [n for n in golden_cases() if not n.startswith(("dupes_abort", "odd_sites"))])
def test_abi_files_through_parser_processes_match_reference_fixture(name):
    """computeCounts with the parsers run as worker PROCESSES (bin/ntsm_parse_worker over the shared, page-locked
    mapping of procpipe.h; forced here, automatic beyond 6 parser threads on plain files): every fixture made by the
    reference binary, byte for byte.  gzip inputs and -m runs fall back to the parser threads by design."""
    d, opts, files = _case_files(name)
    if opts["m"] > 0 and "Reached desired" in open(os.path.join(d, "stderr.txt")).read():
        pytest.skip("-m early stop: the cap keeps the parsers in-process")
    fp = ntsm_b200.FingerPrint(opts["sites"], k=opts["k"], dupes=opts["dupes"], cov_thresh=opts["m"], batch_bases=1 << 14,
                               options={"parser_procs": 1})
    fp.computeCounts(files, threads=max(2, opts["t"]))
    assert fp.counts_text().encode() == open(os.path.join(d, "stdout.txt"), "rb").read()
    assert fp.printInfoSummary() in open(os.path.join(d, "stderr.txt")).read()
    fp.close()


def test_parser_processes_many_files_long_reads_vs_oracle(oracle, tmp_path):
    """Nine plain files (FASTA and FASTQ, reads up to 40 kb, N runs) over 1..8 parser processes and batches from 4 Ki to
    1 Mi positions, two GPUs' worth of contexts: the oracle's per-k-mer counters; then a missing file: the reference's
    error text, and the contexts stay usable."""
    import torch
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(61)
    wins = _windows(sites)
    ofp = oracle.fingerprint(sites, 19, False)
    paths = []
    for f in range(9):
        recs = []
        for i in range(rng.randrange(150, 400)):
            parts = []
            for _ in range(rng.choice([1, 1, 2, 5, 120 if f == 4 else 3])):
                parts.append("".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 300))))
                parts.append(rng.choice(wins))
                if rng.random() < 0.1:
                    parts.append("N" * rng.randrange(1, 9))
            r = "".join(parts).encode()
            r = revcomp(r) if rng.random() < 0.5 else r
            ofp.insert(r)
            recs.append((b"@q%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r))) if f % 2 else (b">q%d\n%s\n" % (i, r)))
        p = tmp_path / ("in%d.%s" % (f, "fq" if f % 2 else "fa"))
        p.write_bytes(b"".join(recs))
        paths.append(str(p))
    ndev = torch.cuda.device_count()
    ss = ntsm_b200.SiteSet(sites, 19)
    for threads, bb, n_ctx in ((1, 1 << 14, 1), (4, 4096, 1), (8, 1 << 20, 1), (3, 30000, 2)):
        fps = [ntsm_b200.FingerPrint(ss, device=i % ndev, batch_bases=bb, n_buffers=4, options={"parser_procs": 1}) for i in range(n_ctx)]
        L = ntsm_b200.lib()
        arr = (ntsm_b200._lib.C.c_char_p * len(paths))(*[os.fsencode(x) for x in paths])
        ctxs = (ntsm_b200._lib.C.c_void_p * n_ctx)(*[f._ctx for f in fps])
        early = ntsm_b200._lib.C.c_int(0)
        assert L.ntsm_count_files(ctxs, n_ctx, arr, len(paths), threads, 0, ntsm_b200._lib.C.byref(early)) == 0
        ntsm_b200.FingerPrint.group_finalize(fps)
        assert fps[0].counts_text() == ofp.counts_text(), (threads, bb, n_ctx)
        assert fps[0].printInfoSummary() == ofp.summary()
        assert np.array_equal(fps[0].kmer_counts(), ofp.lists()[2])
        for f in fps:
            f.close()
    fp = ntsm_b200.FingerPrint(ss, batch_bases=1 << 15, options={"parser_procs": 1})
    with pytest.raises(FileNotFoundError) as e:
        fp.computeCounts(paths[:3] + [str(tmp_path / "nope.fq")] + paths[3:], threads=3)
    assert "file %s cannot be opened" % (tmp_path / "nope.fq") in str(e.value)
    fp.reset()
    fp.computeCounts(paths, threads=2)
    assert fp.counts_text() == ofp.counts_text()
    fp.close()


# ---------------------------------------------------------------- several GPUs of one process, no NCCL
@pytest.mark.parametrize("n_ctx", [2, 3])
def test_group_finalize_equals_one_context(oracle, n_ctx):
    """ntsm_group_finalize: ctx 0 sums the other ctxs' counts out of (peer) device memory inside the per-site
    reduce kernel.  Reads dealt over n_ctx contexts -- on as many GPUs as the box has, wrapping around, so
    a one-GPU box runs the same kernel over plain device pointers -- must give the oracle's counts file,
    summary and per-k-mer counters for ALL the reads (sums before max: src/FingerPrint.hpp:277-294)."""
    import torch
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    rng = random.Random(31 + n_ctx)
    reads = _reads(rng, _windows(sites), 6000, "ACGTN")
    ofp = oracle.fingerprint(sites, 19, False)
    for r in reads:
        ofp.insert(r)
    ndev = torch.cuda.device_count()
    ss = ntsm_b200.SiteSet(sites, 19)
    fps = [ntsm_b200.FingerPrint(ss, device=i % ndev, batch_bases=1 << 14) for i in range(n_ctx)]
    for i, r in enumerate(reads):
        fps[(i * 7) % n_ctx].insertCount(r)
    rows = ntsm_b200.FingerPrint.group_finalize(fps)
    assert (int(rows[4][0]), int(rows[4][1]), int(rows[4][2])) == (ofp.total_kmers, ofp.total_counts, ofp.total_bases)
    assert fps[0].counts_text() == ofp.counts_text() and fps[0].printInfoSummary() == ofp.summary()
    _, _, ocnt = ofp.lists()
    assert np.array_equal(fps[0].kmer_counts(), ocnt)
    # a per-GPU max summed afterwards is NOT the same thing (the ntsmEval --merge trap): make sure the reads were really split
    assert all(int(f.poll_totals()[2]) > 0 for f in fps)
    for f in fps:
        f.close()


# ---------------------------------------------------------------- the counts file's consumer
def test_counts_file_is_accepted_by_the_reference_ntsmEval(tmp_path):
    """the file our binary prints goes through the reference's own consumer
    (ntsmEval, src/CompareCounts.hpp:30-114 parses #@TK/#@KS and the six columns): single-sample QC
    (cov, errorRate, miss, hom, het) equals what it reports for the reference's file, and the two
    files compare as the same sample."""
    evalbin = os.path.join(ROOT, "oracle", "_ref", "ntsmEval")
    if not os.path.exists(evalbin):
        pytest.skip("oracle/_ref/ntsmEval not built (make -C oracle ref, where /root/reference exists)")
    d, opts, files = _case_files("gz_single")
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    p = subprocess.run([NTSMCOUNT] + argv, cwd=d, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    ours, ref = tmp_path / "ours.txt", tmp_path / "ref.txt"
    ours.write_bytes(p.stdout)
    ref.write_bytes(open(os.path.join(d, "stdout.txt"), "rb").read())

    def qc(path):
        o = subprocess.run([evalbin, str(path)], capture_output=True, text=True)
        assert o.returncode == 0, o.stderr
        rows = [l.split("\t") for l in (o.stdout + o.stderr).splitlines() if l.startswith(str(path))]
        return [f.split("Time:")[0] for f in rows[0][1:]]

    assert qc(ours) == qc(ref) and len(qc(ours)) == 5
    o = subprocess.run([evalbin, "-a", str(ours), str(ref)], capture_output=True, text=True)
    assert o.returncode == 0, o.stderr
    hdr, row = [l.split("\t") for l in o.stdout.splitlines() if l.startswith(("sample1", str(ours)))][:2]
    rec = dict(zip(hdr, row))
    assert rec["same"] == "1" and float(rec["relate"]) == 1.0 and rec["ibs0"] == "0"
