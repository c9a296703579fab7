"""Exact shard merge at the k-mer level: `ntsmCount --dump-kmer-counts` per shard,
`ntsmCount --merge-kmer-counts` over the dumps == counting the concatenated input, byte for byte -- which
the reference's own `ntsmEval --merge` (sum of per-site maxima, src/CompareCounts.hpp:626-674, the bug
at :648-657) is not."""
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

NTSMCOUNT = os.path.join(ROOT, "ntsm_b200", "bin", "ntsmCount")
SITES = os.path.join(GOLDEN, "shared", "sites300.fa")
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _windows():
    wins = []
    for line in open(SITES):
        if not line.startswith(">"):
            wins.extend(line.strip().split("N"))
    return wins


def _shards(tmp_path, n_shards=3, n_reads=1500):
    """Shards whose reads favour DIFFERENT k-mers of the same sites, so that per-site maxima do not add up."""
    wins = _windows()
    paths = []
    for s in range(n_shards):
        rng = random.Random(100 + s)
        p = tmp_path / ("shard%d.fa" % s)
        with open(p, "w") as fh:
            for i in range(n_reads):
                w = wins[(i * 7 + s) % len(wins)]
                lo = (s * 4) % max(1, len(w) - 19)            # each shard samples a different stretch of the window
                r = w[lo:lo + 19 + rng.randrange(0, 5)]
                r = "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 20))) + r + "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 20)))
                b = r.encode()
                if rng.random() < 0.5:
                    b = b.translate(COMP)[::-1]
                fh.write(">r%d\n%s\n" % (i, b.decode()))
        paths.append(str(p))
    return paths


def _rows(text):
    out = {}
    for line in text.decode().splitlines():
        if not line.startswith("#"):
            f = line.split("\t")
            out[f[0]] = [int(x) for x in f[1:5]]
    return out


def test_kmer_count_file_format_is_rejected_when_foreign(tmp_path):
    """CPU-side: the loader's checks need no GPU up to the point where a ctx would be touched."""
    import ntsm_b200
    L = ntsm_b200.lib()
    ss = ntsm_b200.SiteSet(SITES, 19)
    bad = tmp_path / "bad.kc"
    bad.write_bytes(b"NOTAKMERCOUNTFILE" * 4)
    # a null ctx is refused before any file is read; a foreign file is refused by its magic
    assert L.ntsm_counts_load_add(None, ss._h, str(bad).encode()) == -1
    assert L.ntsm_counts_save(None, ss._h, str(bad).encode()) == -1


@pytest.mark.gpu
def test_merge_of_shard_dumps_equals_counting_everything(tmp_path):
    shards = _shards(tmp_path)
    whole = subprocess.run([NTSMCOUNT, "-s", SITES] + shards, capture_output=True)
    assert whole.returncode == 0, whole.stderr.decode()
    dumps, per_shard = [], []
    for i, s in enumerate(shards):
        d = str(tmp_path / ("shard%d.kc" % i))
        p = subprocess.run([NTSMCOUNT, "-s", SITES, "--dump-kmer-counts", d, s], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()
        per_shard.append(p.stdout)
        dumps.append(d)
    merged = subprocess.run([NTSMCOUNT, "-s", SITES, "--merge-kmer-counts"] + dumps, capture_output=True)
    assert merged.returncode == 0, merged.stderr.decode()
    assert merged.stdout == whole.stdout                                         # the counts file, byte for byte
    keep = lambda t: [l for l in t.decode().splitlines() if l.startswith(("Total ", "Distinct ", "Sites Covered"))]
    assert keep(merged.stderr) == keep(whole.stderr)                             # and the summary: bases, k-mers, hits
    # merging in another order, and merging a merge, changes nothing
    m2 = str(tmp_path / "m01.kc")
    p = subprocess.run([NTSMCOUNT, "-s", SITES, "--merge-kmer-counts", "--dump-kmer-counts", m2, dumps[1], dumps[0]], capture_output=True)
    assert p.returncode == 0
    again = subprocess.run([NTSMCOUNT, "-s", SITES, "--merge-kmer-counts", dumps[2], m2], capture_output=True)
    assert again.stdout == whole.stdout
    # what `ntsmEval --merge` does instead -- add the counts FILES, i.e. sum the per-site maxima -- is a different answer
    want = _rows(whole.stdout)
    summed = {}
    for t in per_shard:
        for name, v in _rows(t).items():
            acc = summed.setdefault(name, [0, 0, 0, 0])
            for j in range(4):
                acc[j] += v[j]
    assert all(summed[n][2:] == want[n][2:] for n in want)                      # the sums do add up ...
    assert any(summed[n][0] > want[n][0] or summed[n][1] > want[n][1] for n in want)     # ... the maxima do not
    assert all(summed[n][0] >= want[n][0] and summed[n][1] >= want[n][1] for n in want)


@pytest.mark.gpu
def test_merge_refuses_a_dump_of_another_panel(tmp_path):
    shards = _shards(tmp_path, n_shards=1, n_reads=50)
    d = str(tmp_path / "a.kc")
    assert subprocess.run([NTSMCOUNT, "-s", SITES, "--dump-kmer-counts", d, shards[0]], capture_output=True).returncode == 0
    other = os.path.join(GOLDEN, "cases", "mini", "sites.fa")
    p = subprocess.run([NTSMCOUNT, "-s", other, "--merge-kmer-counts", d], capture_output=True)
    assert p.returncode == 1 and b"another site set" in p.stderr
    open(d, "ab").write(b"x")
    p = subprocess.run([NTSMCOUNT, "-s", SITES, "--merge-kmer-counts", d], capture_output=True)
    assert p.returncode == 1 and b"truncated or oversized" in p.stderr


@pytest.mark.gpu
def test_add_counts_abi_vs_oracle(oracle):
    """ntsm_add_counts through the ABI: two contexts count halves, one takes the other's k-mer counts in."""
    import ntsm_b200
    rng = random.Random(3)
    wins = _windows()
    reads = [(rng.choice(wins) + "".join(rng.choice("ACGTN") for _ in range(rng.randrange(0, 30)))).encode() for _ in range(3000)]
    ofp = oracle.fingerprint(SITES, 19, False)
    for r in reads:
        ofp.insert(r)
    ss = ntsm_b200.SiteSet(SITES, 19)
    a, b = ntsm_b200.FingerPrint(ss, batch_bases=1 << 14), ntsm_b200.FingerPrint(ss, batch_bases=1 << 14)
    for i, r in enumerate(reads):
        (a if i % 2 else b).insertCount(r)
    L = ntsm_b200.lib()
    cnt = b.kmer_counts()
    tot = np.zeros(3, np.uint64)
    assert L.ntsm_get_totals(b._ctx, tot.ctypes.data) == 0
    assert L.ntsm_add_counts(a._ctx, np.ascontiguousarray(cnt).ctypes.data, tot.ctypes.data) == 0
    assert a.counts_text() == ofp.counts_text() and a.printInfoSummary() == ofp.summary()
    a.close(); b.close()
