"""The oracle (oracle/ntsm_oracle.c) pinned against the reference's own behaviour:
known-answer vectors, iterator dumps produced by the reference header,
stdout/stderr fixtures produced by the real reference binary, and -- when it is built --
the live reference binary on fresh random inputs."""
import json
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, golden_cases

HASH_KAT = {  # vendor/KseqHashIterator.hpp:129, executed in the survey container
    0x0: 0x1df06f29bc, 0x1: 0x29b794f8ce, 0x2: 0x3f6f2a0674,
    0x3fffffffff: 0x1c5d2677be, 0x0123456789: 0x0a635aa6a1, 0x2aaaaaaaaa: 0x1a2ccac738,
}


def test_hash64_kat(oracle):
    for x, h in HASH_KAT.items():
        assert oracle.hash64(x, 19) == h


def test_nt4_table(oracle):
    valid = {ord(c): v for c, v in zip("ACGTUacgtu", [0, 1, 2, 3, 3, 0, 1, 2, 3, 3])}
    valid.update({0: 0, 1: 1, 2: 2, 3: 3})
    for b in range(256):
        assert oracle.nt4(b) == valid.get(b, 4), b


def test_iterator_kat(oracle):
    got = oracle.iter(b"ACGTTGCATGCATGCAAGCTNACGTTGCATGCATGCAAGCTT", 19)
    assert [(p, fw, rv, h) for p, h, fw, rv in got] == [
        (19, 0x06f9393909, 0x27e4e4e41b, 0x1568c69424), (20, 0x1be4e4e427, 0x09f9393906, 0x09b5ca278b),
        (40, 0x06f9393909, 0x27e4e4e41b, 0x1568c69424), (41, 0x1be4e4e427, 0x09f9393906, 0x09b5ca278b),
        (42, 0x2f9393909f, 0x027e4e4e41, 0x1a36e7e7d5)]


def _unescape(s):
    out = bytearray(); i = 0
    while i < len(s):
        if s[i] == "\\" and s[i + 1] == "x":
            out.append(int(s[i + 2:i + 4], 16)); i += 4
        else:
            out.append(ord(s[i])); i += 1
    return bytes(out)


def test_iterator_vectors_from_reference_header(oracle):
    n = 0
    with open(os.path.join(GOLDEN, "iter_vectors.tsv")) as fh:
        for line in fh:
            k, seq, kmers = line.rstrip("\n").split("\t")
            want = [tuple(int(x, 16) if i else int(x) for i, x in enumerate(t.split(":"))) for t in kmers.split(",") if t]
            got = [(p, h) for p, h, _, _ in oracle.iter(_unescape(seq), int(k))]
            assert got == want, (k, seq)
            n += 1
    assert n > 80


KEEP = ("Warning: ", "Reached desired", "Total ", "Distinct ", "Sites Covered")


def _filter_err(text):
    return [l for l in text.splitlines() if l.startswith(KEEP) and not l.startswith("Warning: site coverage")]


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_vs_reference_fixture(oracle_build, name):
    d = os.path.join(GOLDEN, "cases", name)
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    p = subprocess.run([os.path.join(oracle_build, "ntsm_oracle")] + argv, cwd=d, capture_output=True)
    ref_rc = int(open(os.path.join(d, "rc.txt")).read())
    ref_out = open(os.path.join(d, "stdout.txt"), "rb").read()
    ref_err = open(os.path.join(d, "stderr.txt")).read()
    if ref_rc == 0:
        assert p.returncode == 0
        assert p.stdout == ref_out
    else:  # reference died on std::out_of_range (SIGABRT); stdout is whatever had been flushed
        assert p.returncode == 134
    assert _filter_err(p.stderr.decode()) == _filter_err(ref_err)


def _rand_fastx(rng, windows):
    """A deliberately messy FASTA/FASTQ file."""
    out = []
    for i in range(rng.randrange(1, 60)):
        w = rng.choice(windows)
        s = "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 40))) + w + \
            "".join(rng.choice("ACGTN") for _ in range(rng.randrange(0, 40)))
        if rng.random() < 0.3:
            s = s.lower()
        if rng.random() < 0.2:
            p = rng.randrange(len(s)); s = s[:p] + rng.choice(["N", "R", "-", " ", "\t", "+", ">", "@", "\r"]) + s[p:]
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        wrap = rng.choice([0, 0, 7, 30, 61])
        lines = [s[j:j + wrap] for j in range(0, len(s), wrap)] if wrap else [s]
        if rng.random() < 0.5:
            q = "".join(rng.choice("IJ@+>#5") for _ in s)
            if rng.random() < 0.1:
                q = q[:-1]
            qlines = [q[j:j + wrap] for j in range(0, len(q), wrap)] if wrap else [q]
            out.append("@r%d c" % i + eol + eol.join(lines) + eol + "+" + eol + eol.join(qlines) + eol)
        else:
            out.append(">r%d" % i + eol + eol.join(lines) + (eol if rng.random() < 0.95 else ""))
        if rng.random() < 0.1:
            out.append(eol)
    text = "".join(out)
    if rng.random() < 0.2:
        text = text[:rng.randrange(len(text) + 1)]
    return text


def test_oracle_vs_live_reference_fuzz(oracle_build, ref_bin, tmp_path):
    """Differential fuzz of the whole CPU path (parser + iterator + table + printer)."""
    rng = random.Random(99)
    sites = os.path.join(GOLDEN, "shared", "sites300.fa")
    wins = []
    with open(sites) as fh:
        for line in fh:
            if not line.startswith(">"):
                wins.extend(line.strip().split("N")[:2])
    for it in range(60):
        files = []
        for j in range(rng.randrange(1, 4)):
            f = tmp_path / ("f%d_%d.fx" % (it, j))
            f.write_text(_rand_fastx(rng, wins), newline="")
            files.append(str(f))
        argv = ["-s", sites] + files
        if rng.random() < 0.3:
            argv = ["-m", rng.choice(["0.01", "0.05", "0.2"])] + argv
        a = subprocess.run([ref_bin] + argv, capture_output=True)
        b = subprocess.run([os.path.join(oracle_build, "ntsm_oracle")] + argv, capture_output=True)
        assert a.returncode == 0 and b.returncode == 0
        assert a.stdout == b.stdout, argv
        assert _filter_err(a.stderr.decode()) == _filter_err(b.stderr.decode())
