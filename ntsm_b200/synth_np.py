"""numpy-only generators of the five BASELINE.json workloads, scaled by the caller.

These are bit-reproducible from their seed on any machine with the same numpy (PCG64), which is
what tools/config_parity.py needs: the inputs are generated once in the build container to run the
unmodified reference over them, and regenerated on the GPU box to run this library over the very
same bytes.  (ntsm_b200/synth.py is the torch/CUDA generator bench.py uses for the 100-Gbase shard;
CUDA and CPU random streams differ, so it cannot serve that purpose.)  Input generation only --
nothing here counts anything."""
import gzip
import os
import subprocess

import numpy as np

ASCII = np.frombuffer(b"ACGTN", np.uint8)
COMP_ASCII = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    COMP_ASCII[_a] = _b


# ----------------------------------------------------------------------------- panels
def synthetic_panel(path, n_sites, seed, k=19, flank=15):
    """cfg5 panel: random (2*flank+1)-mers with centre base A/T (ref record) vs C/G (var record),
    every k-mer that covers the centre listed and joined by N like the real panel; sites whose
    canonical k-mers collide with any other site's (or among themselves) are dropped, because the
    reference aborts on duplicates unless -d.  Returns (windows uint8 [n, 2, W] of base codes,
    n_sites kept)."""
    rng = np.random.default_rng(seed)
    W = 2 * flank + 1
    n_k = W - k + 1                                        # k-mers per allele (13 for W=31, k=19)
    assert n_k >= 1 and flank < k
    codes = rng.integers(0, 4, (n_sites, W), dtype=np.uint8)
    at = np.where(rng.random(n_sites) < 0.5, 0, 3).astype(np.uint8)       # A or T
    cg = np.where(rng.random(n_sites) < 0.5, 1, 2).astype(np.uint8)       # C or G
    win = np.stack([codes, codes.copy()], 1)               # [n, 2, W]
    win[:, 0, flank] = at
    win[:, 1, flank] = cg
    # canonical value of every listed k-mer: fw with the first base most significant, rv = revcomp
    fw = np.zeros((n_sites, 2, n_k), np.uint64)
    rv = np.zeros((n_sites, 2, n_k), np.uint64)
    for j in range(k):
        for o in range(n_k):
            c = win[:, :, o + j].astype(np.uint64)
            fw[:, :, o] = (fw[:, :, o] << np.uint64(2)) | c
            rv[:, :, o] |= (np.uint64(3) - c) << np.uint64(2 * j)
    canon = np.minimum(fw, rv).reshape(n_sites, -1)
    flat = canon.reshape(-1)
    order = np.argsort(flat, kind="stable")
    s = flat[order]
    dup = np.zeros(flat.size, bool)
    eq = s[1:] == s[:-1]
    dup[order[1:][eq]] = True
    dup[order[:-1][eq]] = True
    keep = ~dup.reshape(n_sites, -1).any(1)
    win = win[keep]
    n = win.shape[0]
    # FASTA text, vectorised: ">s<idx> ref\n" + k-mers joined by 'N' + "\n"
    body_len = n_k * k + (n_k - 1)
    rec = np.full((n, 2, body_len), ord("N"), np.uint8)
    for o in range(n_k):
        rec[:, :, o * (k + 1): o * (k + 1) + k] = ASCII[win[:, :, o: o + k]]
    with open(path, "wb") as fh:
        step = 100000
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            parts = []
            for i in range(lo, hi):
                parts.append(b">s%d ref\n" % i); parts.append(rec[i, 0].tobytes())
                parts.append(b"\n>s%d var\n" % i); parts.append(rec[i, 1].tobytes()); parts.append(b"\n")
            fh.write(b"".join(parts))
    return win, n


def panel_alleles_from_windows(win):
    """[n, 2, W] windows -> (codes [2n, W], lens [2n]) in the layout synth.panel_windows() uses."""
    n, _, W = win.shape
    return win.reshape(2 * n, W), np.full(2 * n, W, np.int64)


# ----------------------------------------------------------------------------- genome + reads
class Genome:
    """Two haplotypes of `size` random bases with site i's window at the same place on both and the
    allele (ref / var) drawn per haplotype."""

    def __init__(self, size, win_codes, win_lens, seed):
        rng = np.random.default_rng(seed)
        base = rng.integers(0, 4, size, dtype=np.uint8)
        self.size = size
        self.hap = np.stack([base, base.copy()])
        n_sites = win_codes.shape[0] // 2
        if n_sites:
            stride = size // n_sites
            assert stride >= 64, "genome too small for this panel"
            pos0 = np.arange(n_sites, dtype=np.int64) * stride + rng.integers(0, stride - 40, n_sites)
            for h in range(2):
                rows = np.arange(n_sites) * 2 + rng.integers(0, 2, n_sites)
                for ln in np.unique(win_lens[rows]):
                    sel = np.nonzero(win_lens[rows] == ln)[0]
                    idx = pos0[sel, None] + np.arange(ln)[None, :]
                    self.hap[h][idx] = win_codes[rows[sel], :ln]


def short_reads(genome, n_reads, read_len, err, seed, n_frac=0.005):
    """-> uint8 codes [n_reads, read_len] (0-3, 4 = N): uniform start/haplotype/strand, `err`
    substitutions, an N run (1-10) in n_frac of the reads."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_reads, read_len), np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    step = 1 << 18
    for lo in range(0, n_reads, step):
        n = min(step, n_reads - lo)
        hap = rng.integers(0, 2, n)
        start = rng.integers(0, genome.size - read_len, n)
        codes = genome.hap.reshape(-1)[(hap * genome.size + start)[:, None] + ar[None, :]]
        rc = rng.random(n) < 0.5
        codes = np.where(rc[:, None], 3 - codes[:, ::-1], codes)
        if err > 0:
            e = rng.random(codes.shape) < err
            sub = rng.integers(1, 4, codes.shape, dtype=np.uint8)
            codes = np.where(e, (codes + sub) & 3, codes)
        if n_frac > 0:
            has = rng.random(n) < n_frac
            p0 = rng.integers(0, read_len, n)
            ln = rng.integers(1, 11, n)
            inrun = has[:, None] & (ar[None, :] >= p0[:, None]) & (ar[None, :] < (p0 + ln)[:, None])
            codes = np.where(inrun, np.uint8(4), codes)
        out[lo:lo + n] = codes
    return out


def paired_reads(genome, n_pairs, read_len, err, seed, insert_mean=350, insert_sd=30):
    """Paired-end: R1 reads the fragment's forward strand from its start, R2 the reverse strand
    from its end.  -> (codes_r1, codes_r2), each [n_pairs, read_len]."""
    rng = np.random.default_rng(seed)
    ins = np.clip(rng.normal(insert_mean, insert_sd, n_pairs).astype(np.int64), read_len, 4 * insert_mean)
    hap = rng.integers(0, 2, n_pairs)
    start = rng.integers(0, genome.size - 4 * insert_mean - 1, n_pairs)
    ar = np.arange(read_len, dtype=np.int64)
    flat = genome.hap.reshape(-1)
    r1 = flat[(hap * genome.size + start)[:, None] + ar[None, :]]
    r2 = 3 - flat[(hap * genome.size + start + ins - 1)[:, None] - ar[None, :]]
    out = []
    for r in (r1, r2):
        e = rng.random(r.shape) < err
        sub = rng.integers(1, 4, r.shape, dtype=np.uint8)
        out.append(np.where(e, (r + sub) & 3, r).astype(np.uint8))
    return out[0], out[1]


def ont_reads(genome, total_bases, seed, mu=9.6, sigma=0.8, n_frac=0.02):
    """ONT-like long reads (cfg3): log-normal lengths clipped to 200 b - 500 kb (N50 about 20 kb),
    5-10 % error per read split evenly into substitutions, insertions and deletions, an N run
    (geometric, mean 50) in n_frac of the reads.  -> list of ASCII uint8 arrays."""
    rng = np.random.default_rng(seed)
    reads, done = [], 0
    while done < total_bases:
        ln = int(np.clip(rng.lognormal(mu, sigma), 200, 500000))
        ln = min(ln, genome.size - 1)
        hap = int(rng.integers(0, 2))
        st = int(rng.integers(0, genome.size - ln))
        src = genome.hap[hap, st:st + ln]
        if rng.random() < 0.5:
            src = 3 - src[::-1]
        rate = rng.uniform(0.05, 0.10)
        u = rng.random(ln)
        op = np.zeros(ln, np.uint8)                        # 0 keep, 1 substitute, 2 insert after, 3 delete
        op[u < rate] = 1
        op[u < 2 * rate / 3] = 2
        op[u < rate / 3] = 3
        sub = rng.integers(1, 4, ln, dtype=np.uint8)
        base = np.where(op == 1, (src + sub) & 3, src).astype(np.uint8)
        reps = np.where(op == 3, 0, np.where(op == 2, 2, 1))
        seq = np.repeat(base, reps)
        second = np.zeros(seq.size, bool)                  # the inserted copy is the second of each doubled base
        ends = np.cumsum(reps) - 1
        second[ends[op == 2]] = True
        seq[second] = rng.integers(0, 4, int(second.sum()), dtype=np.uint8)
        if rng.random() < n_frac and seq.size > 2:
            p0 = int(rng.integers(0, seq.size - 1))
            seq[p0:p0 + int(rng.geometric(1 / 50.0))] = 4
        reads.append(ASCII[seq])
        done += seq.size
    return reads


# ----------------------------------------------------------------------------- files
def write_fastq_matrix(codes, path, first_index=0, prefix=b"r"):
    """Fixed-length reads [n, L] -> one FASTQ file (constant quality), vectorised."""
    n, L = codes.shape
    rec = np.empty((n, 2 + 9 + 1 + L + 3 + L + 1), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1] = prefix[0]
    idx = np.arange(first_index, first_index + n, dtype=np.int64)
    for d in range(9):
        rec[:, 2 + d] = (idx // 10 ** (8 - d)) % 10 + 48
    rec[:, 11] = 10
    rec[:, 12:12 + L] = ASCII[codes]
    rec[:, 12 + L] = 10; rec[:, 13 + L] = ord("+"); rec[:, 14 + L] = 10
    rec[:, 15 + L:15 + 2 * L] = ord("I")
    rec[:, 15 + 2 * L] = 10
    rec.tofile(path)
    return path


def write_fastq_ragged(reads, path, prefix=b"ont"):
    with open(path, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b"@%s%d\n" % (prefix, i)); fh.write(r.tobytes()); fh.write(b"\n+\n"); fh.write(b"I" * r.size); fh.write(b"\n")
    return path


def gzip_files(paths, level=6):
    """gzip -<level> each file in parallel (keeps nothing but the .gz); returns the new paths."""
    procs = [subprocess.Popen(["gzip", "-%d" % level, "-n", "-f", p]) for p in paths]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("gzip failed")
    return [p + ".gz" for p in paths]


def read_panel_windows(path, limit_sites=None):
    """Windows of a real panel file (overlap-merge of each record's k-mers), cached by synth.panel_windows."""
    from . import synth
    return synth.panel_windows(path, limit_sites)
