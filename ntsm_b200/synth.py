"""Seeded synthetic workloads for tests and bench.py: a diploid genome with every
panel site's window planted once per haplotype, Illumina-like reads sampled from it, and the
library's packed 2-bit + N-mask layout produced directly with torch ops so a 100-Gbase shard
can be generated on the GPU in seconds.  Everything here is input generation -- nothing is counted."""
import gzip
import os

import numpy as np
import torch

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
ASCII = np.frombuffer(b"ACGTN", np.uint8)
TILE, HALO = 8192, 64           # ntsm_padded_positions() contract (pack.h)


def padded_positions(n_pos):
    return (n_pos + TILE - 1) // TILE * TILE + HALO


def merge_kmers(kmers):
    """Overlap-merge the sliding k-mers of one allele record back into its window."""
    w = kmers[0]
    for km in kmers[1:]:
        n = len(km)
        for d in range(1, n + 1):
            if d == n or w[-(n - d):] == km[:n - d]:
                w += km[n - d:]
                break
    return w


def panel_windows(path, limit_sites=None):
    """-> (codes uint8 [n_alleles, 32] padded with 255, lengths int64 [n_alleles]); allele 2i = ref, 2i+1 = var."""
    cache = "/tmp/ntsm_panel_%s_%s.npz" % (os.path.basename(path), limit_sites)
    if os.path.exists(cache) and os.path.getmtime(cache) > os.path.getmtime(path):
        z = np.load(cache)
        return z["codes"], z["lens"]
    op = gzip.open if path.endswith(".gz") else open
    wins = []
    with op(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                continue
            wins.append(merge_kmers(line.strip().upper().split("N")))
            if limit_sites and len(wins) >= 2 * limit_sites:
                break
    width = max(32, max(len(w) for w in wins))
    codes = np.full((len(wins), width), 255, np.uint8)
    lens = np.zeros(len(wins), np.int64)
    lut = np.full(256, 255, np.uint8)
    for ch, v in CODE.items():
        lut[ord(ch)] = v
    for i, w in enumerate(wins):
        codes[i, :len(w)] = lut[np.frombuffer(w.encode(), np.uint8)]
        lens[i] = len(w)
    tmp = "%s.%d.tmp.npz" % (cache, os.getpid())         # several ranks may build the cache at once: publish it atomically
    np.savez(tmp, codes=codes, lens=lens)
    os.replace(tmp, cache)
    return codes, lens


class Genome:
    """Two haplotypes of `size` random bases; site i's window sits at the same place on both, the
    allele (ref/var) drawn per haplotype."""

    def __init__(self, size, win_codes, win_lens, seed, device):
        g = torch.Generator(device=device); g.manual_seed(seed)
        self.size, self.device = size, device
        base = torch.randint(0, 4, (size,), dtype=torch.uint8, device=device, generator=g)
        self.hap = torch.stack([base, base.clone()])                     # [2, size]
        n_sites = win_codes.shape[0] // 2
        if n_sites:
            stride = size // n_sites
            assert stride >= 64, "genome too small for this panel"
            wc = torch.from_numpy(win_codes).to(device)
            wl = torch.from_numpy(win_lens).to(device)
            pos0 = torch.arange(n_sites, device=device, dtype=torch.int64) * stride + \
                torch.randint(0, stride - 40, (n_sites,), device=device, generator=g)
            ar = torch.arange(wc.shape[1], device=device, dtype=torch.int64)
            for h in range(2):
                allele = torch.randint(0, 2, (n_sites,), device=device, generator=g)
                rows = torch.arange(n_sites, device=device) * 2 + allele
                c, l = wc[rows], wl[rows]
                idx = pos0[:, None] + ar[None, :]
                ok = ar[None, :] < l[:, None]
                self.hap[h][idx[ok]] = c[ok]


def sample_reads(genome, n_reads, read_len, err, seed, n_frac=0.005):
    """-> codes uint8 [n_reads, read_len] (0-3, 4 = N): random haplotype/start/strand, `err`
    substitution rate, an N run (1-10) in n_frac of the reads."""
    dev = genome.device
    g = torch.Generator(device=dev); g.manual_seed(seed)
    hap = torch.randint(0, 2, (n_reads,), device=dev, generator=g)
    start = torch.randint(0, genome.size - read_len, (n_reads,), device=dev, generator=g)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    flat = genome.hap.view(-1)
    codes = flat[(hap * genome.size + start)[:, None] + ar[None, :]]
    rc = torch.rand(n_reads, device=dev, generator=g) < 0.5
    codes = torch.where(rc[:, None], 3 - codes.flip(1), codes)
    if err > 0:
        e = torch.rand(codes.shape, device=dev, generator=g) < err
        sub = torch.randint(1, 4, codes.shape, dtype=torch.uint8, device=dev, generator=g)
        codes = torch.where(e, (codes + sub) & 3, codes)
    if n_frac > 0:
        has = torch.rand(n_reads, device=dev, generator=g) < n_frac
        p0 = torch.randint(0, read_len, (n_reads,), device=dev, generator=g)
        ln = torch.randint(1, 11, (n_reads,), device=dev, generator=g)
        inrun = has[:, None] & (ar[None, :] >= p0[:, None]) & (ar[None, :] < (p0 + ln)[:, None])
        codes = torch.where(inrun, torch.full_like(codes, 4), codes)
    return codes


_SH16 = None


def pack_codes(codes, out_bases=None, out_mask=None, word_off=0):
    """codes uint8 [n, L] -> the packed stream (one separator after each read), written into
    out_bases (int32, 16 positions/word) / out_mask (int32, 32 positions/word) at stream offset
    word_off*32 positions.  n*(L+1) must be a multiple of 32.  Returns positions written."""
    n, L = codes.shape
    dev = codes.device
    stream = torch.cat([codes, torch.full((n, 1), 4, dtype=torch.uint8, device=dev)], 1).view(-1)
    n_pos = stream.numel()
    assert n_pos % 32 == 0
    s = stream.to(torch.int64)
    sh16 = 2 * torch.arange(16, device=dev, dtype=torch.int64)
    sh32 = torch.arange(32, device=dev, dtype=torch.int64)
    b = (((s & 3).view(-1, 16)) << sh16).sum(1)
    m = (((s >> 2).view(-1, 32)) << sh32).sum(1)
    b = torch.where(b >= 2 ** 31, b - 2 ** 32, b).to(torch.int32)
    m = torch.where(m >= 2 ** 31, m - 2 ** 32, m).to(torch.int32)
    if out_bases is None:
        return b, m, n_pos
    out_bases[word_off * 2: word_off * 2 + b.numel()] = b
    out_mask[word_off: word_off + m.numel()] = m
    return n_pos


def make_packed_shard(genome, n_reads, read_len, err, seed, chunk_reads=1 << 20):
    """Generate + pack n_reads reads chunk by chunk.  -> (bases int32, mask int32, n_pos, n_bases)."""
    dev = genome.device
    n_reads = n_reads // 32 * 32
    n_pos = n_reads * (read_len + 1)
    pad = padded_positions(n_pos)
    bases = torch.zeros(pad // 16, dtype=torch.int32, device=dev)
    mask = torch.full((pad // 32,), -1, dtype=torch.int32, device=dev)      # padding = invalid
    done = 0
    ci = 0
    while done < n_reads:
        c = min(chunk_reads, n_reads - done)
        codes = sample_reads(genome, c, read_len, err, seed * 1000003 + ci)
        pack_codes(codes, bases, mask, done * (read_len + 1) // 32)
        done += c
        ci += 1
    return bases, mask, n_pos, n_reads * read_len


def codes_to_fastq(codes, prefix="r"):
    """uint8 codes [n, L] (CPU tensor or array) -> FASTQ bytes with constant quality 'I'."""
    c = codes.cpu().numpy() if isinstance(codes, torch.Tensor) else codes
    n, L = c.shape
    seq = ASCII[c]
    out = []
    q = b"I" * L
    for i in range(n):
        out.append(b"@%s%d\n%s\n+\n%s\n" % (prefix.encode(), i, seq[i].tobytes(), q))
    return b"".join(out)


def codes_to_reads(codes):
    c = codes.cpu().numpy() if isinstance(codes, torch.Tensor) else codes
    seq = ASCII[c]
    return [seq[i].tobytes() for i in range(c.shape[0])]


def fastq_records(codes, first_index=0):
    """codes uint8 [n, L] on any device -> FASTQ records as a uint8 tensor [n, 2L + 15] on the same
    device: "@r<8 digits>\\n<seq>\\n+\\n<L x 'I'>\\n" (the layout bench.py's write_fastq_files makes with numpy)."""
    n, L = codes.shape
    dev = codes.device
    rec = torch.empty((n, 2 * L + 15), dtype=torch.uint8, device=dev)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    idx = torch.arange(first_index, first_index + n, device=dev, dtype=torch.int64)
    for d in range(8):
        rec[:, 2 + d] = ((idx // 10 ** (7 - d)) % 10 + 48).to(torch.uint8)
    rec[:, 10] = 10
    lut = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device=dev)
    rec[:, 11:11 + L] = lut[codes.long()]
    rec[:, 11 + L] = 10
    rec[:, 12 + L] = ord("+")
    rec[:, 13 + L] = 10
    rec[:, 14 + L:14 + 2 * L] = ord("I")
    rec[:, 14 + 2 * L] = 10
    return rec


def write_fastq_set(genome, n_reads, read_len, err, seed, n_files, outdir, prefix="part", chunk_reads=1 << 20):
    """n_reads reads of the bench generator as n_files plain FASTQ files (records built on the device,
    written chunk by chunk).  -> list of paths."""
    os.makedirs(outdir, exist_ok=True)
    per = (n_reads + n_files - 1) // n_files
    paths = []
    ci = 0
    for fi in range(n_files):
        p = os.path.join(outdir, "%s_%03d.fq" % (prefix, fi))
        left = min(per, n_reads - fi * per)
        with open(p, "wb") as fh:
            done = 0
            while done < left:
                c = min(chunk_reads, left - done)
                codes = sample_reads(genome, c, read_len, err, seed * 1000003 + ci)
                fastq_records(codes, first_index=(fi * per + done) % 100_000_000).cpu().numpy().tofile(fh)
                done += c
                ci += 1
        paths.append(p)
    return paths
