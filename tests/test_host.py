"""Host-side logic of libntsm_b200.so checked on CPU against the oracle and the reference
fixtures: ABI surface, hash64 and its inverse, the kseq-grammar reader, the site-table builder
(FingerPrint::initCountsHash), the 2-bit/N-mask packer and the counts-file formatter.
No CUDA call is made here."""
import ctypes as C
import json
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases

import ntsm_b200
from ntsm_b200 import _lib


@pytest.fixture(scope="module")
def L():
    return ntsm_b200.lib()


def test_library_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "ntsm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ntsm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 50
    for name in sorted(declared):
        assert hasattr(L, name), "declared in include/ntsm_b200.h but not exported: " + name
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())


def test_no_cpu_fallback(L):
    """Without a device the context cannot be created -- the product never counts on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    cfg = _lib.Cfg(k=19, device=0)
    assert L.ntsm_ctx_create(C.byref(ctx), C.byref(cfg)) == -2
    assert b"no CUDA device" in L.ntsm_last_error(None)


def test_hash64_matches_oracle_and_inverts(L, oracle):
    rng = random.Random(3)
    for k in (1, 2, 5, 11, 15, 16, 17, 19, 25, 31):
        for x in [0, 1, (1 << (2 * k)) - 1] + [rng.getrandbits(2 * k) for _ in range(500)]:
            h = L.ntsm_hash64(x, k)
            assert h == oracle.hash64(x, k)
            assert L.ntsm_hash64_inv(h, k) == x


def test_nt4(L, oracle):
    for b in range(256):
        assert L.ntsm_nt4(b) == oracle.nt4(b)


def _lib_records(L, path):
    r = C.c_void_p()
    assert L.ntsm_reader_open(C.byref(r), os.fsencode(path)) == 0
    out = []
    s = C.c_char_p()
    while True:
        l = L.ntsm_reader_next(r, C.byref(s))
        if l < 0:
            break
        out.append((L.ntsm_reader_name(r), C.string_at(s, l) if l else b""))
    L.ntsm_reader_close(r)
    return out, l


def _case_files(name):
    d = os.path.join(GOLDEN, "cases", name)
    argv = json.load(open(os.path.join(d, "cmd.json")))["argv"]
    files, i = [], 0
    opts = {"k": 19, "dupes": False, "m": 0.0, "t": 1, "sites": None}
    while i < len(argv):
        a = argv[i]
        if a == "-d":
            opts["dupes"] = True
        elif a in ("-s", "-k", "-m", "-t"):
            v = argv[i + 1]; i += 1
            if a == "-s": opts["sites"] = os.path.join(d, v)
            elif a == "-k": opts["k"] = int(v)
            elif a == "-m": opts["m"] = float(v)
            else: opts["t"] = int(v)
        else:
            files.append(os.path.join(d, a))
        i += 1
    return d, opts, files


@pytest.mark.parametrize("name", golden_cases())
def test_reader_matches_oracle_reader(L, oracle, name):
    d, opts, files = _case_files(name)
    for f in files + [opts["sites"]]:
        assert _lib_records(L, f) == oracle.read_records(f), f


@pytest.fixture(params=["memchr", "avx2", "avx512"])
def scan_isa(L, request):
    """every newline scanner of the reader's FASTQ fast path (fastx.cpp), skipping what the CPU lacks"""
    got = L.ntsm_scan_isa(request.param.encode()).decode()
    if got != request.param:
        L.ntsm_scan_isa(b"")
        pytest.skip("CPU has no %s" % request.param)
    yield got
    L.ntsm_scan_isa(b"")


def test_reader_fuzz(L, oracle, tmp_path, scan_isa):
    from test_oracle import _rand_fastx
    rng = random.Random(5)
    wins = ["ACGTTGCATGCATGCAAGCTT", "CCACGTAGCACTGCACCCCCAT"]
    for i in range(300):
        f = tmp_path / ("z%d" % i)
        txt = _rand_fastx(rng, wins)
        if rng.random() < 0.3:   # records larger than the reader's window boundaries do not matter; long lines do
            txt += ">big\n" + "ACGT" * rng.randrange(1, 5000) + "\n"
        f.write_text(txt, newline="")
        assert _lib_records(L, str(f)) == oracle.read_records(str(f))


def test_reader_fast_path_across_window_refills(L, oracle, tmp_path, scan_isa):
    """Several MiB of mostly regular 4-line FASTQ (the zero-copy fast path) with irregular records
    mixed in (CRLF, multi-line, FASTA, short/long quality, '@' quality lines), in plain and gz
    form: records that straddle the 1 MiB window and every hand-over between the fast path and the
    general parser give the oracle's records."""
    import gzip
    rng = random.Random(17)
    parts = []
    for i in range(9000):
        n = rng.choice([1, 2, 19, 75, 150, 151, 250, 1000])
        seq = "".join(rng.choice("ACGTN") for _ in range(n))
        x = rng.random()
        if x < 0.90:
            q = "".join(rng.choice("@+>I#5") for _ in range(n))
            parts.append("@r%d some comment\n%s\n+\n%s\n" % (i, seq, q))
        elif x < 0.92:
            parts.append("@r%d\r\n%s\r\n+\r\n%s\r\n" % (i, seq, "I" * n))
        elif x < 0.94:
            h = n // 2
            parts.append("@r%d\n%s\n%s\n+r%d\n%s\n%s\n" % (i, seq[:h], seq[h:], i, "I" * h, "I" * (n - h)))
        elif x < 0.96:
            parts.append(">f%d\n%s\n" % (i, seq))
        elif x < 0.97:
            parts.append("@r%d\n%s\n+\n%s\n" % (i, seq, "I" * (n + 3)))      # long quality: -2, file ends there
        elif x < 0.98:
            parts.append("@r%d\n\n%s\n+\n%s\n" % (i, seq, "I" * n))         # empty line before the sequence
        else:
            parts.append("\n\n@r%d\t x\n%s\n+\n%s\n\n" % (i, seq, "I" * n))
    for cut in (len(parts), 4000):
        txt = "".join(p for p in parts[:cut] if cut == len(parts) or "III" not in p[-6:] or len(p) < 40)
        plain = tmp_path / ("w%d.fq" % cut)
        plain.write_text(txt, newline="")
        gz = tmp_path / ("w%d.fq.gz" % cut)
        with gzip.open(gz, "wb") as fh:
            fh.write(txt.encode())
        want = oracle.read_records(str(plain))
        assert len(want[0]) > 100
        assert _lib_records(L, str(plain)) == want
        assert _lib_records(L, str(gz)) == want
    # a clean multi-MiB file never leaves the fast path; its last record has no trailing newline
    clean = "".join("@c%d\n%s\n+\n%s\n" % (i, "ACGTTGCA" * 19, "I" * 152) for i in range(12000))[:-1]
    f = tmp_path / "clean.fq"
    f.write_text(clean, newline="")
    want = oracle.read_records(str(f))
    assert len(want[0]) == 12000
    assert _lib_records(L, str(f)) == want


@pytest.mark.parametrize("name", golden_cases())
def test_site_table_matches_oracle(L, oracle, name):
    d, opts, files = _case_files(name)
    s = ntsm_b200.SiteSet(opts["sites"], opts["k"], opts["dupes"])
    fp = oracle.fingerprint(opts["sites"], opts["k"], opts["dupes"], opts["m"])
    hs, off, cnt = fp.lists()
    assert s.n_sites == fp.n_sites and s.table_size == fp.table_size
    assert np.array_equal(s.hashes, hs) and np.array_equal(s.allele_off, off)
    assert np.array_equal(s.erased.astype(bool), cnt == 0xFFFFFFFF)
    assert s.names == fp.names()
    assert s.max_counts(opts["m"]) == fp.max_counts
    # warnings exactly as the reference printed them
    want = [l for l in open(os.path.join(d, "stderr.txt")).read().splitlines() if "k-mer collision" in l]
    assert s.warnings == want
    ref_rc = int(open(os.path.join(d, "rc.txt")).read())
    assert s.printable() == (ref_rc == 0)


def _py_pack(reads, nt4):
    """Independent restatement of the packed layout in include/ntsm_b200.h."""
    codes = []
    for r in reads:
        codes.extend(nt4(b) for b in r)
        codes.extend([4] * (((len(r) + 8) & ~7) - len(r)))        # separator + padding to a multiple of 8
    n = len(codes)
    padded = (n + 8191) // 8192 * 8192 + 64
    codes += [4] * (padded - n)
    c = np.array(codes, np.uint64)
    b2 = ((c & 3) << (2 * (np.arange(padded, dtype=np.uint64) % 16))).reshape(-1, 16).sum(1).astype(np.uint32)
    mk = ((c >> 2) << (np.arange(padded, dtype=np.uint64) % 32)).reshape(-1, 32).sum(1).astype(np.uint32)
    return b2, mk, n


def test_packer_layout(L, oracle):
    rng = random.Random(11)
    alphabet = b"ACGTacgtNnUuRY-\x00\x01\x02\x03\xff "
    for trial in range(40):
        reads = [bytes(rng.choice(alphabet) for _ in range(rng.choice([0, 1, 18, 19, 31, 32, 33, 63, 64, 65, 150, 151, 700])))
                 for _ in range(rng.randrange(1, 30))]
        b2, mk, n_pos, roff = ntsm_b200.pack_reads(reads)
        wb, wm, wn = _py_pack(reads, oracle.nt4)
        assert n_pos == wn == sum((len(r) + 8) & ~7 for r in reads)
        assert np.array_equal(b2[:len(wb)], wb) and np.array_equal(mk[:len(wm)], wm)
        assert list(roff[:-1]) == list(np.cumsum([0] + [(len(r) + 8) & ~7 for r in reads])[:-1])


def test_packer_isa_variants_agree(L):
    """scalar, AVX2 and AVX-512 VBMI packers write the same words (whichever this CPU has)."""
    rng = random.Random(5)
    alphabet = bytes(range(256))
    reads = [bytes(rng.choice(alphabet) if rng.random() < 0.2 else rng.choice(b"ACGTacgtNU") for _ in range(rng.choice([0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 150, 1000])))
             for _ in range(300)]
    outs = {}
    try:
        for isa in (b"scalar", b"avx2", b"avx512"):
            got = L.ntsm_pack_isa(isa)
            outs[got] = ntsm_b200.pack_reads(reads)
    finally:
        L.ntsm_pack_isa(b"")
    assert b"scalar" in outs
    ref = outs[b"scalar"]
    for name, o in outs.items():
        assert o[2] == ref[2] and np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1]), name


def _aligned(n, dtype, fill):
    """numpy array of n items whose data starts on a 64-byte boundary"""
    raw = np.full(n * np.dtype(dtype).itemsize + 64, fill, np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype)


@pytest.mark.parametrize("isa", [b"scalar", b"avx2", b"avx512"])
def test_streaming_packer_writes_the_same_words(L, isa):
    """The mode the pinned batches are filled in (staging area + non-temporal 64-byte lines, pack.h) against
    the plain mode, word for word: many short reads (staging wraps many times), reads around the block and
    staging sizes, a read far longer than the staging area in the middle and at the end, empty reads."""
    rng = random.Random(77)
    alphabet = b"ACGTacgtNnU\x00\x03\xff"
    lens = [rng.choice([0, 1, 7, 8, 31, 150, 151, 152, 500, 503, 504, 511, 512, 513, 1000]) for _ in range(4000)]
    lens[1000] = 200_000
    lens[2500] = 65536 - 1024 - 512 - 8
    lens[2501] = 65536 - 1024 - 512
    lens[2502] = 65536 - 1024 - 512 + 8
    lens[-1] = 70_001
    reads = [bytes(rng.choice(alphabet) for _ in range(n)) if n < 5000 else bytes(rng.choice(b"ACGTN") for _ in range(1000)) * (n // 1000) + b"A" * (n % 1000)
             for n in lens]
    buf = b"".join(reads)
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    n_pos_max = sum((len(r) + 8) & ~7 for r in reads)
    padded = L.ntsm_padded_positions(n_pos_max)
    cbuf = ctypes.create_string_buffer(buf, len(buf) + 1)
    try:
        L.ntsm_pack_isa(isa)
        outs = []
        for streaming in (0, 1):
            b2 = _aligned(padded // 16, np.uint32, 0xAB)
            mk = _aligned(padded // 32, np.uint32, 0xCD)
            n = L.ntsm_pack_reads2(ctypes.cast(cbuf, ctypes.c_void_p), off.ctypes.data, len(reads), b2.ctypes.data, mk.ctypes.data, None, streaming)
            assert n == n_pos_max
            outs.append((b2.copy(), mk.copy()))
    finally:
        L.ntsm_pack_isa(b"")
    assert np.array_equal(outs[0][1], outs[1][1])                     # N-mask plane: every word, padding included
    # base codes under invalid positions are don't-care inside a read's padding; compare them where the mask says valid
    valid = ~np.unpackbits(outs[0][1].view(np.uint8), bitorder="little").astype(bool)
    c0 = np.unpackbits(outs[0][0].view(np.uint8), bitorder="little").reshape(-1, 2)
    c1 = np.unpackbits(outs[1][0].view(np.uint8), bitorder="little").reshape(-1, 2)
    assert np.array_equal(c0[valid], c1[valid])


def test_packer_reads_nothing_past_a_page_edge(L):
    """A read that ends on the last byte before an unmapped page is packed without touching that page."""
    import mmap
    libc = C.CDLL(None, use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    base = libc.mmap(None, 2 * 4096, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
    assert base not in (None, C.c_void_p(-1).value)
    assert libc.mprotect(base + 4096, 4096, 0) == 0                 # PROT_NONE
    try:
        for isa in (b"scalar", b"avx2", b"avx512"):
            L.ntsm_pack_isa(isa)
            for n in (1, 7, 22, 31, 32, 33, 63, 64, 65, 150):
                read = bytes(b"ACGTN"[i % 5] for i in range(n))
                C.memmove(base + 4096 - n, read, n)
                off = np.array([0, n], np.uint64)
                padded = L.ntsm_padded_positions((n + 8) & ~7)
                b2 = np.zeros(padded // 16, np.uint32)
                mk = np.zeros(padded // 32, np.uint32)
                assert L.ntsm_pack_reads(base + 4096 - n, off.ctypes.data, 1, b2.ctypes.data, mk.ctypes.data, None) == (n + 8) & ~7
                wb, wm, _, _ = ntsm_b200.pack_reads([read])
                assert np.array_equal(b2, wb) and np.array_equal(mk, wm), (isa, n)
    finally:
        L.ntsm_pack_isa(b"")
        libc.munmap(C.c_void_p(base), 2 * 4096)


def test_format_counts_matches_reference_fixture(L, oracle):
    """printOptionalHeader/printCountsMax/printInfoSummary text from oracle rows == reference stdout."""
    for name in ("mini", "panel300_fq", "dupes_allowed", "k31"):
        d, opts, files = _case_files(name)
        fp = oracle.fingerprint(opts["sites"], opts["k"], opts["dupes"], opts["m"])
        for f in files:
            fp.count_file(f)
        mr, mv, sr, sv, _, _ = fp.rows()
        s = ntsm_b200.SiteSet(opts["sites"], opts["k"], opts["dupes"])
        n = L.ntsm_format_counts(s._h, mr.ctypes.data, mv.ctypes.data, sr.ctypes.data, sv.ctypes.data, fp.total_kmers, None, 0)
        buf = C.create_string_buffer(n)
        L.ntsm_format_counts(s._h, mr.ctypes.data, mv.ctypes.data, sr.ctypes.data, sv.ctypes.data, fp.total_kmers, buf, n)
        assert buf.raw[:n] == open(os.path.join(d, "stdout.txt"), "rb").read()
        t = np.array([fp.total_kmers, fp.total_counts, fp.total_bases], np.uint64)
        cov = L.ntsm_sites_covered(mr.ctypes.data, mv.ctypes.data, s.n_sites)
        sb = C.create_string_buffer(2048)
        m = L.ntsm_format_summary(s._h, t.ctypes.data, cov, sb, 2048)
        assert sb.raw[:m].decode() in open(os.path.join(d, "stderr.txt")).read()


def test_format_counts_aborts_like_reference(L):
    d, opts, files = _case_files("dupes_abort")
    s = ntsm_b200.SiteSet(opts["sites"], 19, False)
    z = np.zeros(s.n_sites, np.uint32)
    assert L.ntsm_format_counts(s._h, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, 0, None, 0) == -134


@pytest.mark.parametrize("threads", ["1", "3", "8", "32"])
def test_site_table_parallel_first_wins_with_many_duplicates(L, oracle, tmp_path, monkeypatch, threads):
    """ntsm_sites_load finds first occurrences with a parallel min-over-occurrence-number table
    (sites.cpp); the result must not depend on the thread count and must equal the oracle's serial
    insert loop (src/FingerPrint.hpp:506-563) -- dense order, per-site offsets, erased flags and the
    collision warnings in file order -- on a panel full of duplicates: k-mers repeated inside a
    record, between the two alleles of a site, and between sites far apart."""
    import random
    rng = random.Random(31)
    pool = ["".join(rng.choice("ACGT") for _ in range(19)) for _ in range(300)]
    recs = []
    for i in range(1500):
        for tag in ("ref", "var"):
            kms = [rng.choice(pool) if rng.random() < 0.3 else "".join(rng.choice("ACGT") for _ in range(19)) for _ in range(rng.randrange(1, 9))]
            if rng.random() < 0.1:
                kms.append(kms[0])                                   # the same k-mer twice in one record
            body = "N".join(kms)
            if rng.random() < 0.2:
                body = body.lower()
            recs.append(">site%d %s\n%s\n" % (i, tag, body))
    recs.append(">odd ref\n%s\n" % pool[0])                           # odd record count
    p = tmp_path / "dups.fa"
    p.write_text("".join(recs))
    monkeypatch.setenv("NTSM_SITES_THREADS", threads)
    for dupes in (False, True):
        s = ntsm_b200.SiteSet(str(p), 19, dupes)
        fp = oracle.fingerprint(str(p), 19, dupes, 0)
        hs, off, cnt = fp.lists()
        assert s.n_sites == fp.n_sites and s.table_size == fp.table_size
        assert np.array_equal(s.hashes, hs) and np.array_equal(s.allele_off, off)
        assert np.array_equal(s.erased.astype(bool), cnt == 0xFFFFFFFF)
        assert s.names == fp.names()
        assert len(s.warnings) > 1000
    # the warnings are the reference's own, in its order: compare with the real binary when it is here
    ref = os.path.join(ROOT, "oracle", "_ref", "ntsmCount")
    if os.path.exists(ref):
        reads = tmp_path / "r.fa"
        reads.write_text(">r\nACGT\n")
        out = subprocess.run([ref, "-d", "-s", str(p), str(reads)], capture_output=True, text=True)
        want = [l for l in out.stderr.splitlines() if "k-mer collision" in l]
        assert s.warnings == want


def test_numa_helpers_parse_and_stay_out_of_the_way_on_one_node_hosts(L):
    """numa.cpp: the cpulist parser (sysfs format) and the node count; with one node nothing is pinned or bound."""
    import ctypes
    L.ntsm_numa_parse_cpulist.restype = ctypes.c_int
    L.ntsm_numa_parse_cpulist.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    out = (ctypes.c_int * 64)()
    n = L.ntsm_numa_parse_cpulist(b"0-3,8,10-11\n", out, 64)
    assert list(out[:n]) == [0, 1, 2, 3, 8, 10, 11]
    assert L.ntsm_numa_parse_cpulist(b"", out, 64) == 0
    assert L.ntsm_numa_parse_cpulist(b"5", out, 64) == 1 and out[0] == 5
    L.ntsm_numa_nodes.restype = ctypes.c_int
    nodes = L.ntsm_numa_nodes()
    want = len([d for d in os.listdir("/sys/devices/system/node") if re.fullmatch(r"node\d+", d)]) if os.path.isdir("/sys/devices/system/node") else 1
    assert nodes == max(1, want)
    before = os.sched_getaffinity(0)
    # a bulk insert on a box without CUDA fails early; the calling thread's affinity must be untouched either way
    assert os.sched_getaffinity(0) == before


def test_parser_worker_processes_speak_the_shared_memory_protocol(tmp_path):
    """bin/ntsm_parse_worker + procpipe.h without a GPU: tools/procpipe_selftest.cpp plays the owner (shared mapping,
    spawn, slot hand-over) and tallies what the workers packed.  Reads, bases and valid positions must be what the
    files hold -- several files and workers, FASTA + FASTQ, N runs, reads longer than a slot (split with a k-1 overlap),
    an unreadable file (error text of src/FingerPrint.hpp:51-57, the other workers stop)."""
    exe = str(tmp_path / "procpipe_selftest")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "ntsm_b200", "csrc"),
                           os.path.join(ROOT, "tools", "procpipe_selftest.cpp"), "-o", exe])
    worker = os.path.join(ROOT, "ntsm_b200", "bin", "ntsm_parse_worker")
    assert os.access(worker, os.X_OK)
    rng = random.Random(8)
    valid_set = set(b"ACGTUacgtu\x00\x01\x02\x03")
    paths, reads = [], []
    for f in range(5):
        p = tmp_path / ("f%d.%s" % (f, "fq" if f % 2 else "fa"))
        recs = []
        for i in range(rng.randrange(200, 600)):
            n = rng.choice([0, 1, 18, 19, 40, 150, 151, 300, 5000 if f == 2 else 77])
            s = bytes(rng.choice(b"ACGTNacgtR") for _ in range(n))
            reads.append(s)
            recs.append((b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * n)) if f % 2 else (b">r%d\n%s\n" % (i, s)))
        p.write_bytes(b"".join(recs))
        paths.append(str(p))
    want_bases = sum(map(len, reads))
    want_valid = sum(sum(c in valid_set for c in r) for r in reads)

    def run(cap, workers, files):
        out = subprocess.run([exe, worker, "19", str(cap), str(workers)] + files, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        return json.loads(out.stdout)

    for cap, workers in ((1 << 20, 1), (1 << 20, 3), (65536, 7)):
        r = run(cap, workers, paths)
        assert (r["reads"], r["bases"], r["valid"], r["error"], r["bad_exit"]) == (len(reads), want_bases, want_valid, 0, 0), (cap, workers)
        assert r["positions"] == sum((len(x) + 8) & ~7 for x in reads)
    r = run(4096, 4, paths)                # slots shorter than the 5000-base reads: split, k-1 bases packed twice at every cut
    assert (r["reads"], r["bases"], r["error"], r["bad_exit"]) == (len(reads), want_bases, 0, 0)
    assert r["valid"] > want_valid and r["batches"] > 50
    r = run(65536, 3, paths[:2] + [str(tmp_path / "missing.fq")] + paths[2:])
    assert r["error"] == -5 and r["error_text"] == "file %s cannot be opened" % (tmp_path / "missing.fq") and r["bad_exit"] == 0
