"""Host ingest, byte-source level: ntsm_gz_* (ntsm_b200/csrc/gzsource.cpp + inflate.cpp) must deliver
exactly the bytes zlib's gzread delivers -- the reference reads every input through gzopen/gzread
(src/FingerPrint.hpp:50, vendor/kseq.h:68-79).  The checker is zlib itself, twice: Python's zlib
for what valid files contain, and the library's own NTSM_INFLATE=zlib mode (plain gzread) for how
damaged files behave (bytes delivered before the error, then -1)."""
import ctypes as C
import gzip
import os
import random
import struct
import zlib

import pytest

import ntsm_b200

L = ntsm_b200.lib()


def read_all(path, helpers=0, chunk=None, rng=None):
    """-> (bytes delivered, final return code (0 end / -1 error), mode at open, fell_back)"""
    h = C.c_void_p()
    assert L.ntsm_gz_open(C.byref(h), str(path).encode(), helpers) == 0
    mode0 = L.ntsm_gz_mode(h).decode()
    out = bytearray()
    buf = C.create_string_buffer(1 << 20)
    rc = 0
    while True:
        n = chunk if chunk else (rng.choice([1, 7, 100, 4096, 65536, 1 << 20]) if rng else 1 << 20)
        r = L.ntsm_gz_read(h, buf, n)
        if r <= 0:
            rc = r
            break
        out += buf.raw[:r]
        if r < n:                      # short read = end (or an error to be reported by the next call)
            rc = L.ntsm_gz_read(h, buf, n)
            assert rc <= 0
            break
    fb = L.ntsm_gz_fell_back(h)
    L.ntsm_gz_close(h)
    return bytes(out), rc, mode0, fb


def zlib_mode(path, monkeypatch, **kw):
    monkeypatch.setenv("NTSM_INFLATE", "zlib")
    try:
        return read_all(path, **kw)
    finally:
        monkeypatch.delenv("NTSM_INFLATE")


def fastq(rng, n, read_len=150):
    recs = []
    for i in range(n):
        s = "".join(rng.choice("ACGT") for _ in range(read_len))
        q = "".join(rng.choice("FFFFFFFF:,#") for _ in range(read_len))
        recs.append("@r%d lane:%d\n%s\n+\n%s\n" % (i, i % 4, s, q))
    return "".join(recs).encode()


def member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, name=None, comment=None, extra=None, hcrc=False):
    """one gzip member written by hand so every header field can be exercised"""
    flg = (4 if extra is not None else 0) | (8 if name else 0) | (16 if comment else 0) | (2 if hcrc else 0)
    hdr = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\x00\xff"
    if extra is not None:
        hdr += struct.pack("<H", len(extra)) + extra
    if name:
        hdr += name + b"\0"
    if comment:
        hdr += comment + b"\0"
    if hcrc:
        hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    body = co.compress(data) + co.flush()
    return hdr + body + struct.pack("<II", zlib.crc32(data), len(data) & 0xFFFFFFFF)


def bgzf(data, block=0xFF00, level=6, eof_marker=True):
    out = bytearray()
    blocks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
    for b in blocks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(b) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(b), len(b))
    return bytes(out)


def test_crc32_matches_zlib():
    rng = random.Random(5)
    for n in list(range(0, 300)) + [1000, 4095, 4096, 65537, 1 << 20]:
        b = rng.randbytes(n)
        seed = rng.getrandbits(32)
        assert L.ntsm_crc32(seed, b, n) == zlib.crc32(b, seed), n
    # running CRC over uneven pieces
    b = rng.randbytes(300000)
    c = 0
    pos = 0
    while pos < len(b):
        k = rng.choice([1, 15, 16, 255, 256, 257, 5000, 70000])
        c = L.ntsm_crc32(c, b[pos:pos + k], len(b[pos:pos + k]))
        pos += k
    assert c == zlib.crc32(b)


PAYLOADS = {
    "fastq": lambda rng: fastq(rng, 3000),
    "fastq_big": lambda rng: fastq(rng, 12000) * 2,                     # > 1 MiB window, matches 32 KiB back across the slide
    "random": lambda rng: rng.randbytes(300000),                        # incompressible: stored blocks
    "zeros": lambda rng: bytes(2_500_000),                              # one distance code of one bit (incomplete set zlib accepts)
    "runs": lambda rng: b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 700) for _ in range(4000)),
    "short_period": lambda rng: (b"ACGTTGCA" * 5 + b"N") * 30000,       # distances 2..7 and 41
    "one_byte": lambda rng: b"A",
    "empty": lambda rng: b"",
    "text_small": lambda rng: b">s\nACGT\n",                            # fixed-Huffman block
}
VARIANTS = {
    "l1": dict(level=1), "l6": dict(level=6), "l9": dict(level=9), "stored": dict(level=0),
    "fixed": dict(level=6, strategy=zlib.Z_FIXED), "huffman_only": dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY),
    "rle": dict(level=6, strategy=zlib.Z_RLE), "filtered": dict(level=4, strategy=zlib.Z_FILTERED),
}


@pytest.mark.parametrize("payload", sorted(PAYLOADS))
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_valid_members_decode_to_the_original(tmp_path, payload, variant):
    rng = random.Random(hash((payload, variant)) & 0xFFFF)
    data = PAYLOADS[payload](rng)
    p = tmp_path / "x.gz"
    p.write_bytes(member(data, name=b"x.fq", **VARIANTS[variant]))
    got, rc, mode, fb = read_all(p, rng=rng)
    assert (got, rc) == (data, 0)
    assert mode == "fast" and not fb
    assert gzip.decompress(p.read_bytes()) == data


def test_multi_member_headers_and_trailing_garbage(tmp_path, monkeypatch):
    rng = random.Random(11)
    a, b, c = fastq(rng, 500), b"", fastq(rng, 700, 90)
    blob = member(a, name=b"a.fq") + member(b) + member(c, extra=b"XY\x03\0abc", comment=b"hello", level=9)
    for tail in (b"", b"\0" * 100, b"garbage after the last member", b"\x1f", b"\x1f\x8b"):
        p = tmp_path / "m.gz"
        p.write_bytes(blob + tail)
        want = zlib_mode(p, monkeypatch)
        got = read_all(p, rng=rng)
        assert got[:2] == want[:2], tail
        if tail != b"\x1f\x8b":
            assert got[0] == a + b + c and got[1] == 0


def test_header_crc_member_goes_to_zlib(tmp_path):
    rng = random.Random(12)
    a, b = fastq(rng, 300), fastq(rng, 300)
    p = tmp_path / "h.gz"
    p.write_bytes(member(a) + member(b, hcrc=True, name=b"b") + member(a))
    got, rc, mode, fb = read_all(p, chunk=777)
    assert (got, rc) == (a + b + a, 0)
    assert mode == "fast" and fb          # the second member is read by zlib, and so is what follows


@pytest.mark.parametrize("helpers", [0, 1, 3])
def test_bgzf_block_parallel(tmp_path, helpers):
    rng = random.Random(13 + helpers)
    data = fastq(rng, 20000)               # ~6.5 MB -> ~100 blocks, several batches
    p = tmp_path / "b.fq.gz"
    p.write_bytes(bgzf(data))
    got, rc, mode, fb = read_all(p, helpers=helpers, rng=rng)
    assert (got, rc) == (data, 0) and not fb
    assert mode == ("bgzf" if helpers else "fast")
    # blocks of the maximum size, no EOF marker, an ordinary member and garbage appended
    data2 = data[1000:1000 + 65536 * 3 + 17]
    p.write_bytes(bgzf(data2, block=65536, level=1, eof_marker=False) + member(data[:100000]) + b"junk")
    got, rc, mode, fb = read_all(p, helpers=helpers, rng=rng)
    assert (got, rc) == (data2 + data[:100000], 0) and not fb


def test_bgzf_with_a_bad_block_behaves_like_zlib(tmp_path, monkeypatch):
    rng = random.Random(14)
    data = fastq(rng, 9000)
    good = bytearray(bgzf(data))
    for trial in range(12):
        blob = bytearray(good)
        pos = rng.randrange(len(blob))
        if trial % 3 == 0:
            blob = blob[:pos]                                   # truncated
        else:
            blob[pos] ^= 1 << rng.randrange(8)                  # one flipped bit: header, body, CRC or size
        p = tmp_path / "bad.gz"
        p.write_bytes(bytes(blob))
        want = zlib_mode(p, monkeypatch, chunk=65536)
        for helpers in (0, 2):
            got = read_all(p, helpers=helpers, chunk=65536)
            assert got[:2] == want[:2], (trial, helpers, pos)


def test_damaged_members_behave_like_zlib(tmp_path, monkeypatch):
    """bit flips, truncation and junk anywhere in ordinary gzip files: same bytes, same final code as gzread"""
    rng = random.Random(15)
    base = [member(fastq(rng, 400), level=lv) for lv in (1, 6, 9)] + [member(rng.randbytes(70000), level=0),
                                                                     member(fastq(rng, 200), strategy=zlib.Z_FIXED)]
    n_fb = 0
    for trial in range(150):
        blob = bytearray(b"".join(rng.sample(base, rng.randrange(1, 4))))
        kind = trial % 5
        pos = rng.randrange(len(blob))
        if kind == 0:
            blob = blob[:pos]
        elif kind == 1:
            blob[pos] ^= 1 << rng.randrange(8)
        elif kind == 2:
            blob[pos:pos + rng.randrange(1, 40)] = rng.randbytes(rng.randrange(1, 40))
        elif kind == 3:
            blob[pos:pos] = rng.randbytes(rng.randrange(1, 9))
        else:
            del blob[pos:pos + rng.randrange(1, 9)]
        if len(blob) < 18:
            continue
        p = tmp_path / "d.gz"
        p.write_bytes(bytes(blob))
        want = zlib_mode(p, monkeypatch, chunk=50000)
        got = read_all(p, chunk=50000)
        assert got[0] == want[0], (trial, kind, pos)
        assert got[1] == want[1], (trial, kind, pos)
        n_fb += got[3]
    assert n_fb > 30          # the damage was really met by the fast decoder, not only by zlib


def test_plain_and_tiny_files_pass_through(tmp_path):
    p = tmp_path / "plain.fq"
    data = fastq(random.Random(16), 100)
    p.write_bytes(data)
    got, rc, mode, fb = read_all(p, chunk=1000)
    assert (got, rc, mode) == (data, 0, "mapped")        # a plain regular file: mapped, same bytes as gzread's transparent mode
    p.write_bytes(b"\x1f" + data)                        # starts like gzip but is not: still passed through byte for byte
    assert read_all(p, chunk=7)[:2] == (b"\x1f" + data, 0)
    p.write_bytes(b"")
    assert read_all(p)[:3] == (b"", 0, "zlib")
    p.write_bytes(b"\x1f\x8b\x08")                # shorter than any gzip member: zlib's call
    assert read_all(p)[2] == "zlib"


def test_reader_over_fast_source_equals_reader_over_zlib(tmp_path, monkeypatch):
    """the record parser sees the same records whichever way the bytes were inflated"""
    rng = random.Random(17)
    data = fastq(rng, 5000) + b">fa1\nACGTNNNN\nACGT\n"
    files = {"gz": member(data), "bgzf": bgzf(data), "multi": member(data[:300000]) + member(data[300000:])}

    def records(path, helpers):
        h = C.c_void_p()
        assert L.ntsm_reader_open2(C.byref(h), str(path).encode(), helpers) == 0
        seq = C.c_char_p()
        out = []
        while True:
            n = L.ntsm_reader_next(h, C.byref(seq))
            if n < 0:
                out.append(n)
                break
            out.append(C.string_at(seq, n))
        L.ntsm_reader_close(h)
        return out

    for name, blob in files.items():
        p = tmp_path / (name + ".gz")
        p.write_bytes(blob)
        monkeypatch.setenv("NTSM_INFLATE", "zlib")
        want = records(p, 0)
        monkeypatch.delenv("NTSM_INFLATE")
        assert len(want) == 5002
        for helpers in (0, 2):
            assert records(p, helpers) == want, (name, helpers)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cases")
GZ_CASES = sorted(c for c in os.listdir(GOLDEN) if c.startswith(("gz_", "bgzf")) or c == "panel300_gz")


@pytest.mark.parametrize("case", GZ_CASES)
@pytest.mark.parametrize("helpers", [0, 3])
def test_reference_fixture_bases_through_the_reader(case, helpers):
    """tests/golden gzip cases (clean, damaged, truncated, BGZF; outputs of the real reference binary):
    the records our reader produces from the fast / block-parallel sources add up to the reference's
    own 'Total Bases Considered' and 'Total k-mers Considered' (every window of 19 valid bases)."""
    d = os.path.join(GOLDEN, case)
    err = open(os.path.join(d, "stderr.txt")).read()
    want_bases = int(err.split("Total Bases Considered: ")[1].split()[0])
    want_tk = int(err.split("Total k-mers Considered: ")[1].split()[0])
    h = C.c_void_p()
    assert L.ntsm_reader_open2(C.byref(h), os.path.join(d, "reads.fq.gz").encode(), helpers) == 0
    seq = C.c_char_p()
    bases = tk = 0
    valid = set(b"ACGTUacgtu\x00\x01\x02\x03")
    while True:
        n = L.ntsm_reader_next(h, C.byref(seq))
        if n < 0:
            break
        bases += n
        run = 0
        for ch in C.string_at(seq, n):
            run = run + 1 if ch in valid else 0
            tk += run >= 19
    L.ntsm_reader_close(h)
    assert (bases, tk) == (want_bases, want_tk)


def test_bgzf_close_while_helpers_are_busy(tmp_path):
    """-m stops reading in the middle of a file: closing a block-parallel source right after opening it,
    after one byte, or half way must neither hang nor crash, however often it is done"""
    rng = random.Random(21)
    data = fastq(rng, 12000)
    p = tmp_path / "b.fq.gz"
    p.write_bytes(bgzf(data))
    buf = C.create_string_buffer(1 << 16)
    for i in range(120):
        h = C.c_void_p()
        assert L.ntsm_gz_open(C.byref(h), str(p).encode(), 1 + i % 4) == 0
        want = [0, 1, 70000, len(data) // 2][i % 4]
        got = 0
        while got < want:
            r = L.ntsm_gz_read(h, buf, min(len(buf), want - got))
            assert r > 0
            assert buf.raw[:r] == data[got:got + r]
            got += r
        L.ntsm_gz_close(h)


def test_many_sources_at_once(tmp_path):
    """one source per parser thread, as ntsm_count_files runs them: eight files of three kinds read
    concurrently (ctypes releases the GIL), each with helpers, each byte-exact"""
    import threading
    rng = random.Random(22)
    files = []
    for i in range(8):
        data = fastq(rng, 4000 + 500 * i)
        blob = [member(data), bgzf(data), member(data[:len(data) // 3]) + member(data[len(data) // 3:], level=1)][i % 3]
        p = tmp_path / ("f%d.gz" % i)
        p.write_bytes(blob)
        files.append((p, data))
    out = [None] * len(files)

    def work(j):
        out[j] = read_all(files[j][0], helpers=2, chunk=50000 + j)

    th = [threading.Thread(target=work, args=(j,)) for j in range(len(files))]
    for t in th: t.start()
    for t in th: t.join()
    for j, (p, data) in enumerate(files):
        assert out[j][:2] == (data, 0), j


# ---------------------------------------------------------------- one member on several threads (pargz)
@pytest.fixture
def small_parallel_chunks(monkeypatch):
    """make the parallel single-member decoder kick in for test-sized files: 64 KiB chunks (zlib's blocks
    of this kind of data are 30-60 KB compressed, and a worker searches three chunks for a start), no minimum"""
    monkeypatch.setenv("NTSM_PARGZ_MIN", "0")
    monkeypatch.setenv("NTSM_PARGZ_MIN_WORKERS", "1")
    monkeypatch.setenv("NTSM_PARGZ_CHUNK", "65536")


def _read_counting_chunks(path, helpers, rng=None):
    h = C.c_void_p()
    assert L.ntsm_gz_open(C.byref(h), str(path).encode(), helpers) == 0
    out = bytearray()
    buf = C.create_string_buffer(1 << 20)
    while True:
        n = rng.choice([1, 100, 4096, 65536, 1 << 20]) if rng else 1 << 20
        r = L.ntsm_gz_read(h, buf, n)
        if r <= 0:
            rc = r
            break
        out += buf.raw[:r]
    res = bytes(out), rc, L.ntsm_gz_parallel_chunks(h), L.ntsm_gz_fell_back(h)
    L.ntsm_gz_close(h)
    return res


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("helpers", [1, 2, 5])
def test_single_member_on_several_threads(tmp_path, small_parallel_chunks, level, helpers):
    """a FASTQ-like member cut into chunks: workers find block starts on their own, decode with markers
    for the window they cannot know, and the stitcher accepts a chunk only where the previous one
    stopped; the bytes are the original's and most chunks really came from workers"""
    rng = random.Random(40 + level)
    data = fastq(rng, 9000) + b">tail\n" + b"ACGTTGCA" * 3000 + b"\n"
    p = tmp_path / "one.gz"
    p.write_bytes(member(data, level=level, name=b"reads.fq"))
    got, rc, chunks, fb = _read_counting_chunks(p, helpers, rng)
    assert (got, rc, fb) == (data, 0, 0)
    assert chunks >= 0.7 * (p.stat().st_size / 65536) and chunks >= 5


def test_parallel_decoder_hands_back_to_one_thread_where_it_must(tmp_path, small_parallel_chunks):
    """what the workers cannot line up -- stored blocks (incompressible data), fixed-Huffman blocks, one
    block longer than the search range, several members in a row -- is decoded by the ordinary
    decoder from the last confirmed block boundary (a bit offset, with the 32 KiB before it)"""
    rng = random.Random(50)
    fq = fastq(rng, 4000)
    cases = {
        "stored_in_the_middle": fq + rng.randbytes(200000) + fq,                       # deflate stores the random part
        "fixed_blocks": fq,                                                              # Z_FIXED: no dynamic header anywhere
        "zeros_one_long_block": fq[:50000] + bytes(3_000_000) + fq[:50000],
        "tiny": b"@r\nACGT\n+\nFFFF\n",
    }
    for name, data in cases.items():
        p = tmp_path / (name + ".gz")
        strat = zlib.Z_FIXED if name == "fixed_blocks" else zlib.Z_DEFAULT_STRATEGY
        p.write_bytes(member(data, strategy=strat) + member(fq[:30000], level=1) + member(b""))
        got, rc, chunks, fb = _read_counting_chunks(p, 3, rng)
        assert (got, rc, fb) == (data + fq[:30000], 0, 0), name


def test_parallel_decoder_on_damaged_members_behaves_like_zlib(tmp_path, small_parallel_chunks, monkeypatch):
    rng = random.Random(51)
    base = [member(fastq(rng, 1500), level=lv) for lv in (1, 6)]
    for trial in range(120):
        blob = bytearray(b"".join(rng.sample(base, rng.randrange(1, 3))))
        kind, pos = trial % 4, rng.randrange(len(blob))
        if kind == 0:
            blob = blob[:pos]
        elif kind == 1:
            blob[pos] ^= 1 << rng.randrange(8)
        elif kind == 2:
            blob[pos:pos + 20] = rng.randbytes(20)
        else:
            del blob[pos:pos + rng.randrange(1, 5)]
        if len(blob) < 18:
            continue
        p = tmp_path / "d.gz"
        p.write_bytes(bytes(blob))
        want = zlib_mode(p, monkeypatch, chunk=50000)
        got = read_all(p, helpers=3, chunk=50000)
        assert got[:2] == want[:2], (trial, kind, pos)
