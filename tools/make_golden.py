#!/usr/bin/env python3
"""Generate tests/golden/ from the REAL reference binary (oracle/_ref/ntsmCount).

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tools/make_golden.py

For every case it writes  tests/golden/cases/<name>/  holding the input files, `cmd.json`
(argv after the binary, file names relative to the case dir), and what the reference printed:
`stdout.txt`, `stderr.txt` (with the wall-time/RSS line removed) and `rc.txt`.
It also writes tests/golden/iter_vectors.tsv by running the reference's KseqHashIterator
(tools/ref_iter_dump.cpp compiled against /root/reference) over seeded sequences.

The inputs are deterministic (seeded), so re-running reproduces the same fixtures.
Nothing here is imported by the product.
"""
import gzip
import json
import os
import random
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ntsmCount")
REF_SRC = "/root/reference"
PANEL = os.path.join(REF_SRC, "data", "human_sites_n10.fa")
OUT = os.path.join(ROOT, "tests", "golden")

COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def revcomp(s):
    return s.translate(COMP)[::-1]


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def merge_kmers(kmers):
    """Overlap-merge the sliding k-mers of one allele record back into its window."""
    w = kmers[0]
    for km in kmers[1:]:
        for d in range(1, len(km) + 1):
            if w[-(len(km) - d):] == km[:len(km) - d] or d == len(km):
                w += km[len(km) - d:]
                break
    return w


def read_panel(path, n_sites):
    recs = []
    with open(path) as fh:
        for _ in range(2 * n_sites):
            h = fh.readline().rstrip("\n")
            s = fh.readline().rstrip("\n")
            recs.append((h, s))
    return recs


def fastq(reads, qual="I"):
    return "".join("@r%d\n%s\n+\n%s\n" % (i, s, qual * len(s)) for i, s in enumerate(reads))


def fasta(reads, width=0):
    out = []
    for i, s in enumerate(reads):
        out.append(">r%d some comment" % i)
        if width:
            out.extend(s[j:j + width] for j in range(0, len(s), width))
        else:
            out.append(s)
    return "\n".join(out) + "\n"


def mutate(rng, s, err):
    out = []
    for ch in s:
        if rng.random() < err:
            out.append(rng.choice([c for c in "ACGT" if c != ch]))
        else:
            out.append(ch)
    return "".join(out)


def panel_reads(rng, recs, n, length=150, err=0.01):
    """Reads drawn around panel windows (so they hit the table), random strand, with errors."""
    reads = []
    for _ in range(n):
        _, s = recs[rng.randrange(len(recs))]
        w = merge_kmers(s.split("N"))
        left = rng.randrange(0, length - len(w) + 1) if len(w) < length else 0
        r = rand_seq(rng, left) + w + rand_seq(rng, max(0, length - left - len(w)))
        r = r[:length] if rng.random() < 0.7 else r[rng.randrange(1, 40):][:length]
        r = mutate(rng, r, err)
        if rng.random() < 0.5:
            r = revcomp(r)
        if rng.random() < 0.05:
            p = rng.randrange(len(r))
            r = r[:p] + "N" * rng.randrange(1, 6) + r[p + 1:]
        reads.append(r)
    return reads


CASES = []


def case(name, files, argv, binary_files=None):
    CASES.append((name, files, argv, binary_files or {}))


def build_cases():
    rng = random.Random(20261017)
    recs2 = read_panel(PANEL, 2)
    mini_sites = "".join("%s\n%s\n" % r for r in recs2)

    # --- SURVEY 8(c) mini fixture ---------------------------------------------------
    case("mini", {
        "sites.fa": mini_sites,
        "reads.fq": fastq(["CCACGTAGCACTGCACCCCCAT", "taggggtccatctaagtgacNACGT",
                           "ACGTACGTACGTACGTAC", "CCACGTAGCACTGCACCCCCAT"]),
    }, ["-s", "sites.fa", "reads.fq"])

    # --- duplicate k-mers between sites: abort without -d, counted with -d ------------
    dup_sites = (">s1 ref\nACGTACGTTAGCTAGCTAGNCGTACGTTAGCTAGCTAGG\n"
                 ">s1 var\nACGTACGTTCGCTAGCTAGNCGTACGTTCGCTAGCTAGG\n"
                 ">s2 ref\nACGTACGTTAGCTAGCTAGNTTTTTTTTTTGGGGGGGGGG\n"
                 ">s2 var\nGGGGGGGGGGAAAAAAAAAANTTTTTTTTTTGGGGGGGGGC\n")
    dup_reads = ">a\nACGTACGTTAGCTAGCTAGG\n>b\nctagctagctaacgtacgt\n>c\nTTTTTTTTTTGGG\nGGGGGGGC\n"
    case("dupes_abort", {"sites.fa": dup_sites, "reads.fa": dup_reads}, ["-s", "sites.fa", "reads.fa"])
    case("dupes_allowed", {"sites.fa": dup_sites, "reads.fa": dup_reads}, ["-d", "-s", "sites.fa", "reads.fa"])
    # odd number of site records: the last site has no var list -> reference aborts when printing
    case("odd_sites", {"sites.fa": mini_sites + ">s3 ref\nACGTTGCATGCATGCAAGCTT\n", "reads.fa": dup_reads},
         ["-s", "sites.fa", "reads.fa"])

    # --- 300-site slice of the real panel with reads that hit it -----------------------
    recs300 = read_panel(PANEL, 300)
    sites300 = "".join("%s\n%s\n" % r for r in recs300)
    os.makedirs(os.path.join(OUT, "shared"), exist_ok=True)
    with open(os.path.join(OUT, "shared", "sites300.fa"), "w") as fh:
        fh.write(sites300)
    S300 = "../../shared/sites300.fa"
    reads = panel_reads(rng, recs300, 600)
    case("panel300_fq", {"reads.fq": fastq(reads)}, ["-s", S300, "reads.fq"])
    half = len(reads) // 2
    case("panel300_two_files_t2", {
        "a.fq": fastq(reads[:half]), "b.fa": fasta(reads[half:], width=60),
    }, ["-t", "2", "-s", S300, "a.fq", "b.fa"])
    case("panel300_gz", {}, ["-s", "sites.fa.gz", "reads.fq.gz"], binary_files={
        "reads.fq.gz": gzip.compress(fastq(reads[:300]).encode(), 6, mtime=0) +
                       gzip.compress(fastq(reads[300:]).encode(), 6, mtime=0),   # multi-member
        "sites.fa.gz": gzip.compress(sites300.encode(), 6, mtime=0),
    })
    # -m early stop (single thread => deterministic)
    case("panel300_m1", {"reads.fq": fastq(reads)}, ["-m", "1", "-s", S300, "reads.fq"])
    case("panel300_m0.5_two_files", {"a.fq": fastq(reads[:half]), "b.fq": fastq(reads[half:])},
         ["-m", "0.5", "-t", "1", "-s", S300, "a.fq", "b.fq"])
    case("panel300_m0_disabled", {"reads.fq": fastq(reads[:300])},
         ["-m", "0", "-s", S300, "reads.fq"])

    # --- other k (sites = random windows joined by N; -d because short k collides) -----
    for k in (5, 11, 16, 17, 25, 31):
        srng = random.Random(1000 + k)
        recs = []
        for i in range(40):
            w = rand_seq(srng, k + 12)
            c = (k + 12) // 2
            alt = w[:c] + srng.choice([b for b in "ACGT" if b != w[c]]) + w[c + 1:]
            for tag, ww in (("ref", w), ("var", alt)):
                kms = [ww[j:j + k] for j in range(0, 13) if c - k < j <= c and srng.random() < 0.8]
                if not kms:
                    kms = [ww[c - k + 1:c + 1]]
                recs.append((">site%d %s" % (i, tag), "N".join(kms)))
        sites = "".join("%s\n%s\n" % r for r in recs)
        rr = []
        for _ in range(200):
            _, s = recs[srng.randrange(len(recs))]
            w = merge_kmers(s.split("N"))
            r = rand_seq(srng, srng.randrange(0, 30)) + w + rand_seq(srng, srng.randrange(0, 30))
            r = mutate(srng, r, 0.02)
            rr.append(revcomp(r) if srng.random() < 0.5 else r)
        case("k%d" % k, {"sites.fa": sites, "reads.fq": fastq(rr)}, ["-d", "-k", str(k), "-s", "sites.fa", "reads.fq"])

    # --- parser / alphabet edge cases (panel300 sites so some k-mers hit) ----------------
    w0 = merge_kmers(recs300[0][1].split("N"))
    w1 = merge_kmers(recs300[1][1].split("N"))
    w2 = merge_kmers(recs300[4][1].split("N"))
    edge = {
        "crlf.fq": "@a x\r\n" + w0 + "\r\n+\r\n" + "I" * len(w0) + "\r\n@b\r\n" + w1 + "\r\n+\r\n" + "I" * len(w1) + "\r\n",
        "multiline.fa": ">a\n" + w0[:10] + "\n" + w0[10:20] + "\n\n" + w0[20:] + "\n>b desc\n" + w1 + "\n",
        "multiline.fq": "@a\n" + w0[:15] + "\n" + w0[15:] + "\n+a\n" + "I" * 15 + "\n" + "I" * (len(w0) - 15) + "\n@b\n" + w1 + "\n+\n" + "I" * len(w1) + "\n",
        "lower_u_iupac.fa": ">a\n" + w0.lower() + "\n>b\n" + w1.replace("T", "U") + "\n>c\n" + w2[:12] + "R" + w2[13:] + "\n>d\n" + w2.replace("T", "u") + "\n",
        "truncated_seq.fq": fastq([w0, w1, w2])[:-(len(w2) + 3 + len(w2) // 2)],
        "no_qual.fq": "@a\n" + w0 + "\n+\n" + "I" * len(w0) + "\n@b\n" + w1 + "\n+\n",
        "no_qual_noeol.fq": "@a\n" + w0 + "\n+\n" + "I" * len(w0) + "\n@b\n" + w1 + "\n+",
        "qual_short_stops_file.fq": "@a\n" + w0 + "\n+\n" + "I" * (len(w0) - 3) + "\n@b\n" + w1 + "\n+\n" + "I" * len(w1) + "\n",
        "qual_long.fq": "@a\n" + w0 + "\n+\n" + "I" * (len(w0) - 3) + "\n" + "IIIIIIII\n@b\n" + w1 + "\n+\n" + "I" * len(w1) + "\n",
        "qual_at_sign.fq": "@a\n" + w0 + "\n+\n@" + "I" * (len(w0) - 1) + "\n@b\n" + w1 + "\n+\n" + "I" * len(w1) + "\n",
        "short_and_empty.fa": ">e\n\n>s\nACGT\n>k18\n" + w0[:18] + "\n>k19\n" + w0[:19] + "\n>last_no_eol\n" + w1,
        "leading_garbage.fa": "garbage line\nmore\n>a\n" + w0 + "\n>b\n" + w1 + "\n",
        "plus_at_in_line.fa": ">a\n" + w0[:20] + "+" + w0[20:] + "\n>b\n" + w1[:20] + ">" + w1[20:] + "\n>c\n" + w2[:9] + "@" + w2[9:] + "\n",
        "spaces_tabs.fa": ">a\n" + w0[:21] + " " + w0[21:] + "\n>b\n" + w1[:25] + "\t" + w1[25:] + "\n>c\n " + w2 + "\n",
        "header_only_eof.fa": ">a\n" + w0 + "\n>b",
        "header_only_eof2.fa": ">a\n" + w0 + "\n>",
        "empty_file.fa": "",
        "only_newlines.fa": "\n\n\n",
        "cr_only_line.fa": ">a\n" + w0[:25] + "\n\r\n" + w0[25:] + "\n>b\n\r\n" + w1 + "\n",
        "name_tab_comment.fq": "@a\tcomment here\n" + w0 + "\n+\n" + "I" * len(w0) + "\n",
        "fasta_then_fastq.fa": ">a\n" + w0 + "\n@b\n" + w1 + "\n+\n" + "I" * len(w1) + "\n>c\n" + w2 + "\n",
        "long_read.fa": ">long\n" + "".join((w0, rand_seq(rng, 5000), "NNNNN", revcomp(w1), rand_seq(rng, 9000), w2, rand_seq(rng, 12000))) + "\n",
        "long_read_wrapped.fa": fasta(["".join((rand_seq(rng, 3000), w2, rand_seq(rng, 9000), revcomp(w0)))], width=80),
    }
    raw = {
        "raw_bytes_0_3.fa": b">a\n" + bytes("ACGT".index(c) for c in w0) + b"\n>b\n" + w1.encode() + b"\n",
        "high_bytes.fa": b">a\n" + w0[:19].encode() + b"\xff\x80" + w0[19:].encode() + b"\n",
    }
    for fname, text in edge.items():
        nm = "edge_" + fname.replace(".", "_")
        case(nm, {fname: text}, ["-s", S300, fname])
    for fname, blob in raw.items():
        nm = "edge_" + fname.replace(".", "_")
        case(nm, {}, ["-s", S300, fname], binary_files={fname: blob})
    # all edge files in one run, 3 threads
    allfiles = dict(edge)
    case("edge_all_t3", {}, ["-t", "3", "-s", S300] + ["../edge_%s/%s" % (f.replace(".", "_"), f) for f in sorted(allfiles)])

    # sites file itself with awkward formatting (multi-line records, CRLF, lowercase)
    s_multi = ""
    for h, s in recs300[:40]:
        s_multi += h + "\r\n" + s[:30].lower() + "\r\n" + s[30:] + "\r\n"
    case("sites_multiline_crlf", {"sites.fa": s_multi, "reads.fq": fastq(reads[:400])}, ["-s", "sites.fa", "reads.fq"])

    # --- gzip inputs the way gzread treats them: damage, truncation, garbage, BGZF ---------------
    # (own generator so the cases above keep their bytes).  What the reference sees of a damaged
    # stream is decided by zlib's gzread under kseq's 16 KiB reads; these pin it.
    import struct
    import zlib
    grng = random.Random(20261018)
    greads = panel_reads(grng, recs300, 1500)
    gtext = fastq(greads, qual="F").encode()                     # ~470 KB: ~29 reads of 16 KiB
    whole = gzip.compress(gtext, 6, mtime=0)

    def flipped(blob, frac, bit=3):
        b = bytearray(blob)
        b[int(len(b) * frac)] ^= 1 << bit
        return bytes(b)

    def bgzf(data, block=0xFF00, eof_marker=True):
        out = bytearray()
        for i in list(range(0, len(data), block)) + ([len(data)] if eof_marker else []):
            chunk = data[i:i + block]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = co.compress(chunk) + co.flush()
            out += b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(body) + 8 - 1)
            out += body + struct.pack("<II", zlib.crc32(chunk), len(chunk))
        return bytes(out)

    blocks = bgzf(gtext)
    two = gzip.compress(gtext[:200000], 9, mtime=0) + gzip.compress(gtext[200000:], 1, mtime=0)
    gz_cases = {
        "gz_single": whole,
        "gz_bitflip_mid": flipped(whole, 0.55),
        "gz_bitflip_early": flipped(whole, 0.02, 6),
        "gz_bad_crc": whole[:-8] + bytes([whole[-8] ^ 0xFF]) + whole[-7:],
        "gz_bad_isize": whole[:-1] + bytes([whole[-1] ^ 1]),
        "gz_truncated": whole[:int(len(whole) * 0.7)],
        "gz_truncated_in_trailer": whole[:-3],
        "gz_garbage_tail": whole + b"not a gzip member\n" * 3,
        "gz_two_members_second_damaged": flipped(two, 0.8),
        "gz_second_member_bad_method": two[:len(gzip.compress(gtext[:200000], 9, mtime=0)) + 2] + b"\x07" + two[len(gzip.compress(gtext[:200000], 9, mtime=0)) + 3:],
        "bgzf": blocks,
        "bgzf_no_eof_marker": bgzf(gtext, eof_marker=False),
        "bgzf_bitflip_block": flipped(blocks, 0.6, 1),
        "bgzf_truncated": blocks[:int(len(blocks) * 0.45)],
        "bgzf_then_plain_member": bgzf(gtext[:250000], eof_marker=False) + gzip.compress(gtext[250000:], 6, mtime=0),
    }
    for nm, blob in gz_cases.items():
        case(nm, {}, ["-t", "4", "-s", S300, "reads.fq.gz"], binary_files={"reads.fq.gz": blob})


def run_case(name, files, argv, binary_files):
    d = os.path.join(OUT, "cases", name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    for fn, text in files.items():
        with open(os.path.join(d, fn), "w", newline="") as fh:
            fh.write(text)
    for fn, blob in binary_files.items():
        with open(os.path.join(d, fn), "wb") as fh:
            fh.write(blob)
    p = subprocess.run([REF_BIN] + argv, cwd=d, capture_output=True)
    err = "".join(l for l in p.stderr.decode(errors="replace").splitlines(True) if not l.startswith("Time: "))
    with open(os.path.join(d, "stdout.txt"), "wb") as fh:
        fh.write(p.stdout)
    with open(os.path.join(d, "stderr.txt"), "w") as fh:
        fh.write(err)
    with open(os.path.join(d, "rc.txt"), "w") as fh:
        fh.write("%d\n" % p.returncode)
    with open(os.path.join(d, "cmd.json"), "w") as fh:
        json.dump({"argv": argv}, fh)
    # big site files are stored gzipped to keep the fixtures small (gzopen reads both)
    return p.returncode


def iter_vectors():
    exe = "/tmp/ref_iter_dump"
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-I" + REF_SRC, os.path.join(ROOT, "tools", "ref_iter_dump.cpp"), "-o", exe])
    rng = random.Random(7)
    lines = ["19\tACGTTGCATGCATGCAAGCTNACGTTGCATGCATGCAAGCTT"]
    for k in (1, 2, 5, 11, 15, 16, 17, 19, 25, 31):
        for _ in range(6):
            n = rng.randrange(0, 90)
            s = "".join(rng.choice("ACGTACGTACGTacgtNnUuRY-.") for _ in range(n))
            lines.append("%d\t%s" % (k, s))
        lines.append("%d\t%s" % (k, "A" * (k + 3)))
        lines.append("%d\t%s" % (k, "T" * (k + 3)))
        lines.append("%d\t%s" % (k, "\\x00\\x01\\x02\\x03" * 10))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    with open(os.path.join(OUT, "iter_vectors.tsv"), "w") as fh:
        fh.write(out)


def main():
    if not os.path.exists(REF_BIN):
        sys.exit("build the reference first: make -C oracle ref")
    os.makedirs(OUT, exist_ok=True)
    build_cases()
    for name, files, argv, bfiles in CASES:
        rc = run_case(name, files, argv, bfiles)
        print("%-40s rc=%d" % (name, rc))
    iter_vectors()
    print("wrote", len(CASES), "cases to", OUT)


if __name__ == "__main__":
    main()
