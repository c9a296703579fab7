#!/usr/bin/env python3
"""Generate tests/golden/vcf/ from the REAL reference classes (tools/ref_vcf_harness.cpp: the reference's
MultiCount + VCFConvert compiled unmodified from /root/reference; see its header for why the shipped
ntsmVCF binary cannot be used).

Run in the build container (where /root/reference exists):
    make -C oracle ref_vcf && python tools/make_golden_vcf.py

For every case it writes  tests/golden/vcf/<name>/  holding the inputs (`sites.fa`, `ref.fa[.gz]`, `in.vcf`),
`args.json` ({k, multi, window, dupes}) and what the reference produced: `out_matrix.tsv`, `out_center.txt`
(VCFConvert::outputMatrix), `out_counts_<j>.txt` (MultiCount::printCountsMax per sample), `out_mat.bin`
(m_matCounts), `stderr.bin` (the inconsistent-count warnings; raw bytes) and `rc.txt` (134 = the reference
aborted on an uncaught exception / failed assert).  Seeded, so re-running reproduces the same fixtures.
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
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_vcf_harness")
OUT = os.path.join(ROOT, "tests", "golden", "vcf")
SITES300 = os.path.join(ROOT, "tests", "golden", "shared", "sites300.fa")

GT = ["0|0", "0|1", "1|0", "1|1"]


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def site_record(name, window, k, alt):
    """The two records of one site as the real panel spells them: the window's k-mers joined by N."""
    half = len(window) // 2
    var = window[:half] + alt + window[half + 1:]
    n = len(window) - k + 1
    return ">%s ref\n%s\n>%s var\n%s\n" % (name, "N".join(window[j:j + k] for j in range(n)), name,
                                           "N".join(var[j:j + k] for j in range(n)))


def fasta(name, seq, width=0):
    if not width:
        return ">%s\n%s\n" % (name, seq)
    return ">%s\n%s\n" % (name, "\n".join(seq[i:i + width] for i in range(0, len(seq), width)))


def vcf_header(samples):
    return "##fileformat=VCFv4.2\n##source=make_golden_vcf\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" + \
        "".join("\t" + s for s in samples) + "\n"


def vcf_line(chrom, pos, rsid, ref, alt, gts):
    return "%s\t%d\t%s\t%s\t%s\t.\tPASS\t.\tGT%s\n" % (chrom, pos, rsid, ref, alt, "".join("\t" + g for g in gts))


def make_genome(rng, length, positions, ref_base="A"):
    g = list(rand_seq(rng, length))
    for p in positions:
        g[p - 1] = ref_base
    return "".join(g)


def write_case(name, sites, ref, vcf, k=19, multi=20, window=31, dupes=0, gz_ref=False):
    d = os.path.join(OUT, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    open(os.path.join(d, "sites.fa"), "w").write(sites)
    ref_name = "ref.fa.gz" if gz_ref else "ref.fa"
    if gz_ref:
        with gzip.GzipFile(os.path.join(d, ref_name), "wb", mtime=0) as fh:
            fh.write(ref.encode())
    else:
        open(os.path.join(d, ref_name), "w").write(ref)
    open(os.path.join(d, "in.vcf"), "wb").write(vcf.encode() if isinstance(vcf, str) else vcf)
    json.dump({"k": k, "multi": multi, "window": window, "dupes": dupes, "ref": ref_name}, open(os.path.join(d, "args.json"), "w"))
    p = subprocess.run([HARNESS, "sites.fa", ref_name, "in.vcf", "out", str(k), str(multi), str(window), str(dupes)],
                       cwd=d, capture_output=True)
    rc = p.returncode if p.returncode >= 0 else 128 - p.returncode
    # what the reference wrote before dying is not a result: keep outputs of completed runs only
    if rc != 0:
        for f in os.listdir(d):
            if f.startswith("out_"):
                os.remove(os.path.join(d, f))
        # libstdc++'s terminate / glibc's assert text is not part of the contract
        open(os.path.join(d, "stderr.bin"), "wb").write(b"")
    else:
        open(os.path.join(d, "stderr.bin"), "wb").write(p.stderr)
    open(os.path.join(d, "rc.txt"), "w").write("%d\n" % rc)
    print("%-28s rc=%d stderr=%d bytes" % (name, rc, len(p.stderr)))


def basic_inputs(rng, n_sites, length=4000, k=19, window=31, alt="C", spacing=None):
    half = window // 2
    if spacing:
        positions = [200 + i * spacing for i in range(n_sites)]
    else:
        positions = sorted(rng.sample(range(half + 50, length - half - 50, 80), n_sites))
    g = make_genome(rng, length, positions)
    sites = "".join(site_record("rs%d" % (i + 1), g[p - half - 1:p + half], k, alt) for i, p in enumerate(positions))
    return g, positions, sites


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build the harness first: make -C oracle ref_vcf")
    os.makedirs(OUT, exist_ok=True)
    rng = random.Random(20260218)

    # 1. the four genotypes, an unknown genotype (reads as hom1), a non-SNP ALT (site stays missing everywhere)
    g, pos, sites = basic_inputs(rng, 4)
    vcf = vcf_header(["S1", "S2", "S3", "S4"])
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|0", "0|1", "1|1", "./."])
    vcf += vcf_line("chr1", pos[1], "rs2", "A", "C", ["0|1", "1|1", "0|0", "1|0"])
    vcf += vcf_line("chr1", pos[2], "rs3", "A", "CT", ["0|1", "1|1", "0|0", "1|0"])
    vcf += vcf_line("chr1", pos[3], "rs4", "A", "C", ["1|1", "1|1", "1|1", "1|1"])
    write_case("basic", sites, fasta("chr1 some description", g), vcf)

    # 2. missing values in the middle: the matrix stream's precision switches to 19 digits and stays
    g, pos, sites = basic_inputs(rng, 12, length=6000)
    samples = ["HG%05d" % i for i in range(7)]
    vcf = vcf_header(samples)
    for i, p in enumerate(pos):
        if i in (3, 7):
            continue                                     # site absent from the VCF: missing for every sample
        gts = [rng.choice(GT) for _ in samples]
        vcf += vcf_line("chr1", p, "rs%d" % (i + 1), "A", "C", gts)
    write_case("missing_precision", sites, fasta("chr1", g, 60), vcf)

    # 3. thirds and sevenths: values that need all 19 digits, before and after the switch
    g, pos, sites = basic_inputs(rng, 6, length=4000)
    samples = ["a", "b", "c"]
    vcf = vcf_header(samples)
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1", "0|0", "1|1"])
    vcf += vcf_line("chr1", pos[2], "rs3", "A", "C", ["0|1", "0|1", "0|0"])
    vcf += vcf_line("chr1", pos[3], "rs4", "A", "C", ["1|1", "0|1", "0|1"])
    vcf += vcf_line("chr1", pos[5], "rs6", "A", "C", ["0|0", "0|0", "0|1"])
    write_case("fractions", sites, fasta("chr1", g), vcf, multi=7)

    # 4. overlapping sites with -d: shared k-mers, first writer wins, the reference warns about the rest
    g, pos, sites = basic_inputs(rng, 5, length=3000, spacing=6)
    samples = ["S%d" % i for i in range(5)]
    vcf = vcf_header(samples)
    for i, p in enumerate(pos):
        vcf += vcf_line("chr1", p, "rs%d" % (i + 1), "A", "C", [rng.choice(GT) for _ in samples])
    write_case("overlap_dupes", sites, fasta("chr1", g), vcf, dupes=1)
    # ... and without -d: the duplicates leave the table and the printers abort (rc 134)
    write_case("overlap_no_dupes_aborts", sites, fasta("chr1", g), vcf, dupes=0)

    # 5. the same site twice in the VCF with different genotypes: warnings, first line wins
    g, pos, sites = basic_inputs(rng, 3)
    vcf = vcf_header(["x", "y", "z"])
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|0", "0|1", "1|1"])
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1", "0|1", "0|0"])
    vcf += vcf_line("chr1", pos[1], "rs2", "A", "C", ["1|1", "0|0", "0|1"])
    vcf += vcf_line("chr1", pos[1], "rs2", "A", "C", ["0|0", "1|1", "0|1"])
    write_case("repeated_site", sites, fasta("chr1", g), vcf)

    # 6. -m 200: the byte holds 400 mod 256 = 144 for homozygotes, every later insert of 400 "differs"
    g, pos, sites = basic_inputs(rng, 3)
    vcf = vcf_header(["x", "y"])
    for i, p in enumerate(pos):
        vcf += vcf_line("chr1", p, "rs%d" % (i + 1), "A", "C", [GT[i % 4], GT[(i + 3) % 4]])
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|0", "1|1"])
    write_case("multi_wraps_byte", sites, fasta("chr1", g), vcf, multi=200)
    write_case("multi_1", sites, fasta("chr1", g), vcf, multi=1)

    # 7. other k / window
    for k, window in ((21, 31), (19, 35), (25, 41), (31, 31), (11, 21)):
        g, pos, sites = basic_inputs(rng, 6, k=k, window=window, length=5000)
        samples = ["S%d" % i for i in range(4)]
        vcf = vcf_header(samples)
        for i, p in enumerate(pos):
            vcf += vcf_line("chr1", p, "rs%d" % (i + 1), "A", "C", [rng.choice(GT) for _ in samples])
        write_case("k%d_w%d" % (k, window), sites, fasta("chr1", g), vcf, k=k, window=window)

    # 8. several chromosomes, multi-line gzipped reference, lower case and N in the genome, a name used twice (the later record wins)
    chroms = {}
    all_sites = ""
    vcf = vcf_header(["S1", "S2", "S3"])
    n = 0
    for c in ("chr1", "chr2", "chrX"):
        g, pos, _ = basic_inputs(rng, 4, length=3000)
        gl = list(g)
        for j in range(0, len(gl), 7):
            gl[j] = gl[j].lower()
        gl[pos[1] - 5] = "N"                              # an N inside site 2's window: fewer k-mers from the genome than the panel lists
        g2 = "".join(gl)
        chroms[c] = g2
        for p in pos:
            n += 1
            all_sites += site_record("rs%d" % n, g[p - 16:p + 15], 19, "G")
            vcf += vcf_line(c, p, "rs%d" % n, "A", "G", [rng.choice(GT) for _ in range(3)])
    ref = fasta("chr1 first spelling", rand_seq(rng, 3000), 70) + "".join(fasta(c + " x", s, 70) for c, s in chroms.items())
    write_case("multi_chrom_gz_ref", all_sites, ref, vcf, gz_ref=True)

    # 9. line-level oddities: REF ".", REF longer than one base (still taken), unphased and \r-terminated genotypes,
    #    data line before the header line, last line without its newline (dropped)
    g, pos, sites = basic_inputs(rng, 6)
    vcf = "##fileformat=VCFv4.2\n"
    vcf += vcf_line("chr1", pos[5], "rs6", "A", "C", ["1|1", "1|1"])          # before #CHROM: ignored
    vcf += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
    vcf += vcf_line("chr1", pos[0], "rs1", ".", "C", ["0|1", "1|1"])          # skipped
    vcf += vcf_line("chr1", pos[1], "rs2", "AT", "C", ["0|1", "1|1"])         # taken
    vcf += vcf_line("chr1", pos[2], "rs3", "A", "C", ["0/1", "1/1"])          # unphased: both read as hom1
    vcf += vcf_line("chr1", pos[3], "rs4", "A", "C", ["0|1", "1|1"]).replace("\n", "\r\n")   # "1|1\r" reads as hom1
    vcf += vcf_line("chr1", pos[4], "rs5", "A", "C", ["1|0", "0|0"]).rstrip("\n")            # no newline: dropped
    write_case("line_oddities", sites, fasta("chr1", g), vcf)

    # 10. a site whose window runs off the end of its chromosome (strncpy stops at the sequence's NUL)
    rng2 = random.Random(7)
    g1 = make_genome(rng2, 1000, [300, 990])
    g2 = make_genome(rng2, 800, [797])
    sites = site_record("rs1", g1[300 - 16:300 + 15], 19, "C") + site_record("rsEnd", (g1[990 - 16:] + "ACGT" * 8)[:31], 19, "C") + \
        site_record("rsEnd2", (g2[797 - 16:] + "TTGCA" * 8)[:31], 19, "C")
    vcf = vcf_header(["S1", "S2"]) + vcf_line("chr1", 300, "rs1", "A", "C", ["0|1", "1|1"]) + \
        vcf_line("chr1", 990, "rsEnd", "A", "C", ["0|1", "0|0"]) + vcf_line("chr2", 797, "rsEnd2", "A", "C", ["0|1", "1|1"]) + \
        vcf_line("chr1", 1001, "rsPast", "A", "C", ["0|1", "1|1"]) + vcf_line("chr2", 786, "rsShort", "A", "C", ["1|1", "0|1"])
    write_case("window_past_chrom_end", sites, fasta("chr1", g1) + fasta("chr2", g2), vcf)

    # 11. no sample columns at all; no data lines
    g, pos, sites = basic_inputs(rng, 2)
    vcf = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\n" + "chr1\t%d\trs1\tA\tC\t.\tPASS\t.\tGT\n" % pos[0]
    write_case("zero_samples", sites, fasta("chr1", g), vcf)
    write_case("header_only", sites, fasta("chr1", g), vcf_header(["S1", "S2"]))

    # 12. where the reference dies (rc 134): unknown chromosome, too few / too many sample columns, empty line in the header,
    #     unpaired last site record
    g, pos, sites = basic_inputs(rng, 2)
    write_case("abort_unknown_chrom", sites, fasta("chr1", g), vcf_header(["S1"]) + vcf_line("chr9", pos[0], "rs1", "A", "C", ["0|1"]))
    write_case("abort_few_columns", sites, fasta("chr1", g), vcf_header(["S1", "S2"]) + vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1"]))
    write_case("abort_empty_header_line", sites, fasta("chr1", g), "##x\n\n" + vcf_header(["S1"]) + vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1"]))
    write_case("abort_bad_pos", sites, fasta("chr1", g), vcf_header(["S1"]) + "chr1\tabc\trs1\tA\tC\t.\tPASS\t.\tGT\t0|1\n")
    odd = sites + ">rsOdd ref\n" + rand_seq(rng, 19) + "\n"
    write_case("abort_unpaired_site", odd, fasta("chr1", g), vcf_header(["S1"]) + vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1"]))

    # 13. a slice of the real panel (tests/golden/shared/sites300.fa): genome with the 300 windows planted, 24 samples
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import panel_util
    wins = panel_util.panel_windows(open(SITES300).read().splitlines(), rng)      # the panel lists 3-13 of a window's 13 k-mers
    pieces, positions, at = [], [], 0
    for name, w, alt in wins:
        if w is None:
            continue
        pad = rand_seq(rng, rng.randrange(5, 60))
        pieces.append(pad + w)
        at += len(pad)
        positions.append((name, at + 16, w[15], alt))
        at += len(w)
    g = "".join(pieces) + rand_seq(rng, 100)
    samples = ["NA%05d" % (18500 + i) for i in range(24)]
    vcf = vcf_header(samples)
    for name, p, r, a in positions:
        if rng.random() < 0.03:
            continue
        vcf += vcf_line("chr1", p, name, r, a, [rng.choice(GT) if rng.random() > 0.02 else "./." for _ in samples])
    write_case("panel300_24samples", open(SITES300).read(), fasta("chr1", g, 80), vcf)

    # 14. tabs at the end of lines: `while (getline(ss, item, '\t'))` stops on the nothing after a last tab (no empty last
    #     sample ID / column), while an empty field BETWEEN two tabs is a column (an unknown genotype: hom1)
    g, pos, sites = basic_inputs(rng, 5)
    vcf = "##x\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3\t\n"
    vcf += vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1", "1|1", "0|0"]).replace("\n", "\t\n")
    vcf += vcf_line("chr1", pos[1], "rs2", "A", "C", ["1|1", "", "0|1"])
    vcf += vcf_line("chr1", pos[2], "rs3", "A", "C", ["1|1", "0|1", ""]).replace("\n", "\t\n")      # "...0|1\t\t": third column empty
    vcf += vcf_line("chr1", pos[3], "rs4", "A", "C", ["0|1", "1|0", "1|1"])
    write_case("trailing_tabs", sites, fasta("chr1", g), vcf)
    vcf2 = vcf_header(["S1", "S2"]) + vcf_line("chr1", pos[0], "rs1", "A", "C", ["0|1", "1|1"]).replace("\n", "\t\t\n")
    write_case("abort_two_trailing_tabs", sites, fasta("chr1", g), vcf2)
    write_case("zero_samples_trailing_tab", sites, fasta("chr1", g),
               "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t\n" + "chr1\t%d\trs1\tA\tC\t.\tPASS\t.\tGT\n" % pos[0] +
               "chr1\t%d\trs2\tA\tC\t.\tPASS\t.\tGT\t\n" % pos[1])


if __name__ == "__main__":
    main()
