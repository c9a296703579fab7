#!/usr/bin/env python3
"""Measure the multi-sample matrix path (VCFConvert::count + MultiCount::printNormMatrix,
what `ntsmVCF -p` does) on the GPU box, next to the reference's own classes on the host cores.

Workload: the human_sites_n10 panel (96 287 sites, 1 270 317 k-mers) against a synthetic multi-sample VCF with one
SNP line per site and 2 504 phased samples (the 1000 Genomes width): a reference genome with every site's window
planted, genotypes drawn per site from a random allele frequency.  One run = VCF text -> matrix.tsv + center.txt.

  ours       ntsm_vcf_convert + ntsm_vcf_output_matrix through the C ABI (host parse, GPU inserts, GPU norm matrix,
             host formatting), wall clock; the kernels' device time by CUDA events (ntsm_multi_kernel_ms)
  reference  tools/ref_vcf_harness.cpp (the reference's MultiCount + VCFConvert, unmodified) with all host threads on
             the FIRST --ref-sites sites of the same workload (bounded: the full job takes the reference many minutes),
             and our two files for that sample compared byte for byte with the reference's

Prints one JSON object (not a bench.py line).   usage: python tools/bench_matrix.py [--sites N] [--samples S] [--ref-sites R]
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_vcf_harness")
PANEL = os.path.join(ROOT, "data", "human_sites_n10.fa.gz")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def read_panel(n_sites):
    """[(name, ref_window, alt_base)] of the first n_sites sites; the window is rebuilt from the 3-13 k-mers a record lists
    (tests/panel_util.py); (name, None, None) for the few sites whose records do not pin it down."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import random

    import panel_util
    with gzip.open(PANEL, "rt") as fh:
        lines = [l for _, l in zip(range(4 * n_sites), fh)]
    return panel_util.panel_windows(lines, random.Random(7))


def write_inputs(d, sites, n_samples, seed, tag):
    """ref.fa, sites.fa (the panel's own records for these sites), in.vcf under d; returns paths + VCF bytes."""
    rng = np.random.default_rng(seed)
    pads = rng.integers(20, 60, len(sites))
    letters = np.frombuffer(b"ACGT", np.uint8)
    pieces, positions, at = [], [], 0
    for (name, w, alt), pad in zip(sites, pads):
        pieces.append(letters[rng.integers(0, 4, pad)].tobytes().decode())
        at += int(pad)
        positions.append(at + 16)
        if w is not None:
            pieces.append(w)
            at += len(w)
    ref = os.path.join(d, tag + "_ref.fa")
    with open(ref, "w") as fh:
        fh.write(">chr1\n" + "".join(pieces) + "ACGT" * 16 + "\n")
    sites_path = os.path.join(d, tag + "_sites.fa")
    with gzip.open(PANEL, "rt") as src, open(sites_path, "w") as dst:
        for _ in range(4 * len(sites)):
            dst.write(src.readline())
    gt = np.frombuffer(b"0|0\t0|1\t1|0\t1|1\t", np.uint8).view(np.uint32)      # one uint32 per genotype + tab
    vcf = os.path.join(d, tag + "_in.vcf")
    lines = 0
    with open(vcf, "wb") as fh:
        fh.write(("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" +
                  "\t".join("HG%05d" % i for i in range(n_samples)) + "\n").encode())
        for (name, w, alt), pos in zip(sites, positions):
            if alt is None:
                continue
            p = rng.random()
            codes = (rng.random(n_samples) < p).astype(np.uint8) * 2 + (rng.random(n_samples) < p)
            row = gt[codes].view(np.uint8)
            fh.write(("chr1\t%d\t%s\t%s\t%s\t.\tPASS\t.\tGT\t" % (pos, name, w[15], alt)).encode())
            fh.write(row[:-1].tobytes())
            fh.write(b"\n")
            lines += 1
    return sites_path, ref, vcf, lines, os.path.getsize(vcf)


def run_ours(sites_path, ref, vcf, prefix, threads):
    import ctypes as C

    import ntsm_b200
    t0 = time.perf_counter()
    vc = ntsm_b200.VCFConvert(sites_path, ref, threads=threads)
    t1 = time.perf_counter()
    vc.count(vcf)
    t2 = time.perf_counter()
    vc.outputMatrix(prefix)
    t3 = time.perf_counter()
    ms = (C.c_double * 3)()
    cells = C.c_uint64()
    ntsm_b200.lib().ntsm_multi_kernel_ms(vc.counts._h, ms, C.byref(cells))
    return vc, {"setup_s": t1 - t0, "count_s": t2 - t1, "output_matrix_s": t3 - t2, "kernel_ms": {"kmerize_lookup_lists": ms[0], "fill": ms[1], "norm_matrix": ms[2]},
                "cells_per_fill_pass": cells.value, "launches": vc.counts.launches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=96287)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--ref-sites", type=int, default=6000, help="sites of the bounded reference sample")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--profile", action="store_true", help="one run of ours only (under ncu): no warm-up, no one-thread run, no reference")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6454.6))
    panel = read_panel(args.sites)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        t = time.perf_counter()
        sites_path, ref, vcf, lines, vcf_bytes = write_inputs(d, panel, args.samples, 1, "full")
        log("inputs: %d sites, %d SNP lines, %d samples, VCF %.2f GB (%.1f s to write)" % (len(panel), lines, args.samples, vcf_bytes / 1e9, time.perf_counter() - t))
        if args.profile:
            vc, ours = run_ours(sites_path, ref, vcf, os.path.join(d, "full"), args.threads)
            print(json.dumps({"profile_run": ours}))
            return
        run_ours(sites_path, ref, vcf, os.path.join(d, "warm"), args.threads)                 # warm-up: CUDA context, page cache
        vc, ours = run_ours(sites_path, ref, vcf, os.path.join(d, "full"), args.threads)
        _, ours_1t = run_ours(sites_path, ref, vcf, os.path.join(d, "full1"), 1)
        n_kmers = vc._fp.sites.n_kmers
        inserts = lines * 26 * args.samples                                     # (k-mer, sample) pairs VCFConvert::count walks
        k_ms = ours["kernel_ms"]
        # algorithmic bytes: the fill (one pass: multi = 20 fits the byte) reads 2 bits per (line, sample), reads each touched cell and
        # writes it; the norm matrix reads every byte of the matrix once and writes one double per (site, sample), then reads them for the sums
        fill_bytes = 2 * ours["cells_per_fill_pass"] + lines * ((args.samples + 15) // 16) * 4
        norm_bytes = n_kmers * args.samples + 2 * len(panel) * args.samples * 8
        out = {
            "what": "ntsmVCF -p: multi-sample VCF -> PCA matrix + centre file (VCFConvert::count + MultiCount::printNormMatrix)",
            "workload": "human_sites_n10 (%d sites, %d k-mers) x %d samples, %d SNP lines, VCF %.2f GB in tmpfs" % (len(panel), n_kmers, args.samples, lines, vcf_bytes / 1e9),
            "ours": ours,
            "host_threads": args.threads,
            "ours_one_host_thread": {"count_s": ours_1t["count_s"], "output_matrix_s": ours_1t["output_matrix_s"]},
            "ours_total_s": ours["count_s"] + ours["output_matrix_s"],
            "ours_inserts_per_s": inserts / ours["count_s"],
            "matrix_bytes": n_kmers * args.samples,
            "roofline": {
                "bound": "hbm", "peak": hbm, "unit": "GB/s",
                "fill": {"algorithmic_bytes": fill_bytes, "ms": k_ms["fill"], "achieved": fill_bytes / 1e6 / max(k_ms["fill"], 1e-9),
                         "frac": fill_bytes / 1e6 / max(k_ms["fill"], 1e-9) / hbm},
                "norm_matrix": {"algorithmic_bytes": norm_bytes, "ms": k_ms["norm_matrix"], "achieved": norm_bytes / 1e6 / max(k_ms["norm_matrix"], 1e-9),
                                "frac": norm_bytes / 1e6 / max(k_ms["norm_matrix"], 1e-9) / hbm},
            },
            "output_bytes": os.path.getsize(os.path.join(d, "full_matrix.tsv")) + os.path.getsize(os.path.join(d, "full_center.txt")),
        }
        # the bounded reference sample: the first --ref-sites sites, all samples
        if os.path.exists(HARNESS):
            sub = panel[:args.ref_sites]
            s2, r2, v2, l2, b2 = write_inputs(d, sub, args.samples, 1, "sub")
            t = time.perf_counter()
            p = subprocess.run([HARNESS, s2, r2, v2, os.path.join(d, "refsub"), "19", "20", "31", "0", str(args.threads), "1"], capture_output=True, text=True)
            ref_wall = time.perf_counter() - t
            secs = [l for l in p.stderr.splitlines() if l.startswith("harness_seconds")]
            run_ours(s2, r2, v2, os.path.join(d, "oursub_warm"), args.threads)
            _, o2 = run_ours(s2, r2, v2, os.path.join(d, "oursub"), args.threads)
            same = all(open(os.path.join(d, "refsub" + x), "rb").read() == open(os.path.join(d, "oursub" + x), "rb").read() for x in ("_matrix.tsv", "_center.txt"))
            parts = secs[0].split() if secs else []
            ref_count, ref_out = (float(parts[2]), float(parts[4])) if parts else (None, None)
            out["cpu_baseline"] = {
                "kind": "reference", "cores": args.threads,
                "sample": "first %d sites (%d SNP lines) x %d samples through tools/ref_vcf_harness (the reference's classes, OpenMP %d threads)" % (len(sub), l2, args.samples, args.threads),
                "count_s": ref_count, "output_matrix_s": ref_out, "wall_s": ref_wall, "rc": p.returncode,
                "inserts_per_s": (l2 * 26 * args.samples / ref_count) if ref_count else None,
            }
            out["ours_on_the_same_sample"] = {"count_s": o2["count_s"], "output_matrix_s": o2["output_matrix_s"],
                                              "speedup_count_plus_output": ((ref_count + ref_out) / (o2["count_s"] + o2["output_matrix_s"])) if ref_count else None}
            out["files_equal_reference_on_sample"] = same
        else:
            out["cpu_baseline"] = {"unavailable": "oracle/_ref/ref_vcf_harness not built"}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
