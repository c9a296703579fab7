#!/usr/bin/env python3
"""bench.py -- ntsmCount counting hot path on B200: Gbases/s, % of HBM roofline, CPU reference beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): synthetic 30x human short reads -- 150 bp, 1 % substitution
error, 0.5 % of reads with an N run -- sampled from a 3.1 Gb diploid genome with every window of
data/human_sites_n10.fa.gz planted once per haplotype; ~100 Gbases per GPU, packed 2-bit + N-mask.
One step = one whole counting job over the rank's shard: zero counts -> count kernel over the
packed stream -> (N>1: one NCCL u32 all-reduce of counts + one u64 all-reduce of tallies) ->
per-site max/sum kernel.  `value` times that with the packed shard resident in HBM (CUDA events,
max over ranks); `roofline` is the count kernel alone against the measured HBM peak.
`e2e` (headline) is what the reference arm does too: FingerPrint::computeCounts over plain FASTQ FILES
(ntsm_count_files: parse + pack on the host cores into pinned batches, H2D, count) + ntsm_finalize
(all-reduce, per-site reduce, D2H of the rows), wall clock between device syncs, max over ranks, copies
counted by the library.  Beside it: `e2e_gz` (8 .fq.gz; 2 .fq.gz with single members cut between helper
threads), `e2e_ascii*` (insertCount over ASCII reads in pinned host memory: host packers, the device
packer, or both), `e2e_packed` (already packed pinned stream: what PCIe alone allows).  At N > 1 also
`strong_scaling` (one 100-Gbase job cut N ways) and `check.n_gpu_equals_1_gpu`.  `cpu_baseline` and
`--impl reference`: the unmodified reference binary on 1.5 Gbases of the same workload, 16 host threads.
`matrix_path` (rank 0): the multi-sample matrix path of `ntsmVCF -p` on its own workload and clock,
measured by tools/bench_matrix.py next to the reference's MultiCount + VCFConvert classes on the host cores.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PANEL = os.path.join(ROOT, "data", "human_sites_n10.fa.gz")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ntsmCount")
PORT_BIN = os.path.join(ROOT, "oracle", "_build", "ntsm_oracle")
READ_LEN = 150
METRIC = "ntsmCount Gbases/s (synthetic 30x human 150bp reads vs human_sites_n10, k=19)"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries do not know that: NCCL prints its version banner
# to stdout when the host exports NCCL_DEBUG=VERSION (and honours NCCL_DEBUG_FILE only above that
# level).  So the real stdout is put aside, file descriptor 1 is pointed at stderr for the life of the
# process, and the JSON line is the only thing ever written to the saved descriptor.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent sampling (NVML) of SM clocks and throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonSwPowerCap if hasattr(nv, "nvmlClocksEventReasonSwPowerCap") else 0x4: "sw_power_cap",
                     0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that rather than inventing clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- CPU reference
def run_reference(paths, threads):
    """One timed run of the CPU implementation on the sample files -> (seconds, bases, stdout bytes, kind)."""
    if os.path.exists(REF_BIN):
        exe, kind = REF_BIN, "reference"
    else:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        exe, kind = PORT_BIN, "port"
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-t", str(threads), "-s", PANEL] + paths, capture_output=True, env=env)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + p.stderr.decode()[-400:])
    err = p.stderr.decode()
    m = re.search(r"Total Bases Considered: (\d+)", err)
    tm = re.search(r"Time: ([0-9.eE+-]+) s", err)
    secs = float(tm.group(1)) if tm else dt          # the tool's own Time: line, wall as fallback
    return secs, int(m.group(1)), p.stdout, kind


def config_dict(args, per_gpu_gbases):
    return {"workload": "cfg2: synthetic 30x human short reads (150bp, 1% error, 0.5% reads with N run) vs human_sites_n10.fa (96287 sites, 1270317 k-mers), k=19",
            "per_gpu_gbases": per_gpu_gbases, "genome_mb": args.genome_mb, "read_len": READ_LEN,
            "l2": "inputs (%.1f GB packed per GPU) far exceed the 126 MB L2; no flush needed" % (per_gpu_gbases * 0.3775),
            "parallelism": "reads sharded over %d GPU(s), table replicated, one NCCL all-reduce per job" % args.gpus}


def reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference binary (oracle/_ref/ntsmCount, else the C port) on the box's
    host cores, FASTQ files -> counts file, on a bounded sample of the same workload: >= 1.5 Gbases per step
     so that its ~0.7 s site-table build is a few percent of a step, not a quarter."""
    import torch
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 16)
    n_reads = env_int("NTSM_BENCH_REF_READS", 10_000_000)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    tmp = tempfile.mkdtemp(prefix="ntsm_ref_", dir=scratch_dir())
    try:
        paths = make_fastq_files(dev, n_reads, 7, args.genome_mb, cores, tmp)
        times, bases, kind = [], 0, "reference"
        for i in range(args.warmup + args.steps):
            secs, bases, _, kind = run_reference(paths, cores)
            if i >= args.warmup:
                times.append(secs)
            log("reference step %d: %.2f s" % (i, secs))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ms = 1000 * sum(times) / len(times)
    v = bases / (ms / 1000) / 1e9
    sample = "%d x %dbp reads (%.2f Gbases) of the bench workload in %d plain FASTQ files, ntsmCount -t %d (-t only parallelises over files)" % (
        n_reads, READ_LEN, bases / 1e9, cores, cores)
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config_dict(args, args.gbases),
        "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "ntsmCount -t %d -s sites.fa f1.fq .. f%d.fq (FASTQ files -> counts file), the tool's own Time: line" % (cores, cores)},
        "gpu_launches": 0}))


def scratch_dir():
    """Where sample FASTQ files go: tmpfs when there is one (the files are inputs, not what is measured)."""
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None


def make_fastq_files(device, n_reads, seed, genome_mb, n_files, outdir, prefix="sample"):
    """n_reads reads of the bench generator (same genome size => same hit density) as n_files FASTQ files."""
    from ntsm_b200 import synth
    wc, wl = synth.panel_windows(PANEL)
    g = synth.Genome(genome_mb * 1_000_000, wc, wl, seed, device)
    return synth.write_fastq_set(g, n_reads, READ_LEN, 0.01, seed + 1, n_files, outdir, prefix=prefix)


# ----------------------------------------------------------------------------- our arm
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ntsm_b200
    from ntsm_b200 import synth

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the counting path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    options = dict(kv.split("=") for kv in args.opt)
    options = {k: int(v) for k, v in options.items()}

    # ---- FingerPrint object on this rank's GPU ------------------------------------------
    t0 = time.time()
    panel, wc, wl = PANEL, None, None
    if args.synthetic_sites:          # cfg5 (kernel-only sweeps): 10^6-site synthetic panel instead of the human one
        from ntsm_b200 import synth_np
        panel = "/tmp/ntsm_bench_sites_%d_k%d.fa" % (args.synthetic_sites, args.k)
        win, n_kept = synth_np.synthetic_panel(panel, args.synthetic_sites, 5, k=args.k, flank=max(1, args.k - 4))     # k - 6 k-mers per allele (13 at k = 19)
        wc, wl = synth_np.panel_alleles_from_windows(win)
        log("rank %d: synthetic panel of %d sites written in %.1f s" % (rank, n_kept, time.time() - t0))
    sites = ntsm_b200.SiteSet(panel, args.k)
    fp = ntsm_b200.FingerPrint(sites, device=local, batch_bases=1 << 26, n_buffers=3, options=options)
    # a real (non-default) torch stream: the ABI reads a NULL handle as "use the ctx's own stream",
    # and torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    fp.set_stream(stream.cuda_stream)
    from ntsm_b200 import dist as ndist
    if world > 1:
        ndist.attach_comm(fp)          # rank 0's ncclUniqueId travels over torch.distributed; the all-reduce is the library's
    log("rank %d: panel + table ready in %.1f s (filter 2^%d bits, probe tables %.1f MiB, L2 window %d%%)" %
        (rank, time.time() - t0, fp.filter_bits, fp.probe_bytes / 2 ** 20, fp.l2_window))

    # ---- this rank's shard, generated on the device ----------------------------------------
    t0 = time.time()
    read_len = args.read_len
    n_reads = int(args.gbases * 1e9 / read_len) // 32 * 32
    if wc is None:
        wc, wl = synth.panel_windows(PANEL)
    genome = synth.Genome(args.genome_mb * 1_000_000, wc, wl, 2, dev)
    bases, mask, n_pos, n_bases = synth.make_packed_shard(genome, n_reads, read_len, args.err, seed=1000 + rank,
                                                          chunk_reads=(1 << 20) if read_len == READ_LEN else max(32, ((1 << 27) // read_len) // 32 * 32))
    # (the ASCII e2e leg regenerates the first reads of this shard chunk by chunk with the same seeds: 2^20 reads per chunk)
    torch.cuda.synchronize()
    alg_bytes = (3 * n_bases + 7) // 8 + 8 * n_reads          # 2-bit + N-mask per base, one u64 offset per read
    phys_bytes = bases.numel() * 4 + mask.numel() * 4           # what the kernel actually streams (separator instead of offsets)
    log("rank %d: %d reads / %.2f Gbases packed into %.2f GB in %.1f s" % (rank, n_reads, n_bases / 1e9, phys_bytes / 1e9, time.time() - t0))

    def job_resident(ev=None, pos=n_pos, nb=n_bases):
        fp.reset_async()
        if ev:
            ev[0].record(stream)
        fp.count_packed_device(bases.data_ptr(), mask.data_ptr(), pos, nb, stream.cuda_stream)
        if ev:
            ev[1].record(stream)
        fp.reduce_async()

    for _ in range(args.warmup):
        job_resident()
    barrier()
    l0 = fp.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        job_resident(kev[i])
    e1.record(stream)
    barrier()
    clocks = sampler.result() if rank == 0 else None
    launches = fp.launches - l0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    mr, mv, sr, sv, tot = fp.finalize()
    check = {"TK": int(tot[0]), "hits": int(tot[1]), "bases": int(tot[2]), "sites_covered": int(fp.sites_covered())}
    value = world * n_bases / (ms_step / 1000) / 1e9

    if args.kernel_only:       # variant sweeps: device-resident number only, not a bench line for the driver
        if rank == 0:
            emit(({"kernel_only": True, "value": value, "unit": "Gbases/s", "kernel_ms": ms_kernel, "ms_per_step": ms_step,
                              "n_gpus": world, "gbases_per_gpu": n_bases / 1e9, "check": check, "kernel": fp.kernel_name,
                              "read_len": read_len, "err": args.err, "k": args.k,
                              "n_sites": int(sites.n_sites), "n_kmers": int(sites.n_kmers), "filter_bits": int(fp.filter_bits),
                              "probe_mib": fp.probe_bytes / 2 ** 20, "l2_window_pct": fp.l2_window, "options": options}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- other regimes, kernel-resident, one GPU only (BASELINE cfg 1 and cfg 3 shapes; cfg 5 needs its 10^6-site
    #      panel, 30 s to build: `--kernel-only --synthetic-sites 1000000`, numbers in DESIGN.md) ---------------
    regimes = None
    if world == 1 and not args.no_cpu:
        regimes = {}
        for name, gmb, rl, er, gb in (("cfg1_like_dense_hits_30Mb_genome", 30, READ_LEN, 0.01, 10.0), ("cfg3_like_20kb_reads_7.5pct_error", args.genome_mb, 20000, 0.075, 10.0)):
            g2 = genome if gmb == args.genome_mb else synth.Genome(gmb * 1_000_000, wc, wl, 2, dev)
            nr = int(gb * 1e9 / rl) // 32 * 32
            b2, m2, np2, nb2 = synth.make_packed_shard(g2, nr, rl, er, seed=77, chunk_reads=(1 << 20) if rl == READ_LEN else max(32, ((1 << 27) // rl) // 32 * 32))
            for _ in range(3):
                fp.reset_async(); fp.count_packed_device(b2.data_ptr(), m2.data_ptr(), np2, nb2, stream.cuda_stream); fp.reduce_async()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(3):
                fp.reset_async(); fp.count_packed_device(b2.data_ptr(), m2.data_ptr(), np2, nb2, stream.cuda_stream); fp.reduce_async()
            r1.record(stream)
            torch.cuda.synchronize()
            rows2 = fp.finalize()
            regimes[name] = {"value": nb2 / (r0.elapsed_time(r1) / 3 / 1000) / 1e9, "unit": "Gbases/s", "gbases": nb2 / 1e9,
                             "hits_per_kilobase": 1000.0 * float(rows2[4][1]) / nb2}
            del b2, m2, g2
        torch.cuda.synchronize()

    # ---- strong scaling: the SAME 100-Gbase job cut N ways (this rank counts the first 1/N of its shard) ----
    strong = None
    if world > 1:
        s_reads = (n_reads // world) // 32 * 32
        s_pos, s_bases = s_reads * (read_len + 1), s_reads * read_len
        for _ in range(args.warmup):
            job_resident(None, s_pos, s_bases)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for i in range(args.steps):
            job_resident(None, s_pos, s_bases)
        s1.record(stream)
        barrier()
        s_ms = max_over_ranks(s0.elapsed_time(s1) / args.steps)
        s_val = world * s_bases / (s_ms / 1000) / 1e9
        strong = {"scaling": "strong", "value": s_val, "unit": "Gbases/s", "ms_per_step": s_ms, "job_gbases": world * s_bases / 1e9,
                  "gbases_per_gpu": s_bases / 1e9, "fraction_of_weak_value": s_val / value,
                  "note": "one %.0f-Gbase job split over %d GPUs; step = zero counts -> count kernel -> NCCL all-reduce -> per-site reduce, "
                          "CUDA events, max over ranks" % (world * s_bases / 1e9, world)}

    # ---- N > 1: the combined result equals what ONE GPU counts from all the shards ----------------------
    n_equals_1 = None
    if world > 1:
        c_reads = min(n_reads, env_int("NTSM_BENCH_IDENTITY_READS", 2_000_000)) // 32 * 32
        c_pos, c_bases = c_reads * (read_len + 1), c_reads * read_len
        job_resident(None, c_pos, c_bases)
        rows_all = fp.finalize()                       # all-reduced over the N ranks
        cb, cm = bases[:c_pos // 16 + 1024].contiguous(), mask[:c_pos // 32 + 512].contiguous()     # + halo words
        gb = [torch.empty_like(cb) for _ in range(world)] if rank == 0 else None
        gm = [torch.empty_like(cm) for _ in range(world)] if rank == 0 else None
        dist.gather(cb, gb, dst=0)
        dist.gather(cm, gm, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            single = ntsm_b200.FingerPrint(sites, device=local, options=options)       # no communicator: one GPU on its own
            for r in range(world):
                single.count_packed_device(gb[r].data_ptr(), gm[r].data_ptr(), c_pos, c_bases)
            rows_one = single.finalize()
            n_equals_1 = all(np.array_equal(x, y) for x, y in zip(rows_all, rows_one))
            single.close()
            del gb, gm
        del cb, cm

    # ---- e2e legs ---------------------------------------------------------------------------
    avail = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            avail = int(line.split()[1]) * 1024
    local_world = env_int("LOCAL_WORLD_SIZE", world)
    host_threads = max(1, env_int("NTSM_BENCH_THREADS", (os.cpu_count() or 1) // max(1, local_world)))

    def timed_host_job(job, steps=args.steps, warm=max(1, args.warmup)):
        for _ in range(warm):
            job()
        times, rows = [], None
        for _ in range(steps):
            barrier()
            t = time.perf_counter()
            rows = job()
            torch.cuda.synchronize()
            times.append(max_over_ranks(time.perf_counter() - t))
        return 1000 * sum(times) / len(times), rows

    def pcie_per_step(fpx, before, runs):
        """bytes per step over PCIe, as counted by the library where it issues the copies"""
        now = fpx.pcie_bytes
        return (now[0] - before[0]) // runs, (now[1] - before[1]) // runs

    fp_a = ntsm_b200.FingerPrint(sites, device=local, batch_bases=1 << 24, n_buffers=host_threads + 6, options=options)
    if world > 1:
        ndist.attach_comm(fp_a)
    d2h_rows = 4 * 4 * sites.n_sites + 24

    # (a) HEADLINE e2e: the whole FingerPrint::computeCounts path, plain FASTQ files -> parse -> pack -> H2D ->
    #     count -> (all-reduce) -> rows on the host: what the reference arm does with the same kind of files.
    #     Files live in tmpfs; every file is read `f_passes` times per step so that a step lasts >= ~1 s.
    f_total = int(env_int("NTSM_BENCH_FILE_MREADS", 48) * 1e6) // max(1, local_world)        # reads per rank in the file set
    f_total = min(f_total, int(avail * 0.12 / max(1, local_world) / (2 * READ_LEN + 15)))
    n_files = max(host_threads, 4)
    fdir = tempfile.mkdtemp(prefix="ntsm_e2e_%d_" % rank, dir=scratch_dir())
    files = None
    try:
        tg = time.time()
        fpaths = synth.write_fastq_set(genome, f_total, READ_LEN, 0.01, 7000 + rank, n_files, fdir)
        fbytes = sum(os.path.getsize(p) for p in fpaths)
        # every file is read f_passes times per step: the box's whole file set is ~7 Gbases, a step should last >= 1 s
        # (>= 40 Gbases per step per box)
        f_passes = max(1, env_int("NTSM_BENCH_FILE_PASSES", 6 if local_world <= 2 else 12))
        log("rank %d: %d FASTQ files, %.2f GB, written in %.1f s" % (rank, len(fpaths), fbytes / 1e9, time.time() - tg))

        def job_files():
            fp_a.reset()
            fp_a.computeCounts(fpaths * f_passes, threads=host_threads)
            return fp_a.finalize()

        lf, pb = fp_a.launches, fp_a.pcie_bytes
        f_ms, f_rows = timed_host_job(job_files)
        f_launches = (fp_a.launches - lf) // (max(1, args.warmup) + args.steps)
        f_h2d, f_d2h = pcie_per_step(fp_a, pb, max(1, args.warmup) + args.steps)
        f_bases = int(f_rows[4][2]) // world if world > 1 else int(f_rows[4][2])
        f_val = world * f_bases / (f_ms / 1000) / 1e9
        files = {"value": f_val, "unit": "Gbases/s", "h2d_bytes_per_step": int(f_h2d), "d2h_bytes_per_step": int(f_d2h), "ms_per_step": f_ms, "gbases_per_step_per_gpu": f_bases / 1e9,
                 "files": len(fpaths), "passes_over_the_files": f_passes, "fastq_bytes_per_step": fbytes * f_passes,
                 "host_threads": min(host_threads, len(fpaths)), "launches_per_step": int(f_launches),
                 "api": "ntsm_count_files + ntsm_finalize (FingerPrint::computeCounts: plain FASTQ files in tmpfs -> parse -> pack -> "
                        "pinned batches -> H2D -> count -> rows on the host), wall clock between device syncs, max over ranks"}

        # (a') the same path on gzip'd input (cfg4's shape: 8 .fq.gz, -t 16 -> spare threads inflate; N = 1 only)
        gz = None
        if world == 1 and not args.no_cpu:
            import subprocess as sp
            gdir = os.path.join(fdir, "gz")
            gpaths = synth.write_fastq_set(genome, env_int("NTSM_BENCH_GZ_READS", 8_000_000), READ_LEN, 0.01, 9000, 8, gdir, prefix="lane")
            procs = [sp.Popen(["gzip", "-6", "-f", p]) for p in gpaths]
            for pr in procs:
                pr.wait()
            gpaths = [p + ".gz" for p in gpaths]
            gbytes = sum(os.path.getsize(p) for p in gpaths)

            def job_gz():
                fp_a.reset()
                fp_a.computeCounts(gpaths, threads=host_threads)
                return fp_a.finalize()

            g_ms, g_rows = timed_host_job(job_gz, steps=min(3, args.steps), warm=1)
            gz = {"value": float(g_rows[4][2]) / (g_ms / 1000) / 1e9, "unit": "Gbases/s", "ms_per_step": g_ms, "files": len(gpaths),
                  "gz_bytes": gbytes, "threads": host_threads,
                  "api": "ntsm_count_files on 8 .fq.gz (gzip -6), one parser thread per file with our own inflate (1 spare thread each: too few to cut members)"}
            # two whole-run files and 16 threads: the reference (and any one-thread-per-file reader) leaves 14 cores idle;
            # here 7 helper threads per file inflate chunks of the single gzip member in parallel (pargz.cpp).  Timed with and without.
            two = gpaths[:2]

            def job_gz2():
                fp_a.reset()
                fp_a.computeCounts(two, threads=host_threads)
                return fp_a.finalize()

            p_ms, p_rows = timed_host_job(job_gz2, steps=min(3, args.steps), warm=1)
            os.environ["NTSM_PARALLEL_GZ"] = "0"
            try:
                s_ms, s_rows = timed_host_job(job_gz2, steps=min(3, args.steps), warm=1)
            finally:
                del os.environ["NTSM_PARALLEL_GZ"]
            gz["two_files_16_threads"] = {"value": float(p_rows[4][2]) / (p_ms / 1000) / 1e9, "unit": "Gbases/s", "ms_per_step": p_ms,
                                          "one_thread_per_member_value": float(s_rows[4][2]) / (s_ms / 1000) / 1e9,
                                          "same_counts": bool(np.array_equal(p_rows[0], s_rows[0]) and int(p_rows[4][0]) == int(s_rows[4][0])),
                                          "note": "2 .fq.gz, -t 16: each single gzip member inflated by 7 helper threads (pargz) vs one thread per member"}
    finally:
        shutil.rmtree(fdir, ignore_errors=True)

    # (b) ASCII reads in PINNED host memory -> ntsm_insert_reads_fixed -> ntsm_finalize: host packers and the
    #     device packer (DMA of the ASCII bytes + pack_ascii_kernel) work the same queue of read blocks.
    #     The sample is the first a_reads reads of this rank's shard (same generator, same seeds).
    chunk = 1 << 20
    a_budget = min(int(avail * 0.2 / max(1, local_world)), env_int("NTSM_BENCH_E2E_ASCII_GB", 24) << 30)
    a_reads = max(chunk, min(n_reads, a_budget // READ_LEN) // chunk * chunk)
    ascii_lut = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device=dev)
    host_ascii = torch.empty((a_reads, READ_LEN), dtype=torch.uint8, pin_memory=True)
    for ci in range(a_reads // chunk):
        codes = synth.sample_reads(genome, chunk, READ_LEN, 0.01, (1000 + rank) * 1000003 + ci)
        host_ascii[ci * chunk:(ci + 1) * chunk].copy_(ascii_lut[codes.long()])
    del genome, codes
    torch.cuda.synchronize()
    a_bases = a_reads * READ_LEN
    a_pos = a_reads * ((READ_LEN + 8) & ~7)                  # the packers start every read at a multiple of 8 positions

    def job_ascii():
        fp_a.reset_async()
        fp_a.insertReadsFixed(host_ascii.data_ptr(), READ_LEN, READ_LEN, a_reads, threads=host_threads)
        return fp_a.finalize()

    def ascii_leg(device_pack, threads):
        nonlocal host_threads
        fp_a.set_option("device_pack", device_pack)
        keep, host_threads = host_threads, threads
        la, pb = fp_a.launches, fp_a.pcie_bytes
        try:
            ms, rows = timed_host_job(job_ascii, steps=min(3, args.steps), warm=1)
        finally:
            host_threads = keep
        runs = 1 + min(3, args.steps)
        return ms, rows, ((fp_a.launches - la) // runs,) + pcie_per_step(fp_a, pb, runs)

    a_ms, a_rows, a_io = ascii_leg(-1, host_threads)                # the library's default: feeders when < 8 host packers serve each GPU
    y_ms, y_rows, y_io = ascii_leg(1, host_threads)                 # hybrid forced: host packers + device packer
    h_ms, h_rows, h_io = ascii_leg(0, host_threads)                 # host packers only (round 1's path)
    d_ms, d_rows, d_io = ascii_leg(1, 0)                            # device packer only: no host core touches a base
    fp_a.set_option("device_pack", -1)
    log("rank %d: e2e ascii %.2f Gbases: default %.1f ms, hybrid %.1f ms, host packers only %.1f ms, device packer only %.1f ms (%d pack threads, %s)" %
        (rank, a_bases / 1e9, a_ms, y_ms, h_ms, d_ms, host_threads, ntsm_b200.lib().ntsm_pack_isa(None).decode()))

    # (c) the same reads, already packed, in pinned host memory (what PCIe alone allows)
    budget = min(phys_bytes, int(avail * 0.2 / max(1, local_world)), env_int("NTSM_BENCH_E2E_GB", 32) << 30)
    e_reads = min(n_reads, int(budget / (0.375 * (READ_LEN + 1)))) // 32 * 32
    e_pos, e_bases = e_reads * (READ_LEN + 1), e_reads * READ_LEN
    e_pad = synth.padded_positions(e_pos)
    hb = torch.zeros(e_pad // 16, dtype=torch.int32).pin_memory()
    hm = torch.full((e_pad // 32,), -1, dtype=torch.int32).pin_memory()
    hb[:e_pos // 16].copy_(bases[:e_pos // 16]); hm[:e_pos // 32].copy_(mask[:e_pos // 32])
    torch.cuda.synchronize()
    slice_pos = (1 << 26) // 8192 * 8192
    n_slices = (e_pos + slice_pos - 1) // slice_pos
    p_h2d = (e_pos // 32 + 2 * n_slices) * 12

    def job_host():
        fp.reset_async()
        fp.count_packed_host(hb.data_ptr(), hm.data_ptr(), e_pos, e_bases)
        return fp.finalize()

    e2e_ms, rows = timed_host_job(job_host, steps=min(3, args.steps), warm=1)
    e2e_val = world * e_bases / (e2e_ms / 1000) / 1e9
    e2e_check = int(rows[4][0])
    # cross-check: the ASCII legs (all three) and the packed leg agree on their common prefix of reads
    ascii_check = None
    if world == 1:
        c_reads = min(a_reads, e_reads) // 32 * 32
        fp.reset_async()
        fp.count_packed_host(hb.data_ptr(), hm.data_ptr(), c_reads * (READ_LEN + 1), c_reads * READ_LEN)
        want = fp.finalize()
        ascii_check = True
        for dp, th in ((-1, host_threads), (1, host_threads), (0, host_threads), (1, 0)):
            fp_a.set_option("device_pack", dp)
            fp_a.reset_async()
            fp_a.insertReadsFixed(host_ascii.data_ptr(), READ_LEN, READ_LEN, c_reads, threads=th)
            got = fp_a.finalize()
            ascii_check = ascii_check and all(np.array_equal(x, y) for x, y in zip(want, got))
        fp_a.set_option("device_pack", -1)
    del hb, hm, host_ascii

    # ---- rank 0: CPU reference on a bounded sample + byte-for-byte parity of our whole pipeline on it ----
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        cores = min(os.cpu_count() or 1, 16)
        n_s = env_int("NTSM_BENCH_CPU_READS", 10_000_000 if world == 1 else 2_000_000)
        tmp = tempfile.mkdtemp(prefix="ntsm_cpu_", dir=scratch_dir())
        try:
            paths = make_fastq_files(dev, n_s, 7, args.genome_mb, cores, tmp)
            secs, cb, ref_stdout, kind = run_reference(paths, cores)
            single = ntsm_b200.FingerPrint(sites, device=local, options=options)     # no communicator: this is a one-GPU check
            single.computeCounts(paths, threads=cores)
            parity = single.counts_text().encode() == ref_stdout
            single.close()
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        cpu = {"value": cb / secs / 1e9, "unit": "Gbases/s", "cores": cores, "kind": kind,
               "sample": "%d x %dbp reads (%.2f Gbases) of the same generator in %d plain FASTQ files, ntsmCount -t %d; %.1f s" % (
                   n_s, READ_LEN, cb / 1e9, cores, cores, secs)}
    # the multi-sample matrix path (`ntsmVCF -p`): its own workload and clock, measured by
    # tools/bench_matrix.py in a process of its own on this rank's GPU, next to the reference's classes on the host cores
    matrix_path = None
    if rank == 0 and not args.no_cpu:
        try:
            visible = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=visible[local] if local < len(visible) else str(local))
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_matrix.py"), "--threads", str(cores)],
                               capture_output=True, text=True, timeout=600, env=env)
            matrix_path = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-400:]}
        except Exception as e:      # the bench line must not depend on it
            matrix_path = {"error": repr(e)}
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak = 6650.0
        achieved = alg_bytes / (ms_kernel / 1000) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass

        def ascii_obj(ms, io, what, threads):
            return {"value": world * a_bases / (ms / 1000) / 1e9, "unit": "Gbases/s",
                    "h2d_bytes_per_step": int(io[1]), "d2h_bytes_per_step": int(io[2]),
                    "ms_per_step": ms, "gbases_per_step_per_gpu": a_bases / 1e9, "host_pack_threads": threads, "launches_per_step": int(io[0]),
                    "api": what}
        out = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": config_dict(args, args.gbases),      # the same dict the reference arm prints (exact shard size: roofline.gbases_per_launch)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic["dram_bytes_per_base"] * n_bases) if traffic else None,
                         "traffic_source": (traffic or {}).get("source"), "peak_source": peak_src,
                         "kernel": fp.kernel_name, "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                         "packed_bytes_per_launch": phys_bytes, "gbases_per_launch": n_bases / 1e9, "l2_window_pct": fp.l2_window,
                         "note": "HBM fraction as BASELINE asks; the kernel is bound by L1->L2 probe requests, see DESIGN.md"},
            "cpu_baseline": cpu,
            "e2e": files,
            "e2e_gz": gz,
            "e2e_ascii": ascii_obj(a_ms, a_io, "ntsm_insert_reads_fixed + ntsm_finalize, ASCII reads in pinned host memory, library defaults: %d host "
                                   "packer threads per GPU; the device packer (DMA of ASCII + pack_ascii_kernel) joins them on the same queue of "
                                   "read blocks when fewer than 8 packers serve a GPU" % host_threads, host_threads),
            "e2e_ascii_hybrid": ascii_obj(y_ms, y_io, "same, host packers AND the device packer forced on", host_threads),
            "e2e_ascii_host_pack_only": ascii_obj(h_ms, h_io, "same, device packer off (round 1's path)", host_threads),
            "e2e_ascii_device_pack_only": ascii_obj(d_ms, d_io, "same, zero host packers: every base goes over PCIe as ASCII and is packed on the GPU", 0),
            "e2e_packed": {"value": e2e_val, "unit": "Gbases/s", "h2d_bytes_per_step": p_h2d, "d2h_bytes_per_step": d2h_rows,
                           "ms_per_step": e2e_ms, "gbases_per_step_per_gpu": e_bases / 1e9,
                           "api": "ntsm_count_packed_host + ntsm_finalize (pinned host stream already packed)"},
            "strong_scaling": strong,
            "other_regimes_kernel_resident": regimes,
            "gpu_launches": int(launches), "clocks": clocks,
            "check": dict(check, e2e_TK=int(f_rows[4][0]), e2e_ascii_TK=int(a_rows[4][0]), e2e_packed_TK=e2e_check,
                          ascii_equals_packed=ascii_check, n_gpu_equals_1_gpu=n_equals_1),
            "parity_vs_reference_on_cpu_sample": parity,
            "matrix_path": matrix_path,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gbases", type=float, default=float(os.environ.get("NTSM_BENCH_GBASES", 100)), help="Gbases per GPU")
    ap.add_argument("--genome-mb", type=int, default=int(os.environ.get("NTSM_BENCH_GENOME_MB", 3100)))
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and gz legs")
    ap.add_argument("--kernel-only", action="store_true", help="device-resident leg only (kernel variant sweeps)")
    ap.add_argument("--synthetic-sites", type=int, default=0, help="with --kernel-only: cfg5's synthetic panel of this many sites")
    ap.add_argument("--read-len", type=int, default=READ_LEN, help="with --kernel-only: fixed read length (cfg3-like long reads)")
    ap.add_argument("--err", type=float, default=0.01, help="with --kernel-only: substitution rate")
    ap.add_argument("--k", type=int, default=19, help="with --kernel-only: k-mer size (the panel is re-cut at that k)")
    ap.add_argument("--opt", action="append", default=[], help="name=value for ntsm_ctx_set_option (kernel sweeps)")
    args = ap.parse_args()
    protect_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.k != 19 and not args.synthetic_sites:
        ap.error("--k other than 19 needs --synthetic-sites N (the human panel file lists 19-mers)")
    if (args.synthetic_sites or args.read_len != READ_LEN or args.err != 0.01 or args.k != 19) and not args.kernel_only:
        ap.error("--synthetic-sites / --read-len / --err / --k are --kernel-only measurements; the bench line is BASELINE configs[1]")
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
