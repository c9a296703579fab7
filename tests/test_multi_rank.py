"""N>1 path.  CPU (gloo, world_size 2): the sharding helpers, the unique-id rendezvous and the
combine ORDER (sum the k-mer counts across ranks first, per-site max afterwards) with the oracle
standing in for each rank's private counts.  GPU (nccl, needs >= 2 devices): the real thing
through ntsm_comm_init / ntsm_allreduce."""
import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

from ntsm_b200 import dist as ndist

SITES = os.path.join(GOLDEN, "shared", "sites300.fa")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _reads(seed, n):
    rng = random.Random(seed)
    wins = []
    for line in open(SITES):
        if not line.startswith(">"):
            wins.extend(line.strip().split("N"))
    return [("".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 50))) + rng.choice(wins) +
             "".join(rng.choice("ACGTN") for _ in range(rng.randrange(0, 50)))).encode() for _ in range(n)]


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            cuts = [ndist.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    files = ["f%d" % i for i in range(11)]
    assert sorted(sum((ndist.shard_files(files, r, 4) for r in range(4)), [])) == sorted(files)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load(os.path.join(ROOT, "oracle", "_build", "libntsm_oracle.so"))
    reads = _reads(123, 2000)
    lo, hi = ndist.shard_bounds(len(reads), rank, world)
    fp = orc.fingerprint(SITES, 19)
    for r in reads[lo:hi]:
        fp.insert(r)
    _, off, cnt = fp.lists()
    # private tallies -> one integer sum over ranks (the library does this with ncclAllReduce)
    c = torch.from_numpy(cnt.astype(np.int64)); dist.all_reduce(c)
    t = torch.tensor([fp.total_kmers, fp.total_counts, fp.total_bases], dtype=torch.int64); dist.all_reduce(t)
    # per-site max AFTER the sum
    c = c.numpy()
    S = fp.n_sites
    mx = np.array([[c[off[2 * i]:off[2 * i + 1]].max(initial=0), c[off[2 * i + 1]:off[2 * i + 2]].max(initial=0)] for i in range(S)])
    # the wrong order (max per rank, then sum) for contrast
    mr, mv = fp.rows()[:2]
    w = torch.from_numpy(np.stack([mr, mv], 1).astype(np.int64)); dist.all_reduce(w)
    # rendezvous of an opaque 128-byte id, as attach_comm does
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank == 0:
        q.put((c, t.numpy(), mx, w.numpy(), box[0]))
    else:
        assert box[0] == bytes(range(128))
    dist.destroy_process_group()


def test_two_rank_combine_matches_single_rank(oracle):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    c, t, mx, wrong, uid = q.get(timeout=120)
    for p in procs:
        p.join(60); assert p.exitcode == 0
    whole = oracle.fingerprint(SITES, 19)
    for r in _reads(123, 2000):
        whole.insert(r)
    _, off, cnt = whole.lists()
    assert np.array_equal(c, cnt.astype(np.int64))
    assert tuple(t) == (whole.total_kmers, whole.total_counts, whole.total_bases)
    mr, mv = whole.rows()[:2]
    assert np.array_equal(mx[:, 0], mr) and np.array_equal(mx[:, 1], mv)
    assert not np.array_equal(wrong[:, 0], mr)      # sum of per-rank maxima is NOT the answer
    assert uid == bytes(range(128))


# ------------------------------------------------------------------------------- GPU, NCCL
def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import ntsm_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # rendezvous only; the reduce is the library's NCCL
    reads = _reads(77, 6000)
    lo, hi = ndist.shard_bounds(len(reads), rank, world)
    fp = ntsm_b200.FingerPrint(SITES, device=rank, batch_bases=1 << 14)
    ndist.attach_comm(fp)
    for r in reads[lo:hi]:
        fp.insertCount(r)
    text = fp.counts_text()          # finalize(): drain -> all-reduce -> per-site reduce
    summ = fp.printInfoSummary()
    cnt = fp.kmer_counts()
    q.put((rank, text, summ, cnt))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_nccl_allreduce_matches_oracle(oracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)])
    for p in procs:
        p.join(120); assert p.exitcode == 0
    whole = oracle.fingerprint(SITES, 19)
    for r in _reads(77, 6000):
        whole.insert(r)
    for rank, text, summ, cnt in outs:                    # every rank ends with the global result
        assert text == whole.counts_text()
        assert summ == whole.summary()
        assert np.array_equal(cnt, whole.lists()[2])


@pytest.mark.gpu
@pytest.mark.parametrize("nccl_debug", ["VERSION", "WARN", None])
def test_cli_all_gpus_of_one_process_equals_one_gpu(nccl_debug):
    """ntsmCount --gpus N (one process, one NCCL communicator over its GPUs) must print the very same
    counts file as --gpus 1 -- including when the host exports NCCL_DEBUG=VERSION, which makes NCCL
    print its banner to stdout unless it is kept away from the counts file."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "ntsm_b200", "bin", "ntsmCount")
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i in range(3):
            p = os.path.join(tmp, "r%d.fa" % i)
            with open(p, "w") as fh:
                for j, r in enumerate(_reads(500 + i, 3000)):
                    fh.write(">r%d\n%s\n" % (j, r.decode()))
            paths.append(p)
        env = {k: v for k, v in os.environ.items() if k != "NCCL_DEBUG"}
        if nccl_debug:
            env["NCCL_DEBUG"] = nccl_debug
        outs = []
        for g in ("1", str(torch.cuda.device_count())):
            p = subprocess.run([exe, "--gpus", g, "--batch-bases", "100000", "-t", "3", "-s", SITES] + paths, capture_output=True, env=env)
            assert p.returncode == 0, p.stderr.decode()
            outs.append(p.stdout)
        assert outs[0].startswith(b"#@TK\t")
        assert outs[0] == outs[1]
