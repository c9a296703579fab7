"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for the rendezvous.

The data path has no collective while counting: reads are sharded, every rank keeps private
uint32 counts and uint64 tallies.  At the end ONE NCCL all-reduce (inside libntsm_b200.so, on the
ctx's stream) sums the k-mer counts and one sums the three tallies; the per-site max/sum kernel
runs AFTER that, because max-of-sums != sum-of-maxes (the trap ntsmEval --merge falls into,
src/CompareCounts.hpp:648-657).
"""
import os


def shard_bounds(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for this rank; shards differ in size by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_files(paths, rank, world):
    """File-level sharding for the file-driven path: rank r takes files r, r+world, ..."""
    return list(paths[rank::world])


def attach_comm(fp, group=None):
    """Give every rank's FingerPrint the same NCCL communicator: rank 0 draws the unique id,
    torch.distributed broadcasts the 128 bytes, each rank joins with its own rank number."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    box = [type(fp).nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    fp.comm_init(box[0], rank, world)


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
