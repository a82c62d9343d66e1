"""Multi-GPU plumbing (one process per GPU, torch.distributed): read sharding, tie-break ids, index broadcast.

The data path has no collective: reads are split into contiguous equal shards by global read index,
the index image is broadcast once, every rank aligns its shard.  These helpers are backend agnostic (NCCL on the GPU
box, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank):
    """Contiguous shard [b, e) of rank; the first n_total % world ranks get one extra read.
    The same rule as b200_shard_bounds in the C ABI (seqlib_b200/csrc/dist.cu), which b200_reads_scatter and bench.py use;
    tests/test_dist_cpu.py checks the two against each other."""
    q, r = divmod(n_total, world)
    b = rank * q + min(rank, r)
    return b, b + q + (1 if rank < r else 0)


def read_ids(b, e, stride=7919, base=13):
    """Deterministic per-read tie-break ids (stand-in for the caller's lrand48() draws), a function of the GLOBAL read index."""
    return np.arange(b, e, dtype=np.int64) * stride + base


def scatter_fixed_len_reads(seqs, n_total, read_len, device, src=0):
    """Rank `src` holds all reads (uint8, n_total*read_len); every rank receives its contiguous shard as a uint8 tensor."""
    world, rank = dist.get_world_size(), dist.get_rank()
    b, e = shard_bounds(n_total, world, rank)
    mine = torch.empty((e - b) * read_len, dtype=torch.uint8, device=device)
    if rank == src:
        full = torch.as_tensor(seqs)
        parts = []
        for r in range(world):
            rb, re_ = shard_bounds(n_total, world, r)
            parts.append(full[rb * read_len:re_ * read_len].to(device).contiguous())
    # dist.scatter needs equal sizes; shards may differ by one read, so send point-to-point
    if rank == src:
        for r in range(world):
            if r == src:
                mine.copy_(parts[r])
            else:
                dist.send(parts[r], dst=r)
    else:
        dist.recv(mine, src=src)
    return mine, b, e


def broadcast_bytes(buf, nbytes, device, src=0):
    """Broadcast a uint8 device buffer of known size (the FM-index image)."""
    t = buf if buf is not None else torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return t


def gather_counts(local_count, device):
    """All ranks learn every rank's result count (for placing shards in a global output)."""
    world = dist.get_world_size()
    t = torch.zeros(world, dtype=torch.int64, device=device)
    t[dist.get_rank()] = local_count
    dist.all_reduce(t)
    return t.cpu().numpy()
