"""world_size-2 gloo tests of the host-side multi-GPU plumbing (sharding, ids, scatter, broadcast); no CUDA."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, read_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seqlib_b200 import shard
    rng = np.random.default_rng(5)
    full = rng.integers(65, 85, size=n_total * read_len, dtype=np.uint8) if rank == 0 else None
    mine, b, e = shard.scatter_fixed_len_reads(full, n_total, read_len, torch.device("cpu"))
    blob = torch.arange(1000, dtype=torch.int64).to(torch.uint8) if rank == 0 else None
    got = shard.broadcast_bytes(blob, 1000, torch.device("cpu"))
    counts = shard.gather_counts(e - b, torch.device("cpu"))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), mine=mine.numpy(), b=b, e=e, ids=shard.read_ids(b, e), blob=got.numpy(), counts=counts)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_scatter_broadcast_gloo(tmp_path):
    world, n_total, read_len = 2, 1001, 150
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, read_len, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    full = rng.integers(65, 85, size=n_total * read_len, dtype=np.uint8)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert parts[0]["b"] == 0 and parts[0]["e"] == parts[1]["b"] and parts[1]["e"] == n_total
    assert np.array_equal(np.concatenate([p["mine"] for p in parts]), full)
    ids = np.concatenate([p["ids"] for p in parts])
    assert np.array_equal(ids, np.arange(n_total, dtype=np.int64) * 7919 + 13)
    for p in parts:
        assert np.array_equal(p["blob"], (np.arange(1000) % 256).astype(np.uint8))
        assert list(p["counts"]) == [501, 500]


def test_shard_bounds_cover():
    from seqlib_b200 import shard
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_c_abi_shard_bounds_match_python():
    """b200_shard_bounds (the rule b200_reads_scatter and bench.py shard by) == seqlib_b200.shard.shard_bounds; no CUDA needed."""
    from seqlib_b200 import shard, capi
    for n in (0, 1, 7, 1000, 1001, 10_000_000, 100_000_003):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                assert capi.shard_bounds(n, w, r) == shard.shard_bounds(n, w, r)
