"""N>1 path on CPU: world_size-2 gloo run of the shard/gather plumbing (the extract itself is
replaced by a deterministic stand-in; the GPU kernels are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfd2_b200.shard import shard_indices, gather_table

K = 16


def _fake_extract(item: int):
    g = torch.Generator().manual_seed(item)
    n = 3 + item % (K - 3)
    kp = torch.zeros(K, 2)
    sc = torch.zeros(K)
    kp[:n] = torch.randint(0, 1000, (n, 2), generator=g).float()
    sc[:n] = torch.sort(torch.rand(n, generator=g), descending=True).values
    return kp, sc, n


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    res = [_fake_extract(i) for i in mine]
    kp = torch.stack([r[0] for r in res]) if res else torch.zeros(0, K, 2)
    sc = torch.stack([r[1] for r in res]) if res else torch.zeros(0, K)
    cnt = torch.tensor([r[2] for r in res], dtype=torch.int32)
    table, counts = gather_table(kp, sc, cnt, n_items)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), table=table.numpy(), counts=counts.numpy())
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_indices_partition():
    for n, w in [(10, 2), (7, 4), (3, 8), (0, 2)]:
        seen = sorted(i for r in range(w) for i in shard_indices(n, r, w))
        assert seen == list(range(n))


def test_two_rank_gather_equals_single_rank(tmp_path):
    n_items, world = 7, 2       # ragged: rank 0 owns 4 items, rank 1 owns 3
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    ref_t = torch.zeros(n_items, K, 3)
    ref_c = torch.zeros(n_items, dtype=torch.int32)
    for i in range(n_items):
        kp, sc, n = _fake_extract(i)
        ref_t[i, :, :2], ref_t[i, :, 2], ref_c[i] = kp, sc, n
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["table"], ref_t.numpy())
        assert np.array_equal(z["counts"], ref_c.numpy())
