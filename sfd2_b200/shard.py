"""Multi-GPU plumbing for the extract / match sweeps: images (or pairs) are independent units
(extract_localization.py:240 and hloc/match_features.py:90 are batch-1 loops), so rank r of W
owns items[r::W] and there is NO collective on the data path.  The only exchange is one
all_gather of the fixed-capacity (x, y, score) table + counts at the end, for reporting."""
import torch
import torch.distributed as dist

__all__ = ["shard_indices", "gather_table"]


def shard_indices(n_items: int, rank: int, world: int):
    """Indices of the items rank `rank` processes: rank, rank+W, rank+2W, ..."""
    return list(range(rank, n_items, world))


def gather_table(kpts: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, n_items: int, group=None):
    """kpts [n_local, K, 2], scores [n_local, K], counts [n_local] of this rank's shard (items
    rank::world) -> (table [n_items, K, 3], counts [n_items]) in the ORIGINAL item order, on every rank.
    Shards are padded to ceil(n_items / world) rows so the collective has a fixed shape."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    K = kpts.shape[1]
    per = (n_items + world - 1) // world
    table = torch.zeros(per, K, 3, dtype=torch.float32, device=kpts.device)
    cnt = torch.zeros(per, dtype=torch.int32, device=kpts.device)
    nl = kpts.shape[0]
    table[:nl, :, :2] = kpts
    table[:nl, :, 2] = scores
    cnt[:nl] = counts.to(torch.int32)
    if world == 1:
        return table[:n_items], cnt[:n_items]
    all_t = [torch.empty_like(table) for _ in range(world)]
    all_c = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(all_t, table, group=group)
    dist.all_gather(all_c, cnt, group=group)
    out_t = torch.zeros(n_items, K, 3, dtype=torch.float32, device=kpts.device)
    out_c = torch.zeros(n_items, dtype=torch.int32, device=kpts.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out_t[idx] = all_t[r][:len(idx)]
        out_c[idx] = all_c[r][:len(idx)]
    return out_t, out_c
