"""Batch-sharded data-parallel inference.

The path shards naturally: every sample is independent through every layer, so B is cut into G contiguous
shards (one process per GPU, weights replicated and packed per rank) and the ONLY collective is one all-gather
of the fp32 logits.  No data-path collective exists inside the layers.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous [lo, hi) rows of rank `rank`; the first batch % world ranks take one extra row."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world=None):
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(y_local, batch=None):
    """All-gather per-rank logits [B_r, C] into [B, C] on every rank (ragged shards are padded to the largest)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return y_local
    world = dist.get_world_size()
    if batch is None or batch % world == 0:
        out = torch.empty((world * y_local.shape[0],) + tuple(y_local.shape[1:]), dtype=y_local.dtype,
                          device=y_local.device)
        dist.all_gather_into_tensor(out, y_local.contiguous())
        return out
    rows = [shard_bounds(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in rows)
    pad = torch.zeros((mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[:y_local.shape[0]] = y_local
    out = torch.empty((world * mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(rows)], 0)


class ShardedInference(torch.nn.Module):
    """Wrap a replicated network: forward(x_full) runs this rank's shard and returns the gathered logits."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return gather_logits(self.net(shard_batch(x)), batch=x.shape[0])
