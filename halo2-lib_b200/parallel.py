"""Multi-GPU host logic for the sharded hot path (one process per GPU, torch.distributed for the plumbing).

MSM shards by contiguous point-scalar range: rank g owns [g*n/G, (g+1)*n/G) of both SRS base arrays (resident on its
GPU) and of every scalar column; the only exchange is an all-reduce of the G partial sums under EC addition, done
as an all-gather of 96-byte Jacobian points followed by G-1 local additions (h2b_g1_sum) because EC addition is not
an NCCL reduction op.  NTT does not shard (the butterfly network couples all elements): one polynomial per device,
round-robin.  Witness assignment shards per physical column onto the device that commits it."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, begin+count) of rank `rank`; n must be divisible by world (n = 2^k, world in {1,2,4,8})."""
    if n % world:
        raise ValueError("shard_range: n must be divisible by the world size")
    count = n // world
    return rank * count, count


def ntt_owner(poly_index: int, world: int) -> int:
    """one column polynomial per device, round-robin"""
    return poly_index % world


def all_gather_points(partials, group=None):
    """partials: int64 tensor [m, 12] (this rank's m partial commitments) -> [m, world, 12] (contiguous per point),
    ready for h2b_g1_sum(_dev) over axis 1.  Works with nccl (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    m = partials.shape[0]
    out = torch.empty((world * m,) + tuple(partials.shape[1:]), dtype=partials.dtype, device=partials.device)
    dist.all_gather_into_tensor(out, partials.contiguous(), group=group)  # concatenation along dim 0 (gloo and nccl)
    return out.view((world, m) + tuple(partials.shape[1:])).transpose(0, 1).contiguous()
