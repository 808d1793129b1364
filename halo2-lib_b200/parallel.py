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


def ntt_owners_balanced(costs, world: int):
    """owner of every transform of a proof when their costs differ (an iNTT over 2^k against a coset NTT over 2^(k+2)):
    longest-processing-time-first — transforms in decreasing cost, each to the least loaded device; deterministic, the
    same on every rank.  NTT does not shard (one polynomial per device): this only deals whole transforms."""
    load = [0.0] * world
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


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


def connect_peers(ctx, group=None):
    """Creates this rank's NVLink mailbox, exchanges the CUDA IPC handles of all ranks through torch.distributed and
    connects them (h2b_peer_create / h2b_peer_connect).  After this `allreduce_points` needs no NCCL call."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._capi import lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = (C.c_uint8 * 64)()
    ctx.check(lib.h2b_peer_create(ctx.h, rank, world, mine))
    dev = torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    out = torch.empty(world * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    blob = bytes(out.cpu().tolist())
    ctx.check(lib.h2b_peer_connect(ctx.h, C.c_char_p(blob)))
    dist.barrier(group)


def allreduce_points(ctx, d_points_ptr: int, m: int):
    """in place: m Jacobian partials (device, m x 12 limbs) -> the m sums over all ranks; one fused kernel"""
    import ctypes as C
    from ._capi import lib

    ctx.check(lib.h2b_g1_allreduce_dev(ctx.h, C.c_void_p(d_points_ptr), m))
