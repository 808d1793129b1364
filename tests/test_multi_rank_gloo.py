"""world_size-2 `gloo` test of the N>1 path on CPU: point-range sharding of an MSM + all-gather of the partial sums +
local additions equals the full MSM; NTT polynomials are dealt round-robin.  The partial MSMs are computed by the
CPU oracle here (there is no GPU in this container) — what is under test is the host-side sharding/exchange logic
of halo2-lib_b200/parallel.py that bench.py uses with NCCL."""
import os
import socket
import sys
import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import halo2_lib_b200 as h
    from oracle import oracle as orc, pyref
    from util import mont, rand_ints, witness_like_ints, affine_to_limbs
    n = 256
    rng = np.random.default_rng(123)  # same inputs on every rank
    g = affine_to_limbs([pyref.G1])[0]
    bases = orc.g1_fixed_base_mul(mont([3 + 5 * i for i in range(n)], pyref.R), g)
    cols = [mont(rand_ints(rng, n, pyref.R), pyref.R), mont(witness_like_ints(rng, n), pyref.R)]
    begin, count = h.shard_range(n, rank, world)
    partials = np.stack([orc.msm_pippenger(c[begin:begin + count], bases[begin:begin + count], 2) for c in cols])
    gathered = h.all_gather_points(torch.from_numpy(partials.view(np.int64)))
    assert gathered.shape == (2, world, 12)
    got = []
    for j in range(2):
        acc = gathered[j, 0].numpy().view(np.uint64)
        for r in range(1, world):
            acc = orc.g1_add(acc, gathered[j, r].numpy().view(np.uint64))
        got.append(orc.g1_normalize(acc))
    want = [orc.msm_pippenger(c, bases, 2) for c in cols]
    ok = all(np.array_equal(a, b) for a, b in zip(got, want))
    owners = [h.ntt_owner(i, world) for i in range(5)]
    ok = ok and owners == [0, 1, 0, 1, 0]
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_allgather_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_range():
    sys.path.insert(0, ROOT)
    import halo2_lib_b200 as h
    assert [h.shard_range(1 << 19, r, 8) for r in (0, 7)] == [(0, 65536), (458752, 65536)]
    import pytest
    with pytest.raises(ValueError):
        h.shard_range(10, 0, 3)


def test_transforms_are_dealt_by_cost():
    """ntt_owners_balanced: whole transforms, longest first to the least loaded device; deterministic on every rank"""
    import halo2_lib_b200 as h
    costs = [1.0] * 5 + [4.0] * 6  # the ECDSA schedule: 5 iNTT(2^19), 5 coset NTT(2^21) + extended_to_coeff(2^21)
    for world in (1, 2, 4, 8):
        own = h.ntt_owners_balanced(costs, world)
        assert own == h.ntt_owners_balanced(costs, world) and len(own) == len(costs) and set(own) <= set(range(world))
        load = [sum(c for c, o in zip(costs, own) if o == r) for r in range(world)]
        assert max(load) <= sum(costs) / world + max(costs)  # LPT bound
        if world == 8:
            assert max(load) == 5.0 or max(load) == 4.0  # never two large transforms on one device
            assert sum(1 for r in range(8) if any(c == 4.0 and o == r for c, o in zip(costs, own))) == 6
