"""Randomised soak of the C ABI against the CPU oracle (not part of pytest): many sizes / seeds / scalar mixes.
Usage on the GPU box: python tools/soak.py [seconds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import halo2_lib_b200 as h
from oracle import oracle as orc, pyref
from util import mont, rand_ints, witness_like_ints, affine_to_limbs

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
ctx = h.Context(0)
R = pyref.R
g = affine_to_limbs([pyref.G1])[0]
rng = np.random.default_rng(int(time.time()))
norm = lambda x: ctx.g1_normalize(np.asarray(x, dtype=np.uint64).reshape(1, 12))[0]
stats = {"msm_adhoc": 0, "msm_srs": 0, "ntt": 0, "assign": 0, "scan": 0}
t0 = time.time()
pool = ctx.g1_fixed_base_mul(g, mont([int(x) for x in rng.integers(1, 1 << 40, size=1 << 13)], R))
while time.time() - t0 < budget:
    # ---- MSM, ad-hoc bases, odd sizes, adversarial scalar mixes
    n = int(rng.integers(1, 3000))
    idx = rng.integers(0, len(pool), size=n)
    B = pool[idx].copy()
    if n > 3 and rng.random() < 0.5:
        B[rng.integers(0, n)] = 0                      # identity base
        B[rng.integers(0, n)] = B[rng.integers(0, n)]  # repeated base
    mix = rng.random()
    if mix < 0.3: sc = rand_ints(rng, n, R)
    elif mix < 0.6: sc = witness_like_ints(rng, n)
    elif mix < 0.8: sc = [int(rng.choice([0, 1, R - 1, 2, R - 2, 1 << 17, (1 << 17) - 1, 1 << 16]))] * n
    else: sc = [int(v) for v in rng.integers(0, 4, size=n)]
    S = mont(sc, R)
    assert np.array_equal(norm(h.best_multiexp(ctx, S, B)), orc.msm_pippenger(S, B, 4)), ("adhoc", n, mix)
    stats["msm_adhoc"] += 1
    # ---- MSM through SRS tables with shards
    k = int(rng.integers(1, 13)); N = 1 << k
    Bk = pool[rng.integers(0, len(pool), size=N)].copy()
    begin = int(rng.integers(0, N)); count = int(rng.integers(1, N - begin + 1))
    p = h.ParamsKZG(ctx, k, g=Bk, g_lagrange=Bk[::-1].copy(), begin=begin, count=count)
    cols = [mont(witness_like_ints(rng, count) if rng.random() < 0.5 else rand_ints(rng, count, R), R) for _ in range(int(rng.integers(1, 6)))]
    bs = [int(rng.integers(0, 2)) for _ in cols]
    outs = p.commit_batch(bs, cols)
    for b, c, o in zip(bs, cols, outs):
        base = Bk if b == 0 else Bk[::-1].copy()
        assert np.array_equal(norm(o), orc.msm_pippenger(c, base[begin:begin + count], 4)), ("srs", k, begin, count)
    p.close(); stats["msm_srs"] += 1
    # ---- NTT family
    k = int(rng.integers(0, 17)); N = 1 << k
    A = rng.integers(0, 1 << 62, size=(N, 4), dtype=np.int64).astype(np.uint64); A[:, 3] &= np.uint64((1 << 60) - 1)
    j = int(rng.integers(3, 6))
    dom = h.EvaluationDomain(ctx, j, k)
    co = dom.lagrange_to_coeff(A)
    assert np.array_equal(co, orc.lagrange_to_coeff(A, k, 4)), ("intt", k)
    ext = dom.coeff_to_extended(co)
    assert np.array_equal(ext, orc.coeff_to_extended(co, dom.extended_k, 4)), ("coset", k, j)
    assert np.array_equal(dom.extended_to_coeff(ext)[:N], co), ("coset_inv", k, j)
    stats["ntt"] += 1
    # ---- assignment with random break points
    k = int(rng.integers(3, 12)); rows = 1 << k; ncols = int(rng.integers(1, 7))
    N = int(rng.integers(0, ncols * (rows - 2)))
    V = rng.integers(0, 1 << 62, size=(N, 4), dtype=np.int64).astype(np.uint64)
    bps = sorted(int(x) for x in rng.integers(0, rows, size=int(rng.integers(0, ncols))))
    rc, want = orc.assign_witnesses(V, np.array(bps, dtype=np.uint64), k, ncols)
    try:
        got = h.assign_witnesses(ctx, [V], bps, k, ncols)
        assert rc == 0 and np.array_equal(got, want), ("assign", k, ncols, N, bps)
    except h.LayoutError:
        assert rc != 0, ("assign should not fail", k, ncols, N, bps)
    stats["assign"] += 1
    # ---- batch inversion / grand product
    n = int(rng.integers(1, 50000))
    A = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64); A[:, 3] &= np.uint64((1 << 60) - 1)
    A[rng.random(n) < 0.1] = 0
    assert np.array_equal(ctx.batch_invert(A), orc.batch_invert(A)), ("binv", n)
    st = mont([int(rng.integers(1, 1 << 60))], R)[0]
    assert np.array_equal(ctx.grand_product(A, st), orc.grand_product(A, st)), ("gp", n)
    stats["scan"] += 1
print("soak OK", stats, f"{time.time() - t0:.0f}s")
