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
stats = {"msm_adhoc": 0, "msm_srs": 0, "ntt": 0, "assign": 0, "scan": 0, "opening": 0, "lookup_perm": 0, "quotient": 0, "srs": 0}
from halo2_lib_b200 import evaluation as ev
def rnd_fr(m):
    x = rng.integers(0, 1 << 62, size=(m, 4), dtype=np.int64).astype(np.uint64); x[:, 3] &= np.uint64((1 << 60) - 1); return x
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
    # ---- opening arithmetic, odd sizes around the tile boundaries
    n = int(rng.choice([1, 2, 2047, 2048, 2049, 4096, 4097])) if rng.random() < 0.3 else int(rng.integers(1, 200000))
    A, z = rnd_fr(n), rnd_fr(1)[0]
    if rng.random() < 0.1: z = mont([int(rng.choice([0, 1, R - 1]))], R)[0]
    assert np.array_equal(h.eval_polynomial(ctx, A, z), orc.eval_polynomial(A, z)), ("eval", n)
    assert np.array_equal(h.kate_division(ctx, A, z), orc.kate_division(A, z)), ("kate", n)
    m = int(rng.integers(1, 6)); nn = int(rng.integers(1, 5000))
    ps, sc = [rnd_fr(nn) for _ in range(m)], rnd_fr(m)
    assert np.array_equal(h.poly_lincomb(ctx, ps, sc), orc.poly_lincomb(ps, sc)), ("lincomb", m, nn)
    stats["opening"] += 1
    # ---- lookup permutation: random table (with duplicates), inputs drawn from it; sometimes one value outside
    k = int(rng.integers(3, 13)); bf = int(rng.integers(0, 6)); rows = 1 << k
    if bf + 1 < rows:
        u = rows - (bf + 1)
        wide = rng.random() < 0.4
        tab = rand_ints(rng, u, R) if wide else [int(x) for x in rng.integers(0, max(2, u // int(rng.integers(1, 5))), size=u)]
        inp = [tab[int(j)] for j in rng.integers(0, u, size=u)]
        bad = rng.random() < 0.15
        if bad: inp[int(rng.integers(0, u))] = (max(tab) + 1) % R if not wide else (tab[0] + 1) % R
        pad = rand_ints(rng, bf + 1, R)
        Am, Tm = mont(inp + pad, R), mont(tab + pad, R)
        rc, wa, wt = orc.permute_expression_pair(Am, Tm, k, bf)
        try:
            pa, pt = h.permute_expression_pair(ctx, Am, Tm, k, bf)
            assert rc == 0 and np.array_equal(pa, wa) and np.array_equal(pt, wt), ("permute", k, bf, wide)
        except h.ConstraintSystemFailure:
            assert rc != 0, ("permute should not fail", k, bf, wide)
        stats["lookup_perm"] += 1
    # ---- quotient terms on random columns: gate program + permutation sets + lookup, folded into one accumulator
    k = int(rng.integers(2, 9)); ext_k = k + int(rng.integers(1, 4)); ne = 1 << ext_k
    ncols = int(rng.integers(1, 7)); chunk = int(rng.integers(1, 4)); nsets = (ncols + chunk - 1) // chunk
    cols = [rnd_fr(ne) for _ in range(2 * ncols + nsets + 6)]
    ch = rnd_fr(4); bfq = int(rng.integers(0, 6))
    zs, cs, ss = cols[:nsets], cols[nsets:nsets + ncols], cols[nsets + ncols:nsets + 2 * ncols]
    l0, ll, la, q, a, t = cols[nsets + 2 * ncols:]
    acc = rnd_fr(ne)
    g1 = ev.GraphEvaluator(); adv = [("advice", 0, r) for r in range(4)]
    gate = g1.add_gates([("product", ("fixed", 0, 0), ("sum", ("sum", adv[0], ("product", adv[1], adv[2])), ("negated", adv[3])))])
    kw = dict(beta=ch[0], gamma=ch[1], theta=ch[2], y=ch[3])
    b1 = ev.BoundGraph(g1, gate, fixed=[q], advice=[a], **kw)
    g2 = ev.GraphEvaluator(); lk = g2.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])
    b2 = ev.BoundGraph(g2, lk, fixed=[q, t], advice=[a], **kw)
    got = h.quotient_graph(ctx, b1, k, ext_k, acc); want = orc.quotient_graph(b1.struct, k, ext_k, acc)
    got = h.permutation_fold(ctx, zs, cs, ss, chunk, l0, ll, la, ch[0], ch[1], ch[3], bfq, k, ext_k, got)
    want = orc.permutation_fold(zs, cs, ss, chunk, l0, ll, la, ch[0], ch[1], ch[3], bfq, k, ext_k, want)
    got = h.lookup_fold(ctx, b2, zs[0], cs[0], ss[0], l0, ll, la, k, ext_k, got)
    want = orc.lookup_fold(b2.struct, zs[0], cs[0], ss[0], l0, ll, la, k, ext_k, want)
    got = h.divide_by_vanishing_poly(ctx, got, k, ext_k); want = orc.divide_by_vanishing_poly(want, k, ext_k)
    assert np.array_equal(got, want), ("quotient", k, ext_k, ncols, chunk, bfq)
    stats["quotient"] += 1
    # ---- G1 FFT on small random point sets (identities included)
    if rng.random() < 0.3:
        k = int(rng.integers(0, 7))
        pts = pool[rng.integers(0, len(pool), size=1 << k)].copy()
        if k > 1 and rng.random() < 0.5: pts[int(rng.integers(0, 1 << k))] = 0
        import ctypes as C
        from halo2_lib_b200._capi import lib
        out = np.empty_like(pts)
        ctx.check(lib.h2b_g_to_lagrange(ctx.h, C.c_void_p(pts.ctypes.data), k, C.c_void_p(out.ctypes.data)))
        assert np.array_equal(out, orc.g_to_lagrange(pts, k)), ("g_to_lagrange", k)
        stats["srs"] += 1
print("soak OK", stats, f"{time.time() - t0:.0f}s")
