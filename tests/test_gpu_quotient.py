"""GPU parity tests (-m gpu) of the "next" rows: the general quotient evaluation
(GraphEvaluator programs, permutation and lookup argument terms) and the opening arithmetic (eval_polynomial,
kate_division, linear combinations).  Every call goes through the C ABI and is compared bit-exactly with the CPU oracle;
larger sizes are covered by identities (a = q (X - z) + a(z); satisfied arguments vanish on the domain)."""
import ctypes as C
import numpy as np
import pytest
from oracle import pyref, oracle as orc
from util import mont, unmont, rand_ints
import quotient_cases as qc

pytestmark = pytest.mark.gpu
R = pyref.R


@pytest.fixture(scope="module")
def h2b():
    import halo2_lib_b200 as h
    return h


@pytest.fixture(scope="module")
def ctx(h2b):
    c = h2b.Context(0)
    yield c
    c.close()


def rnd_fr(rng, n):
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    return x


# ------------------------------------------------------------------ opening arithmetic
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 2047, 2048, 2049, 4097, 70001, (1 << 19) + 3])
def test_eval_polynomial_and_kate_division(ctx, h2b, n):
    rng = np.random.default_rng(1200 + n % 101)
    a, z = rnd_fr(rng, n), rnd_fr(rng, 1)[0]
    assert np.array_equal(h2b.eval_polynomial(ctx, a, z), orc.eval_polynomial(a, z))
    assert np.array_equal(h2b.kate_division(ctx, a, z), orc.kate_division(a, z))


def test_eval_polynomial_edges(ctx, h2b):
    one, zero = mont([1], R)[0], mont([0], R)[0]
    a = mont([5, 7, 11], R)
    assert unmont(h2b.eval_polynomial(ctx, a, zero), R)[0] == 5
    assert unmont(h2b.eval_polynomial(ctx, a, one), R)[0] == 23
    assert unmont(h2b.eval_polynomial(ctx, np.zeros((0, 4), dtype=np.uint64), one), R)[0] == 0
    assert unmont(h2b.kate_division(ctx, a, zero), R) == [7, 11]  # division by X
    with pytest.raises(h2b.H2BError):
        h2b.kate_division(ctx, np.zeros((0, 4), dtype=np.uint64), one)


def test_kate_division_identity_full_size(ctx, h2b):
    """2^22 coefficients: a(x) == q(x) * (x - z) + a(z) at a random x, all evaluations on the GPU"""
    rng = np.random.default_rng(1300)
    n = 1 << 22
    a, z, x = rnd_fr(rng, n), rnd_fr(rng, 1)[0], rnd_fr(rng, 1)[0]
    q = h2b.kate_division(ctx, a, z)
    ax, az, qx = (h2b.eval_polynomial(ctx, p, pt) for p, pt in ((a, x), (a, z), (q, x)))
    rhs = orc.f_add(orc.FR, orc.f_mul(orc.FR, qx, orc.f_sub(orc.FR, x, z)), az)
    assert np.array_equal(ax.reshape(1, 4), rhs)
    # and against the oracle on a slice-sized problem embedded at the top (carry path across many tiles)
    assert np.array_equal(q[-5000:], orc.kate_division(a[-5001:], z))


@pytest.mark.parametrize("m,n", [(1, 5), (3, 1000), (32, 4099)])
def test_poly_lincomb(ctx, h2b, m, n):
    rng = np.random.default_rng(1400 + m)
    polys, sc = [rnd_fr(rng, n) for _ in range(m)], rnd_fr(rng, m)
    assert np.array_equal(h2b.poly_lincomb(ctx, polys, sc), orc.poly_lincomb(polys, sc))
    with pytest.raises(h2b.H2BError):
        h2b.poly_lincomb(ctx, [polys[0]] * 33, rnd_fr(rng, 33))


def test_opening_arithmetic_device_pointers(ctx, h2b):
    import torch
    from halo2_lib_b200._capi import lib
    rng = np.random.default_rng(1500)
    n = 10000
    a, b, z = rnd_fr(rng, n), rnd_fr(rng, n), rnd_fr(rng, 1)[0]
    sc = rnd_fr(rng, 2)
    da, db = (torch.from_numpy(x.view(np.int64)).cuda() for x in (a, b))
    dq = torch.zeros((n - 1, 4), dtype=torch.int64, device="cuda")
    vp = C.c_void_p
    ctx.check(lib.h2b_kate_division_dev(ctx.h, vp(da.data_ptr()), n, vp(z.ctypes.data), vp(dq.data_ptr())))
    out = np.empty(4, dtype=np.uint64)
    ctx.check(lib.h2b_eval_polynomial_dev(ctx.h, vp(da.data_ptr()), n, vp(z.ctypes.data), vp(out.ctypes.data)))
    ctx.synchronize()
    assert np.array_equal(dq.cpu().numpy().view(np.uint64), orc.kate_division(a, z))
    assert np.array_equal(out, orc.eval_polynomial(a, z))
    ptrs = (C.c_void_p * 2)(da.data_ptr(), db.data_ptr())
    ctx.check(lib.h2b_poly_lincomb_dev(ctx.h, ptrs, vp(sc.ctypes.data), 2, n, vp(da.data_ptr())))  # in place over polys[0]
    ctx.synchronize()
    assert np.array_equal(da.cpu().numpy().view(np.uint64), orc.poly_lincomb([a, b], sc))
    assert lib.h2b_kate_division_dev(ctx.h, vp(da.data_ptr()), n, vp(z.ctypes.data), vp(da.data_ptr())) != 0  # aliasing refused


# ------------------------------------------------------------------ GraphEvaluator programs
def test_graph_flex_gate(ctx, h2b):
    from halo2_lib_b200 import evaluation as ev
    k, ext_k = 9, 11
    n = 1 << ext_k
    rng = np.random.default_rng(1600)
    q, a, acc, y = rnd_fr(rng, n), rnd_fr(rng, n), rnd_fr(rng, n), rnd_fr(rng, 1)[0]
    g = ev.GraphEvaluator()
    adv = [("advice", 0, r) for r in range(4)]
    res = g.add_gates([("product", ("fixed", 0, 0), ("sum", ("sum", adv[0], ("product", adv[1], adv[2])), ("negated", adv[3])))])
    bound = ev.BoundGraph(g, res, fixed=[q], advice=[a], y=y)
    got = h2b.quotient_graph(ctx, bound, k, ext_k, acc)
    assert np.array_equal(got, orc.quotient_graph(bound.struct, k, ext_k, acc))
    assert np.array_equal(got, ctx.flex_gate_fold(q, a, y, k, ext_k, acc))  # the dedicated kernel computes the same term


def test_graph_every_opcode_and_source(ctx, h2b):
    from halo2_lib_b200 import evaluation as ev
    k, ext_k = 7, 10
    n = 1 << ext_k
    rng = np.random.default_rng(1700)
    cols = [rnd_fr(rng, n) for _ in range(4)]
    g = ev.GraphEvaluator()
    c7 = g.add_constant(mont([7], R)[0])
    s0 = g.add_calculation((ev.STORE, ev.src(ev.FIXED, 0, g.add_rotation(-1))))
    s1 = g.add_calculation((ev.ADD, s0, ev.src(ev.ADVICE, 1, g.add_rotation(2))))
    s2 = g.add_calculation((ev.SUB, s1, ev.src(ev.INSTANCE, 0, g.add_rotation(0))))
    s3 = g.add_calculation((ev.MUL, s2, ev.src(ev.CHALLENGE, 1)))
    s4 = g.add_calculation((ev.SQUARE, s3))
    s5 = g.add_calculation((ev.DOUBLE, s4))
    s6 = g.add_calculation((ev.NEGATE, s5))
    parts = (s6, c7, ev.src(ev.BETA), ev.src(ev.GAMMA), ev.src(ev.Y), ev.src(ev.ADVICE, 0, g.add_rotation(-(1 << k) - 3)))
    s7 = g.add_calculation((ev.HORNER, ev.src(ev.PREVIOUS), ev.src(ev.THETA), parts))
    ch = rnd_fr(rng, 4)
    bound = ev.BoundGraph(g, s7, fixed=[cols[0]], advice=[cols[1], cols[2]], instance=[cols[3]], challenges=rnd_fr(rng, 2), beta=ch[0],
                          gamma=ch[1], theta=ch[2], y=ch[3])
    prev = rnd_fr(rng, n)
    assert np.array_equal(h2b.quotient_graph(ctx, bound, k, ext_k, prev), orc.quotient_graph(bound.struct, k, ext_k, prev))


def test_graph_random_programs(ctx, h2b):
    """seeded random straight-line programs up to the calculation limit"""
    from halo2_lib_b200 import evaluation as ev
    k, ext_k = 6, 8
    n = 1 << ext_k
    rng = np.random.default_rng(1800)
    for trial in range(6):
        cols = [rnd_fr(rng, n) for _ in range(6)]
        g = ev.GraphEvaluator()
        rots = [g.add_rotation(int(r)) for r in rng.integers(-70, 70, size=5)]
        extra = [g.add_constant(rnd_fr(rng, 1)[0]) for _ in range(3)]

        def leaf():
            kind = int(rng.integers(0, 9))
            if kind < 3:
                return ev.src((ev.FIXED, ev.ADVICE, ev.INSTANCE)[kind], int(rng.integers(0, 2)), rots[int(rng.integers(0, len(rots)))])
            if kind == 3:
                return ev.src(ev.CHALLENGE, int(rng.integers(0, 3)))
            if kind == 4:
                return extra[int(rng.integers(0, 3))]
            return ev.src((ev.BETA, ev.GAMMA, ev.THETA, ev.Y, ev.PREVIOUS)[kind - 5])

        ncalc = (5, 17, 33, 48, 64, 64)[trial]
        last = None
        while len(g.calculations) < ncalc:
            def operand():
                t = len(g.calculations)
                return ev.src(ev.INTERMEDIATE, int(rng.integers(0, t))) if t and rng.random() < 0.6 else leaf()
            op = int(rng.integers(0, 8))
            if op == ev.HORNER:
                calc = (ev.HORNER, operand(), operand(), tuple(operand() for _ in range(int(rng.integers(0, 5)))))
            elif op <= ev.MUL:
                calc = (op, operand(), operand())
            else:
                calc = (op, operand())
            last = g.add_calculation(calc)
        ch = rnd_fr(rng, 4)
        bound = ev.BoundGraph(g, last, fixed=cols[0:2], advice=cols[2:4], instance=cols[4:6], challenges=rnd_fr(rng, 3), beta=ch[0],
                              gamma=ch[1], theta=ch[2], y=ch[3])
        prev = rnd_fr(rng, n)
        assert np.array_equal(h2b.quotient_graph(ctx, bound, k, ext_k, prev), orc.quotient_graph(bound.struct, k, ext_k, prev)), trial


def test_graph_malformed_programs_are_refused(ctx, h2b):
    from halo2_lib_b200 import evaluation as ev
    k, ext_k = 4, 6
    col = rnd_fr(np.random.default_rng(1), 1 << ext_k)
    vals = np.zeros((1 << ext_k, 4), dtype=np.uint64)

    def run(calcs, result, **kw):
        g = ev.GraphEvaluator()
        g.rotations = [0]
        g.calculations = list(calcs)
        return h2b.quotient_graph(ctx, ev.BoundGraph(g, result, **kw), k, ext_k, vals)

    run([(ev.STORE, ev.src(ev.ADVICE, 0, 0))], ev.src(ev.INTERMEDIATE, 0), advice=[col])  # well-formed
    for calcs, result, kw in [
        ([(ev.STORE, ev.src(ev.ADVICE, 1, 0))], ev.src(ev.INTERMEDIATE, 0), dict(advice=[col])),   # column out of range
        ([(ev.STORE, ev.src(ev.ADVICE, 0, 3))], ev.src(ev.INTERMEDIATE, 0), dict(advice=[col])),   # rotation slot out of range
        ([(ev.ADD, ev.src(ev.INTERMEDIATE, 0), ev.src(ev.Y))], ev.src(ev.INTERMEDIATE, 0), {}),    # uses itself
        ([(ev.STORE, ev.src(ev.CONSTANT, 9))], ev.src(ev.INTERMEDIATE, 0), {}),                    # constant out of range
        ([(ev.STORE, ev.src(ev.Y))], ev.src(ev.INTERMEDIATE, 1), {}),                              # result not computed
        ([(9, ev.src(ev.Y))], ev.src(ev.INTERMEDIATE, 0), {}),                                     # unknown opcode
        ([(ev.STORE, ev.src(13))], ev.src(ev.INTERMEDIATE, 0), {}),                                # unknown source kind
    ]:
        with pytest.raises(h2b.H2BError):
            run(calcs, result, **kw)


# ------------------------------------------------------------------ permutation / lookup argument terms
def gpu_to_ext(dom, col_lagrange):
    return dom.coeff_to_extended(dom.lagrange_to_coeff(mont(col_lagrange, R)))


def on_domain_values(ctx, h2b, values_ext, k, ext_k, rows):
    """interpolate the numerator from its 2^ext_k coset evaluations and evaluate it at omega^row (all on the GPU)"""
    dom = h2b.EvaluationDomain(ctx, 2, k)
    dom.extended_k = ext_k
    dom.quotient_poly_degree = 1 << (ext_k - k)
    coeffs = dom.extended_to_coeff(values_ext)
    w = pyref.omega_for(k)
    return [unmont(h2b.eval_polynomial(ctx, coeffs, mont([pow(w, r, R)], R)[0]), R)[0] for r in rows]


@pytest.mark.parametrize("n_cols,chunk_len,ext_bits", [(1, 2, 2), (3, 2, 2), (5, 3, 3), (7, 3, 3)])
def test_permutation_fold(ctx, h2b, n_cols, chunk_len, ext_bits):
    k, bf = 7, 5
    ext_k = k + ext_bits
    rng = np.random.default_rng(1900 + n_cols)
    beta, gamma, y = rand_ints(rng, 3, R)
    bm, gm, ym = (mont([v], R)[0] for v in (beta, gamma, y))
    l0, l_last, l_active, u = qc.lagrange_basis_columns(k, bf)
    dom = h2b.EvaluationDomain(ctx, 2, k)
    dom.extended_k = ext_k
    rows = [0, 1, 2, u - 1, u, u + 1, (1 << k) - 1] + [int(r) for r in rng.integers(0, 1 << k, size=8)]
    for broken in (False, True):
        cols, sigma, z_sets = qc.permutation_case(k, n_cols, chunk_len, bf, beta, gamma, seed=11, break_copy=broken)
        e = lambda c: gpu_to_ext(dom, c)
        zs, cs, ss, ls = [e(z) for z in z_sets], [e(c) for c in cols], [e(s) for s in sigma], [e(l0), e(l_last), e(l_active)]
        start = rnd_fr(rng, 1 << ext_k)
        got = h2b.permutation_fold(ctx, zs, cs, ss, chunk_len, *ls, bm, gm, ym, bf, k, ext_k, start)
        assert np.array_equal(got, orc.permutation_fold(zs, cs, ss, chunk_len, *ls, bm, gm, ym, bf, k, ext_k, start))
        zero = np.zeros((1 << ext_k, 4), dtype=np.uint64)
        num = h2b.permutation_fold(ctx, zs, cs, ss, chunk_len, *ls, bm, gm, ym, bf, k, ext_k, zero)
        vals = on_domain_values(ctx, h2b, num, k, ext_k, rows if not broken else range(1 << k))
        assert all(v == 0 for v in vals) != broken


def test_permutation_fold_no_sets_is_a_no_op(ctx, h2b):
    from halo2_lib_b200._capi import lib
    v = rnd_fr(np.random.default_rng(3), 64)
    w = v.copy()
    ch = mont([1, 2, 3], R)
    vp = C.c_void_p
    ctx.check(lib.h2b_permutation_fold(ctx.h, None, 0, None, None, 0, 2, None, None, None, vp(ch[0].ctypes.data), vp(ch[1].ctypes.data),
                                       vp(ch[2].ctypes.data), 5, 4, 6, vp(w.ctypes.data)))
    assert np.array_equal(v, w)


def test_lookup_fold(ctx, h2b):
    from halo2_lib_b200 import evaluation as ev
    k, ext_k, bf = 7, 10, 5  # the selector-gated lookup has degree 5: interpolate from 8n points
    rng = np.random.default_rng(2000)
    beta, gamma, theta, y = rand_ints(rng, 4, R)
    m1 = lambda v: mont([v], R)[0]
    l0, l_last, l_active, u = qc.lagrange_basis_columns(k, bf)
    dom = h2b.EvaluationDomain(ctx, 2, k)
    dom.extended_k = ext_k
    e = lambda c: gpu_to_ext(dom, c)
    for broken in (False, True):
        q, a, table, a_perm, s_perm, z = qc.lookup_case(k, bf, beta, gamma, seed=13, break_lookup=broken)
        g = ev.GraphEvaluator()
        res = g.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])  # range/mod.rs:131-140
        bound = ev.BoundGraph(g, res, fixed=[e(q), e(table)], advice=[e(a)], beta=m1(beta), gamma=m1(gamma), theta=m1(theta), y=m1(y))
        ze, ape, spe, ls = e(z), e(a_perm), e(s_perm), [e(l0), e(l_last), e(l_active)]
        start = rnd_fr(rng, 1 << ext_k)
        got = h2b.lookup_fold(ctx, bound, ze, ape, spe, *ls, k, ext_k, start)
        assert np.array_equal(got, orc.lookup_fold(bound.struct, ze, ape, spe, *ls, k, ext_k, start))
        num = h2b.lookup_fold(ctx, bound, ze, ape, spe, *ls, k, ext_k, np.zeros((1 << ext_k, 4), dtype=np.uint64))
        vals = on_domain_values(ctx, h2b, num, k, ext_k, range(1 << k))
        assert all(v == 0 for v in vals) != broken


def test_quotient_device_pointers_full_size(ctx, h2b):
    """k = 17, extended 2^19 rows, resident columns (`_dev` entry points): gate program + permutation + lookup terms folded
    into one accumulator, compared with the oracle"""
    import torch
    from halo2_lib_b200 import evaluation as ev
    from halo2_lib_b200._capi import lib
    k, ext_k, bf = 17, 19, 5
    n = 1 << ext_k
    rng = np.random.default_rng(2100)
    host = [rnd_fr(rng, n) for _ in range(12)]
    dev = [torch.from_numpy(x.view(np.int64)).cuda() for x in host]
    ch = rnd_fr(rng, 4)
    g = ev.GraphEvaluator()
    adv = [("advice", 0, r) for r in range(4)]
    gate = g.add_gates([("product", ("fixed", 0, 0), ("sum", ("sum", adv[0], ("product", adv[1], adv[2])), ("negated", adv[3])))])
    g2 = ev.GraphEvaluator()
    lk = g2.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])
    kw = dict(beta=ch[0], gamma=ch[1], theta=ch[2], y=ch[3])
    vp = C.c_void_p
    acc_h = rnd_fr(rng, n)
    acc_d = torch.from_numpy(acc_h.view(np.int64)).cuda()
    # device side
    bd = ev.BoundGraph(g, gate, fixed=[dev[0].data_ptr()], advice=[dev[1].data_ptr()], **kw)
    ctx.check(lib.h2b_quotient_graph_dev(ctx.h, C.byref(bd.struct), k, ext_k, vp(acc_d.data_ptr())))
    tz = (C.c_void_p * 2)(dev[2].data_ptr(), dev[3].data_ptr())
    tc = (C.c_void_p * 3)(dev[1].data_ptr(), dev[4].data_ptr(), dev[5].data_ptr())
    ts = (C.c_void_p * 3)(dev[6].data_ptr(), dev[7].data_ptr(), dev[8].data_ptr())
    ctx.check(lib.h2b_permutation_fold_dev(ctx.h, tz, 2, tc, ts, 3, 2, vp(dev[9].data_ptr()), vp(dev[10].data_ptr()), vp(dev[11].data_ptr()),
                                           vp(ch[0].ctypes.data), vp(ch[1].ctypes.data), vp(ch[3].ctypes.data), bf, k, ext_k, vp(acc_d.data_ptr())))
    bl = ev.BoundGraph(g2, lk, fixed=[dev[0].data_ptr(), dev[5].data_ptr()], advice=[dev[1].data_ptr()], **kw)
    ctx.check(lib.h2b_lookup_fold_dev(ctx.h, C.byref(bl.struct), vp(dev[2].data_ptr()), vp(dev[4].data_ptr()), vp(dev[6].data_ptr()),
                                      vp(dev[9].data_ptr()), vp(dev[10].data_ptr()), vp(dev[11].data_ptr()), k, ext_k, vp(acc_d.data_ptr())))
    ctx.synchronize()
    # oracle side, same order
    bh = ev.BoundGraph(g, gate, fixed=[host[0]], advice=[host[1]], **kw)
    want = orc.quotient_graph(bh.struct, k, ext_k, acc_h)
    want = orc.permutation_fold([host[2], host[3]], [host[1], host[4], host[5]], [host[6], host[7], host[8]], 2, host[9], host[10], host[11],
                                ch[0], ch[1], ch[3], bf, k, ext_k, want)
    blh = ev.BoundGraph(g2, lk, fixed=[host[0], host[5]], advice=[host[1]], **kw)
    want = orc.lookup_fold(blh.struct, host[2], host[4], host[6], host[9], host[10], host[11], k, ext_k, want)
    assert np.array_equal(acc_d.cpu().numpy().view(np.uint64), want)


# ------------------------------------------------------------------ lookup argument: permute_expression_pair
@pytest.mark.parametrize("kind,k", [("range", 4), ("range", 7), ("dup_table", 9), ("wide", 10), ("all_same", 8), ("perm", 9), ("range", 16), ("wide", 15),
                                    ("wide", 18), ("dup_table", 17)])
def test_permute_expression_pair(ctx, h2b, kind, k):
    from test_oracle_quotient import lookup_columns
    bf = 5
    rng = np.random.default_rng(2200 + k)
    u = (1 << k) - (bf + 1)
    inputs, table = lookup_columns(rng, k, bf, kind)
    pad = rand_ints(rng, bf + 1, R)
    A, T = mont(inputs + pad, R), mont(table + pad, R)
    pa, pt = h2b.permute_expression_pair(ctx, A, T, k, bf)
    rc, wa, wt = orc.permute_expression_pair(A, T, k, bf)
    assert rc == 0
    assert np.array_equal(pa, wa) and np.array_equal(pt, wt)
    assert not pa[u:].any() and not pt[u:].any()  # blinding rows are the caller's


def test_permute_expression_pair_zcash_order_option(h2b):
    """option "lookup.leftover_order" = 1: left-over table values go to the repeated rows from the back (zcash halo2)"""
    from test_oracle_quotient import lookup_columns
    c = h2b.Context(0)
    try:
        c.set_option("lookup.leftover_order", 1)
        for kind, k in (("range", 8), ("dup_table", 9), ("wide", 10)):
            bf = 5
            rng = np.random.default_rng(2250 + k)
            inputs, table = lookup_columns(rng, k, bf, kind)
            pad = rand_ints(rng, bf + 1, R)
            A, T = mont(inputs + pad, R), mont(table + pad, R)
            pa, pt = h2b.permute_expression_pair(c, A, T, k, bf)
            rc, wa, wt = orc.permute_expression_pair(A, T, k, bf, zcash_order=True)
            assert rc == 0 and np.array_equal(pa, wa) and np.array_equal(pt, wt)
        c.set_option("lookup.leftover_order", 0)
        pa, pt = h2b.permute_expression_pair(c, A, T, k, bf)
        rc, wa, wt = orc.permute_expression_pair(A, T, k, bf)
        assert np.array_equal(pa, wa) and np.array_equal(pt, wt)
    finally:
        c.close()


def test_permute_expression_pair_async_keeps_the_verdict_on_the_device(ctx, h2b):
    """h2b_permute_expression_pair_async_dev: no host synchronisation, the verdict is a device word"""
    import torch
    from halo2_lib_b200._capi import lib
    from test_oracle_quotient import lookup_columns
    k, bf = 11, 5
    rng = np.random.default_rng(2290)
    u = (1 << k) - (bf + 1)
    inputs, table = lookup_columns(rng, k, bf, "range")
    pad = rand_ints(rng, bf + 1, R)
    A, T = mont(inputs + pad, R), mont(table + pad, R)
    rc, wa, wt = orc.permute_expression_pair(A, T, k, bf)
    assert rc == 0
    vp = C.c_void_p
    dA, dT = torch.from_numpy(A.view(np.int64)).cuda(), torch.from_numpy(T.view(np.int64)).cuda()
    dpa, dpt = torch.zeros_like(dA), torch.zeros_like(dT)
    st = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    ctx.check(lib.h2b_permute_expression_pair_async_dev(ctx.h, vp(dA.data_ptr()), vp(dT.data_ptr()), k, bf, vp(dpa.data_ptr()), vp(dpt.data_ptr()), vp(st.data_ptr())))
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    assert np.array_equal(dpa.cpu().numpy().view(np.uint64)[:u], wa[:u]) and np.array_equal(dpt.cpu().numpy().view(np.uint64)[:u], wt[:u])
    A2 = A.copy()
    A2[3] = mont([(1 << 200) + 12345], R)[0]  # not in the table
    dA2 = torch.from_numpy(A2.view(np.int64)).cuda()
    ctx.check(lib.h2b_permute_expression_pair_async_dev(ctx.h, vp(dA2.data_ptr()), vp(dT.data_ptr()), k, bf, vp(dpa.data_ptr()), vp(dpt.data_ptr()), vp(st.data_ptr())))
    torch.cuda.synchronize()
    assert int(st.item()) & 1


def test_permute_expression_pair_missing_value_and_bad_arguments(ctx, h2b):
    k, bf = 8, 5
    u = (1 << k) - (bf + 1)
    table = list(range(u))
    inputs = [3] * (u - 1) + [u + 9]
    with pytest.raises(h2b.ConstraintSystemFailure):
        h2b.permute_expression_pair(ctx, mont(inputs + [0] * (bf + 1), R), mont(table + [0] * (bf + 1), R), k, bf)
    from halo2_lib_b200._capi import lib
    a = mont(table + [0] * (bf + 1), R)
    assert lib.h2b_permute_expression_pair(ctx.h, C.c_void_p(a.ctypes.data), C.c_void_p(a.ctypes.data), 2, 5, C.c_void_p(a.ctypes.data),
                                           C.c_void_p(a.ctypes.data)) == -1  # 2^2 rows, 6 of them blinding: no usable rows
    # the context stays usable after the failures
    pa, pt = h2b.permute_expression_pair(ctx, a, a, k, bf)
    assert unmont(pa[:u], R) == table and unmont(pt[:u], R) == table


def test_lookup_argument_end_to_end(ctx, h2b):
    """the GPU pipeline a prover runs for one lookup: permute_expression_pair -> denominators batch-inverted -> grand
    product -> the five quotient terms; a satisfied lookup must vanish on the 2^k domain."""
    from halo2_lib_b200 import evaluation as ev
    k, ext_k, bf = 7, 10, 5
    n = 1 << k
    rng = np.random.default_rng(2300)
    beta, gamma, theta, y = rand_ints(rng, 4, R)
    m1 = lambda v: mont([v], R)[0]
    l0, l_last, l_active, u = qc.lagrange_basis_columns(k, bf)
    q = [int(rng.integers(0, 2)) for _ in range(u)] + [0] * (n - u)
    a = [int(rng.integers(0, u)) for _ in range(u)] + rand_ints(rng, n - u, R)
    table = list(range(u)) + rand_ints(rng, n - u, R)
    inputs = [qq * aa % R for qq, aa in zip(q, a)]
    pa, pt = h2b.permute_expression_pair(ctx, mont(inputs, R), mont(table, R), k, bf)
    blind = rnd_fr(rng, 2 * (bf + 1))
    pa[u:], pt[u:] = blind[:bf + 1], blind[bf + 1:]
    # z[i+1] = z[i] * (A_i + beta)(S_i + gamma) / ((A'_i + beta)(S'_i + gamma)); all field work on the GPU
    B, G = np.tile(m1(beta), (n, 1)), np.tile(m1(gamma), (n, 1))
    add = lambda x, yv: ctx.field_op(1, 1, x, yv)
    mul = lambda x, yv: ctx.field_op(1, 0, x, yv)
    num = mul(add(mont(inputs, R), B), add(mont(table, R), G))
    den = ctx.batch_invert(mul(add(pa, B), add(pt, G)))
    z = ctx.grand_product(mul(num, den), m1(1))
    assert unmont(z[u:u + 1], R)[0] == 1  # the product over the usable rows closes
    z[u + 1:] = rnd_fr(rng, n - u - 1)
    dom = h2b.EvaluationDomain(ctx, 2, k)
    dom.extended_k = ext_k
    ext = lambda col: dom.coeff_to_extended(dom.lagrange_to_coeff(col))
    g = ev.GraphEvaluator()
    res = g.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])
    bound = ev.BoundGraph(g, res, fixed=[ext(mont(q, R)), ext(mont(table, R))], advice=[ext(mont(a, R))], beta=m1(beta), gamma=m1(gamma),
                          theta=m1(theta), y=m1(y))
    numer = h2b.lookup_fold(ctx, bound, ext(z), ext(pa), ext(pt), ext(mont(l0, R)), ext(mont(l_last, R)), ext(mont(l_active, R)), k, ext_k,
                            np.zeros((1 << ext_k, 4), dtype=np.uint64))
    assert all(v == 0 for v in on_domain_values(ctx, h2b, numer, k, ext_k, range(n)))


@pytest.mark.parametrize("k,ext_k", [(4, 5), (9, 11), (12, 15), (17, 19)])
def test_divide_by_vanishing_poly(ctx, h2b, k, ext_k):
    v = rnd_fr(np.random.default_rng(2700 + k), 1 << ext_k)
    assert np.array_equal(h2b.divide_by_vanishing_poly(ctx, v, k, ext_k), orc.divide_by_vanishing_poly(v, k, ext_k))
    with pytest.raises(h2b.H2BError):
        h2b.divide_by_vanishing_poly(ctx, v[: 1 << k], k, k)
