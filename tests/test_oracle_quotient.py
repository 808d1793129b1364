"""CPU tests of the oracle's quotient-evaluation and opening-arithmetic restatements:
C oracle == independent Python big-int formulas, the two formulations of the flex gate agree, and — the check that
pins the term formulas to the protocol — a satisfied permutation / lookup argument makes the folded terms vanish on
the 2^k domain while a broken one does not."""
import numpy as np
import pytest
from oracle import oracle as orc, pyref
from util import mont, unmont, rand_ints
import quotient_cases as qc
import halo2_lib_b200 as h
from halo2_lib_b200 import evaluation as ev

R = pyref.R


def to_ext(col_lagrange, k, ext_k):
    """Lagrange ints -> extended-domain Montgomery limbs through the C oracle's transforms"""
    return orc.coeff_to_extended(orc.lagrange_to_coeff(mont(col_lagrange, R), k), ext_k)


def vanishes_on_domain(values_ext, k, ext_k):
    """interpolate the extended-domain evaluations and evaluate at every omega^i of the 2^k domain"""
    coeffs = unmont(orc.extended_to_coeff(values_ext, ext_k), R)
    w = pyref.omega_for(k)
    return [pyref.eval_polynomial(coeffs, pow(w, i, R)) for i in range(1 << k)]


def test_eval_polynomial_and_kate_division_vs_python():
    rng = np.random.default_rng(41)
    for n in (1, 2, 7, 8, 9, 255, 2049):
        a = rand_ints(rng, n, R)
        z = rand_ints(rng, 1, R)[0]
        am, zm = mont(a, R), mont([z], R)[0]
        assert unmont(orc.eval_polynomial(am, zm), R)[0] == pyref.eval_polynomial(a, z)
        q = unmont(orc.kate_division(am, zm), R)
        assert q == pyref.kate_division(a, z)
        # a(X) = q(X) (X - z) + a(z): compare at a random point
        x = rand_ints(rng, 1, R)[0]
        assert (pyref.eval_polynomial(q, x) * (x - z) + pyref.eval_polynomial(a, z)) % R == pyref.eval_polynomial(a, x)
    assert unmont(orc.eval_polynomial(np.zeros((0, 4), dtype=np.uint64), mont([5], R)[0]), R)[0] == 0


def test_poly_lincomb_vs_python():
    rng = np.random.default_rng(42)
    polys = [rand_ints(rng, 33, R) for _ in range(5)]
    sc = rand_ints(rng, 5, R)
    got = unmont(orc.poly_lincomb([mont(p, R) for p in polys], mont(sc, R)), R)
    assert got == [sum(s * p[i] for s, p in zip(sc, polys)) % R for i in range(33)]


def flex_gate_graph():
    """halo2-base's gate q * (a + b*c - out), rotations 0..3 of advice column 0 (flex_gate/mod.rs:80-91)"""
    g = ev.GraphEvaluator()
    a, b, c, out = [("advice", 0, r) for r in range(4)]
    poly = ("product", ("fixed", 0, 0), ("sum", ("sum", a, ("product", b, c)), ("negated", out)))
    return g, g.add_gates([poly])


def test_graph_flex_gate_equals_dedicated_fold_and_python():
    k, ext_k = 4, 6
    rng = np.random.default_rng(43)
    n = 1 << ext_k
    q, a, acc = (rand_ints(rng, n, R) for _ in range(3))
    y = rand_ints(rng, 1, R)[0]
    qm, am, accm, ym = mont(q, R), mont(a, R), mont(acc, R), mont([y], R)[0]
    g, res = flex_gate_graph()
    assert len(g.calculations) <= 10 and g.rotations == [0, 1, 2, 3]
    bound = ev.BoundGraph(g, res, fixed=[qm], advice=[am], y=ym)
    got = orc.quotient_graph(bound.struct, k, ext_k, accm)
    assert np.array_equal(got, orc.flex_gate_fold(qm, am, ym, k, ext_k, accm))
    consts = unmont(np.array(g.constants, dtype=np.uint64), R)
    prog = [int(x) for x in g.program()]
    for idx in (0, 1, n - 1, 17):
        want = pyref.graph_row(prog, len(g.calculations), ev._word(res), consts, g.rotations, [q], [a], [], [], 0, 0, 0, y, acc[idx], idx, k, ext_k)
        assert unmont(got[idx:idx + 1], R)[0] == want == (acc[idx] * y + q[idx] * (a[idx] + a[(idx + 4) % n] * a[(idx + 8) % n] - a[(idx + 12) % n])) % R


def test_graph_every_opcode_vs_python():
    k, ext_k = 3, 5
    n = 1 << ext_k
    rng = np.random.default_rng(44)
    cols = [rand_ints(rng, n, R) for _ in range(4)]
    chal = rand_ints(rng, 2, R)
    beta, gamma, theta, y = rand_ints(rng, 4, R)
    prev = rand_ints(rng, n, R)
    g = ev.GraphEvaluator()
    c7 = g.add_constant(mont([7], R)[0])
    s0 = g.add_calculation((ev.STORE, ev.src(ev.FIXED, 0, g.add_rotation(-1))))
    s1 = g.add_calculation((ev.ADD, s0, ev.src(ev.ADVICE, 1, g.add_rotation(2))))
    s2 = g.add_calculation((ev.SUB, s1, ev.src(ev.INSTANCE, 0, g.add_rotation(0))))
    s3 = g.add_calculation((ev.MUL, s2, ev.src(ev.CHALLENGE, 1)))
    s4 = g.add_calculation((ev.SQUARE, s3))
    s5 = g.add_calculation((ev.DOUBLE, s4))
    s6 = g.add_calculation((ev.NEGATE, s5))
    s7 = g.add_calculation((ev.HORNER, ev.src(ev.PREVIOUS), ev.src(ev.THETA), (s6, c7, ev.src(ev.BETA), ev.src(ev.GAMMA), ev.src(ev.Y), ev.src(ev.ADVICE, 0, g.add_rotation(-3)))))
    m = lambda c: mont(c, R)
    bound = ev.BoundGraph(g, s7, fixed=[m(cols[0])], advice=[m(cols[1]), m(cols[2])], instance=[m(cols[3])], challenges=m(chal),
                          beta=m([beta])[0], gamma=m([gamma])[0], theta=m([theta])[0], y=m([y])[0])
    got = unmont(orc.quotient_graph(bound.struct, k, ext_k, m(prev)), R)
    consts = unmont(np.array(g.constants, dtype=np.uint64), R)
    prog = [int(x) for x in g.program()]
    for idx in range(n):
        assert got[idx] == pyref.graph_row(prog, len(g.calculations), ev._word(s7), consts, g.rotations, [cols[0]], [cols[1], cols[2]],
                                           [cols[3]], chal, beta, gamma, theta, y, prev[idx], idx, k, ext_k)


@pytest.mark.parametrize("n_cols,chunk_len", [(1, 2), (3, 2), (4, 3)])
def test_permutation_terms_vs_python_and_vanish_when_satisfied(n_cols, chunk_len):
    k, ext_k, bf = 4, 6 if chunk_len <= 2 else 7, 5
    rng = np.random.default_rng(45 + n_cols)
    beta, gamma, y = rand_ints(rng, 3, R)
    l0, l_last, l_active, _ = qc.lagrange_basis_columns(k, bf)
    for broken in (False, True):
        cols, sigma, z_sets = qc.permutation_case(k, n_cols, chunk_len, bf, beta, gamma, seed=7, break_copy=broken)
        e = lambda c: to_ext(c, k, ext_k)
        zs, cs, ss, ls = [e(z) for z in z_sets], [e(c) for c in cols], [e(s) for s in sigma], [e(l0), e(l_last), e(l_active)]
        start = rand_ints(rng, 1 << ext_k, R) if not broken else [0] * (1 << ext_k)
        got = orc.permutation_fold(zs, cs, ss, chunk_len, *ls, mont([beta], R)[0], mont([gamma], R)[0], mont([y], R)[0], bf, k, ext_k, mont(start, R))
        if not broken:  # C oracle == Python big-int formulas (non-zero start exercises the fold order)
            u = lambda arrs: [unmont(a, R) for a in arrs]
            want = pyref.permutation_terms(u(zs), u(cs), u(ss), chunk_len, *u(ls), beta, gamma, y, bf, k, ext_k, start)
            assert unmont(got, R) == want
            got = orc.permutation_fold(zs, cs, ss, chunk_len, *ls, mont([beta], R)[0], mont([gamma], R)[0], mont([y], R)[0], bf, k, ext_k,
                                       np.zeros((1 << ext_k, 4), dtype=np.uint64))
        on_domain = vanishes_on_domain(got, k, ext_k)
        assert all(v == 0 for v in on_domain) != broken


def test_lookup_terms_vs_python_and_vanish_when_satisfied():
    # the selector-gated lookup has degree 5: its numerator only interpolates from a domain of >= 5n points
    k, ext_k, bf = 4, 7, 5
    rng = np.random.default_rng(46)
    beta, gamma, theta, y = rand_ints(rng, 4, R)
    l0, l_last, l_active, _ = qc.lagrange_basis_columns(k, bf)
    e = lambda c: to_ext(c, k, ext_k)
    m1 = lambda v: mont([v], R)[0]
    for broken in (False, True):
        q, a, table, a_perm, s_perm, z = qc.lookup_case(k, bf, beta, gamma, seed=9, break_lookup=broken)
        g = ev.GraphEvaluator()
        res = g.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])
        qe, ae, te = e(q), e(a), e(table)
        bound = ev.BoundGraph(g, res, fixed=[qe, te], advice=[ae], beta=m1(beta), gamma=m1(gamma), theta=m1(theta), y=m1(y))
        ze, ape, spe, ls = e(z), e(a_perm), e(s_perm), [e(l0), e(l_last), e(l_active)]
        zero = np.zeros((1 << ext_k, 4), dtype=np.uint64)
        got = orc.lookup_fold(bound.struct, ze, ape, spe, *ls, k, ext_k, zero)
        if not broken:
            u = lambda arr: unmont(arr, R)
            tv = [(qq * aa + beta) * (tt + gamma) % R for qq, aa, tt in zip(u(qe), u(ae), u(te))]
            want = pyref.lookup_terms(tv, u(ze), u(ape), u(spe), *[u(x) for x in ls], beta, gamma, y, k, ext_k, [0] * (1 << ext_k))
            assert unmont(got, R) == want
        on_domain = vanishes_on_domain(got, k, ext_k)
        assert all(v == 0 for v in on_domain) != broken


def lookup_columns(rng, k, bf, kind):
    """(inputs, table) integers over the usable rows for a few shapes of lookup"""
    u = (1 << k) - (bf + 1)
    if kind == "range":      # range check: table 0..u-1, inputs small with many repeats
        table = list(range(u))
        inputs = [int(x) for x in rng.integers(0, min(u, 37), size=u)]
    elif kind == "dup_table":  # the table itself holds duplicates (zero padding), inputs hit a few values
        table = [int(x) for x in rng.integers(0, 16, size=u)]
        inputs = [table[int(j)] for j in rng.integers(0, u, size=u)]
    elif kind == "wide":     # full-width field elements: every limb takes part in the order
        table = rand_ints(rng, u, R)
        inputs = [table[int(j)] for j in rng.integers(0, u, size=u)]
    elif kind == "all_same":
        table = list(range(u))
        inputs = [5] * u
    else:                    # a permutation of the table: no repeated row at all
        table = rand_ints(rng, u, R)
        inputs = [table[int(j)] for j in rng.permutation(u)]
    return inputs, table


@pytest.mark.parametrize("kind", ["range", "dup_table", "wide", "all_same"])
def test_permute_expression_pair_both_upstream_orders(kind):
    """the two walks that exist upstream (zcash: BTreeMap + pop from the back; PSE / axiom: two cursors front to back) give
    the same A', the same multiset on the repeated rows, and mirror-image orders there; the C walks agree with the
    closed forms in pyref for both"""
    k, bf = 7, 5
    rng = np.random.default_rng(48)
    u = (1 << k) - (bf + 1)
    inputs, table = lookup_columns(rng, k, bf, kind)
    pad = rand_ints(rng, bf + 1, R)
    A, T = mont(inputs + pad, R), mont(table + pad, R)
    rc0, pa0, pt0 = orc.permute_expression_pair(A, T, k, bf)
    rc1, pa1, pt1 = orc.permute_expression_pair(A, T, k, bf, zcash_order=True)
    assert rc0 == 0 and rc1 == 0 and np.array_equal(pa0, pa1)
    w0, w1 = pyref.permute_expression_pair(inputs, table), pyref.permute_expression_pair(inputs, table, zcash_order=True)
    assert unmont(pt0[:u], R) == w0[1] and unmont(pt1[:u], R) == w1[1]
    rep = [i for i in range(u) if i > 0 and w0[0][i] == w0[0][i - 1]]
    assert [w0[1][i] for i in rep] == [w1[1][i] for i in rep][::-1]
    if len(set(w0[1][i] for i in rep)) > 1:
        assert w0[1] != w1[1]


@pytest.mark.parametrize("kind", ["range", "dup_table", "wide", "all_same", "perm"])
def test_permute_expression_pair_vs_python(kind):
    k, bf = 7, 5
    rng = np.random.default_rng(47)
    u = (1 << k) - (bf + 1)
    inputs, table = lookup_columns(rng, k, bf, kind)
    pad = rand_ints(rng, bf + 1, R)
    rc, pa, pt = orc.permute_expression_pair(mont(inputs + pad, R), mont(table + pad, R), k, bf)
    assert rc == 0
    want_a, want_s = pyref.permute_expression_pair(inputs, table)
    assert unmont(pa[:u], R) == want_a and unmont(pt[:u], R) == want_s
    assert not pa[u:].any() and not pt[u:].any()
    # the properties the lookup argument needs: S' is a permutation of S, and A'[i] is S'[i] or A'[i-1]
    assert sorted(want_s) == sorted(table)
    assert all(want_a[i] == want_s[i] or (i and want_a[i] == want_a[i - 1]) for i in range(u))


def test_permute_expression_pair_missing_value():
    k, bf = 5, 5
    u = (1 << k) - (bf + 1)
    table = list(range(u))
    inputs = [3] * (u - 1) + [u + 9]
    rc, _, _ = orc.permute_expression_pair(mont(inputs + [0] * (bf + 1), R), mont(table + [0] * (bf + 1), R), k, bf)
    assert rc == -1 and pyref.permute_expression_pair(inputs, table) is None


def test_divide_by_vanishing_poly_vs_python():
    k, ext_k = 3, 5
    n, ne = 1 << k, 1 << ext_k
    rng = np.random.default_rng(48)
    vals = rand_ints(rng, ne, R)
    got = unmont(orc.divide_by_vanishing_poly(mont(vals, R), k, ext_k), R)
    w = pyref.omega_for(ext_k)
    want = [v * pow((pow(pyref.ZETA * pow(w, i, R) % R, n, R) - 1) % R, -1, R) % R for i, v in enumerate(vals)]
    assert got == want
    # a multiple of t(X) = X^n - 1 of degree < 2^ext_k divides exactly: (X^n - 1) * g(X) evaluated on the coset, divided, gives g
    gcoef = rand_ints(rng, n, R)
    prod = [0] * (2 * n)
    for i, c in enumerate(gcoef):
        prod[i] = (prod[i] - c) % R
        prod[i + n] = (prod[i + n] + c) % R
    ext = orc.coeff_to_extended(mont(prod, R), ext_k)
    back = unmont(orc.extended_to_coeff(orc.divide_by_vanishing_poly(ext, k, ext_k), ext_k), R)
    assert back[:n] == gcoef and not any(back[n:])


def test_graph_evaluator_builder_rules():
    """the add_expression simplifications the upstream GraphEvaluator applies: shared calculations and constants,
    x + 0, x * 1, x * 2 -> Double, x * x -> Square, a + (-b) -> Sub, scaling by 0 / 1"""
    g = ev.GraphEvaluator()
    zero, one, two = ("constant", ev.FR_ZERO), ("constant", ev.FR_ONE), ("constant", ev.FR_TWO)
    a, b = ("advice", 0, 0), ("advice", 1, -1)
    sa, sb = g.add_expression(a), g.add_expression(b)
    assert sa == ev.src(ev.INTERMEDIATE, 0) and sb == ev.src(ev.INTERMEDIATE, 1) and g.rotations == [0, -1]
    assert g.add_expression(a) == sa and len(g.calculations) == 2                      # shared
    assert g.add_expression(("sum", a, zero)) == sa and g.add_expression(("sum", zero, b)) == sb
    assert g.add_expression(("product", a, one)) == sa and g.add_expression(("product", zero, b)) == ev.src(ev.CONSTANT, 0)
    assert g.calculations[g.add_expression(("product", two, a))[1]] == (ev.DOUBLE, sa)
    assert g.calculations[g.add_expression(("product", a, a))[1]] == (ev.SQUARE, sa)
    assert g.calculations[g.add_expression(("sum", a, ("negated", b)))[1]] == (ev.SUB, sa, sb)
    assert g.calculations[g.add_expression(("sum", zero, ("negated", b)))[1]] == (ev.NEGATE, sb)
    assert g.add_expression(("product", b, a)) == g.add_expression(("product", a, b))  # commutative operands are ordered
    assert g.add_expression(("scaled", a, ev.FR_ZERO)) == ev.src(ev.CONSTANT, 0) and g.add_expression(("scaled", a, ev.FR_ONE)) == sa
    seven = mont([7], R)[0]
    s7 = g.add_expression(("scaled", a, seven))
    assert g.calculations[s7[1]] == (ev.MUL, sa, ev.src(ev.CONSTANT, 3)) and g.constants[3] == tuple(int(x) for x in seven)
    assert g.add_constant(seven) == ev.src(ev.CONSTANT, 3) and len(g.constants) == 4
    n_before = len(g.calculations)
    res = g.add_gates([("product", ("fixed", 0, 0), a)])
    assert g.calculations[res[1]][0] == ev.HORNER and g.calculations[res[1]][1] == ev.src(ev.PREVIOUS) and len(g.calculations) == n_before + 3
    # the encoded program round-trips through the oracle on a tiny domain
    k, ext_k = 2, 3
    rng = np.random.default_rng(49)
    cols = [rand_ints(rng, 1 << ext_k, R) for _ in range(3)]
    y = rand_ints(rng, 1, R)[0]
    prev = rand_ints(rng, 1 << ext_k, R)
    bound = ev.BoundGraph(g, res, fixed=[mont(cols[0], R)], advice=[mont(cols[1], R), mont(cols[2], R)], y=mont([y], R)[0])
    got = unmont(orc.quotient_graph(bound.struct, k, ext_k, mont(prev, R)), R)
    assert got == [(prev[i] * y + cols[0][i] * cols[1][i]) % R for i in range(1 << ext_k)]
