"""Committed golden vectors (tests/golden/*.json, made by tests/golden/make_golden.py from pure-Python integers):
the C oracle must reproduce them on the CPU, the CUDA path must reproduce them on the GPU."""
import json
import os
import numpy as np
import pytest
from oracle import pyref, oracle as orc
from util import *

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P, R = pyref.P, pyref.R
ints = lambda xs: [int(x, 16) for x in xs]
point = lambda a: None if a is None else (int(a[0], 16), int(a[1], 16))


def _msm_cases():
    d = json.load(open(os.path.join(G, "msm_g1.json")))
    bases = [point(a) for a in d["bases"]]
    yield bases, ints(d["scalars_uniform"]), point(d["result_uniform"])
    yield bases, ints(d["scalars_witness_like"]), point(d["result_witness_like"])
    s = d["sum_to_infinity"]
    yield [point(a) for a in s["bases"]], ints(s["scalars"]), None


def test_golden_regenerates_identically():
    """the generator is deterministic: the committed files are what pyref produces today"""
    import subprocess, sys, hashlib
    before = {f: hashlib.sha256(open(os.path.join(G, f), "rb").read()).hexdigest() for f in ("msm_g1.json", "ntt_fr.json", "assign.json", "next_rows.json")}
    subprocess.check_call([sys.executable, os.path.join(G, "make_golden.py")], stdout=subprocess.DEVNULL)
    after = {f: hashlib.sha256(open(os.path.join(G, f), "rb").read()).hexdigest() for f in before}
    assert before == after


def test_oracle_matches_golden():
    for bases, sc, want in _msm_cases():
        B, S = affine_to_limbs(bases), mont(sc, R)
        assert jac_limbs_to_affine(orc.msm_naive(S, B)) == want
        assert jac_limbs_to_affine(orc.msm_pippenger(S, B, 3)) == want
    d = json.load(open(os.path.join(G, "ntt_fr.json")))
    a, k, ek = ints(d["a"]), d["k"], d["extended_k"]
    assert unmont(orc.omega(k), R) == [int(d["omega"], 16)]
    assert unmont(orc.ntt(mont(a, R), k, orc.omega(k)), R) == ints(d["best_fft"])
    assert unmont(orc.ntt_fast(mont(a, R), k, orc.omega(k), 2), R) == ints(d["best_fft"])
    co = orc.lagrange_to_coeff(mont(a, R), k)
    assert unmont(co, R) == ints(d["lagrange_to_coeff"])
    assert unmont(orc.coeff_to_extended(mont(a, R), ek), R) == ints(d["coeff_to_extended"])
    d = json.load(open(os.path.join(G, "assign.json")))
    flat = [int(v, 16) for t in d["threads"] for v in t]
    rc, cols = orc.assign_witnesses(ints_to_limbs(flat), np.array(d["break_points"], dtype=np.uint64), d["k"], 3)
    assert rc == 0 and [limbs_to_ints(c) for c in cols] == [ints(c) for c in d["columns"]]
    rc, lk = orc.assign_lookups(ints_to_limbs(ints(d["lookup_values"])), d["k"], 3)
    assert rc == 0 and [limbs_to_ints(c) for c in lk] == [ints(c) for c in d["lookup_columns"]]


def _next_rows(lib_eval, lib_kate, lib_permute, lib_perm_fold, lib_lookup_fold, lib_vanish=None, lib_setup=None, lib_to_lagrange=None):
    """shared by the CPU (oracle) and GPU (CUDA) checks of tests/golden/next_rows.json"""
    d = json.load(open(os.path.join(G, "next_rows.json")))
    k, ek, bf = d["k"], d["extended_k"], d["blinding_factors"]
    m = lambda xs: mont(ints(xs), R)
    m1 = lambda x: mont([int(x, 16)], R)[0]
    assert unmont(lib_eval(m(d["poly"]), m1(d["point"])).reshape(1, 4), R) == [int(d["eval_polynomial"], 16)]
    assert unmont(lib_kate(m(d["poly"]), m1(d["point"])), R) == ints(d["kate_division"])
    pad = [0] * (bf + 1)
    pa, pt = lib_permute(mont(ints(d["lookup_inputs"]) + pad, R), mont(ints(d["lookup_table"]) + pad, R), k, bf)
    u = (1 << k) - (bf + 1)
    assert unmont(pa[:u], R) == ints(d["permuted_input"]) and unmont(pt[:u], R) == ints(d["permuted_table"])
    got = lib_perm_fold([m(c) for c in d["z_sets"]], [m(c) for c in d["columns"]], [m(c) for c in d["sigma"]], d["chunk_len"], m(d["l0"]),
                        m(d["l_last"]), m(d["l_active"]), m1(d["beta"]), m1(d["gamma"]), m1(d["y"]), bf, k, ek, m(d["start"]))
    assert unmont(got, R) == ints(d["permutation_fold"])
    # the lookup's table value comes from a one-calculation program: Store(fixed column 0)
    from halo2_lib_b200 import evaluation as ev
    g = ev.GraphEvaluator()
    res = g.add_calculation((ev.STORE, ev.src(ev.FIXED, 0, g.add_rotation(0))))
    bound = ev.BoundGraph(g, res, fixed=[m(d["table_values"])], beta=m1(d["beta"]), gamma=m1(d["gamma"]), y=m1(d["y"]))
    got = lib_lookup_fold(bound, m(d["lookup_z"]), m(d["lookup_a"]), m(d["lookup_s"]), m(d["l0"]), m(d["l_last"]), m(d["l_active"]), k, ek,
                          m(d["start"]))
    assert unmont(got, R) == ints(d["lookup_fold"])
    if lib_vanish:
        assert unmont(lib_vanish(m(d["vanishing_in"]), k, ek), R) == ints(d["vanishing_out"])
    if lib_setup:
        srs = d["srs"]
        want_g, want_gl = affine_to_limbs([point(a) for a in srs["g"]]), affine_to_limbs([point(a) for a in srs["g_lagrange"]])
        g, gl = lib_setup(m1(srs["tau"]), affine_to_limbs([pyref.G1])[0], srs["k"])
        assert np.array_equal(g, want_g) and np.array_equal(gl, want_gl)
        assert np.array_equal(lib_to_lagrange(want_g, srs["k"]), want_gl)


def test_oracle_matches_golden_next_rows():
    def permute(a, t, k, bf):
        rc, pa, pt = orc.permute_expression_pair(a, t, k, bf)
        assert rc == 0
        return pa, pt
    _next_rows(orc.eval_polynomial, orc.kate_division, permute, orc.permutation_fold,
               lambda b, *a: orc.lookup_fold(b.struct, *a), orc.divide_by_vanishing_poly, orc.srs_setup, orc.g_to_lagrange)


@pytest.mark.gpu
def test_cuda_matches_golden_next_rows():
    import halo2_lib_b200 as h
    ctx = h.Context(0)
    import ctypes as C
    from halo2_lib_b200._capi import lib

    def setup(tau, base, k):
        g, gl = np.empty((1 << k, 8), dtype=np.uint64), np.empty((1 << k, 8), dtype=np.uint64)
        ctx.check(lib.h2b_srs_setup(ctx.h, C.c_void_p(tau.ctypes.data), C.c_void_p(base.ctypes.data), k, C.c_void_p(g.ctypes.data), C.c_void_p(gl.ctypes.data)))
        return g, gl

    def to_lagrange(g, k):
        g = np.ascontiguousarray(g)
        out = np.empty_like(g)
        ctx.check(lib.h2b_g_to_lagrange(ctx.h, C.c_void_p(g.ctypes.data), k, C.c_void_p(out.ctypes.data)))
        return out
    _next_rows(lambda a, x: h.eval_polynomial(ctx, a, x), lambda a, x: h.kate_division(ctx, a, x),
               lambda a, t, k, bf: h.permute_expression_pair(ctx, a, t, k, bf),
               lambda *a: h.permutation_fold(ctx, *a), lambda *a: h.lookup_fold(ctx, *a),
               lambda v, k, ek: h.divide_by_vanishing_poly(ctx, v, k, ek), setup, to_lagrange)
    ctx.close()


@pytest.mark.gpu
def test_cuda_matches_golden():
    import halo2_lib_b200 as h
    ctx = h.Context(0)
    norm = lambda x: ctx.g1_normalize(np.asarray(x, dtype=np.uint64).reshape(1, 12))[0]
    for bases, sc, want in _msm_cases():
        B, S = affine_to_limbs(bases), mont(sc, R)
        assert jac_limbs_to_affine(norm(h.best_multiexp(ctx, S, B))) == want
    d = json.load(open(os.path.join(G, "msm_g1.json")))
    bases = [point(a) for a in d["bases"]]
    full = affine_to_limbs(bases + [pyref.G1] * 8)  # pad to 2^5 for the SRS path; padded scalars are zero
    p = h.ParamsKZG(ctx, 5, g=full)
    S = mont(ints(d["scalars_uniform"]) + [0] * 8, R)
    assert jac_limbs_to_affine(norm(p.commit(S))) == point(d["result_uniform"])
    p.close()
    d = json.load(open(os.path.join(G, "ntt_fr.json")))
    a, k, ek = ints(d["a"]), d["k"], d["extended_k"]
    assert unmont(h.omega(k).reshape(1, 4), R) == [int(d["omega"], 16)]
    assert unmont(h.best_fft(ctx, mont(a, R), h.omega(k), k), R) == ints(d["best_fft"])
    dom = h.EvaluationDomain(ctx, 5, k)
    assert dom.extended_k == ek
    assert unmont(dom.lagrange_to_coeff(mont(a, R)), R) == ints(d["lagrange_to_coeff"])
    ext = dom.coeff_to_extended(mont(a, R))
    assert unmont(ext, R) == ints(d["coeff_to_extended"])
    assert unmont(dom.extended_to_coeff(ext), R) == a + [0] * (3 << k)
    d = json.load(open(os.path.join(G, "assign.json")))
    threads = [ints_to_limbs(ints(t)) for t in d["threads"]]
    cols = h.assign_witnesses(ctx, threads, d["break_points"], d["k"], 3)
    assert [limbs_to_ints(c) for c in cols] == [ints(c) for c in d["columns"]]
    lk = h.assign_lookups(ctx, ints_to_limbs(ints(d["lookup_values"])), d["k"], 3)
    assert [limbs_to_ints(c) for c in lk] == [ints(c) for c in d["lookup_columns"]]
    ctx.close()
