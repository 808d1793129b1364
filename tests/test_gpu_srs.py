"""GPU parity tests (-m gpu) of the keygen-side SRS utilities: g_to_lagrange (FFT over G1),
srs_setup for a given tau, the on-curve check — against the oracle at small k, and against each other at a size the CPU
cannot reach in test time (two independent routes to the Lagrange basis must agree, and must commit consistently)."""
import ctypes as C
import numpy as np
import pytest
from oracle import pyref, oracle as orc
from util import mont, unmont, rand_ints, affine_to_limbs

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


BASE = affine_to_limbs([pyref.G1])[0]


def setup(ctx, tau, k):
    from halo2_lib_b200._capi import lib
    g, gl = np.empty((1 << k, 8), dtype=np.uint64), np.empty((1 << k, 8), dtype=np.uint64)
    t = mont([tau], R)[0]
    ctx.check(lib.h2b_srs_setup(ctx.h, C.c_void_p(t.ctypes.data), C.c_void_p(BASE.ctypes.data), k, C.c_void_p(g.ctypes.data), C.c_void_p(gl.ctypes.data)))
    return g, gl


def to_lagrange(ctx, g, k):
    from halo2_lib_b200._capi import lib
    g = np.ascontiguousarray(g, dtype=np.uint64)
    out = np.empty_like(g)
    ctx.check(lib.h2b_g_to_lagrange(ctx.h, C.c_void_p(g.ctypes.data), k, C.c_void_p(out.ctypes.data)))
    return out


@pytest.mark.parametrize("k", [0, 1, 2, 3, 6, 9])
def test_srs_setup_and_g_to_lagrange_vs_oracle(ctx, h2b, k):
    tau = rand_ints(np.random.default_rng(2400 + k), 1, R)[0]
    g, gl = setup(ctx, tau, k)
    wg, wgl = orc.srs_setup(mont([tau], R)[0], BASE, k)
    assert np.array_equal(g, wg) and np.array_equal(gl, wgl)
    assert np.array_equal(to_lagrange(ctx, g, k), wgl)


def test_g_to_lagrange_arbitrary_points_with_identities(ctx, h2b):
    k = 7
    rng = np.random.default_rng(2500)
    sc = mont(rand_ints(rng, 1 << k, R), R)
    pts = ctx.g1_fixed_base_mul(BASE, sc)
    pts[3] = 0
    pts[100] = 0  # identities (0,0) among the inputs
    pts[5] = pts[4]  # repeated point
    assert np.array_equal(to_lagrange(ctx, pts, k), orc.g_to_lagrange(pts, k))


def test_two_routes_to_the_lagrange_basis_agree_at_2_16(ctx, h2b):
    """k = 16: the Lagrange basis from tau (closed form) == the G1 FFT of the monomial basis; both on the curve; and the
    KZG identity commit_lagrange(evals) == commit(coefficients) holds with them"""
    from halo2_lib_b200._capi import lib
    k = 16
    n = 1 << k
    rng = np.random.default_rng(2600)
    tau = rand_ints(rng, 1, R)[0]
    g, gl = setup(ctx, tau, k)
    assert np.array_equal(to_lagrange(ctx, g, k), gl)
    bad = C.c_size_t(99)
    ctx.check(lib.h2b_g1_check_on_curve(ctx.h, C.c_void_p(gl.ctypes.data), n, C.byref(bad)))
    assert bad.value == 0
    gl2 = gl.copy()
    gl2[17, 0] ^= np.uint64(1)
    gl2[4000, 5] ^= np.uint64(4)
    ctx.check(lib.h2b_g1_check_on_curve(ctx.h, C.c_void_p(gl2.ctypes.data), n, C.byref(bad)))
    assert bad.value == 2
    params = h2b.ParamsKZG(ctx, k, g=g, g_lagrange=gl)
    evals = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    evals[:, 3] &= np.uint64((1 << 60) - 1)
    coeffs = h2b.EvaluationDomain(ctx, 2, k).lagrange_to_coeff(evals)
    a = ctx.g1_normalize(params.commit_lagrange(evals).reshape(1, 12))
    b = ctx.g1_normalize(params.commit(coeffs).reshape(1, 12))
    assert np.array_equal(a, b)
    params.close()
