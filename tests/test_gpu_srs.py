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


def test_g1_decompress_and_processed_params_image(ctx, h2b):
    """SerdeFormat::Processed: the Fq square-root kernel against Python integers (both parities, identity, invalid encodings)
    and the device side of `ParamsKZG::read` on a whole params image (halo2-base/src/utils/mod.rs:401-435)"""
    from halo2_lib_b200._capi import lib
    from util import limbs_to_ints
    rng = np.random.default_rng(2900)
    P = pyref.P
    pts = [pyref.g1_mul(s, pyref.G1) for s in rand_ints(rng, 40, R)] + [None, pyref.G1, pyref.g1_neg(pyref.G1)]
    enc = [pyref.g1_compress(p) for p in pts]
    for e, p in zip(enc, pts):
        assert pyref.g1_decompress(e) == (p, True)
    # invalid encodings: x >= p, x^3 + 3 a non-residue, infinity flag with x != 0, infinity flag with the sign bit
    nonres = next(x for x in range(2, 100) if pow((x ** 3 + 3) % P, (P - 1) // 2, P) != 1)
    bad = [P.to_bytes(32, "little"), nonres.to_bytes(32, "little"), (5).to_bytes(31, "little") + bytes([0x80]), bytes(31) + bytes([0xC0])]
    for e in bad:
        assert pyref.g1_decompress(e)[1] is False
    blob = np.frombuffer(b"".join(enc + bad), dtype=np.uint8).copy()
    n = len(enc) + len(bad)
    out = np.empty((n, 8), dtype=np.uint64)
    inv = C.c_size_t()
    ctx.check(lib.h2b_g1_decompress(ctx.h, C.c_void_p(blob.ctypes.data), n, C.c_void_p(out.ctypes.data), C.byref(inv)))
    assert inv.value == len(bad)
    assert np.array_equal(out[: len(enc)], affine_to_limbs(pts)) and not out[len(enc):].any()
    # a whole image: k, g, g_lagrange, g2, s_g2 (G2 parts are opaque to this library)
    k = 6
    tau = rand_ints(rng, 1, R)[0]
    g, gl = setup(ctx, tau, k)

    def aff(arr):
        v = [pyref.from_mont(x, P) for x in limbs_to_ints(arr.reshape(-1, 4))]
        return [None if (v[2 * i] == 0 and v[2 * i + 1] == 0) else (v[2 * i], v[2 * i + 1]) for i in range(len(arr))]
    image = (k).to_bytes(4, "little") + b"".join(pyref.g1_compress(p) for p in aff(g)) + b"".join(pyref.g1_compress(p) for p in aff(gl)) + bytes(128)
    img = np.frombuffer(image, dtype=np.uint8).copy()
    kk, og, ol = C.c_uint32(), C.c_size_t(), C.c_size_t()
    assert lib.h2b_params_processed_view(C.c_void_p(img.ctypes.data), len(img), C.byref(kk), C.byref(og), C.byref(ol), None, None) == 0
    assert (kk.value, og.value, ol.value) == (k, 4, 4 + 32 * (1 << k))
    assert lib.h2b_params_processed_view(C.c_void_p(img.ctypes.data), len(img) - 1, C.byref(kk), None, None, None, None) == -1
    hsrs = C.c_void_p()
    ctx.check(lib.h2b_srs_read_processed(ctx.h, C.c_void_p(img.ctypes.data), len(img), 0, 0, C.byref(hsrs)))
    sc = mont(rand_ints(rng, 1 << k, R), R)
    for basis, bases in ((0, g), (1, gl)):
        got = np.empty(12, dtype=np.uint64)
        ctx.check(lib.h2b_msm_g1(ctx.h, hsrs, basis, C.c_void_p(sc.ctypes.data), 1 << k, C.c_void_p(got.ctypes.data)))
        assert np.array_equal(ctx.g1_normalize(got.reshape(1, 12))[0], orc.msm_pippenger(sc, bases))
    lib.h2b_srs_destroy(ctx.h, hsrs)
    # a corrupted point makes the read fail (H2B_ERR_ARG), as ParamsKZG::read rejects it
    img2 = img.copy()
    img2[4:36] = np.frombuffer(nonres.to_bytes(32, "little"), dtype=np.uint8)
    assert lib.h2b_srs_read_processed(ctx.h, C.c_void_p(img2.ctypes.data), len(img2), 0, 0, C.byref(hsrs)) == -1
