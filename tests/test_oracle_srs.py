"""CPU tests of the oracle's keygen-side SRS restatements and of the host-only params view."""
import ctypes as C
import numpy as np
from oracle import oracle as orc, pyref
from util import mont, unmont, rand_ints, affine_to_limbs

R, P = pyref.R, pyref.P


def affine_ints(xy):
    xy = np.asarray(xy, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in xy:
        x, y = unmont(row.reshape(2, 4), P)
        out.append(None if x == 0 and y == 0 else (x, y))
    return out


def lagrange_at(tau, k):
    n = 1 << k
    w = pyref.omega_for(k)
    c = (pow(tau, n, R) - 1) * pow(n, -1, R) % R
    return [c * pow(w, i, R) * pow((tau - pow(w, i, R)) % R, -1, R) % R for i in range(n)]


def test_srs_setup_vs_python():
    k, tau = 3, 0x1234567890ABCDEF1234567
    g, gl = orc.srs_setup(mont([tau], R)[0], affine_to_limbs([pyref.G1])[0], k)
    assert affine_ints(g) == [pyref.g1_mul(pow(tau, i, R), pyref.G1) for i in range(1 << k)]
    assert affine_ints(gl) == [pyref.g1_mul(l, pyref.G1) for l in lagrange_at(tau, k)]
    assert sum(lagrange_at(tau, k)) % R == 1  # partition of unity


def test_g_to_lagrange_matches_setup_and_definition():
    rng = np.random.default_rng(51)
    base = affine_to_limbs([pyref.G1])[0]
    for k in (0, 1, 2, 5):
        tau = rand_ints(rng, 1, R)[0]
        g, gl = orc.srs_setup(mont([tau], R)[0], base, k)
        assert np.array_equal(orc.g_to_lagrange(g, k), gl)
    # definition on arbitrary points (not powers of tau), n = 4: out[i] = 1/n * sum_j omega^(-i j) g[j]
    k, n = 2, 4
    pts = [pyref.g1_mul(s, pyref.G1) for s in (5, 77, 1234, 99999)]
    got = affine_ints(orc.g_to_lagrange(affine_to_limbs(pts), k))
    w_inv, n_inv = pow(pyref.omega_for(k), -1, R), pow(n, -1, R)
    for i in range(n):
        acc = None
        for j in range(n):
            acc = pyref.g1_add(acc, pyref.g1_mul(pow(w_inv, i * j, R) * n_inv % R, pts[j]))
        assert got[i] == acc
    # an identity among the inputs
    pts[2] = None
    got = affine_ints(orc.g_to_lagrange(affine_to_limbs(pts), k))
    acc = None
    for j in range(n):
        acc = pyref.g1_add(acc, pyref.g1_mul(n_inv, pts[j]) if pts[j] else None)
    assert got[0] == acc


def test_params_raw_view_is_host_only():
    from halo2_lib_b200._capi import lib
    k = 3
    n = 1 << k
    blob = np.zeros(4 + 2 * n * 64 + 256, dtype=np.uint8)
    blob[0] = k
    kk, o = C.c_uint32(), [C.c_size_t() for _ in range(4)]
    assert lib.h2b_params_raw_view(C.c_void_p(blob.ctypes.data), len(blob), C.byref(kk), *[C.byref(x) for x in o]) == 0
    assert kk.value == k and [x.value for x in o] == [4, 4 + n * 64, 4 + 2 * n * 64, 4 + 2 * n * 64 + 128]
    assert lib.h2b_params_raw_view(C.c_void_p(blob.ctypes.data), len(blob) - 1, C.byref(kk), *[C.byref(x) for x in o]) == -1  # truncated
    blob[0] = 29
    assert lib.h2b_params_raw_view(C.c_void_p(blob.ctypes.data), len(blob), C.byref(kk), *[C.byref(x) for x in o]) == -1


def test_compressed_point_roundtrip_python():
    """SerdeFormat::Processed encoding restated in oracle/pyref.py: round trip, parity flag, invalid encodings"""
    rng = np.random.default_rng(77)
    for s in rand_ints(rng, 30, pyref.R):
        p = pyref.g1_mul(s, pyref.G1)
        e = pyref.g1_compress(p)
        assert len(e) == 32 and (e[31] >> 6) & 1 == p[1] & 1 and e[31] >> 7 == 0
        assert pyref.g1_decompress(e) == (p, True)
        assert pyref.g1_decompress(pyref.g1_compress(pyref.g1_neg(p))) == (pyref.g1_neg(p), True)
    assert pyref.g1_decompress(pyref.g1_compress(None)) == (None, True)
    assert pyref.g1_decompress(pyref.P.to_bytes(32, "little"))[1] is False


def test_glv_constants_of_the_g1_fft():
    """csrc/srs.cu multiplies by the twiddles through the endomorphism phi(x, y) = (beta x, y) = lambda (x, y): the
    constants hard-coded there, re-derived with plain integers — lambda and beta are matching cube roots of unity, the
    lattice vectors annihilate (1, lambda), the approximate division yields |k1|, |k2| < 2^127 and k = k1 + k2 lambda."""
    import random
    R, P = pyref.R, pyref.P
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe
    assert (lam * lam + lam + 1) % R == 0 and (beta * beta + beta + 1) % P == 0
    gx, gy = pyref.G1
    assert pyref.g1_mul(lam, pyref.G1) == (beta * gx % P, gy)
    a1, b1 = 0x89d3256894d213e3, -0x6f4d8248eeb859fc8211bbeb7d4f1128
    a2, b2 = 0x6f4d8248eeb859fd0be4e1541221250b, 0x89d3256894d213e3
    assert (a1 + b1 * lam) % R == 0 and (a2 + b2 * lam) % R == 0 and a1 * b2 - a2 * b1 == R
    g1, g2 = (b2 << 256) // R, (-b1 << 256) // R
    assert g1 == 0x2d91d232ec7e0b3d7 and g2 == 0x24ccef014a773d2cf7a7bd9d4391eb18d
    rng = random.Random(5)
    for k in [0, 1, R - 1, lam, R - lam] + [rng.randrange(R) for _ in range(2000)]:
        c1, c2 = (k * g1) >> 256, (k * g2) >> 256
        k1, k2 = k - c1 * a1 - c2 * a2, c1 * (-b1) - c2 * b2
        assert (k1 + k2 * lam - k) % R == 0 and abs(k1) < 1 << 127 and abs(k2) < 1 << 127
    # Montgomery form of beta as the kernel holds it
    bm = beta * (1 << 256) % P
    assert [(bm >> (32 * i)) & 0xffffffff for i in range(8)] == [0xd782e155, 0x71930c11, 0xffbe3323, 0xa6bb947c, 0xd4741444, 0xaa303344, 0x26594943, 0x2c3b3f0d]
