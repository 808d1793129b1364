"""Protocol-level checks of one resident proof (halo2-lib_b200/prover.py), plain Python integers only — no oracle/, no
library calls: (1) the quotient identity  sum of the y-folded gate / permutation / lookup terms at x  ==  h(x) (x^n - 1)
from the evaluations the prover wrote; (2) Horner re-evaluation of downloaded coefficient arrays.  Used by
tests/test_gpu_prover.py and by bench.py's self-verification (outside the timed region)."""
from __future__ import annotations
import numpy as np

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
RINV = pow(1 << 256, -1, R)
ROOT = pow(7, (R - 1) >> 28, R)
DELTA = pow(7, 1 << 28, R)


def fr(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(l, dtype=np.uint64).reshape(4))) * RINV % R


def horner(coeff_limbs: np.ndarray, x: int) -> int:
    acc = 0
    for row in np.asarray(coeff_limbs, dtype=np.uint64).reshape(-1, 4)[::-1]:
        acc = (acc * x + (int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192))) % R
    return acc * RINV % R


def lagrange_at(k: int, i: int, x: int) -> int:
    """L_i(x) on the 2^k domain: omega^i (x^n - 1) / (n (x - omega^i))"""
    n = 1 << k
    w = pow(ROOT, 1 << (28 - k), R)
    wi = pow(w, i % n, R)
    return wi * (pow(x, n, R) - 1) % R * pow(n * (x - wi) % R, -1, R) % R


def quotient_identity(res: dict, k: int, blinding_factors: int, degree: int = 5) -> tuple[int, int]:
    """(left, right) of  fold(terms)(x) == h(x) * (x^n - 1); term order as documented in include/h2b200.h:
    gates by Horner in y, then the permutation terms, then the lookup's five terms."""
    n = 1 << k
    u = n - (blinding_factors + 1)
    ch = res["challenges"]
    beta, gamma, y, x = ch["beta"], ch["gamma"], ch["y"], ch["x"]
    e = lambda name, r=0: fr(res["evals"][(name, r)])
    last = -(blinding_factors + 1)
    l0, l_last = lagrange_at(k, 0, x), lagrange_at(k, u, x)
    l_blind = sum(lagrange_at(k, i, x) for i in range(u + 1, n)) % R
    l_active = (1 - l_last - l_blind) % R
    a0, a1, a2, a3 = e("a", 0), e("a", 1), e("a", 2), e("a", 3)
    q, qlk, t, c, sc, sa = e("q"), e("q_lookup"), e("table"), e("c"), e("sigma_c"), e("sigma_a")
    zp, zp_n, zp_l = e("zp", 0), e("zp", 1), e("zp", last)
    pa, pa_p, ps = e("pa", 0), e("pa", -1), e("ps", 0)
    zl, zl_n = e("zl", 0), e("zl", 1)
    v = 0
    v = (v * y + q * (a0 + a1 * a2 - a3)) % R                                    # the vertical gate
    v = (v * y + (1 - zp) * l0) % R                                               # permutation argument
    v = (v * y + (zp * zp - zp) * l_last) % R
    left = zp_n * (c + beta * sc + gamma) % R * (a0 + beta * sa + gamma) % R
    right = zp * (c + beta * x + gamma) % R * (a0 + beta * DELTA % R * x + gamma) % R
    v = (v * y + (left - right) * l_active) % R
    v = (v * y + (1 - zl) * l0) % R                                               # lookup argument
    v = (v * y + (zl * zl - zl) * l_last) % R
    v = (v * y + (zl_n * (pa + beta) % R * (ps + gamma) - zl * (qlk * a0 + beta) % R * (t + gamma)) * l_active) % R
    v = (v * y + (pa - ps) * l0) % R
    v = (v * y + (pa - ps) * (pa - pa_p) % R * l_active) % R
    xn = pow(x, n, R)
    h = sum(e("h%d" % j) * pow(xn, j, R) for j in range(degree - 1)) % R
    return v % R, h * (xn - 1) % R
