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


def quotient_identity(res: dict, k: int, blinding_factors: int, A: int = 1, L: int = 0, selector_lookup: bool = True) -> tuple[int, int]:
    """(left, right) of  fold(terms)(x) == h(x) * (x^n - 1)  for the halo2-base shape with A gate-advice and L lookup-advice
    columns (halo2-lib_b200/prover.py `Circuit`); term order as documented in include/h2b200.h: the gates by Horner in y,
    then the permutation terms (first set, last set, the links between consecutive sets, one product term per set), then
    every lookup's five terms."""
    n = 1 << k
    u = n - (blinding_factors + 1)
    n_lookups = L if L else (1 if selector_lookup else 0)
    degree = 4 if L else (5 if n_lookups else 3)
    chunk = degree - 2
    ch = res["challenges"]
    beta, gamma, y, x = ch["beta"], ch["gamma"], ch["y"], ch["x"]
    e = lambda name, r=0: fr(res["evals"][(name, r)])
    last = -(blinding_factors + 1)
    l0, l_last = lagrange_at(k, 0, x), lagrange_at(k, u, x)
    l_blind = sum(lagrange_at(k, i, x) for i in range(u + 1, n)) % R
    l_active = (1 - l_last - l_blind) % R
    v = 0
    for j in range(A):                                                             # the vertical gates
        a = "a%d" % j
        v = (v * y + e("q%d" % j) * (e(a, 0) + e(a, 1) * e(a, 2) - e(a, 3))) % R
    perm = ["c"] + ["a%d" % j for j in range(A)] + ["l%d" % t for t in range(L)]   # permutation argument
    n_sets = (len(perm) + chunk - 1) // chunk
    v = (v * y + (1 - e("zp0")) * l0) % R
    zl_ = e("zp%d" % (n_sets - 1))
    v = (v * y + (zl_ * zl_ - zl_) * l_last) % R
    for s in range(1, n_sets):
        v = (v * y + (e("zp%d" % s) - e("zp%d" % (s - 1), last)) * l0) % R
    for s in range(n_sets):
        left, right = e("zp%d" % s, 1), e("zp%d" % s, 0)
        for cidx in range(s * chunk, min(len(perm), (s + 1) * chunk)):
            val = e(perm[cidx])
            left = left * (val + beta * e("sigma_" + perm[cidx]) + gamma) % R
            right = right * (val + beta * pow(DELTA, cidx, R) % R * x + gamma) % R
        v = (v * y + (left - right) * l_active) % R
    for t in range(n_lookups):                                                     # lookup arguments
        pa, pa_p, ps = e("pa%d" % t, 0), e("pa%d" % t, -1), e("ps%d" % t, 0)
        zl, zl_n = e("zl%d" % t, 0), e("zl%d" % t, 1)
        inp = e("q_lookup") * e("a0") % R if L == 0 else e("l%d" % t)
        v = (v * y + (1 - zl) * l0) % R
        v = (v * y + (zl * zl - zl) * l_last) % R
        v = (v * y + (zl_n * (pa + beta) % R * (ps + gamma) - zl * (inp + beta) % R * (e("table") + gamma)) * l_active) % R
        v = (v * y + (pa - ps) * l0) % R
        v = (v * y + (pa - ps) * (pa - pa_p) % R * l_active) % R
    xn = pow(x, n, R)
    h = sum(e("h%d" % j) * pow(xn, j, R) for j in range(degree - 1)) % R
    return v % R, h * (xn - 1) % R
