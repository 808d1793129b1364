"""TEST INFRASTRUCTURE — a CPU restatement of the data flow of halo2-axiom 0.5.3 `create_proof` (sole halo2-lib call site:
halo2-base/src/utils/testing.rs:40-48) for the constraint system halo2-base builds, on plain Python integers.

Only tests may import this file.  It is the checker of the RESIDENT prover (halo2-lib_b200/prover.py and its compiled twin
include/h2b200_prover.hpp): given the same instance, SRS, random polynomial and blinding rows it must produce the same
commitments and evaluations byte for byte, although nothing here shares code with them — polynomials are lists of integers,
transforms are the recursive definitions of oracle/pyref.py, commitments are naive sums of scalar multiples, the quotient is
formed row by row on the extended coset from the textbook terms (pyref.permutation_terms / lookup_terms), divisions are
schoolbook.  Sizes: k <= 6 finishes in seconds.

The prover crate is not vendored: phase order, term order and the transcript are the restated ones
the resident prover documents; what this file pins is that the device pipeline computes exactly that flow.

Circuit shape (halo2-base `BaseCircuitParams`): A gate-advice columns with selectors q{j} and the vertical gate
q (a0 + a1 a2 - a3) (flex_gate/mod.rs:80-91); L lookup-advice columns looked up in `table` (range/mod.rs:131-150), or the
selector lookup q_lookup * a0 (range/mod.rs:92-94), or none; one constants column; equality on [c, a0.., l0..]."""
from __future__ import annotations
import hashlib
from . import pyref

R, P = pyref.R, pyref.P
R_MONT, P_MONT = (1 << 256) % R, (1 << 256) % P
BLINDING_FACTORS = 6


def fr_bytes(v: int) -> bytes:
    """an Fr value as the provers hold it: Montgomery form, 32 little-endian bytes"""
    return (v % R * R_MONT % R).to_bytes(32, "little")


def g1_bytes(pt) -> bytes:
    """a commitment as it enters the transcript: affine (x, y, 1) in Montgomery form, 96 bytes; the identity is all zero"""
    if pt is None:
        return bytes(96)
    return b"".join((c * P_MONT % P).to_bytes(32, "little") for c in (pt[0], pt[1], 1))


class Transcript:
    def __init__(self):
        self.h = hashlib.blake2b(digest_size=64)

    def absorb(self, b: bytes):
        self.h.update(b)

    def squeeze(self) -> int:
        d = self.h.digest()
        self.h.update(b"\x00")
        return int.from_bytes(d, "little") % R


def create_proof(k: int, A: int, L: int, selector_lookup: bool, fixed: dict, sigma: list, virtual: list, break_points: list,
                 lookup_cells: list, random_poly: list, blind, bases_m: list, bases_l: list) -> dict:
    """all field values canonical integers; `fixed`: name -> 2^k values (q0.., [q_lookup], [table], c); `sigma`: one column per
    permutation column [c, a0.., l0..]; `blind(rows)`: the next `rows` blinding scalars; bases: affine points (None = identity).
    Returns {"commitments": [96-byte strings], "evals": [(name, rotation, value)], "challenges": {...}}."""
    n = 1 << k
    selector_lookup = selector_lookup and L == 0
    n_lookups = L if L else (1 if selector_lookup else 0)
    degree = 4 if L else (5 if selector_lookup else 3)
    chunk = degree - 2
    ext_k = k + (1 if degree == 3 else 2)
    ne = 1 << ext_k
    bf = BLINDING_FACTORS
    u = n - (bf + 1)
    adv_names = ["a%d" % j for j in range(A)] + ["l%d" % t for t in range(L)]
    perm_cols = ["c"] + adv_names
    n_sets = (len(perm_cols) + chunk - 1) // chunk
    fixed_names = ["q%d" % j for j in range(A)] + (["q_lookup"] if selector_lookup else []) + (["table"] if n_lookups else []) + ["c"]
    w = pyref.omega_for(k)
    tr = Transcript()
    commitments, lagr, coef, ext = [], {}, {}, {}

    def commit(items):
        """items: (basis, values); basis 0 = monomial (coefficients), 1 = lagrange"""
        out = []
        for basis, vals in items:
            cm = g1_bytes(pyref.msm_naive(vals, bases_l if basis else bases_m))
            commitments.append(cm)
            out.append(cm)
        return b"".join(out)

    def transforms(names):
        for nm in names:
            coef[nm] = pyref.lagrange_to_coeff(lagr[nm], k)
            ext[nm] = pyref.coeff_to_extended(coef[nm], k, ext_k)

    def blind_rows(col, first_row):
        col[first_row:] = blind(n - first_row)

    # the fixed side in its three forms
    fx = {nm: list(fixed[nm]) for nm in fixed_names}
    fx.update({"sigma_" + nm: list(sg) for nm, sg in zip(perm_cols, sigma)})
    fx["l0"] = [1] + [0] * (n - 1)
    fx["l_last"] = [1 if i == u else 0 for i in range(n)]
    fx["l_active"] = [1 if i < u else 0 for i in range(n)]
    fx_coef = {nm: pyref.lagrange_to_coeff(v, k) for nm, v in fx.items()}
    fx_ext = {nm: pyref.coeff_to_extended(c, k, ext_k) for nm, c in fx_coef.items()}

    # ---- phase 0: assignment (single_phase.rs:273-312, lookups.rs:130-155), blinding rows, advice commitments
    cols = pyref.assign_witnesses([list(virtual)], [int(b) for b in break_points], A, n)
    if L:
        cols += pyref.assign_lookups(list(lookup_cells), L, n)
    for nm, col in zip(adv_names, cols):
        lagr[nm] = col
        blind_rows(col, u)
    tr.absorb(commit([(1, lagr[nm]) for nm in adv_names]))
    theta = tr.squeeze()
    transforms(adv_names)
    # ---- lookups: compressed input, permuted pair
    lk_in = []
    for t in range(n_lookups):
        inp = [q * a % R for q, a in zip(fx["q_lookup"], lagr["a0"])] if L == 0 else lagr["l%d" % t]
        lk_in.append(inp)
        pair = pyref.permute_expression_pair(inp[:u], fx["table"][:u])
        if pair is None:
            raise ValueError("ConstraintSystemFailure: a lookup input is not in the table")
        for nm, vals in zip(("pa%d" % t, "ps%d" % t), pair):
            lagr[nm] = list(vals) + [0] * (n - u)
            blind_rows(lagr[nm], u)
    perm_names = [nm % t for t in range(n_lookups) for nm in ("pa%d", "ps%d")]
    if n_lookups:
        tr.absorb(commit([(1, lagr[nm]) for nm in perm_names]))
    beta, gamma = tr.squeeze(), tr.squeeze()
    transforms(perm_names)
    # ---- product columns
    col_of = lambda nm: fx["c"] if nm == "c" else lagr[nm]
    start = 1
    for s in range(n_sets):
        z = [start]
        for i in range(u):
            num = den = 1
            for cidx in range(s * chunk, min(len(perm_cols), (s + 1) * chunk)):
                v = col_of(perm_cols[cidx])[i]
                num = num * (v + beta * pow(pyref.DELTA, cidx, R) % R * pow(w, i, R) + gamma) % R
                den = den * (v + beta * fx["sigma_" + perm_cols[cidx]][i] + gamma) % R
            z.append(z[-1] * num % R * pow(den, -1, R) % R)
        start = z[u]
        lagr["zp%d" % s] = z + [0] * (n - u - 1)
    for t in range(n_lookups):
        z = [1]
        pa, ps = lagr["pa%d" % t], lagr["ps%d" % t]
        for i in range(u):
            z.append(z[-1] * (lk_in[t][i] + beta) % R * (fx["table"][i] + gamma) % R * pow((pa[i] + beta) * (ps[i] + gamma) % R, -1, R) % R)
        lagr["zl%d" % t] = z + [0] * (n - u - 1)
    prod_names = ["zp%d" % s for s in range(n_sets)] + ["zl%d" % t for t in range(n_lookups)]
    for nm in prod_names:
        blind_rows(lagr[nm], u + 1)
    transforms(prod_names)
    rnd = [c % R for c in random_poly]
    tr.absorb(commit([(1, lagr[nm]) for nm in prod_names] + [(0, rnd)]))
    y = tr.squeeze()
    # ---- quotient on the extended coset: gates (Horner in y), permutation terms, lookup terms, division by X^n - 1
    rot = lambda col, idx, r: pyref.rotate(col, idx, r, k, ext_k)
    values = []
    for idx in range(ne):
        v = 0
        for j in range(A):
            a = ext["a%d" % j]
            v = (v * y + fx_ext["q%d" % j][idx] * (a[idx] + rot(a, idx, 1) * rot(a, idx, 2) - rot(a, idx, 3))) % R
        values.append(v)
    ext_of = lambda nm: fx_ext["c"] if nm == "c" else ext[nm]
    values = pyref.permutation_terms([ext["zp%d" % s] for s in range(n_sets)], [ext_of(nm) for nm in perm_cols],
                                     [fx_ext["sigma_" + nm] for nm in perm_cols], chunk, fx_ext["l0"], fx_ext["l_last"], fx_ext["l_active"],
                                     beta, gamma, y, bf, k, ext_k, values)
    for t in range(n_lookups):
        if L == 0:
            inp_e = [q * a % R for q, a in zip(fx_ext["q_lookup"], ext["a0"])]
        else:
            inp_e = ext["l%d" % t]
        tv = [(i_ + beta) * (t_ + gamma) % R for i_, t_ in zip(inp_e, fx_ext["table"])]
        values = pyref.lookup_terms(tv, ext["zl%d" % t], ext["pa%d" % t], ext["ps%d" % t], fx_ext["l0"], fx_ext["l_last"], fx_ext["l_active"],
                                    beta, gamma, y, k, ext_k, values)
    we = pyref.omega_for(ext_k)
    for idx in range(ne):
        x_row = pyref.ZETA * pow(we, idx, R) % R
        values[idx] = values[idx] * pow(pow(x_row, n, R) - 1, -1, R) % R
    h = pyref.extended_to_coeff(values, k, ext_k)
    pieces = degree - 1
    assert not any(h[pieces * n:]), "the quotient has degree (degree - 1) n at most"
    tr.absorb(commit([(0, h[j * n:(j + 1) * n]) for j in range(pieces)]))
    x = tr.squeeze()
    # ---- evaluations
    point = lambda r: x * pow(w, r % n, R) % R
    last = -(bf + 1)
    queries = [("a%d" % j, coef["a%d" % j], r) for j in range(A) for r in (0, 1, 2, 3)]
    queries += [("l%d" % t, coef["l%d" % t], 0) for t in range(L)]
    queries += [(nm, fx_coef[nm], 0) for nm in fixed_names + ["sigma_" + nm for nm in perm_cols]]
    for s in range(n_sets):
        queries += [("zp%d" % s, coef["zp%d" % s], r) for r in ((0, 1, last) if s < n_sets - 1 else (0, 1))]
    for t in range(n_lookups):
        queries += [("pa%d" % t, coef["pa%d" % t], 0), ("pa%d" % t, coef["pa%d" % t], -1), ("ps%d" % t, coef["ps%d" % t], 0),
                    ("zl%d" % t, coef["zl%d" % t], 0), ("zl%d" % t, coef["zl%d" % t], 1)]
    queries += [("h%d" % j, h[j * n:(j + 1) * n], 0) for j in range(pieces)] + [("rnd", rnd, 0)]
    evals = [(nm, r, pyref.eval_polynomial(poly, point(r))) for nm, poly, r in queries]
    tr.absorb(b"".join(fr_bytes(v) for _, _, v in evals))
    # ---- SHPLONK-shaped opening: per rotation set sum_i v^i p_i divided by every (X - point) of the set
    v_ch, mu = tr.squeeze(), tr.squeeze()
    by_poly = {}
    for nm, poly, r in queries:
        by_poly.setdefault(id(poly), (poly, []))[1].append(r)
    groups = {}
    for poly, rots in by_poly.values():
        groups.setdefault(tuple(rots), []).append(poly)
    sets = sorted(groups.items(), key=lambda kv: (len(kv[0]), kv[0]))
    total = [0] * n
    for si, (rots, plist) in enumerate(sets):
        f = [sum(pow(v_ch, i, R) * p[c] for i, p in enumerate(plist)) % R for c in range(n)]
        for r in rots:
            f = pyref.kate_division(f, point(r)) + [0]  # n - 1 quotient coefficients, kept as an n-coefficient polynomial
        ms = pow(mu, si, R)
        total = [(a + ms * b) % R for a, b in zip(total, f)]
    tr.absorb(commit([(0, total)]))
    u_ch = tr.squeeze()
    commit([(0, pyref.kate_division(total, u_ch) + [0])])
    return {"commitments": commitments, "evals": evals, "challenges": dict(theta=theta, beta=beta, gamma=gamma, y=y, x=x)}
