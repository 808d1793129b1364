"""CPU: the oracle's restatement of the whole create_proof flow (oracle/prover_ref.py, plain Python integers) is itself
checked — on a SATISFIED tiny halo2-base circuit built here with integers only, the quotient it forms is a polynomial
(its coefficients beyond (degree - 1) n vanish, asserted inside create_proof), the quotient identity holds at the
challenge point (tests/prover_check.py, an independent formula sheet), and a broken witness violates both.
The GPU twin of this test (tests/test_gpu_prover.py::test_resident_prover_matches_the_oracle_prover) compares the resident
prover with this oracle byte for byte."""
import random
import numpy as np
import pytest
from oracle import pyref, prover_ref
import prover_check as pc

R = pyref.R


def int_instance(k, A, L, sel, seed):
    """the integer twin of halo2_lib_b200.prover.synthetic_circuit: gates on rows 4i..4i+3, looked-up small operands, bit cells
    tied to the constants column, lookup copies tied to their sources"""
    rng = random.Random(seed)
    n = 1 << k
    usable = n - 20
    G = usable // 4 if A == 1 else (usable - 4) // 4
    bits = min(8, k - 2)
    cols, a1s, a2s = [], [], []
    for j in range(A):
        col = [0] * n
        a1, a2 = [], []
        for i in range(G):
            x0, x1, x2 = rng.randrange(1 << 62), rng.randrange(1 << bits), rng.randrange(2)
            col[4 * i: 4 * i + 4] = [x0, x1, x2, (x0 + x1 * x2) % R]
            a1.append(x1); a2.append(x2)
        cols.append(col); a1s.append(a1); a2s.append(a2)
    virtual = [v for c in cols for v in c[: 4 * G]] if A > 1 else cols[0][:usable]
    break_points = [4 * G] * (A - 1)
    fixed = {}
    for j in range(A):
        fixed["q%d" % j] = [1 if (i % 4 == 0 and i < 4 * G) else 0 for i in range(n)]
    fixed["table"] = list(range(1 << bits)) + [0] * (n - (1 << bits))
    fixed["c"] = [0, 1] + [0] * (n - 2)
    lookup, lk_src = [], []
    if L == 0:
        if sel:
            fixed["q_lookup"] = [1 if (i % 4 == 1 and i < 4 * G) else 0 for i in range(n)]
    else:
        per_col = min(G, L * (usable - 7) // A)
        for j in range(A):
            for i in range(per_col):
                lookup.append(a1s[j][i]); lk_src.append((j, 4 * i + 1))
    w = pyref.omega_for(k)
    ids = [[pow(pyref.DELTA, c, R) * pow(w, i, R) % R for i in range(n)] for c in range(1 + A + L)]
    sig = [list(c) for c in ids]

    def tie(cells):
        for (c0, r0), (c1, r1) in zip(cells, cells[1:] + cells[:1]):
            sig[c0][r0] = ids[c1][r1]
    for bit in (0, 1):
        tie([(0, bit)] + [(1 + j, 4 * i + 2) for j in range(A) for i in range(G) if a2s[j][i] == bit])
    for i, (j, r) in enumerate(lk_src):
        tie([(1 + j, r), (1 + A + i % L, i // L)])
    return dict(fixed=fixed, sigma=sig, virtual=virtual, break_points=break_points, lookup=lookup)


def small_bases(n, a0, d):
    return [pyref.g1_mul(a0 + d * i, pyref.G1) for i in range(n)]


def run(k, A, L, sel, seed, inst=None):
    rng = random.Random(seed + 1)
    n = 1 << k
    inst = inst or int_instance(k, A, L, sel, seed)
    blind = lambda rows: [rng.randrange(R) for _ in range(rows)]
    res = prover_ref.create_proof(k, A, L, sel, inst["fixed"], inst["sigma"], inst["virtual"], inst["break_points"], inst["lookup"],
                                  [rng.randrange(R) for _ in range(n)], blind, small_bases(n, 3, 5), small_bases(n, 7, 11))
    as_limbs = lambda v: np.frombuffer(prover_ref.fr_bytes(v), dtype=np.uint64)
    return {"evals": {(nm, r): as_limbs(v) for nm, r, v in res["evals"]}, "challenges": res["challenges"], "commitments": res["commitments"]}


@pytest.mark.parametrize("A,L,sel", [(1, 0, True), (1, 0, False), (2, 1, True), (3, 2, True)])
def test_oracle_prover_satisfies_the_quotient_identity(A, L, sel):
    k = 5
    res = run(k, A, L, sel, 900 + 10 * A + L)
    nlk = L if L else (1 if sel else 0)
    deg = 4 if L else (5 if nlk else 3)
    n_sets = -(-(1 + A + L) // (deg - 2))
    assert len(res["commitments"]) == (A + L) + 2 * nlk + (n_sets + nlk + 1) + (deg - 1) + 2
    left, right = pc.quotient_identity(res, k, prover_ref.BLINDING_FACTORS, A, L, sel)
    assert left == right
    # a broken gate: the folded terms are no longer divisible by X^n - 1 — either the interpolated "quotient" has
    # coefficients beyond (degree - 1) n (asserted inside create_proof) or, when the extended domain has exactly (degree - 1) n
    # points, the identity fails at the challenge point
    inst = int_instance(k, A, L, sel, 900 + 10 * A + L)
    inst["virtual"][3] = (inst["virtual"][3] + 1) % R
    try:
        bad = run(k, A, L, sel, 900 + 10 * A + L, inst)
    except AssertionError:
        bad = None
    if bad is not None:
        l2, r2 = pc.quotient_identity(bad, k, prover_ref.BLINDING_FACTORS, A, L, sel)
        assert l2 != r2


def test_oracle_prover_rejects_a_value_outside_the_table():
    inst = int_instance(5, 1, 0, True, 77)
    inst["virtual"][1] = 1 << 40
    inst["virtual"][3] = (inst["virtual"][0] + inst["virtual"][1] * inst["virtual"][2]) % R  # the gate still holds
    with pytest.raises(ValueError):
        run(5, 1, 0, True, 77, inst)


def test_oracle_prover_reproduces_the_committed_golden_proof():
    """tests/golden/prover_k5.json (tests/golden/make_golden_prover.py): the frozen answer of the oracle prover"""
    import json, os
    from golden import make_golden_prover as g
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prover_k5.json")))
    assert g.proof() == json.loads(json.dumps(want))
