"""Generates tests/golden/*.json with oracle/pyref.py (pure Python integers — independent of the C oracle and of the
CUDA code).  The reference's own prover cannot run here (Rust, un-vendored crates), so these are restatement goldens:
they freeze today's agreed answers so that the C oracle, the CUDA path and future rounds are compared with FIXED
files rather than with each other.  Run: python tests/golden/make_golden.py"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref as p

rnd = random.Random(0xB200)
hexs = lambda v: hex(v)


def pt(a):
    return None if a is None else [hexs(a[0]), hexs(a[1])]


# ---- MSM: 24 points a_i*G with edge vectors (identity base, repeated base, scalars 0 / 1 / r-1)
n = 24
pts = [p.g1_mul(3 + 5 * i, p.G1) for i in range(n)]
pts[7] = None
pts[9] = pts[8]
sc = [rnd.randrange(p.R) for _ in range(n)]
sc[0], sc[1], sc[2] = 0, 1, p.R - 1
wit = [0, 1, 1, 0, 5, (1 << 88) - 3] + [rnd.randrange(1 << 88) for _ in range(n - 6)]
msm = {
    "curve": "BN254 G1, y^2 = x^3 + 3, G = (1, 2); values are canonical integers (not Montgomery)",
    "bases": [pt(a) for a in pts],
    "scalars_uniform": [hexs(s) for s in sc],
    "result_uniform": pt(p.msm_naive(sc, pts)),
    "scalars_witness_like": [hexs(s) for s in wit],
    "result_witness_like": pt(p.msm_naive(wit, pts)),
    "sum_to_infinity": {"bases": [pt(pts[3]), pt(pts[3])], "scalars": [hexs(5), hexs(p.R - 5)], "result": None},
}
json.dump(msm, open(os.path.join(HERE, "msm_g1.json"), "w"), indent=1)

# ---- NTT over Fr: k = 4, EvaluationDomain with j = 5 (extended_k = 6)
k, ext_k = 4, 6
a = [rnd.randrange(p.R) for _ in range(1 << k)]
ntt = {
    "field": "BN254 Fr; canonical integers; omega_k = ROOT_OF_UNITY^(2^(28-k))",
    "k": k, "extended_k": ext_k, "omega": hexs(p.omega_for(k)), "zeta": hexs(p.ZETA),
    "a": [hexs(v) for v in a],
    "best_fft": [hexs(v) for v in p.dft(a, p.omega_for(k))],
    "lagrange_to_coeff": [hexs(v) for v in p.lagrange_to_coeff(a, k)],
    "coeff_to_extended": [hexs(v) for v in p.coeff_to_extended(a, k, ext_k)],
}
back = p.extended_to_coeff(p.coeff_to_extended(a, k, ext_k), k, ext_k, 4)
assert back == a + [0] * (3 << k)
json.dump(ntt, open(os.path.join(HERE, "ntt_fr.json"), "w"), indent=1)

# ---- witness assignment: the walk of single_phase.rs:273-312 on 3 threads, 2 break points, k = 4
threads = [[rnd.randrange(1, 1 << 60) for _ in range(ln)] for ln in (5, 0, 9, 7)]
bps = [6, 8]
cols = p.assign_witnesses(threads, bps, 3, 1 << 4)
lk = p.assign_lookups([v for t in threads for v in t][:11], 3, 1 << 4)
json.dump({"k": 4, "threads": [[hexs(v) for v in t] for t in threads], "break_points": bps,
           "columns": [[hexs(v) for v in c] for c in cols],
           "lookup_values": [hexs(v) for t in threads for v in t][:11], "lookup_columns": [[hexs(v) for v in c] for c in lk]},
          open(os.path.join(HERE, "assign.json"), "w"), indent=1)

# ---- next rows: opening arithmetic, the lookup permutation, quotient terms — k = 3, extended 2^5
k, ext_k, bf = 3, 5, 2
n, ne, u = 1 << k, 1 << ext_k, (1 << k) - 3
poly = [rnd.randrange(p.R) for _ in range(11)]
z = rnd.randrange(p.R)
table = [4, 9, 9, 2, 7]
inputs = [9, 2, 9, 9, 4]
a_perm, s_perm = p.permute_expression_pair(inputs, table)
col = lambda: [rnd.randrange(p.R) for _ in range(ne)]
zs, cs, ss = [col(), col()], [col(), col(), col()], [col(), col(), col()]
l0, l_last, l_active, start = col(), col(), col(), col()
beta, gamma, y = (rnd.randrange(p.R) for _ in range(3))
perm = p.permutation_terms(zs, cs, ss, 2, l0, l_last, l_active, beta, gamma, y, bf, k, ext_k, start)
tv, zc, ap, sp = col(), col(), col(), col()
lookup = p.lookup_terms(tv, zc, ap, sp, l0, l_last, l_active, beta, gamma, y, k, ext_k, start)
H = lambda xs: [hexs(v) for v in xs]
# keygen side at k = 2: g[i] = tau^i G, g_lagrange[i] = L_i(tau) G (closed form), and the vanishing-polynomial division
tau = rnd.randrange(p.R)
k2, n2 = 2, 4
w2 = p.omega_for(k2)
c2 = (pow(tau, n2, p.R) - 1) * pow(n2, -1, p.R) % p.R
lag = [c2 * pow(w2, i, p.R) * pow((tau - pow(w2, i, p.R)) % p.R, -1, p.R) % p.R for i in range(n2)]
srs = {"k": k2, "tau": hexs(tau), "g": [pt(p.g1_mul(pow(tau, i, p.R), p.G1)) for i in range(n2)],
       "g_lagrange": [pt(p.g1_mul(l, p.G1)) for l in lag]}
we = p.omega_for(ext_k)
vanish_in = col()
vanish_out = [v * pow((pow(p.ZETA * pow(we, i, p.R) % p.R, n, p.R) - 1) % p.R, -1, p.R) % p.R for i, v in enumerate(vanish_in)]
json.dump({
    "note": "BN254 Fr, canonical integers; formulas of oracle/pyref.py (halo2 evaluate_h / arithmetic restated)",
    "k": k, "extended_k": ext_k, "blinding_factors": bf,
    "poly": H(poly), "point": hexs(z), "eval_polynomial": hexs(p.eval_polynomial(poly, z)), "kate_division": H(p.kate_division(poly, z)),
    "lookup_inputs": H(inputs), "lookup_table": H(table), "permuted_input": H(a_perm), "permuted_table": H(s_perm),
    "z_sets": [H(c) for c in zs], "columns": [H(c) for c in cs], "sigma": [H(c) for c in ss], "chunk_len": 2,
    "l0": H(l0), "l_last": H(l_last), "l_active": H(l_active), "start": H(start), "beta": hexs(beta), "gamma": hexs(gamma), "y": hexs(y),
    "permutation_fold": H(perm),
    "srs": srs, "vanishing_in": H(vanish_in), "vanishing_out": H(vanish_out),
    "table_values": H(tv), "lookup_z": H(zc), "lookup_a": H(ap), "lookup_s": H(sp), "lookup_fold": H(lookup),
}, open(os.path.join(HERE, "next_rows.json"), "w"), indent=1)
print("wrote msm_g1.json ntt_fr.json assign.json next_rows.json")
