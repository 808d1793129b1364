"""Builders of small, VALID permutation / lookup argument instances (plain integers, Lagrange form) for the quotient
tests: a satisfied argument makes every folded term vanish on the 2^k domain, which pins the term formulas to the
protocol itself rather than to a second copy of the same code (halo2 book, "Permutation argument" / "Lookup argument";
halo2-axiom 0.5.3 plonk/{permutation,lookup}/prover.rs build the same columns)."""
from __future__ import annotations
import numpy as np
from oracle import pyref
from util import rand_ints

R = pyref.R


def lagrange_basis_columns(k: int, blinding_factors: int):
    """l_0, l_last, l_active = 1 - (l_last + l_blind) as evaluations on the 2^k domain (usable rows u = n - bf - 1)"""
    n = 1 << k
    u = n - (blinding_factors + 1)
    l0 = [1] + [0] * (n - 1)
    l_last = [1 if i == u else 0 for i in range(n)]
    l_active = [1 if i < u else 0 for i in range(n)]
    return l0, l_last, l_active, u


def permutation_case(k: int, n_cols: int, chunk_len: int, blinding_factors: int, beta: int, gamma: int, seed: int, break_copy: bool = False):
    """random copy constraints over the usable rows of n_cols columns; returns columns, sigma, z_sets (Lagrange form)"""
    rng = np.random.default_rng(seed)
    n = 1 << k
    _, _, _, u = lagrange_basis_columns(k, blinding_factors)
    w = pyref.omega_for(k)
    cells = [(j, i) for j in range(n_cols) for i in range(u)]
    perm = rng.permutation(len(cells))
    image = {cells[a]: cells[int(perm[a])] for a in range(len(cells))}  # sigma(cell): cycles of this map share a value
    cols = [rand_ints(rng, n, R) for _ in range(n_cols)]
    seen = set()
    for c in cells:  # equal values along every cycle
        if c in seen:
            continue
        val = rand_ints(rng, 1, R)[0]
        cur = c
        while cur not in seen:
            seen.add(cur)
            cols[cur[0]][cur[1]] = val
            cur = image[cur]
    if break_copy:
        cols[0][1] = (cols[0][1] + 1) % R
    ident = lambda j, i: pow(pyref.DELTA, j, R) * pow(w, i, R) % R
    sigma = [[ident(*image.get((j, i), (j, i))) for i in range(n)] for j in range(n_cols)]
    n_sets = (n_cols + chunk_len - 1) // chunk_len
    z_sets = []
    carry = 1
    for s in range(n_sets):
        z = [carry]
        for i in range(u):
            num = den = 1
            for j in range(s * chunk_len, min((s + 1) * chunk_len, n_cols)):
                num = num * (cols[j][i] + beta * ident(j, i) + gamma) % R
                den = den * (cols[j][i] + beta * sigma[j][i] + gamma) % R
            z.append(z[-1] * num % R * pow(den, -1, R) % R)
        carry = z[u]
        z += rand_ints(rng, n - u - 1, R)  # blinding rows
        z_sets.append(z)
    return cols, sigma, z_sets


def lookup_case(k: int, blinding_factors: int, beta: int, gamma: int, seed: int, break_lookup: bool = False):
    """one advice column gated by a selector looked up in one table column: input expression q*a (halo2-base
    gates/range/mod.rs:131-140).  Returns q, a, table, permuted input A', permuted table S', product z (Lagrange form)."""
    rng = np.random.default_rng(seed)
    n = 1 << k
    _, _, _, u = lagrange_basis_columns(k, blinding_factors)
    table = list(range(u)) + rand_ints(rng, n - u, R)  # table rows 0..u-1 hold the range [0, u)
    q = [int(rng.integers(0, 2)) for _ in range(u)] + [0] * (n - u)
    a = [int(rng.integers(0, u)) if q[i] else rand_ints(rng, 1, R)[0] for i in range(u)] + rand_ints(rng, n - u, R)
    inputs = [q[i] * a[i] % R for i in range(u)]  # rows with q = 0 look up 0, which is in the table
    if break_lookup:
        inputs[2] = (u + 7) % R
        q[2], a[2] = 1, inputs[2]
    a_perm = sorted(inputs)
    # S': the table value where A' starts a new run, the left-over table values elsewhere
    leftover = sorted(set(table[:u]) - set(a_perm))
    s_perm = []
    for i in range(u):
        if i == 0 or a_perm[i] != a_perm[i - 1]:
            s_perm.append(a_perm[i])
        else:
            s_perm.append(leftover.pop(0))
    z = [1]
    for i in range(u):
        num = (inputs[i] + beta) * (table[i] + gamma) % R
        den = (a_perm[i] + beta) * (s_perm[i] + gamma) % R
        z.append(z[-1] * num % R * pow(den, -1, R) % R)
    z += rand_ints(rng, n - u - 1, R)
    a_perm += rand_ints(rng, n - u, R)
    s_perm += rand_ints(rng, n - u, R)
    return q, a, table, a_perm, s_perm, z
