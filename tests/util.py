"""Shared helpers for the tests: int <-> [u64;4] limb arrays, seeded inputs."""
from __future__ import annotations
import numpy as np
from oracle import pyref

MASK = (1 << 64) - 1


def ints_to_limbs(vals) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & MASK
    return out


def limbs_to_ints(arr) -> list[int]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(arr[i, j]) << (64 * j) for j in range(4)) for i in range(len(arr))]


def mont(vals, m) -> np.ndarray:
    return ints_to_limbs([pyref.to_mont(v % m, m) for v in vals])


def unmont(arr, m) -> list[int]:
    return [pyref.from_mont(v, m) for v in limbs_to_ints(arr)]


def rand_ints(rng: np.random.Generator, n: int, m: int) -> list[int]:
    """uniform mod m from 4 x u64 (+ a 5th to kill bias)"""
    raw = rng.integers(0, 1 << 63, size=(n, 5), dtype=np.int64).astype(object)
    return [int((r[0] | (r[1] << 63) | (r[2] << 126) | (r[3] << 189) | (r[4] << 252)) % m) for r in raw]


def witness_like_ints(rng: np.random.Generator, n: int) -> list[int]:
    """SURVEY.md §8(d) distribution (W): 35% zero, 25% one, 30% uniform < 2^88, 10% uniform Fr."""
    cls = rng.random(n)
    small = rng.integers(0, 1 << 62, size=(n, 2), dtype=np.int64).astype(object)
    full = rand_ints(rng, n, pyref.R)
    out = []
    for i in range(n):
        c = cls[i]
        if c < 0.35:
            out.append(0)
        elif c < 0.60:
            out.append(1)
        elif c < 0.90:
            out.append(int((small[i][0] | (small[i][1] << 62)) & ((1 << 88) - 1)))
        else:
            out.append(full[i])
    return out


def affine_to_limbs(pts) -> np.ndarray:
    """list of affine points (ints, None = identity) -> n x 8 Montgomery limbs, identity = (0,0)"""
    out = np.zeros((len(pts), 8), dtype=np.uint64)
    for i, pt in enumerate(pts):
        if pt is None:
            continue
        out[i, :4] = mont([pt[0]], pyref.P)[0]
        out[i, 4:] = mont([pt[1]], pyref.P)[0]
    return out


def jac_limbs_to_affine(xyz):
    """normalised Jacobian (12 limbs, Montgomery) -> affine ints or None"""
    xyz = np.asarray(xyz, dtype=np.uint64).reshape(12)
    x, y, z = unmont(xyz.reshape(3, 4), pyref.P)
    if z == 0:
        return None
    zi = pow(z, -1, pyref.P)
    return (x * zi * zi % pyref.P, y * zi * zi * zi % pyref.P)


def progression_points(n: int, a0: int = 1, delta: int = 1):
    """b_i = (a0 + i*delta)*G by repeated affine addition (SURVEY.md §8c closed-form check)."""
    step = pyref.g1_mul(delta, pyref.G1)
    cur = pyref.g1_mul(a0, pyref.G1)
    pts = []
    for _ in range(n):
        pts.append(cur)
        cur = pyref.g1_add(cur, step)
    return pts
