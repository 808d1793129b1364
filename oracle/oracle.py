"""ctypes loader for the C oracle (oracle/bn254_oracle.c).  TEST INFRASTRUCTURE ONLY — see the header of
bn254_oracle.c: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference may use it."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
U64P = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bn254_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_g1_is_on_curve.restype = C.c_int
        _lib.orc_assign_witnesses.restype = C.c_int
        _lib.orc_assign_lookups.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_permute_expression_pair.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U64P)


def _binop(name, which, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    r = np.empty_like(a)
    getattr(lib(), name)(C.c_int(which), _p(a), _p(b), _p(r), C.c_size_t(len(a)))
    return r


def _unop(name, which, a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    r = np.empty_like(a)
    getattr(lib(), name)(C.c_int(which), _p(a), _p(r), C.c_size_t(len(a)))
    return r


FQ, FR = 0, 1
f_mul = lambda w, a, b: _binop("orc_f_mul", w, a, b)
f_add = lambda w, a, b: _binop("orc_f_add", w, a, b)
f_sub = lambda w, a, b: _binop("orc_f_sub", w, a, b)
f_inv = lambda w, a: _unop("orc_f_inv", w, a)
to_mont = lambda w, a: _unop("orc_f_to_mont", w, a)
from_mont = lambda w, a: _unop("orc_f_from_mont", w, a)


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    """override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to every rank)"""
    lib().orc_set_threads(C.c_int(n))


def g1_is_on_curve(xy) -> bool:
    xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(8)
    return bool(lib().orc_g1_is_on_curve(_p(xy)))


def g1_scalar_mul(s, base_xy):
    s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
    b = np.ascontiguousarray(base_xy, dtype=np.uint64).reshape(8)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_scalar_mul(_p(s), _p(b), _p(out))
    return out


def g1_add(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(12)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def g1_normalize(a):
    a = np.array(a, dtype=np.uint64).reshape(12).copy()
    lib().orc_g1_normalize(_p(a))
    return a


def g1_fixed_base_mul(scalars, base_xy):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(base_xy, dtype=np.uint64).reshape(8)
    out = np.empty((len(s), 8), dtype=np.uint64)
    lib().orc_g1_fixed_base_mul(_p(s), C.c_size_t(len(s)), _p(b), _p(out))
    return out


def msm_naive(scalars, bases):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert len(s) == len(b)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_msm_naive(_p(s), _p(b), C.c_size_t(len(s)), _p(out))
    return out


def msm_pippenger(scalars, bases, threads: int | None = None):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert len(s) == len(b)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_msm_pippenger(_p(s), _p(b), C.c_size_t(len(s)), C.c_int(threads or max_threads()), _p(out))
    return out


def omega(k: int):
    out = np.empty(4, dtype=np.uint64)
    lib().orc_omega(C.c_uint(k), _p(out))
    return out


def ntt(a, log_n: int, omega_m, threads: int | None = None):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(a) == 1 << log_n
    w = np.ascontiguousarray(omega_m, dtype=np.uint64).reshape(4)
    lib().orc_ntt(_p(a), C.c_uint(log_n), _p(w), C.c_int(threads or max_threads()))
    return a


def ntt_fast(a, log_n: int, omega_m, threads: int | None = None):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(a) == 1 << log_n
    w = np.ascontiguousarray(omega_m, dtype=np.uint64).reshape(4)
    lib().orc_ntt_fast(_p(a), C.c_uint(log_n), _p(w), C.c_int(threads or max_threads()))
    return a


def use_fast_ntt(on: bool) -> None:
    """route the EvaluationDomain wrappers through the many-core four-step transform (timed CPU baseline)"""
    lib().orc_use_fast_ntt(C.c_int(1 if on else 0))


def lagrange_to_coeff(a, k, threads=None):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(a) == 1 << k
    lib().orc_lagrange_to_coeff(_p(a), C.c_uint(k), C.c_int(threads or max_threads()))
    return a


def coeff_to_lagrange(a, k, threads=None):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(a) == 1 << k
    lib().orc_coeff_to_lagrange(_p(a), C.c_uint(k), C.c_int(threads or max_threads()))
    return a


def coeff_to_extended(coeffs, ext_k, threads=None):
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((1 << ext_k, 4), dtype=np.uint64)
    lib().orc_coeff_to_extended(_p(c), C.c_size_t(len(c)), C.c_uint(ext_k), _p(out), C.c_int(threads or max_threads()))
    return out


def extended_to_coeff(a, ext_k, threads=None):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(a) == 1 << ext_k
    lib().orc_extended_to_coeff(_p(a), C.c_uint(ext_k), C.c_int(threads or max_threads()))
    return a


def assign_witnesses(vcol, break_points, k, ncols):
    v = np.ascontiguousarray(vcol, dtype=np.uint64).reshape(-1, 4)
    bp = np.ascontiguousarray(break_points, dtype=np.uint64).reshape(-1)
    cols = np.empty((ncols, 1 << k, 4), dtype=np.uint64)
    rc = lib().orc_assign_witnesses(_p(v), C.c_size_t(len(v)), _p(bp) if len(bp) else None, C.c_size_t(len(bp)),
                                    C.c_uint(k), C.c_size_t(ncols), _p(cols) if ncols else None)
    return rc, cols


def assign_lookups(vals, k, L):
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    cols = np.empty((L, 1 << k, 4), dtype=np.uint64)
    rc = lib().orc_assign_lookups(_p(v), C.c_size_t(len(v)), C.c_uint(k), C.c_size_t(L), _p(cols) if L else None)
    return rc, cols


def batch_invert(a):
    a = np.array(a, dtype=np.uint64).reshape(-1, 4).copy()
    lib().orc_batch_invert(_p(a), C.c_size_t(len(a)))
    return a


def grand_product(f, start):
    f = np.ascontiguousarray(f, dtype=np.uint64).reshape(-1, 4)
    st = np.ascontiguousarray(start, dtype=np.uint64).reshape(4)
    z = np.empty_like(f)
    lib().orc_grand_product(_p(f), _p(st), C.c_size_t(len(f)), _p(z))
    return z


def flex_gate_fold(q, a, y, k, ext_k, acc):
    q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 4)
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    acc = np.array(acc, dtype=np.uint64).reshape(-1, 4).copy()
    yy = np.ascontiguousarray(y, dtype=np.uint64).reshape(4)
    lib().orc_flex_gate_fold(_p(q), _p(a), _p(yy), C.c_uint(k), C.c_uint(ext_k), _p(acc))
    return acc


def eval_rational(num, den):
    a = np.ascontiguousarray(num, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(den, dtype=np.uint64).reshape(-1, 4)
    r = np.empty_like(a)
    lib().orc_eval_rational(_p(a), _p(b), C.c_size_t(len(a)), _p(r))
    return r


# ---- general quotient evaluation / opening arithmetic -----------------------------
# `graph` is a ctypes structure with the field order of orc_graph (== h2b_graph); tests pass the very structure they
# hand to the CUDA library, with HOST column pointers.
def quotient_graph(graph, k, ext_k, values):
    v = np.array(values, dtype=np.uint64).reshape(-1, 4).copy()
    lib().orc_quotient_graph(C.byref(graph), C.c_uint(k), C.c_uint(ext_k), _p(v))
    return v


def lookup_fold(graph, z, permuted_input, permuted_table, l0, l_last, l_active, k, ext_k, values):
    a = [np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4) for x in (z, permuted_input, permuted_table, l0, l_last, l_active)]
    v = np.array(values, dtype=np.uint64).reshape(-1, 4).copy()
    lib().orc_lookup_fold(C.byref(graph), *[_p(x) for x in a], C.c_uint(k), C.c_uint(ext_k), _p(v))
    return v


def _ptr_table(cols):
    return (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])


def permutation_fold(z_sets, columns, sigma, chunk_len, l0, l_last, l_active, beta, gamma, y, blinding_factors, k, ext_k, values):
    conv = lambda xs: [np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4) for x in xs]
    zs, cs, ss, ls, ch = conv(z_sets), conv(columns), conv(sigma), conv([l0, l_last, l_active]), conv([beta, gamma, y])
    v = np.array(values, dtype=np.uint64).reshape(-1, 4).copy()
    lib().orc_permutation_fold(_ptr_table(zs), C.c_size_t(len(zs)), _ptr_table(cs), _ptr_table(ss), C.c_size_t(len(cs)),
                               C.c_size_t(chunk_len), *[_p(x) for x in ls], *[_p(x) for x in ch], C.c_uint(blinding_factors),
                               C.c_uint(k), C.c_uint(ext_k), _p(v))
    return v


def eval_polynomial(coeffs, x):
    a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    xx = np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_eval_polynomial(_p(a) if len(a) else None, C.c_size_t(len(a)), _p(xx), _p(out))
    return out


def kate_division(a, z):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    assert len(a) >= 1
    zz = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
    q = np.empty((len(a) - 1, 4), dtype=np.uint64)
    if len(a) > 1:
        lib().orc_kate_division(_p(a), C.c_size_t(len(a)), _p(zz), _p(q))
    return q


def poly_lincomb(polys, scalars):
    ps = [np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4) for x in polys]
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(ps[0])
    lib().orc_poly_lincomb(_ptr_table(ps), _p(s), C.c_size_t(len(ps)), C.c_size_t(len(ps[0])), _p(out))
    return out


def permute_expression_pair(input_col, table_col, k, blinding_factors, zcash_order=False):
    """returns (rc, permuted_input, permuted_table); rows >= usable are zero-filled here (the prover writes randoms).
    zcash_order: left-over table values to the repeated rows from the back (zcash halo2) instead of front to back"""
    a = np.ascontiguousarray(input_col, dtype=np.uint64).reshape(-1, 4)
    t = np.ascontiguousarray(table_col, dtype=np.uint64).reshape(-1, 4)
    pa, pt = np.zeros_like(a), np.zeros_like(t)
    rc = lib().orc_permute_expression_pair_ordered(_p(a), _p(t), C.c_uint(k), C.c_uint(blinding_factors), _p(pa), _p(pt),
                                                   C.c_int(1 if zcash_order else 0))
    return rc, pa, pt


def g_to_lagrange(g_xy, k):
    g = np.ascontiguousarray(g_xy, dtype=np.uint64).reshape(-1, 8)
    assert len(g) == 1 << k
    out = np.empty_like(g)
    lib().orc_g_to_lagrange(_p(g), C.c_uint(k), _p(out))
    return out


def srs_setup(tau, base_xy, k):
    t = np.ascontiguousarray(tau, dtype=np.uint64).reshape(4)
    b = np.ascontiguousarray(base_xy, dtype=np.uint64).reshape(8)
    g, gl = np.empty((1 << k, 8), dtype=np.uint64), np.empty((1 << k, 8), dtype=np.uint64)
    lib().orc_srs_setup(_p(t), _p(b), C.c_uint(k), _p(g), _p(gl))
    return g, gl


def divide_by_vanishing_poly(values, k, ext_k):
    v = np.array(values, dtype=np.uint64).reshape(-1, 4).copy()
    lib().orc_divide_by_vanishing_poly(_p(v), C.c_uint(k), C.c_uint(ext_k))
    return v
