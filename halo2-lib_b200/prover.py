"""Resident prover core: the device-side data flow of halo2-axiom 0.5.3 `create_proof` for the constraint system
halo2-base builds (one vertical gate per gate-advice column, halo2-base/src/gates/flex_gate/mod.rs:80-91; a range lookup
`q_lookup * a in table`, gates/range/mod.rs:92-94,131-141; equality on the constants column and the gate column,
flex_gate/mod.rs:69,124-129), with every column kept in HBM behind `h2b_poly` handles between the phases:

    witness (host) --H2D--> assign_witnesses --> commit advice                                   
    theta:  q_lookup * a, permute_expression_pair --> commit A', S'                              (step 3)
    beta, gamma:  permutation product, lookup product, random polynomial (host) --> commit       (steps 4, 5)
    y:  lagrange_to_coeff + coeff_to_extended of every column, gate / permutation / lookup terms folded on the
        extended coset, divide_by_vanishing_poly, extended_to_coeff, h pieces --> commit         (step 6)
    x:  evaluations                                                                              (step 7)
    SHPLONK-shaped opening: per rotation set a linear combination and kate divisions, two commitments   (step 8)

Only the witness cells, the random polynomial and the blinding scalars go up; only commitments and evaluations come
down.  The transcript stays on the host (as it stays in Rust): challenges are squeezed from Blake2b over the commitment
bytes.  The prover crate is not vendored, so phase order and term order are restated; what the tests and
bench.py check is protocol-level: every commitment equals the closed form of the polynomial it commits, and the quotient
identity  sum of folded terms (x) == h(x) * (x^n - 1)  holds at the challenge point.

No field arithmetic happens here on the hot path: everything is computed by the kernels behind include/h2b200.h."""
from __future__ import annotations
import ctypes as C
import hashlib
import numpy as np
from ._capi import lib, BASIS_MONOMIAL, BASIS_LAGRANGE
from .host import Context, ParamsKZG, H2BError
from . import evaluation as ev

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
MONT_R = (1 << 256) % R_MOD
MONT_RINV = pow(1 << 256, -1, R_MOD)
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
DELTA = pow(7, 1 << 28, R_MOD)
BLINDING_FACTORS = 6  # max(3, queries of the gate column = 4) + 2  


def to_limbs(x: int) -> np.ndarray:
    """canonical integer -> Montgomery [u64;4]"""
    v = x % R_MOD * MONT_R % R_MOD
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def from_limbs(l) -> int:
    """Montgomery [u64;4] -> canonical integer"""
    return sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(l, dtype=np.uint64).reshape(4))) * MONT_RINV % R_MOD


class Poly:
    """h2b_poly: a device-resident column / polynomial"""

    def __init__(self, ctx: Context, n: int):
        self.ctx, self.n = ctx, n
        h = C.c_void_p()
        ctx.check(lib.h2b_poly_alloc(ctx.h, n, C.byref(h)))
        self.h = h
        self.ptr = int(lib.h2b_poly_device_ptr(h))

    def upload(self, host: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(host, dtype=np.uint64).reshape(-1, 4)
        self.ctx.check(lib.h2b_poly_upload(self.ctx.h, self.h, offset, C.c_void_p(a.ctypes.data), len(a)))

    def upload_ptr(self, host_ptr: int, n: int, offset: int = 0):
        self.ctx.check(lib.h2b_poly_upload(self.ctx.h, self.h, offset, C.c_void_p(host_ptr), n))

    def download(self, offset: int = 0, n: int | None = None) -> np.ndarray:
        n = self.n - offset if n is None else n
        out = np.empty((n, 4), dtype=np.uint64)
        self.ctx.check(lib.h2b_poly_download(self.ctx.h, self.h, offset, C.c_void_p(out.ctypes.data), n))
        return out

    def at(self, elem_offset: int) -> int:
        return self.ptr + 32 * elem_offset

    def free(self):
        if self.h:
            lib.h2b_poly_free(self.ctx.h, self.h)
            self.h = None


class Transcript:
    """Blake2b over what the prover writes; `squeeze` yields an Fr challenge (host side, as the Rust transcript)"""

    def __init__(self):
        self.h = hashlib.blake2b(digest_size=64)

    def absorb(self, arr):
        self.h.update(np.ascontiguousarray(arr).tobytes())

    def squeeze(self) -> int:
        d = self.h.digest()
        self.h.update(b"\x00")
        return int.from_bytes(d, "little") % R_MOD


class Circuit:
    """The fixed side of the synthetic halo2-base circuit (what keygen_pk would hold), resident on the GPU in the three
    forms create_proof needs: Lagrange values, coefficients, extended-coset evaluations.
    Columns: fixed q (gate selector), q_lookup, table t, constants c; advice a.  Permutation over [c, a]."""

    def __init__(self, ctx: Context, k: int, fixed_lagrange: dict, sigma_lagrange: list):
        self.ctx, self.k, self.n = ctx, k, 1 << k
        self.degree = 5
        self.ext_k = k + 2
        self.bf = BLINDING_FACTORS
        self.u = self.n - (self.bf + 1)
        n, ne = self.n, 1 << self.ext_k
        vp = C.c_void_p
        l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = to_limbs(1)
        ll = np.zeros((n, 4), dtype=np.uint64); ll[self.u] = to_limbs(1)
        la = np.zeros((n, 4), dtype=np.uint64); la[: self.u] = to_limbs(1)
        cols = dict(fixed_lagrange)
        cols.update({"sigma_c": sigma_lagrange[0], "sigma_a": sigma_lagrange[1], "l0": l0, "l_last": ll, "l_active": la})
        self.lagr, self.coeff, self.ext = {}, {}, {}
        for name, arr in cols.items():
            lg, cf, ex = Poly(ctx, n), Poly(ctx, n), Poly(ctx, ne)
            lg.upload(arr)
            cf.upload(arr)
            ctx.check(lib.h2b_lagrange_to_coeff_dev(ctx.h, vp(cf.ptr), k))
            ctx.check(lib.h2b_coeff_to_extended_dev(ctx.h, vp(cf.ptr), n, self.ext_k, vp(ex.ptr)))
            self.lagr[name], self.coeff[name], self.ext[name] = lg, cf, ex
        ctx.synchronize()
        # the gate program (one vertical gate on the advice column) and the lookup's table-side program
        g = ev.GraphEvaluator()
        a = lambda r: ("advice", 0, r)
        gate = ("product", ("fixed", 0, 0), ("sum", ("sum", a(0), ("product", a(1), a(2))), ("negated", a(3))))
        self.gate_graph, self.gate_res = g, g.add_gates([gate])
        g2 = ev.GraphEvaluator()
        self.lk_graph, self.lk_res = g2, g2.add_lookup([("product", ("fixed", 0, 0), ("advice", 0, 0))], [("fixed", 1, 0)])

    def free(self):
        for d in (self.lagr, self.coeff, self.ext):
            for p in d.values():
                p.free()


def synthetic_circuit(ctx: Context, k: int, rng: np.random.Generator, lookup_bits: int = 8):
    """A SATISFIED instance: witness column `a` (canonical ints as limbs, Montgomery) + the fixed columns.
    Gates on rows 4j..4j+3 (a3 = a0 + a1*a2, computed on the GPU), lookups on the small operands, copy constraints
    between equal cells (cycles through the rows that hold the same constant)."""
    n = 1 << k
    usable = n - 20
    ngates = usable // 4
    lookup_bits = min(lookup_bits, k - 2)  # the table's 2^bits rows must fit the usable rows
    mont_small = lambda v: ctx.field_op(1, 5, np.stack([v.astype(np.uint64), np.zeros_like(v, dtype=np.uint64), np.zeros_like(v, dtype=np.uint64), np.zeros_like(v, dtype=np.uint64)], axis=1))
    a0c = rng.integers(0, 1 << 62, size=ngates, dtype=np.int64).astype(np.uint64)
    a1c = rng.integers(0, 1 << lookup_bits, size=ngates, dtype=np.int64).astype(np.uint64)  # looked up
    a2c = rng.integers(0, 2, size=ngates, dtype=np.int64).astype(np.uint64)                # bits: many equal cells
    A0, A1, A2 = mont_small(a0c), mont_small(a1c), mont_small(a2c)
    A3 = ctx.field_op(1, 1, A0, ctx.field_op(1, 0, A1, A2))
    a = np.zeros((n, 4), dtype=np.uint64)
    rows = 4 * np.arange(ngates)
    a[rows], a[rows + 1], a[rows + 2], a[rows + 3] = A0, A1, A2, A3
    one = to_limbs(1)
    q = np.zeros((n, 4), dtype=np.uint64); q[rows] = one
    qlk = np.zeros((n, 4), dtype=np.uint64); qlk[rows + 1] = one
    t = np.zeros((n, 4), dtype=np.uint64)
    t[: 1 << lookup_bits] = mont_small(np.arange(1 << lookup_bits, dtype=np.uint64))
    c = np.zeros((n, 4), dtype=np.uint64)
    c[0], c[1] = to_limbs(0), one
    # copy constraints: every bit cell a2 is tied into one cycle with the constant cell of its value (c[0] = 0, c[1] = 1)
    w = pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)
    # identity permutation values delta^j * omega^i; cheap through the GPU: omega^i as a geometric progression
    idx = np.arange(n, dtype=np.uint64)
    wp = _geometric(ctx, w, n)                       # omega^i, Montgomery limbs
    dl = np.tile(to_limbs(DELTA), (n, 1))
    id_c, id_a = wp, ctx.field_op(1, 0, wp, dl)
    sig_c, sig_a = id_c.copy(), id_a.copy()
    for bit in (0, 1):
        cells = rows[a2c == bit] + 2                  # rows of the advice column holding `bit`
        if len(cells) == 0:
            continue
        # cycle: c[bit] -> a[cells[0]] -> a[cells[1]] -> ... -> c[bit]
        sig_c[bit] = id_a[cells[0]]
        sig_a[cells[:-1]] = id_a[cells[1:]]
        sig_a[cells[-1]] = id_c[bit]
    fixed = {"q": q, "q_lookup": qlk, "table": t, "c": c}
    return a, fixed, [sig_c, sig_a], usable


def _geometric(ctx: Context, w: int, n: int) -> np.ndarray:
    """[w^0, w^1, ..., w^(n-1)] as Montgomery limbs, by doubling with the GPU's element-wise multiplier"""
    out = np.zeros((n, 4), dtype=np.uint64)
    out[0] = to_limbs(1)
    have, step = 1, w
    while have < n:
        m = min(have, n - have)
        out[have:have + m] = ctx.field_op(1, 0, out[:m], np.tile(to_limbs(pow(w, have, R_MOD)), (m, 1)))
        have += m
    return out


class ProverSession:
    """One proof at a time on one context; owns the resident working set (allocated once, reused for every proof)."""

    def __init__(self, ctx: Context, params: ParamsKZG, circuit: Circuit):
        self.ctx, self.params, self.cs = ctx, params, circuit
        n, ne = circuit.n, 1 << circuit.ext_k
        P = lambda m: Poly(ctx, m)
        self.v = P(n)                                   # virtual column (witness cells)
        self.a = P(n)                                   # advice column: Lagrange, then coefficients
        self.inp, self.pa, self.ps = P(n), P(n), P(n)   # compressed lookup input, permuted input / table
        self.zp, self.zl, self.rnd = P(n), P(n), P(n)   # product columns, random polynomial
        self.ext = {name: P(ne) for name in ("a", "pa", "ps", "zp", "zl")}
        self.h = P(ne)                                  # quotient values, then its coefficients (d - 1 pieces of n)
        # coefficient forms (the Lagrange forms stay alive for the product columns / the commitments running beside)
        self.ac, self.pac, self.psc, self.zpc, self.zlc = P(n), P(n), P(n), P(n), P(n)
        self.tmp = [P(n) for _ in range(4)]
        self.tmp_side = [P(n) for _ in range(3)]
        self.d_out = Poly(ctx, 16)                      # commitments of a phase: up to 4 x 12 limbs (12 elements of 32 B)
        self.h2d_bytes = self.d2h_bytes = 0
        self.begin, self.n_loc, self.allreduce = 0, n, None
        self.keep = None  # verification runs: dict that receives the committed polynomials (downloaded, untimed)

    def shard(self, begin: int, n_loc: int, allreduce):
        """multi-GPU: this rank commits rows [begin, begin + n_loc) of every polynomial and `allreduce(ptr, m)` combines the
        partial commitments of all ranks in place on the device (h2b_g1_allreduce_dev); everything else is replicated"""
        self.begin, self.n_loc, self.allreduce = begin, n_loc, allreduce

    # ---- helpers
    def _commit(self, items) -> np.ndarray:
        """items: list of (basis, device pointer); one batched launch, the commitments come down in one copy"""
        m = len(items)
        ctx = self.ctx
        ptrs = (C.c_void_p * m)(*[p + 32 * self.begin for _, p in items])
        bs = (C.c_int * m)(*[b for b, _ in items])
        ctx.check(lib.h2b_msm_g1_batch_dev(ctx.h, self.params.h, bs, ptrs, m, self.n_loc, C.c_void_p(self.d_out.ptr)))
        if self.allreduce is not None:
            self.allreduce(self.d_out.ptr, m)
        if self.keep is not None:  # untimed verification run: remember what was committed
            for b, p in items:
                arr = np.empty((self.cs.n, 4), dtype=np.uint64)
                ctx.synchronize()
                self._raw_download(p, arr)
                self.keep.setdefault("committed", []).append((b, arr))
        out = np.empty((m * 3, 4), dtype=np.uint64)
        ctx.check(lib.h2b_poly_download(ctx.h, self.d_out.h, 0, C.c_void_p(out.ctypes.data), m * 3))
        self.d2h_bytes += m * 96
        return out.reshape(m, 12)

    def _raw_download(self, dev_ptr: int, arr: np.ndarray):
        """device pointer inside one of the session's polynomials -> host (verification only)"""
        for p in [self.a, self.pa, self.ps, self.zp, self.zl, self.rnd, self.h, self.ac, self.pac, self.psc, self.zpc, self.zlc] + self.tmp + self.tmp_side:
            if p.ptr <= dev_ptr < p.ptr + 32 * p.n:
                self.ctx.check(lib.h2b_poly_download(self.ctx.h, p.h, (dev_ptr - p.ptr) // 32, C.c_void_p(arr.ctypes.data), len(arr)))
                return
        raise ValueError("pointer outside the session's polynomials")

    def _blind(self, poly: Poly, first_row: int, rng: np.random.Generator):
        cnt = self.cs.n - first_row
        b = rng.integers(0, 1 << 62, size=(cnt, 4), dtype=np.int64).astype(np.uint64)
        b[:, 3] &= np.uint64((1 << 60) - 1)
        poly.upload(b, first_row)
        self.h2d_bytes += cnt * 32

    def prove(self, witness_ptr: int, n_cells: int, random_poly_ptr: int, seed: int = 0) -> dict:
        """witness_ptr / random_poly_ptr: host pointers (pinned) to n_cells / n Montgomery Fr elements"""
        ctx, cs, vp = self.ctx, self.cs, C.c_void_p
        k, n, ext_k, bf, u = cs.k, cs.n, cs.ext_k, cs.bf, cs.u
        rng = np.random.default_rng(seed)
        tr = Transcript()
        self.h2d_bytes = self.d2h_bytes = 0
        res = {"commitments": []}
        import os, time
        trace = [] if os.environ.get("H2B_PROVER_TRACE") else None

        def mark(label):  # diagnostic: wall clock per phase with a full synchronisation (changes the overlap: not for timing runs)
            if trace is not None:
                ctx.synchronize()
                trace.append((label, time.perf_counter()))
        mark("start")
        def side_transforms(pairs):
            """beside the main queue: copy Lagrange -> coefficient buffer, lagrange_to_coeff, coeff_to_extended"""
            ctx.check(lib.h2b_ctx_side_begin(ctx.h))
            try:
                for src, dst, name in pairs:
                    if src is not dst:
                        ctx.check(lib.h2b_poly_copy_dev(ctx.h, vp(dst.ptr), vp(src.ptr), n))
                    ctx.check(lib.h2b_lagrange_to_coeff_dev(ctx.h, vp(dst.ptr), k))
                    ctx.check(lib.h2b_coeff_to_extended_dev(ctx.h, vp(dst.ptr), n, ext_k, vp(self.ext[name].ptr)))
            finally:
                ctx.check(lib.h2b_ctx_side_end(ctx.h))

        # ---- phase 0: witness up, assignment, advice commitment (the random polynomial goes up beside it)
        self.v.upload_ptr(witness_ptr, n_cells)
        self.h2d_bytes += n_cells * 32
        ctx.check(lib.h2b_ctx_side_begin(ctx.h))
        ctx.check(lib.h2b_poly_upload_async(ctx.h, self.rnd.h, 0, vp(random_poly_ptr), n))
        ctx.check(lib.h2b_ctx_side_end(ctx.h))
        self.h2d_bytes += n * 32
        ctx.check(lib.h2b_assign_columns_dev(ctx.h, vp(self.v.ptr), n_cells, None, 0, k, 1, vp(self.a.ptr)))
        self._blind(self.a, u, rng)
        cm = self._commit([(BASIS_LAGRANGE, self.a.ptr)])
        res["commitments"] += list(cm); tr.absorb(cm)
        theta = tr.squeeze()
        mark("phase0 advice")
        ctx.check(lib.h2b_ctx_side_join(ctx.h))  # the random polynomial arrived while phase 0 ran
        side_transforms([(self.a, self.ac, "a")])
        # ---- lookup: compressed input q_lookup * a, permuted pair
        ctx.check(lib.h2b_fr_mul_elementwise_dev(ctx.h, vp(cs.lagr["q_lookup"].ptr), vp(self.a.ptr), n, vp(self.inp.ptr)))
        # enqueue only: the verdict ("an input value is not in the table") lands in the last element of d_out and is read
        # right after the commitments of this phase, whose download synchronises anyway
        ctx.check(lib.h2b_permute_expression_pair_async_dev(ctx.h, vp(self.inp.ptr), vp(cs.lagr["table"].ptr), k, bf, vp(self.pa.ptr),
                                                            vp(self.ps.ptr), vp(self.d_out.at(15))))
        self._blind(self.pa, u, rng)
        self._blind(self.ps, u, rng)
        cm = self._commit([(BASIS_LAGRANGE, self.pa.ptr), (BASIS_LAGRANGE, self.ps.ptr)])
        if int(self.d_out.download(15, 1)[0, 0]) & 0xffffffff:
            raise H2BError(-5, "permute_expression_pair: an input value is not in the table (ConstraintSystemFailure)")
        self.d2h_bytes += 32
        res["commitments"] += list(cm); tr.absorb(cm)
        beta, gamma = tr.squeeze(), tr.squeeze()
        bl, gl = to_limbs(beta), to_limbs(gamma)
        mark("phase1 lookup permuted")
        side_transforms([(self.pa, self.pac, "pa"), (self.ps, self.psc, "ps")])
        # ---- product columns + the vanishing argument's random polynomial
        cols = (C.c_void_p * 2)(cs.lagr["c"].ptr, self.a.ptr)
        sig = (C.c_void_p * 2)(cs.lagr["sigma_c"].ptr, cs.lagr["sigma_a"].ptr)
        ctx.check(lib.h2b_permutation_product_dev(ctx.h, cols, sig, 2, 0, vp(bl.ctypes.data), vp(gl.ctypes.data), k, bf, None, vp(self.zp.ptr)))
        ctx.check(lib.h2b_lookup_product_dev(ctx.h, vp(self.inp.ptr), vp(cs.lagr["table"].ptr), vp(self.pa.ptr), vp(self.ps.ptr),
                                             vp(bl.ctypes.data), vp(gl.ctypes.data), k, bf, vp(self.zl.ptr)))
        self._blind(self.zp, u + 1, rng)
        self._blind(self.zl, u + 1, rng)
        side_transforms([(self.zp, self.zpc, "zp"), (self.zl, self.zlc, "zl")])  # beside the commitments below
        cm = self._commit([(BASIS_LAGRANGE, self.zp.ptr), (BASIS_LAGRANGE, self.zl.ptr), (BASIS_MONOMIAL, self.rnd.ptr)])
        res["commitments"] += list(cm); tr.absorb(cm)
        y = tr.squeeze()
        yl = to_limbs(y)
        mark("phase2 products+random")
        ctx.check(lib.h2b_ctx_side_join(ctx.h))  # all five columns are now in coefficient and extended form
        mark("transforms")
        # ---- quotient: gate, permutation and lookup terms folded with y on the extended coset
        kw = dict(beta=bl, gamma=gl, theta=to_limbs(theta), y=yl)
        ctx.check(lib.h2b_poly_zero(ctx.h, self.h.h))
        bg = ev.BoundGraph(cs.gate_graph, cs.gate_res, fixed=[cs.ext["q"].ptr], advice=[self.ext["a"].ptr], **kw)
        ctx.check(lib.h2b_quotient_graph_dev(ctx.h, C.byref(bg.struct), k, ext_k, vp(self.h.ptr)))
        tz = (C.c_void_p * 1)(self.ext["zp"].ptr)
        tc = (C.c_void_p * 2)(cs.ext["c"].ptr, self.ext["a"].ptr)
        ts = (C.c_void_p * 2)(cs.ext["sigma_c"].ptr, cs.ext["sigma_a"].ptr)
        ctx.check(lib.h2b_permutation_fold_dev(ctx.h, tz, 1, tc, ts, 2, cs.degree - 2, vp(cs.ext["l0"].ptr), vp(cs.ext["l_last"].ptr),
                                               vp(cs.ext["l_active"].ptr), vp(bl.ctypes.data), vp(gl.ctypes.data), vp(yl.ctypes.data), bf, k, ext_k,
                                               vp(self.h.ptr)))
        blk = ev.BoundGraph(cs.lk_graph, cs.lk_res, fixed=[cs.ext["q_lookup"].ptr, cs.ext["table"].ptr], advice=[self.ext["a"].ptr], **kw)
        ctx.check(lib.h2b_lookup_fold_dev(ctx.h, C.byref(blk.struct), vp(self.ext["zl"].ptr), vp(self.ext["pa"].ptr), vp(self.ext["ps"].ptr),
                                          vp(cs.ext["l0"].ptr), vp(cs.ext["l_last"].ptr), vp(cs.ext["l_active"].ptr), k, ext_k, vp(self.h.ptr)))
        ctx.check(lib.h2b_divide_by_vanishing_poly_dev(ctx.h, vp(self.h.ptr), k, ext_k))
        ctx.check(lib.h2b_extended_to_coeff_dev(ctx.h, vp(self.h.ptr), ext_k))
        mark("quotient")
        pieces = cs.degree - 1
        cm = self._commit([(BASIS_MONOMIAL, self.h.at(j * n)) for j in range(pieces)])
        res["commitments"] += list(cm); tr.absorb(cm)
        x = tr.squeeze()
        mark("phase3 h pieces")
        # ---- evaluations at x and its rotations
        w = pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)
        rot = lambda r: x * pow(w, r % n, R_MOD) % R_MOD
        last = -(bf + 1)
        queries = ([("a", self.ac.ptr, r) for r in (0, 1, 2, 3)]
                   + [(nm, cs.coeff[nm].ptr, 0) for nm in ("q", "q_lookup", "table", "c", "sigma_c", "sigma_a")]
                   + [("zp", self.zpc.ptr, r) for r in (0, 1, last)]
                   + [("pa", self.pac.ptr, 0), ("pa", self.pac.ptr, -1), ("ps", self.psc.ptr, 0)]
                   + [("zl", self.zlc.ptr, 0), ("zl", self.zlc.ptr, 1)]
                   + [("h%d" % j, self.h.at(j * n), 0) for j in range(pieces)] + [("rnd", self.rnd.ptr, 0)])
        m = len(queries)
        polys = (C.c_void_p * m)(*[p for _, p, _ in queries])
        xs = np.stack([to_limbs(rot(r)) for _, _, r in queries])
        ev_out = np.empty((m, 4), dtype=np.uint64)
        ctx.check(lib.h2b_eval_polynomial_batch_dev(ctx.h, polys, vp(xs.ctypes.data), m, n, vp(ev_out.ctypes.data)))
        self.d2h_bytes += m * 32
        tr.absorb(ev_out)
        res["evals"] = {(nm, r): ev_out[i] for i, (nm, _, r) in enumerate(queries)}
        res["challenges"] = dict(theta=theta, beta=beta, gamma=gamma, y=y, x=x)
        mark("evaluations")
        # ---- SHPLONK-shaped opening: per rotation set sum_i v^i p_i, divided by (X - point) for every point of the set
        v_ch, mu = tr.squeeze(), tr.squeeze()
        sets = [
            ([0], [cs.coeff[nm].ptr for nm in ("q", "q_lookup", "table", "c", "sigma_c", "sigma_a")] + [self.psc.ptr, self.rnd.ptr]
             + [self.h.at(j * n) for j in range(pieces)]),
            ([0, 1, 2, 3], [self.ac.ptr]),
            ([0, 1, last], [self.zpc.ptr]),
            ([0, -1], [self.pac.ptr]),
            ([0, 1], [self.zlc.ptr]),
        ]
        def run_sets(which, bufs):
            """sum over the given rotation sets of mu^s * (sum_i v^i p_i) / prod (X - point); result in bufs[2]"""
            f, qd, acc = bufs
            first = True
            for si in which:
                rots, plist = sets[si]
                mm = len(plist)
                pp = (C.c_void_p * mm)(*plist)
                sc = np.stack([to_limbs(pow(v_ch, i, R_MOD)) for i in range(mm)])
                ctx.check(lib.h2b_poly_lincomb_dev(ctx.h, pp, vp(sc.ctypes.data), mm, n, vp(f.ptr)))
                src, dst = f, qd
                for r in rots:  # successive divisions by (X - point): the quotient by the set's vanishing polynomial
                    z = to_limbs(rot(r))
                    ctx.check(lib.h2b_kate_division_dev(ctx.h, vp(src.ptr), n, vp(z.ctypes.data), vp(dst.ptr)))
                    src, dst = dst, src
                mu_s = to_limbs(pow(mu, si, R_MOD))
                if first:
                    p1 = (C.c_void_p * 1)(src.ptr)
                    ctx.check(lib.h2b_poly_lincomb_dev(ctx.h, p1, vp(np.stack([mu_s]).ctypes.data), 1, n, vp(acc.ptr)))
                    first = False
                else:
                    p2 = (C.c_void_p * 2)(acc.ptr, src.ptr)
                    ctx.check(lib.h2b_poly_lincomb_dev(ctx.h, p2, vp(np.stack([to_limbs(1), mu_s]).ctypes.data), 2, n, vp(acc.ptr)))

        # the rotation sets are independent: three of them on the side queue (own scratch), two on the main queue
        ctx.check(lib.h2b_ctx_side_begin(ctx.h))
        try:
            run_sets([0, 2, 4], self.tmp_side)
        finally:
            ctx.check(lib.h2b_ctx_side_end(ctx.h))
        run_sets([1, 3], self.tmp[:3])
        ctx.check(lib.h2b_ctx_side_join(ctx.h))
        p2 = (C.c_void_p * 2)(self.tmp[2].ptr, self.tmp_side[2].ptr)
        ones = np.stack([to_limbs(1), to_limbs(1)])
        ctx.check(lib.h2b_poly_lincomb_dev(ctx.h, p2, vp(ones.ctypes.data), 2, n, vp(self.tmp[2].ptr)))
        mark("shplonk arithmetic")
        cm = self._commit([(BASIS_MONOMIAL, self.tmp[2].ptr)])
        res["commitments"] += list(cm); tr.absorb(cm)
        u_ch = tr.squeeze()
        # final quotient: L(X) = h_spl-weighted combination, W' = L / (X - u) (the remainder is dropped by kate_division)
        ul = to_limbs(u_ch)
        ctx.check(lib.h2b_kate_division_dev(ctx.h, vp(self.tmp[2].ptr), n, vp(ul.ctypes.data), vp(self.tmp[3].ptr)))
        cm = self._commit([(BASIS_MONOMIAL, self.tmp[3].ptr)])
        res["commitments"] += list(cm)
        res["h2d_bytes"], res["d2h_bytes"] = self.h2d_bytes, self.d2h_bytes
        mark("phase4-5 openings")
        if trace is not None:
            import sys
            print("prover trace (ms): " + ", ".join("%s=%.2f" % (l, 1e3 * (t - trace[i][1])) for i, (l, t) in enumerate(trace[1:])), file=sys.stderr)
        return res

    def free(self):
        for p in [self.v, self.a, self.inp, self.pa, self.ps, self.zp, self.zl, self.rnd, self.h, self.d_out, self.ac, self.pac, self.psc, self.zpc, self.zlc] + self.tmp + self.tmp_side + list(self.ext.values()):
            p.free()
