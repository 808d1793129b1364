// curve.cuh — BN254 G1 (y^2 = x^3 + 3) group law for the MSM kernels.
//
// Bases arrive as halo2curves `G1Affine` (x||y Montgomery, identity = (0,0)); results leave as `G1`
// (Jacobian x||y||z; identity = (0,1,0)) — the types `best_multiexp(coeffs, bases) -> C::Curve` uses
// (halo2curves-axiom 0.7.3; reached from ParamsKZG::commit / commit_lagrange).
// Internally buckets are XYZZ (x = X/ZZ, y = Y/ZZZ): a mixed add is 8M + 2S, no inversion.
// All formulas are complete in the sense MSM needs: identity operands, P + P and P + (-P) are handled
// (repeated bases and sums to infinity are edge vectors of halo2-ecc/src/bn254/tests/msm_sum_infinity.rs).
#pragma once
#include "field.cuh"

namespace h2b {

struct Affine {
    Fq x, y;
    __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
    __device__ __forceinline__ static Affine load(const void* p) {  // 64-byte record, four 128-bit loads
        Affine a;
        a.x = Fq::load_nc(p);
        a.y = Fq::load_nc(reinterpret_cast<const char*>(p) + 32);
        return a;
    }
    __device__ __forceinline__ void store(void* p) const {
        x.store(p);
        y.store(reinterpret_cast<char*>(p) + 32);
    }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ identity() {
        XYZZ r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    __device__ __forceinline__ static XYZZ from_affine(const Affine& a) {
        XYZZ r;
        if (a.is_identity()) return identity();
        r.x = a.x; r.y = a.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }
    __device__ __forceinline__ static XYZZ load(const void* p) {
        const char* c = reinterpret_cast<const char*>(p);
        XYZZ r;
        r.x = Fq::load(c); r.y = Fq::load(c + 32); r.zz = Fq::load(c + 64); r.zzz = Fq::load(c + 96);
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        char* c = reinterpret_cast<char*>(p);
        x.store(c); y.store(c + 32); zz.store(c + 64); zzz.store(c + 96);
    }
    __device__ __forceinline__ XYZZ neg() const {
        XYZZ r = *this;
        r.y = y.neg();
        return r;
    }
};

// 2 * affine (never the identity unless y = 0, which cannot happen on a prime-order curve) — mdbl-2008-s-1
__device__ __forceinline__ XYZZ xyzz_dbl_affine(const Affine& p) {
    XYZZ r;
    Fq u = p.y.dbl();
    Fq v = u.sqr();
    Fq w = u * v;
    Fq s = p.x * v;
    Fq xx = p.x.sqr();
    Fq m = xx.dbl() + xx;
    r.x = m.sqr() - s.dbl();
    r.y = Fq::mul_sub_mul(m, s - r.x, w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * XYZZ — dbl-2008-s-1 (a = 0)
__device__ __forceinline__ XYZZ xyzz_dbl(const XYZZ& p) {
    if (p.is_identity()) return p;
    XYZZ r;
    Fq u = p.y.dbl();
    Fq v = u.sqr();
    Fq w = u * v;
    Fq s = p.x * v;
    Fq xx = p.x.sqr();
    Fq m = xx.dbl() + xx;
    r.x = m.sqr() - s.dbl();
    r.y = Fq::mul_sub_mul(m, s - r.x, w, p.y);
    r.zz = v * p.zz;
    r.zzz = w * p.zzz;
    return r;
}

// acc += q (affine, non-identity, acc non-identity in the common path) — madd-2008-s, 8M + 2S
__device__ __forceinline__ void xyzz_madd(XYZZ& acc, const Affine& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) {
        acc = XYZZ::from_affine(q);
        return;
    }
    Fq u2 = q.x * acc.zz;
    Fq s2 = q.y * acc.zzz;
    Fq p = u2 - acc.x;
    Fq r = s2 - acc.y;
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl_affine(q);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = p.sqr();
    Fq ppp = p * pp;
    Fq qq = acc.x * pp;
    Fq x3 = r.sqr() - ppp - qq.dbl();
    acc.y = Fq::mul_sub_mul(r, qq - x3, acc.y, ppp);  // one interleaved reduction for both products
    acc.x = x3;
    acc.zz = acc.zz * pp;
    acc.zzz = acc.zzz * ppp;
}

// acc += q (XYZZ) — add-2008-s, 12M + 2S
__device__ __forceinline__ void xyzz_add(XYZZ& acc, const XYZZ& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) {
        acc = q;
        return;
    }
    Fq u1 = acc.x * q.zz;
    Fq u2 = q.x * acc.zz;
    Fq s1 = acc.y * q.zzz;
    Fq s2 = q.y * acc.zzz;
    Fq p = u2 - u1;
    Fq r = s2 - s1;
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl(acc);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = p.sqr();
    Fq ppp = p * pp;
    Fq qq = u1 * pp;
    Fq x3 = r.sqr() - ppp - qq.dbl();
    acc.y = Fq::mul_sub_mul(r, qq - x3, s1, ppp);
    acc.x = x3;
    acc.zz = acc.zz * q.zz * pp;
    acc.zzz = acc.zzz * q.zzz * ppp;
}

// XYZZ -> canonical Jacobian written as 12 x u64: identity -> (0, R, 0) (halo2curves G1::identity()),
// otherwise the affine-normalised representative (x, y, R).  One field inversion; cold path.
static __device__ __noinline__ void xyzz_store_jacobian_normalised(const XYZZ& p, void* out) {
    char* c = reinterpret_cast<char*>(out);
    if (p.is_identity()) {
        Fq::zero().store(c);
        Fq::one().store(c + 32);
        Fq::zero().store(c + 64);
        return;
    }
    Fq t = (p.zz * p.zzz).inv();   // 1 / (ZZ * ZZZ)
    Fq izz = t * p.zzz;            // 1 / ZZ
    Fq izzz = t * p.zz;            // 1 / ZZZ
    (p.x * izz).store(c);
    (p.y * izzz).store(c + 32);
    Fq::one().store(c + 64);
}
static __device__ __noinline__ Affine xyzz_to_affine(const XYZZ& p) {
    Affine a;
    if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); return a; }
    Fq t = (p.zz * p.zzz).inv();
    a.x = p.x * (t * p.zzz);
    a.y = p.y * (t * p.zz);
    return a;
}
// Jacobian (x,y,z) -> XYZZ: ZZ = z^2, ZZZ = z^3
__device__ __forceinline__ XYZZ xyzz_from_jacobian(const void* in) {
    const char* c = reinterpret_cast<const char*>(in);
    XYZZ r;
    Fq z = Fq::load(c + 64);
    if (z.is_zero()) return XYZZ::identity();
    r.x = Fq::load(c);
    r.y = Fq::load(c + 32);
    r.zz = z.sqr();
    r.zzz = r.zz * z;
    return r;
}

}  // namespace h2b
