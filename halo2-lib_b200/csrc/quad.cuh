// quad.cuh — G1 point operations executed by a QUAD of 4 adjacent lanes.
//
// The MSM tail (bucket reduction) is a dependency chain of a few dozen point operations on very little data:
// it is bound by the latency of one warp on the integer-multiplier pipe (one Montgomery product ≈ 834 cycles for
// a lone warp, 14 products per XYZZ addition).  Here the 14 (9) products of an addition (doubling) are scheduled
// as 4 (3) rounds of 4 independent products, one per lane of the quad, and exchanged with width-4 shuffles: the
// latency of a point operation drops from 14 (9) product latencies to 4 (3).
//
// Convention: on entry all 4 lanes of the quad hold IDENTICAL copies of the operands; on exit all 4 lanes hold
// identical copies of the result.  Different quads of a warp may diverge (every shuffle uses the quad's own mask).
#pragma once
#include "curve.cuh"

namespace h2b {

__device__ __forceinline__ unsigned quad_mask() { return 0xFu << (threadIdx.x & 28); }
__device__ __forceinline__ int quad_role() { return threadIdx.x & 3; }

__device__ __forceinline__ Fq quad_bcast(const Fq& v, int src_role) {
    const unsigned m = quad_mask();
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(m, v.l[i], src_role, 4);
    return r;
}
__device__ __forceinline__ Fq sel4(int role, const Fq& a, const Fq& b, const Fq& c, const Fq& d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u32 lo = (role & 1) ? b.l[i] : a.l[i];
        u32 hi = (role & 1) ? d.l[i] : c.l[i];
        r.l[i] = (role & 2) ? hi : lo;
    }
    return r;
}

// acc = 2 * acc   (dbl-2008-s-1, a = 0): 3 product rounds
__device__ __forceinline__ void quad_dbl(XYZZ& p) {
    if (p.is_identity()) return;
    const int role = quad_role();
    const Fq u = p.y.dbl();
    // round 1: V = U^2 | XX = X^2
    Fq a = sel4(role, u, p.x, u, p.x);
    Fq m = a * a;
    const Fq v = quad_bcast(m, 0), xx = quad_bcast(m, 1);
    const Fq mm3 = xx.dbl() + xx;  // M = 3 X^2
    // round 2: W = U*V | S = X*V | MM = M^2
    a = sel4(role, u, p.x, mm3, mm3);
    Fq b = sel4(role, v, v, mm3, mm3);
    m = a * b;
    const Fq w = quad_bcast(m, 0), s = quad_bcast(m, 1), msq = quad_bcast(m, 2);
    const Fq x3 = msq - s.dbl();
    // round 3: M*(S - X3) | W*Y | V*ZZ | W*ZZZ
    a = sel4(role, mm3, w, v, w);
    b = sel4(role, s - x3, p.y, p.zz, p.zzz);
    m = a * b;
    const Fq t0 = quad_bcast(m, 0), t1 = quad_bcast(m, 1);
    p.zz = quad_bcast(m, 2);
    p.zzz = quad_bcast(m, 3);
    p.x = x3;
    p.y = t0 - t1;
}

// acc += q   (add-2008-s): 4 product rounds; complete (identity operands, P + P, P + (-P))
__device__ __forceinline__ void quad_add(XYZZ& acc, const XYZZ& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) { acc = q; return; }
    const int role = quad_role();
    // round 1: U1 = X1*ZZ2 | U2 = X2*ZZ1 | S1 = Y1*ZZZ2 | S2 = Y2*ZZZ1
    Fq a = sel4(role, acc.x, q.x, acc.y, q.y);
    Fq b = sel4(role, q.zz, acc.zz, q.zzz, acc.zzz);
    Fq m = a * b;
    const Fq u1 = quad_bcast(m, 0), u2 = quad_bcast(m, 1), s1 = quad_bcast(m, 2), s2 = quad_bcast(m, 3);
    const Fq p = u2 - u1, r = s2 - s1;
    if (p.is_zero()) {  // same x: doubling or cancellation (uniform across the quad: operands are replicated)
        if (r.is_zero()) quad_dbl(acc);
        else acc = XYZZ::identity();
        return;
    }
    // round 2: PP = P^2 | RR = R^2 | ZZ1*ZZ2 | ZZZ1*ZZZ2
    a = sel4(role, p, r, acc.zz, acc.zzz);
    b = sel4(role, p, r, q.zz, q.zzz);
    m = a * b;
    const Fq pp = quad_bcast(m, 0), rr = quad_bcast(m, 1), zzm = quad_bcast(m, 2), zzzm = quad_bcast(m, 3);
    // round 3: PPP = P*PP | Q = U1*PP | ZZ3 = ZZm*PP
    a = sel4(role, p, u1, zzm, zzm);
    m = a * pp;
    const Fq ppp = quad_bcast(m, 0), qq = quad_bcast(m, 1);
    acc.zz = quad_bcast(m, 2);
    const Fq x3 = rr - ppp - qq.dbl();
    // round 4: R*(Q - X3) | S1*PPP | ZZZ3 = ZZZm*PPP
    a = sel4(role, r, s1, zzzm, zzzm);
    b = sel4(role, qq - x3, ppp, ppp, ppp);
    m = a * b;
    const Fq t0 = quad_bcast(m, 0), t1 = quad_bcast(m, 1);
    acc.zzz = quad_bcast(m, 2);
    acc.x = x3;
    acc.y = t0 - t1;
}

// Out-of-line copies for code that runs only a handful of times per launch (block sums, the final combine): every
// inlined call site would be ~1200 cold instructions that a lone warp has to pull through the instruction cache
// (k_weighted_final: 123 us inlined, 87 us with these); hot loops keep the inlined forms above.
static __device__ __noinline__ void quad_add_nl(XYZZ& acc, const XYZZ& q) { quad_add(acc, q); }
static __device__ __noinline__ void quad_dbl_nl(XYZZ& p) { quad_dbl(p); }

// value of the quad `delta` quads further up the warp (all 32 lanes must call; operands replicated per quad)
__device__ __forceinline__ XYZZ quad_shfl_down(const XYZZ& v, int delta_quads) {
    XYZZ r;
    const int d = delta_quads * 4;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], d);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], d);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], d);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], d);
    }
    return r;
}
// sum over the 8 quads of a warp; valid in quad 0 (lanes 0..3).  All lanes must call.
// INL = true inlines the point operations: right for kernels with hundreds of CTAs running the same code
// (k_rowcol_sums: 90 us inlined, 127 us through the out-of-line copies); the lone-warp kernels use INL = false.
template <bool INL = false>
__device__ __forceinline__ XYZZ quad_warp_sum(XYZZ v) {
#pragma unroll 1
    for (int dq = 4; dq >= 1; dq >>= 1) {
        __syncwarp();
        XYZZ o = quad_shfl_down(v, dq);
        if (INL) quad_add(v, o);
        else quad_add_nl(v, o);
    }
    __syncwarp();
    return v;
}
// sum over all quads of the CTA (blockDim.x multiple of 32, <= 1024); valid in quad 0 of warp 0.
// sh must hold one XYZZ per warp.
template <bool INL = false>
__device__ __forceinline__ XYZZ quad_block_sum(XYZZ v, XYZZ* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = quad_warp_sum<INL>(v);
    __syncthreads();  // sh may still be read from a previous call
    if (lane == 0) v.store(sh + wid);
    __syncthreads();
    XYZZ r = XYZZ::identity();
    if (wid == 0) {
        // quad k of warp 0 takes the results of warps k, k+8, k+16, k+24 (serial), then the 8 quads are summed
        const int qid = lane >> 2;
        for (int w = qid; w < nw; w += 8) {
            if (INL) quad_add(r, XYZZ::load(sh + w));
            else quad_add_nl(r, XYZZ::load(sh + w));
        }
        r = quad_warp_sum<INL>(r);
    }
    return r;
}

// k * p by left-to-right double-and-add, k < 2^bits
template <bool INL = false>
__device__ __forceinline__ XYZZ quad_small_mul(const XYZZ& p, u32 k, int bits) {
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (int b = bits - 1; b >= 0; b--) {
        if (INL) {
            quad_dbl(acc);
            if ((k >> b) & 1) quad_add(acc, p);
        } else {
            quad_dbl_nl(acc);
            if ((k >> b) & 1) quad_add_nl(acc, p);
        }
    }
    return acc;
}

}  // namespace h2b
