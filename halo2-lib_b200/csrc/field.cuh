// field.cuh — BN254 Fq / Fr arithmetic for sm_100a: 254-bit Montgomery (R = 2^256), 8 x 32-bit limbs in registers.
//
// Memory layout is the `[u64;4]` little-endian Montgomery contract halo2-lib exposes
// (halo2-base/src/utils/mod.rs:332-377; halo2curves-axiom 0.7.3 bn256::{Fq,Fr}): on a little-endian
// machine 4 x u64 == 8 x u32, so host arrays are consumed as two 128-bit loads per element.
//
// The multiplier is written for the Blackwell integer pipe: every (mad.lo.cc, madc.hi.cc) pair on the same
// multiplicands is fused by ptxas into ONE `IMAD.WIDE.U32[.X]` with predicate carry, so a product row is
// 4 wide-MADs for the even limbs of `a` + 4 for the odd limbs.  Even-limb and odd-limb partial products are
// kept in two accumulators whose 64-bit register pairs never move; the Montgomery shift by one limb per
// row is absorbed by swapping the roles of the two accumulators (see mont_mul below).
// 16 IMAD.WIDE + 4 adds per row, 8 rows.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace h2b {

typedef uint32_t u32;
typedef uint64_t u64;

// ---------------------------------------------------------------- carry-chain primitives
// PTX has one carry flag; nvcc never emits .cc instructions itself, and volatile asm statements keep
// their relative order, so chains may be composed from these one-instruction helpers.
#define H2B_ASM_3(name, op)                                                                        \
    __device__ __forceinline__ void name(u32& acc, u32 a, u32 b) {                                 \
        asm volatile(op " %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));                       \
    }
H2B_ASM_3(mad_lo_cc, "mad.lo.cc.u32")
H2B_ASM_3(madc_lo_cc, "madc.lo.cc.u32")
H2B_ASM_3(madc_hi_cc, "madc.hi.cc.u32")
H2B_ASM_3(madc_hi, "madc.hi.u32")
#undef H2B_ASM_3
__device__ __forceinline__ void mul_wide(u32& lo, u32& hi, u32 a, u32 b) {
    asm volatile("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ void add_cc(u32& r, u32 a, u32 b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ __forceinline__ void addc_cc(u32& r, u32 a, u32 b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ __forceinline__ void addc(u32& r, u32 a, u32 b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ __forceinline__ void sub_cc(u32& r, u32 a, u32 b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ __forceinline__ void subc_cc(u32& r, u32 a, u32 b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ __forceinline__ void subc(u32& r, u32 a, u32 b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }

// ---------------------------------------------------------------- field parameters
struct FqParams {
    __host__ __device__ static constexpr u32 MOD(int i) {
        constexpr u32 v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static constexpr u32 INV = 0xe4866389u;  // -p^-1 mod 2^32
    __host__ __device__ static constexpr u32 ONE(int i) {
        constexpr u32 v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }  // R mod p
    __host__ __device__ static constexpr u32 R2(int i) {
        constexpr u32 v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
};
struct FrParams {
    __host__ __device__ static constexpr u32 MOD(int i) {
        constexpr u32 v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static constexpr u32 INV = 0xefffffffu;  // -r^-1 mod 2^32
    __host__ __device__ static constexpr u32 ONE(int i) {
        constexpr u32 v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }  // R mod r
    __host__ __device__ static constexpr u32 R2(int i) {
        constexpr u32 v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
};

template <class P>
struct __align__(16) Fp {
    u32 l[8];

    __device__ __forceinline__ static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    __device__ __forceinline__ static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::ONE(i);
        return r;
    }
    __device__ __forceinline__ static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::R2(i);
        return r;
    }
    // two 128-bit loads / stores (pointer must be 16-byte aligned: element stride is 32 B)
    __device__ __forceinline__ static Fp load(const void* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = q[0], b = q[1];
        Fp r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ static Fp load_nc(const void* p) {  // read-only path
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        Fp r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(l[0], l[1], l[2], l[3]);
        q[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
    __device__ __forceinline__ bool is_zero() const {
        return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& o) const {
        u32 d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= l[i] ^ o.l[i];
        return d == 0;
    }

    // r = a - p if a >= p else a   (a < 2p)
    __device__ __forceinline__ static Fp reduce_once(const Fp& a) {
        Fp t;
        u32 borrow;
        sub_cc(t.l[0], a.l[0], P::MOD(0));
#pragma unroll
        for (int i = 1; i < 8; i++) subc_cc(t.l[i], a.l[i], P::MOD(i));
        subc(borrow, 0, 0);  // 0xffffffff if a < p
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }
    __device__ __forceinline__ friend Fp operator+(const Fp& a, const Fp& b) {
        Fp s;
        add_cc(s.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) addc_cc(s.l[i], a.l[i], b.l[i]);
        addc(s.l[7], a.l[7], b.l[7]);  // p < 2^254: a + b < 2^255, no carry out
        return reduce_once(s);
    }
    __device__ __forceinline__ friend Fp operator-(const Fp& a, const Fp& b) {
        Fp d;
        u32 borrow;
        sub_cc(d.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) subc_cc(d.l[i], a.l[i], b.l[i]);
        subc(borrow, 0, 0);  // all-ones if a < b
        Fp r;
        add_cc(r.l[0], d.l[0], P::MOD(0) & borrow);
#pragma unroll
        for (int i = 1; i < 7; i++) addc_cc(r.l[i], d.l[i], P::MOD(i) & borrow);
        addc(r.l[7], d.l[7], P::MOD(7) & borrow);
        return r;
    }
    __device__ __forceinline__ Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod p, fully reduced.
    //
    // E[k] / O[k] hold the limb at absolute position k of the even-start / odd-start accumulator: partial
    // product a_j*b_i lives at positions (i+j, i+j+1), so for a fixed row i the even-j products tile one
    // accumulator and the odd-j products the other, each as ONE carry chain of 4 wide MADs.  In row i the
    // accumulator whose pairs start at position i ("X") owns the limb the reduction must clear; the other
    // ("Y") still holds one live limb at position i (top half of its consumed lowest pair) which is folded
    // into X[i], the carry of that add entering the Y chain at position i+1 — exactly where it belongs.
    // Bounds: running total < 2p before a row and < 2^288 * 2^(32 i) after the products, so the chains
    // that would carry into position i+9 cannot, and the final sum is < 2p.
#ifdef H2B_MUL_NOINLINE
    __device__ __noinline__ friend Fp operator*(const Fp& a, const Fp& b) {
#else
    __device__ __forceinline__ friend Fp operator*(const Fp& a, const Fp& b) {
#endif
        u32 E[17], O[17];
#pragma unroll
        for (int k = 0; k < 17; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u32* X = (i & 1) ? O : E;
            u32* Y = (i & 1) ? E : O;
            const u32 bi = b.l[i];
            if (i == 0) {
                // fresh accumulators: plain wide multiplies, no carries
                mul_wide(Y[1], Y[2], a.l[1], bi);
                mul_wide(Y[3], Y[4], a.l[3], bi);
                mul_wide(Y[5], Y[6], a.l[5], bi);
                mul_wide(Y[7], Y[8], a.l[7], bi);
                mul_wide(X[0], X[1], a.l[0], bi);
                mul_wide(X[2], X[3], a.l[2], bi);
                mul_wide(X[4], X[5], a.l[4], bi);
                mul_wide(X[6], X[7], a.l[6], bi);
            } else {
                // chain 1: fold Y's live limb into X[i]; odd-j products into Y at (i+1 .. i+8)
                add_cc(X[i], X[i], Y[i]);
                madc_lo_cc(Y[i + 1], a.l[1], bi); madc_hi_cc(Y[i + 2], a.l[1], bi);
                madc_lo_cc(Y[i + 3], a.l[3], bi); madc_hi_cc(Y[i + 4], a.l[3], bi);
                madc_lo_cc(Y[i + 5], a.l[5], bi); madc_hi_cc(Y[i + 6], a.l[5], bi);
                madc_lo_cc(Y[i + 7], a.l[7], bi); madc_hi(Y[i + 8], a.l[7], bi);
                // chain 2: even-j products into X at (i .. i+7), carry limb X[i+8]
                mad_lo_cc(X[i], a.l[0], bi);      madc_hi_cc(X[i + 1], a.l[0], bi);
                madc_lo_cc(X[i + 2], a.l[2], bi); madc_hi_cc(X[i + 3], a.l[2], bi);
                madc_lo_cc(X[i + 4], a.l[4], bi); madc_hi_cc(X[i + 5], a.l[4], bi);
                madc_lo_cc(X[i + 6], a.l[6], bi); madc_hi_cc(X[i + 7], a.l[6], bi);
                addc(X[i + 8], X[i + 8], 0);
            }
            const u32 m = X[i] * P::INV;
            // chain 3: m * (p1,p3,p5,p7) into Y
            mad_lo_cc(Y[i + 1], m, P::MOD(1));  madc_hi_cc(Y[i + 2], m, P::MOD(1));
            madc_lo_cc(Y[i + 3], m, P::MOD(3)); madc_hi_cc(Y[i + 4], m, P::MOD(3));
            madc_lo_cc(Y[i + 5], m, P::MOD(5)); madc_hi_cc(Y[i + 6], m, P::MOD(5));
            madc_lo_cc(Y[i + 7], m, P::MOD(7)); madc_hi(Y[i + 8], m, P::MOD(7));
            // chain 4: m * (p0,p2,p4,p6) into X; X[i] becomes 0 and is dropped
            mad_lo_cc(X[i], m, P::MOD(0));      madc_hi_cc(X[i + 1], m, P::MOD(0));
            madc_lo_cc(X[i + 2], m, P::MOD(2)); madc_hi_cc(X[i + 3], m, P::MOD(2));
            madc_lo_cc(X[i + 4], m, P::MOD(4)); madc_hi_cc(X[i + 5], m, P::MOD(4));
            madc_lo_cc(X[i + 6], m, P::MOD(6)); madc_hi_cc(X[i + 7], m, P::MOD(6));
            addc(X[i + 8], X[i + 8], 0);
        }
        // both accumulators are live at positions 8..15
        Fp s;
        add_cc(s.l[0], E[8], O[8]);
#pragma unroll
        for (int k = 1; k < 7; k++) addc_cc(s.l[k], E[8 + k], O[8 + k]);
        addc(s.l[7], E[15], O[15]);
        return reduce_once(s);
    }
    __device__ __forceinline__ Fp sqr() const { return (*this) * (*this); }

    // Montgomery form -> canonical integer (what `to_repr()` yields; best_multiexp slices these bits)
    __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero();
        o.l[0] = 1;
        return (*this) * o;
    }
    __device__ __forceinline__ Fp to_mont() const { return (*this) * r2(); }

    // a^(p-2); inv(0) = 0.  Cold path (one call per MSM / per batch-inversion block).
    __device__ __noinline__ Fp inv() const {
        Fp acc = one(), base = *this;
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            u32 e = P::MOD(i);
            if (i == 0) e -= 2;  // MOD[0] >= 2 for both fields: no borrow
#pragma unroll 1
            for (int bit = 0; bit < 32; bit++) {
                if ((e >> bit) & 1) acc = acc * base;
                base = base.sqr();
            }
        }
        return acc;
    }
};

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

}  // namespace h2b
