// batch_affine.cuh — pairwise AFFINE additions with one shared inversion per CTA (Montgomery's trick), the first stage
// of bucket accumulation (msm.cu).
//
// After the counting sort every bucket's run of entries is padded to a multiple of G = 2^R slots, so the sorted array
// splits into aligned groups of G entries that all belong to ONE bucket.  R passes of this kernel halve the array:
// pass r adds the points of every aligned pair (2p, 2p+1) in affine coordinates,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// which costs 5 products + 1 squaring per addition once the inversion is shared (an XYZZ mixed addition costs the
// equivalent of 9.06 products).  After R passes every group is ONE affine point and k_accumulate only has M / G
// mixed additions left.
//
// One CTA = BA_T threads x K pairs.  Forward sweep: every thread multiplies the denominators of its K pairs into a
// running prefix product kept in shared memory (only the x coordinates are read).  The BA_T thread totals are
// combined by a product tree in shared memory, ONE lane inverts the root with the product-free binary Euclid
// (Fp::inv_bgcd, ~35 us of add/shift work that does not occupy the multiplier pipe the other resident CTAs are using),
// the tree is walked back down to per-thread inverses, and the backward sweep turns prefix products into the
// individual inverses, finishes the additions and stores the sums (coalesced 64-byte records).
//
// Completeness: padding slots and identity bases (0,0) are copied through; P + P uses the tangent (3 x^2 / 2 y);
// P + (-P) yields the identity.  Those pairs contribute the factor 1 to the shared product.
#pragma once
#include "curve.cuh"

namespace h2b {

static constexpr u32 BA_PAD = 0xFFFFFFFFu;  // padding slot of the sorted entry array (never a valid index | sign word)
static constexpr int BA_T = 128;            // threads per CTA
static constexpr u32 BA_SIGN = 0x80000000u;

__device__ __forceinline__ Fq ba_smem_load(const u32* base, int stride) {
    Fq r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.l[l] = base[l * stride];
    return r;
}
__device__ __forceinline__ void ba_smem_store(u32* base, int stride, const Fq& v) {
#pragma unroll
    for (int l = 0; l < 8; l++) base[l * stride] = v.l[l];
}

// kind of a pair: 0 = chord addition, 1 = tangent (P == Q), 2 = trivial (an operand is the identity, or P == -Q)
struct BaPair {
    Affine p, q;
    int kind;
    Fq den;
};

template <bool LEVEL1>
__device__ __forceinline__ void ba_load_x(const u32* __restrict__ vals, const Affine* __restrict__ table,
                                          const Affine* __restrict__ in, u32 pair, bool live, BaPair& r, u32& va, u32& vb) {
    // x coordinates only (one 32-byte sector per point); y is fetched on the rare paths that need it
    r.p.x = Fq::zero(); r.q.x = Fq::zero();
    bool pid = true, qid = true;
    va = BA_PAD; vb = BA_PAD;
    if (live) {
        if (LEVEL1) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(vals) + pair);
            va = v.x; vb = v.y;
            if (va != BA_PAD) { r.p.x = Fq::load_nc(table + (va & ~BA_SIGN)); pid = false; }
            if (vb != BA_PAD) { r.q.x = Fq::load_nc(table + (vb & ~BA_SIGN)); qid = false; }
        } else {
            r.p.x = Fq::load_nc(in + 2 * (size_t)pair);
            r.q.x = Fq::load_nc(in + 2 * (size_t)pair + 1);
            pid = qid = false;
        }
    }
    // x == 0 may be the identity (0,0): look at y
    auto y_of = [&](bool first) -> Fq {
        const Affine* src = LEVEL1 ? table + ((first ? va : vb) & ~BA_SIGN) : in + 2 * (size_t)pair + (first ? 0 : 1);
        return Fq::load_nc(reinterpret_cast<const char*>(src) + 32);
    };
    if (!pid && r.p.x.is_zero()) pid = y_of(true).is_zero();
    if (!qid && r.q.x.is_zero()) qid = y_of(false).is_zero();
    if (pid || qid) { r.kind = 2; return; }
    if (!(r.p.x == r.q.x)) { r.kind = 0; r.den = r.q.x - r.p.x; return; }
    Fq yp = y_of(true), yq = y_of(false);
    if (LEVEL1) {
        if (va & BA_SIGN) yp = yp.neg();
        if (vb & BA_SIGN) yq = yq.neg();
    }
    if (yp == yq && !yp.is_zero()) { r.kind = 1; r.den = yp.dbl(); return; }
    r.kind = 2;
}

template <bool LEVEL1>
__device__ __forceinline__ void ba_load_full(const u32* __restrict__ vals, const Affine* __restrict__ table,
                                             const Affine* __restrict__ in, u32 pair, bool live, BaPair& r) {
    const Affine idn = {Fq::zero(), Fq::zero()};
    r.p = idn; r.q = idn;
    if (live) {
        if (LEVEL1) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(vals) + pair);
            if (v.x != BA_PAD) { r.p = Affine::load(table + (v.x & ~BA_SIGN)); if (v.x & BA_SIGN) r.p.y = r.p.y.neg(); }
            if (v.y != BA_PAD) { r.q = Affine::load(table + (v.y & ~BA_SIGN)); if (v.y & BA_SIGN) r.q.y = r.q.y.neg(); }
        } else {
            r.p = Affine::load(in + 2 * (size_t)pair);
            r.q = Affine::load(in + 2 * (size_t)pair + 1);
        }
    }
    if (r.p.is_identity() || r.q.is_identity()) { r.kind = 2; return; }
    if (!(r.p.x == r.q.x)) { r.kind = 0; r.den = r.q.x - r.p.x; return; }
    if (r.p.y == r.q.y && !r.p.y.is_zero()) { r.kind = 1; r.den = r.p.y.dbl(); return; }
    r.kind = 2;
}

// n_entries_ptr: number of (padded) sorted entries M' on the device; this pass handles M' >> level pairs.
// shared memory: K * 8 * BA_T words of prefix products + 2 * BA_T * 8 words for the product tree (inverted in place).
template <bool LEVEL1, int K>
__global__ void __launch_bounds__(BA_T, 3) k_batch_affine(const u32* __restrict__ vals, const Affine* __restrict__ table,
                                                          const Affine* __restrict__ in, Affine* __restrict__ out,
                                                          const u32* __restrict__ n_entries_ptr, int level) {
    extern __shared__ u32 ba_sh[];
    u32* pref = ba_sh;                          // [K][8][BA_T]
    u32* node = ba_sh + K * 8 * BA_T;           // [8][2 * BA_T]  heap order: leaves at BA_T + tid, root at 1
    const u32 npairs = __ldg(n_entries_ptr) >> level;
    const u32 base = blockIdx.x * (u32)(BA_T * K);
    if (base >= npairs) return;
    const int tid = threadIdx.x;

    // ---- forward sweep: prefix products of the denominators of this thread's pairs
    Fq run = Fq::one();
#pragma unroll 1
    for (int i = 0; i < K; i++) {
        const u32 pair = base + (u32)i * BA_T + tid;
        BaPair pr;
        u32 va, vb;
        ba_load_x<LEVEL1>(vals, table, in, pair, pair < npairs, pr, va, vb);
        if (pr.kind != 2) run = run * pr.den;
        ba_smem_store(pref + (size_t)i * 8 * BA_T + tid, BA_T, run);
    }
    ba_smem_store(node + BA_T + tid, 2 * BA_T, run);
    __syncthreads();
    // ---- product tree of the thread totals, one inversion, back down to per-thread inverses
#pragma unroll 1
    for (int s = BA_T / 2; s >= 1; s >>= 1) {
        if (tid < s) {
            const int j = s + tid;
            const Fq a = ba_smem_load(node + 2 * j, 2 * BA_T), b = ba_smem_load(node + 2 * j + 1, 2 * BA_T);
            ba_smem_store(node + j, 2 * BA_T, a * b);
        }
        __syncthreads();
    }
    if (tid == 0) ba_smem_store(node + 1, 2 * BA_T, ba_smem_load(node + 1, 2 * BA_T).inv_bgcd());
    __syncthreads();
    // in place: node j already holds the inverse of its subtree product; its children (still products) are replaced by
    // their inverses 1/l = (1/(l r)) r, 1/r = (1/(l r)) l — each thread touches only its own three nodes
#pragma unroll 1
    for (int s = 1; s < BA_T; s <<= 1) {
        if (tid < s) {
            const int j = s + tid;
            const Fq a = ba_smem_load(node + j, 2 * BA_T);
            const Fq l = ba_smem_load(node + 2 * j, 2 * BA_T), r = ba_smem_load(node + 2 * j + 1, 2 * BA_T);
            ba_smem_store(node + 2 * j, 2 * BA_T, a * r);
            ba_smem_store(node + 2 * j + 1, 2 * BA_T, a * l);
        }
        __syncthreads();
    }
    // ---- backward sweep: individual inverses, the additions, coalesced stores
    Fq inv_run = ba_smem_load(node + BA_T + tid, 2 * BA_T);  // 1 / (product of all denominators of this thread)
#pragma unroll 1
    for (int i = K - 1; i >= 0; i--) {
        const u32 pair = base + (u32)i * BA_T + tid;
        const bool live = pair < npairs;
        BaPair pr;
        ba_load_full<LEVEL1>(vals, table, in, pair, live, pr);
        Affine sum;
        if (pr.kind == 2) {
            // identity operand -> the other one; P + (-P) (both non-identity) -> identity
            if (pr.p.is_identity()) sum = pr.q;
            else if (pr.q.is_identity()) sum = pr.p;
            else { sum.x = Fq::zero(); sum.y = Fq::zero(); }
        } else {
            Fq inv = inv_run;
            if (i > 0) inv = inv * ba_smem_load(pref + (size_t)(i - 1) * 8 * BA_T + tid, BA_T);
            inv_run = inv_run * pr.den;
            Fq num;
            if (pr.kind == 0) num = pr.q.y - pr.p.y;
            else { const Fq xx = pr.p.x.sqr(); num = xx.dbl() + xx; }
            const Fq lam = num * inv;
            sum.x = lam.sqr() - pr.p.x - pr.q.x;
            sum.y = lam * (pr.p.x - sum.x) - pr.p.y;
        }
        if (live) sum.store(out + pair);
    }
}

}  // namespace h2b
