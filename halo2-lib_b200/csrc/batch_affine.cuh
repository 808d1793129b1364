// batch_affine.cuh — pairwise AFFINE additions with one shared inversion per CTA (Montgomery's trick), the first stage
// of bucket accumulation (msm.cu).
//
// After the counting sort every bucket's run of entries is padded to a multiple of G = 2^R slots, so the sorted array
// splits into aligned groups of G entries that all belong to ONE bucket.  R levels halve the array: level l adds the
// points of every aligned pair (2p, 2p+1) in affine coordinates,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// which costs 5 products + 1 squaring per addition once the inversion is shared (an XYZZ mixed addition costs the
// equivalent of 9.06 products).  After R levels every group is ONE affine point and k_accumulate only has M / G
// mixed additions left.
//
// One launch, persistent CTAs.  A CTA repeatedly claims a TILE of BA_T x k aligned level-1 pairs (an atomic cursor hands
// out contiguous ranges; the first tile of a CTA is shortened by blockIdx so that the CTAs of an SM do not reach their
// inversions in lock step, and tiles shrink towards the end of the array so that all CTAs finish together) and takes it
// through all R levels (level l has BA_T x k / 2^(l-1) pairs; the sums of one level are read back by the same CTA for
// the next one, so they mostly stay in L2).  Per level:
//   forward sweep — every thread multiplies the denominators of its pairs into a running prefix product (only the x
//     coordinates are read; the prefixes go to a global scratch array, 32 B per pair, read back by the same thread);
//   the BA_T thread totals are combined by a product tree in shared memory; ONE lane inverts the root with the
//     product-free binary Euclid (Fp::inv_bgcd: add/shift work that does not occupy the multiplier pipe the other
//     resident CTAs are using); the tree is walked back down to per-thread inverses;
//   backward sweep — prefix products become the individual inverses, the additions are finished and the sums stored
//     (coalesced 64-byte records).
// Both sweeps are software-pipelined: the operands of the next pair (and, at level 1, the index words of the pair after
// it) are in flight while the products of the current pair execute.  Only the 8 KB tree lives in shared memory.
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

// Operand addresses of a pair.  Level 1: table points selected by the (index | sign) words; later levels: the sums of
// the previous level, written by this kernel (coherent loads, not the read-only path).
template <bool LEVEL1>
struct BaSrc {
    const u32* vals;
    const Affine* table;
    const Affine* in;
    __device__ __forceinline__ uint2 words(u32 pair, bool live) const {
        if (!LEVEL1) return make_uint2(live ? 0u : BA_PAD, live ? 0u : BA_PAD);
        return live ? __ldg(reinterpret_cast<const uint2*>(vals) + pair) : make_uint2(BA_PAD, BA_PAD);
    }
    __device__ __forceinline__ const char* addr(u32 pair, u32 word, int which) const {
        if (LEVEL1) return reinterpret_cast<const char*>(table + (word & ~BA_SIGN));
        return reinterpret_cast<const char*>(in + 2 * (size_t)pair + which);
    }
    __device__ __forceinline__ Fq ld(const char* a) const { return LEVEL1 ? Fq::load_nc(a) : Fq::load(a); }
    // y with the sign of the digit applied
    __device__ __forceinline__ Fq ld_y(u32 pair, u32 word, int which) const {
        Fq y = ld(addr(pair, word, which) + 32);
        if (LEVEL1 && (word & BA_SIGN)) y = y.neg();
        return y;
    }
};

// kind of a pair: 0 = chord addition, 1 = tangent (P == Q), 2 = trivial (an operand is the identity, or P == -Q)
// Forward sweep: from the x coordinates alone in the common case; y is fetched on the rare paths that need it.
template <bool LEVEL1>
__device__ __forceinline__ int ba_classify_x(const BaSrc<LEVEL1>& src, u32 pair, uint2 w, const Fq& px, const Fq& qx, Fq& den) {
    bool pid = w.x == BA_PAD, qid = w.y == BA_PAD;
    // x == 0 may be the identity (0,0): look at y
    if (!pid && px.is_zero()) pid = src.ld(src.addr(pair, w.x, 0) + 32).is_zero();
    if (!qid && qx.is_zero()) qid = src.ld(src.addr(pair, w.y, 1) + 32).is_zero();
    if (pid || qid) return 2;
    if (!(px == qx)) { den = qx - px; return 0; }
    const Fq yp = src.ld_y(pair, w.x, 0), yq = src.ld_y(pair, w.y, 1);
    if (yp == yq && !yp.is_zero()) { den = yp.dbl(); return 1; }
    return 2;
}
__device__ __forceinline__ int ba_classify_full(const Affine& p, const Affine& q, Fq& den) {
    if (p.is_identity() || q.is_identity()) return 2;
    if (!(p.x == q.x)) { den = q.x - p.x; return 0; }
    if (p.y == q.y && !p.y.is_zero()) { den = p.y.dbl(); return 1; }
    return 2;
}

// One level of one tile: `kk` pairs per thread, pair index = base + i * BA_T + tid, npairs = total pairs of the level.
// PT = true: every thread inverts the product of its own denominators with the constant-time safegcd routine
// (Fp::inv_safegcd: mostly add / shift work on the ALU pipe the products leave idle) — no product tree, no barrier, no
// warp waits for one lane.  PT = false: one inversion per CTA tile (product tree in shared memory + a single-lane inv_bgcd).
template <bool LEVEL1, bool PT>
__device__ __forceinline__ void ba_level(const BaSrc<LEVEL1> src, Affine* __restrict__ out, Fq* __restrict__ pref,
                                         u32 base, u32 npairs, int kk, u32* node) {
    const int tid = threadIdx.x;
    const Fq zero = Fq::zero();
    Fq inv_total;
    // ---- forward sweep: prefix products of the denominators of this thread's pairs
    {
        Fq run = Fq::one();
        u32 pair = base + tid;
        uint2 w = src.words(pair, pair < npairs);
        uint2 w_next = src.words(pair + BA_T, kk > 1 && pair + BA_T < npairs);
        Fq px = w.x != BA_PAD ? src.ld(src.addr(pair, w.x, 0)) : zero;
        Fq qx = w.y != BA_PAD ? src.ld(src.addr(pair, w.y, 1)) : zero;
#pragma unroll 1
        for (int i = 0; i < kk; i++) {
            // operands of pair i + 1 and index words of pair i + 2 in flight during the product of pair i
            const u32 pair_n = pair + BA_T;
            const uint2 wn = w_next;
            Fq pxn = zero, qxn = zero;
            if (i + 1 < kk) {
                if (wn.x != BA_PAD) pxn = src.ld(src.addr(pair_n, wn.x, 0));
                if (wn.y != BA_PAD) qxn = src.ld(src.addr(pair_n, wn.y, 1));
                w_next = src.words(pair_n + BA_T, i + 2 < kk && pair_n + BA_T < npairs);
            }
            Fq den;
            const int kind = ba_classify_x<LEVEL1>(src, pair, w, px, qx, den);
            if (kind != 2) run = run * den;
            if (pair < npairs) run.store(pref + pair);
            pair = pair_n; w = wn; px = pxn; qx = qxn;
        }
        if (PT) inv_total = run.inv_safegcd();
        else ba_smem_store(node + BA_T + tid, 2 * BA_T, run);
    }
    if (!PT) {
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
    inv_total = ba_smem_load(node + BA_T + tid, 2 * BA_T);
    }
    // ---- backward sweep: individual inverses, the additions, coalesced stores
    {
        Fq inv_run = inv_total;  // 1 / (product of all denominators of this thread)
        auto fetch = [&](u32 pr, uint2 ww, Affine& p, Affine& q) {
            p.x = zero; p.y = zero; q.x = zero; q.y = zero;
            if (ww.x != BA_PAD) { const char* a = src.addr(pr, ww.x, 0); p.x = src.ld(a); p.y = src.ld(a + 32); }
            if (ww.y != BA_PAD) { const char* a = src.addr(pr, ww.y, 1); q.x = src.ld(a); q.y = src.ld(a + 32); }
        };
        u32 pair = base + (u32)(kk - 1) * BA_T + tid;
        uint2 w = src.words(pair, pair < npairs);
        uint2 w_next = src.words(pair - BA_T, kk > 1 && pair - BA_T < npairs);
        Affine p, q;
        fetch(pair, w, p, q);
        Fq pf = (kk > 1 && pair - BA_T < npairs) ? Fq::load(pref + pair - BA_T) : zero;  // prefix after this thread's pair i - 1
#pragma unroll 1
        for (int i = kk - 1; i >= 0; i--) {
            const u32 pair_n = pair - BA_T;  // this thread's pair i - 1 (only used when i > 0)
            const uint2 wn = w_next;
            Affine pn, qn;
            Fq pfn = zero;
            pn.x = zero; pn.y = zero; qn.x = zero; qn.y = zero;
            if (i > 0) {
                fetch(pair_n, wn, pn, qn);
                if (i > 1 && pair_n - BA_T < npairs) pfn = Fq::load(pref + pair_n - BA_T);
                w_next = src.words(pair_n - BA_T, i > 1 && pair_n - BA_T < npairs);
            }
            if (LEVEL1) {
                if (w.x != BA_PAD && (w.x & BA_SIGN)) p.y = p.y.neg();
                if (w.y != BA_PAD && (w.y & BA_SIGN)) q.y = q.y.neg();
            }
            Fq den;
            const int kind = ba_classify_full(p, q, den);
            Affine sum;
            if (kind == 2) {
                // identity operand -> the other one; P + (-P) (both non-identity) -> identity
                if (p.is_identity()) sum = q;
                else if (q.is_identity()) sum = p;
                else { sum.x = zero; sum.y = zero; }
            } else {
                Fq inv = inv_run;
                if (i > 0) inv = inv * pf;
                inv_run = inv_run * den;
                Fq num;
                if (kind == 0) num = q.y - p.y;
                else { const Fq xx = p.x.sqr(); num = xx.dbl() + xx; }
                const Fq lam = num * inv;
                sum.x = lam.sqr() - p.x - q.x;
                sum.y = lam * (p.x - sum.x) - p.y;
            }
            if (pair < npairs) sum.store(out + pair);
            pair = pair_n; w = wn; p = pn; q = qn; pf = pfn;
        }
    }
    __syncthreads();  // the sums of this level are read by other threads of the CTA at the next level; node is reused
}

// n_entries_ptr: number of (padded) sorted entries M' on the device (a multiple of 2^R).  cursor: level-1 pair cursor,
// zero at launch.  buf_a / buf_b / buf_c: sums of levels 1, 2, 3 (M'/2, M'/4, M'/8 points); pref: M'/2 prefix products.
// k_nominal: level-1 pairs per thread and tile (multiple of 4).  R <= 3.
template <bool PT>
__global__ void __launch_bounds__(BA_T, 4) k_batch_affine(const u32* __restrict__ vals, const Affine* __restrict__ table,
                                                          Affine* __restrict__ buf_a, Affine* __restrict__ buf_b,
                                                          Affine* __restrict__ buf_c, Fq* __restrict__ pref,
                                                          const u32* __restrict__ n_entries_ptr, u32* __restrict__ cursor,
                                                          int R, int k_nominal) {
    __shared__ u32 node[8 * 2 * BA_T];  // [8][2 * BA_T]  heap order: leaves at BA_T + tid, root at 1
    __shared__ u32 sh_start, sh_k;
    const u32 m_entries = __ldg(n_entries_ptr);
    const u32 npairs1 = m_entries >> 1;
#pragma unroll 1
    for (int it = 0;; it++) {
        if (threadIdx.x == 0) {
            u32 k = (u32)k_nominal;
            // stagger: the CTAs of an SM start with tiles of different length
            if (it == 0) k = ((k * (1u + (blockIdx.x & 3u))) >> 2) & ~3u;
            // guided scheduling: tiles shrink towards the end so that all CTAs finish together
            const u32 cur = *reinterpret_cast<volatile u32*>(cursor);
            if (cur < npairs1) {
                const u32 share = ((npairs1 - cur) / (BA_T * gridDim.x) + 3u) & ~3u;
                if (share < k) k = share;
            }
            if (k < 8) k = 8;
            sh_k = k;
            sh_start = atomicAdd(cursor, (u32)BA_T * k);
        }
        __syncthreads();
        const u32 base1 = sh_start;
        const int k = (int)sh_k;
        if (base1 >= npairs1) break;
        const BaSrc<true> s1{vals, table, nullptr};
        ba_level<true, PT>(s1, buf_a, pref, base1, npairs1, k, node);
        // the prefix scratch of every level stays inside the tile's own level-1 region [base1, base1 + BA_T * k)
        if (R >= 2) {
            const BaSrc<false> s2{nullptr, nullptr, buf_a};
            ba_level<false, PT>(s2, buf_b, pref + (base1 - (base1 >> 1)), base1 >> 1, m_entries >> 2, k >> 1, node);
        }
        if (R >= 3) {
            const BaSrc<false> s3{nullptr, nullptr, buf_b};
            ba_level<false, PT>(s3, buf_c, pref + (base1 - (base1 >> 2)), base1 >> 2, m_entries >> 3, k >> 2, node);
        }
    }
}

}  // namespace h2b
