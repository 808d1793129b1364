// msm.cu — multi-scalar multiplication over BN254 G1 for sm_100a.
//
// Replaces halo2curves-axiom 0.7.3 `msm::best_multiexp(coeffs, bases)` as reached from
// ParamsKZG::commit / commit_lagrange inside create_proof (sole halo2-lib call site of create_proof:
// halo2-base/src/utils/testing.rs:40-48).  The result is the same group element.
//
// Pipeline (all on the context's stream, no host synchronisation):
//   k_digits          scalars (Montgomery) -> canonical -> signed base-2^c digits -> (bucket key, table index|sign)
//   cub radix sort    by bucket key (<= 22 key bits)
//   k_bucket_offsets  first sorted position of every bucket (binary search)
//   k_accumulate      every thread owns EXACTLY L consecutive sorted entries (perfect balance under any
//                     scalar distribution, witness columns are dominated by 0/1/88-bit limbs), gathers the
//                     64-byte affine points with 128-bit loads, XYZZ mixed adds; buckets that end inside
//                     the chunk are written directly, runs cut by a chunk border go to a partial array
//   k_collect(_big)   per bucket: add the partials of the chunks it spans
//   k_reduce_level    running sums over slices of 8 buckets, hierarchical (weights folded as doublings)
//   k_sum_level/k_finalize  plain tree sums (shared memory + per-thread serial), Horner over bucket sets
//
// Fixed bases (the SRS): `table[w*n + i] = 2^(c*w) * P_i` is built once per SRS (k_precompute_level), so all
// windows of a scalar fall into ONE bucket set: no per-window reduction and no final doublings.
#include <cub/device/device_radix_sort.cuh>

#include "curve.cuh"
#include "h2b_internal.cuh"

namespace h2b {

static constexpr u32 SIGN_BIT = 0x80000000u;
static constexpr int ACC_L = 16;       // sorted entries per accumulate thread
static constexpr int BIG_PARTIALS = 64;  // buckets spanning more chunks than this are summed by a whole CTA
static constexpr int RED_S = 8;        // buckets per reduce slice

// ------------------------------------------------------------------------------------------------ digits
// One thread per scalar.  keys/vals are window-major (index w*n + i) so that stores coalesce.
// Window w belongs to bucket set w / q and table level w % q.
__global__ void __launch_bounds__(256) k_digits(const uint64_t* __restrict__ scalars, u32 n, int c, int W, int q,
                                                u32 nbw, u32 invalid_key, u32* __restrict__ keys,
                                                u32* __restrict__ vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = Fr::load_nc(scalars + 4 * (size_t)i).from_mont();  // canonical integer, as `to_repr()` gives
    const u32 half = 1u << (c - 1);
    const u32 mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1u);
    u32 carry = 0;
    for (int w = 0; w < W; w++) {
        int bit = w * c;
        u32 raw = 0;
        if (bit < 256) {
            int limb = bit >> 5, off = bit & 31;
            // dynamic limb index on a register array: select through a small unrolled scan
            u32 lo = 0, hi = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                if (t == limb) lo = s.l[t];
                if (t == limb + 1) hi = s.l[t];
            }
            u64 v = ((u64)hi << 32) | lo;
            raw = (u32)(v >> off) & mask;
        }
        u32 d = raw + carry;
        u32 neg = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = SIGN_BIT;
            carry = 1;
        } else {
            carry = 0;
        }
        size_t o = (size_t)w * n + i;
        if (d == 0) {
            keys[o] = invalid_key;
            vals[o] = 0;
        } else {
            keys[o] = (u32)(w / q) * nbw + (d - 1);
            vals[o] = ((u32)(w % q) * n + i) | neg;
        }
    }
}

// off[b] = first sorted position with key >= b, for b in [0, nb_total]; off[nb_total] = number of valid entries
__global__ void __launch_bounds__(256) k_bucket_offsets(const u32* __restrict__ keys, u32 M, u32 nb_total,
                                                        u32* __restrict__ off) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb_total) return;
    u32 lo = 0, hi = M;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < b) lo = mid + 1; else hi = mid;
    }
    off[b] = lo;
}

// ------------------------------------------------------------------------------------------------ accumulate
__device__ __forceinline__ Affine load_signed(const Affine* __restrict__ table, u32 val) {
    Affine p = Affine::load(table + (val & ~SIGN_BIT));
    if (val & SIGN_BIT) p.y = p.y.neg();
    return p;
}

template <int L>
__global__ void __launch_bounds__(128) k_accumulate(const u32* __restrict__ keys, const u32* __restrict__ vals,
                                                    const u32* __restrict__ off, u32 nb_total,
                                                    const Affine* __restrict__ table, XYZZ* __restrict__ buckets,
                                                    XYZZ* __restrict__ partials) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 mv = __ldg(off + nb_total);  // valid (non-zero-digit) entries
    const u64 cs64 = (u64)t * L;
    if (cs64 >= mv) return;
    const u32 cs = (u32)cs64;
    const u32 ce = (mv - cs < (u32)L) ? mv : cs + L;

    XYZZ acc = XYZZ::identity();
    u32 cur = __ldg(keys + cs);
    Affine p = load_signed(table, __ldg(vals + cs));
    for (u32 i = cs; i < ce; i++) {
        Affine pn;
        u32 kn = cur;
        const bool more = (i + 1 < ce);
        if (more) {  // prefetch the next gather while this add runs
            kn = __ldg(keys + i + 1);
            pn = load_signed(table, __ldg(vals + i + 1));
        }
        xyzz_madd(acc, p);
        if (!more || kn != cur) {
            // the run of bucket `cur` ends here (inside this chunk or at its border)
            const u32 s = __ldg(off + cur), e = __ldg(off + cur + 1);
            if (s >= cs && e - cs <= (u32)L) acc.store(buckets + cur);          // bucket lies inside the chunk
            else if (s <= cs) acc.store(partials + 2 * (size_t)t);              // covers the chunk start
            else acc.store(partials + 2 * (size_t)t + 1);                        // starts inside, runs past the end
            acc = XYZZ::identity();
            cur = kn;
        }
        if (more) p = pn;
    }
}

__device__ __forceinline__ const XYZZ* partial_of(const XYZZ* partials, u32 s, u32 t, int L) {
    return (s <= t * (u32)L) ? partials + 2 * (size_t)t : partials + 2 * (size_t)t + 1;
}

// one thread per bucket: empty -> identity; spans several chunks -> add their partials
__global__ void __launch_bounds__(128) k_collect(const u32* __restrict__ off, u32 nb_total, int L,
                                                 const XYZZ* __restrict__ partials, XYZZ* __restrict__ buckets,
                                                 u32* __restrict__ big_list, u32* __restrict__ big_count) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb_total) return;
    const u32 s = __ldg(off + b), e = __ldg(off + b + 1);
    if (s == e) {
        XYZZ::identity().store(buckets + b);
        return;
    }
    const u32 t_lo = s / L, t_hi = (e - 1) / L;
    if (t_lo == t_hi) return;  // written by k_accumulate
    if (t_hi - t_lo + 1 > (u32)BIG_PARTIALS) {
        big_list[atomicAdd(big_count, 1u)] = b;
        return;
    }
    XYZZ acc = XYZZ::load(partial_of(partials, s, t_lo, L));
    for (u32 t = t_lo + 1; t <= t_hi; t++) xyzz_add(acc, XYZZ::load(partials + 2 * (size_t)t));
    acc.store(buckets + b);
}

// block-wide sum of one XYZZ per thread (256 threads); result valid in thread 0
__device__ __forceinline__ XYZZ block_sum_256(XYZZ v, XYZZ* sh) {
    const int tid = threadIdx.x;
    v.store(sh + tid);
    __syncthreads();
#pragma unroll 1
    for (int stride = 128; stride >= 1; stride >>= 1) {
        if (tid < stride) {
            XYZZ a = XYZZ::load(sh + tid), b2 = XYZZ::load(sh + tid + stride);
            xyzz_add(a, b2);
            a.store(sh + tid);
        }
        __syncthreads();
    }
    return XYZZ::load(sh);
}

// hot buckets (e.g. digit 1 of a bit-valued witness column): one CTA per bucket
__global__ void __launch_bounds__(256) k_collect_big(const u32* __restrict__ off, int L,
                                                     const XYZZ* __restrict__ partials, XYZZ* __restrict__ buckets,
                                                     const u32* __restrict__ big_list,
                                                     const u32* __restrict__ big_count) {
    __shared__ XYZZ sh[256];
    const u32 nbig = *big_count;
    for (u32 j = blockIdx.x; j < nbig; j += gridDim.x) {
        const u32 b = big_list[j];
        const u32 s = __ldg(off + b), e = __ldg(off + b + 1);
        const u32 t_lo = s / L, t_hi = (e - 1) / L;
        XYZZ acc = XYZZ::identity();
        for (u32 t = t_lo + threadIdx.x; t <= t_hi; t += 256) xyzz_add(acc, XYZZ::load(partial_of(partials, s, t, L)));
        XYZZ r = block_sum_256(acc, sh);
        if (threadIdx.x == 0) r.store(buckets + b);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ bucket reduce
// One level of V = sum_b (b+1) B_b = sum_b B_b + sum_b b*B_b.  With slices of S consecutive inputs,
//   sum_j j*T_j = sum_J [ sum_i i*T_{JS+i} ] + S * sum_J J*T'_J ,  T'_J = sum_i T_{JS+i}
// so each level emits T' (input of the next level) and R_J = sum_i i*T_{JS+i} scaled by the product of the
// slice sizes of the levels below (2^dbl) into a pool that is summed plainly at the end.
__global__ void __launch_bounds__(128) k_reduce_level(const XYZZ* __restrict__ tin, u32 n_in, int S, int dbl,
                                                      u32 nsets, XYZZ* __restrict__ tout,
                                                      XYZZ* __restrict__ pool, u32 pool_stride, u32 pool_off) {
    const u32 n_out = n_in / S;
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_out * nsets) return;
    const u32 set = g / n_out, J = g % n_out;
    const XYZZ* src = tin + (size_t)set * n_in + (size_t)J * S;
    XYZZ run = XYZZ::identity(), acc = XYZZ::identity();
    for (int i = S - 1; i >= 1; i--) {
        xyzz_add(run, XYZZ::load(src + i));
        xyzz_add(acc, run);
    }
    xyzz_add(run, XYZZ::load(src));
    run.store(tout + (size_t)set * n_out + J);
    for (int d = 0; d < dbl; d++) acc = xyzz_dbl(acc);
    acc.store(pool + (size_t)set * pool_stride + pool_off + J);
}

// plain sums: out[set][j] = sum of up to G consecutive in[set][*]
__global__ void __launch_bounds__(128) k_sum_level(const XYZZ* __restrict__ in, u32 n_in, u32 in_stride, int G,
                                                   u32 nsets, XYZZ* __restrict__ out, u32 out_stride) {
    const u32 n_out = (n_in + G - 1) / G;
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_out * nsets) return;
    const u32 set = g / n_out, j = g % n_out;
    const XYZZ* src = in + (size_t)set * in_stride;
    u32 lo = j * G, hi = min(lo + (u32)G, n_in);
    XYZZ acc = XYZZ::load(src + lo);
    for (u32 i = lo + 1; i < hi; i++) xyzz_add(acc, XYZZ::load(src + i));
    acc.store(out + (size_t)set * out_stride + j);
}

__device__ __forceinline__ void store_jacobian(const XYZZ& p, void* out) {
    // (X, Y, ZZ, ZZZ) -> Jacobian with Z = ZZ*ZZZ:  X_j = X*ZZ*ZZZ^2, Y_j = Y*ZZ^3*ZZZ^2   (no inversion)
    char* o = reinterpret_cast<char*>(out);
    if (p.is_identity()) {
        Fq::zero().store(o);
        Fq::one().store(o + 32);
        Fq::zero().store(o + 64);
        return;
    }
    Fq z = p.zz * p.zzz;
    Fq a = p.zz * p.zzz.sqr();  // ZZ * ZZZ^2
    (p.x * a).store(o);
    (p.y * a * p.zz.sqr()).store(o + 32);
    z.store(o + 64);
}

// single CTA: V_set = top[set] + sum(pool[set][0..n_pool)), then Horner over sets with `shift` doublings
__global__ void __launch_bounds__(256) k_finalize(const XYZZ* __restrict__ top, const XYZZ* __restrict__ pool,
                                                  u32 n_pool, u32 pool_stride, u32 nsets, int shift,
                                                  void* __restrict__ out) {
    __shared__ XYZZ sh[256];
    __shared__ XYZZ vset;
    XYZZ total = XYZZ::identity();  // meaningful in thread 0
    for (int set = (int)nsets - 1; set >= 0; set--) {
        XYZZ acc = XYZZ::identity();
        for (u32 i = threadIdx.x; i < n_pool; i += 256) xyzz_add(acc, XYZZ::load(pool + (size_t)set * pool_stride + i));
        XYZZ r = block_sum_256(acc, sh);
        if (threadIdx.x == 0) {
            xyzz_add(r, XYZZ::load(top + set));
            if (set != (int)nsets - 1)
                for (int d = 0; d < shift; d++) total = xyzz_dbl(total);
            xyzz_add(total, r);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) store_jacobian(total, out);
    (void)vset;
}

// ------------------------------------------------------------------------------------------------ SRS table
// next[i] = 2^c * prev[i], normalised to affine (one inversion per point; runs once per SRS upload)
__global__ void __launch_bounds__(128) k_precompute_level(const Affine* __restrict__ prev, Affine* __restrict__ next,
                                                          u32 count, int c) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Affine a = Affine::load(prev + i);
    if (a.is_identity()) {
        a.store(next + i);
        return;
    }
    XYZZ p = xyzz_dbl_affine(a);
    for (int d = 1; d < c; d++) p = xyzz_dbl(p);
    xyzz_to_affine(p).store(next + i);
}

// ------------------------------------------------------------------------------------------------ small group ops
__global__ void __launch_bounds__(256) k_g1_sum(const uint64_t* __restrict__ pts, u32 m, void* __restrict__ out) {
    __shared__ XYZZ sh[256];
    XYZZ acc = XYZZ::identity();
    for (u32 i = threadIdx.x; i < m; i += 256) xyzz_add(acc, xyzz_from_jacobian(pts + 12 * (size_t)i));
    XYZZ r = block_sum_256(acc, sh);
    if (threadIdx.x == 0) store_jacobian(r, out);
}
__global__ void __launch_bounds__(128) k_g1_normalize(uint64_t* __restrict__ pts, u32 m) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    XYZZ p = xyzz_from_jacobian(pts + 12 * (size_t)i);
    xyzz_store_jacobian_normalised(p, pts + 12 * (size_t)i);
}
// out[i] = s_i * base: plain double-and-add over the canonical scalar bits (setup-side utility)
__global__ void __launch_bounds__(128) k_fixed_base_mul(Affine base, const uint64_t* __restrict__ scalars, u32 n,
                                                        Affine* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = Fr::load_nc(scalars + 4 * (size_t)i).from_mont();
    XYZZ acc = XYZZ::identity();
    for (int limb = 7; limb >= 0; limb--) {
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (t == limb) v = s.l[t];
        for (int bit = 31; bit >= 0; bit--) {
            acc = xyzz_dbl(acc);
            if ((v >> bit) & 1) xyzz_madd(acc, base);
        }
    }
    xyzz_to_affine(acc).store(out + i);
}

template <class F>
__global__ void __launch_bounds__(128) k_field_op(int op, const uint64_t* __restrict__ a,
                                                  const uint64_t* __restrict__ b, u32 n, uint64_t* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = F::load(a + 4 * (size_t)i), y = F::zero(), r;
    if (op <= 2) y = F::load(b + 4 * (size_t)i);
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.inv(); break;
        case 4: r = x.from_mont(); break;
        default: r = x.to_mont(); break;
    }
    r.store(out + 4 * (size_t)i);
}

// ------------------------------------------------------------------------------------------------ host side
int msm_choose_c_fixed(size_t n) {
    // one bucket set of 2^(c-1) buckets, ceil(255/c) table levels: cost ~ 10*n*W + 32*2^(c-1) field mults
    int lg = ceil_log2(n ? n : 1);
    int c = lg - 2;
    if (c < 8) c = 8;
    if (c > 22) c = 22;
    return c;
}
static int msm_choose_c_adhoc(size_t n) {
    int lg = ceil_log2(n ? n : 1);
    int c = lg - 4;
    if (c < 6) c = 6;
    if (c > 16) c = 16;
    return c;
}

void msm_build_table(h2b_ctx* ctx, const void* d_bases, size_t count, int c, int W, void* d_table) {
    Affine* t = reinterpret_cast<Affine*>(d_table);
    H2B_CUDA(cudaMemcpyAsync(t, d_bases, count * sizeof(Affine), cudaMemcpyDeviceToDevice, ctx->stream));
    for (int w = 1; w < W; w++)
        H2B_LAUNCH(ctx, k_precompute_level, ceil_div(count, 128), 128, 0, t + (size_t)(w - 1) * count,
                   t + (size_t)w * count, (u32)count, c);
}

void msm_run(h2b_ctx* ctx, const void* d_table, size_t n, int c, int W, int q, const void* d_scalars, void* d_out) {
    H2B_REQUIRE(n >= 1 && n <= ((size_t)1 << 27), "msm: n out of range");
    H2B_REQUIRE((size_t)W * n < ((size_t)1 << 31), "msm: n * windows exceeds the 31-bit table index");
    const u32 nbw = 1u << (c - 1);
    const u32 nsets = (u32)((W + q - 1) / q);
    const u32 nb_total = nsets * nbw;
    const size_t M = (size_t)W * n;
    cudaStream_t st = ctx->stream;

    u32* keys_a = (u32*)ctx->get(WS_KEYS_A, M * 4);
    u32* keys_b = (u32*)ctx->get(WS_KEYS_B, M * 4);
    u32* vals_a = (u32*)ctx->get(WS_VALS_A, M * 4);
    u32* vals_b = (u32*)ctx->get(WS_VALS_B, M * 4);
    u32* off = (u32*)ctx->get(WS_OFFSETS, ((size_t)nb_total + 2) * 4);
    XYZZ* buckets = (XYZZ*)ctx->get(WS_BUCKETS, (size_t)nb_total * sizeof(XYZZ));
    const size_t n_chunks = (M + ACC_L - 1) / ACC_L;
    XYZZ* partials = (XYZZ*)ctx->get(WS_PARTIALS, 2 * n_chunks * sizeof(XYZZ));
    u32* big = (u32*)ctx->get(WS_BIGLIST, ((size_t)nb_total + 1) * 4);  // [0] = counter, list follows

    H2B_LAUNCH(ctx, k_digits, ceil_div(n, 256), 256, 0, (const uint64_t*)d_scalars, (u32)n, c, W, q, nbw, nb_total,
               keys_a, vals_a);

    const int key_bits = ceil_log2((size_t)nb_total + 1);
    size_t tmp_bytes = 0;
    H2B_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, vals_a, vals_b, (int)M, 0, key_bits, st));
    void* tmp = ctx->get(WS_SORT_TMP, tmp_bytes);
    H2B_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, vals_a, vals_b, (int)M, 0, key_bits, st));

    H2B_LAUNCH(ctx, k_bucket_offsets, ceil_div((size_t)nb_total + 1, 256), 256, 0, keys_b, (u32)M, nb_total, off);
    H2B_CUDA(cudaMemsetAsync(big, 0, 4, st));
    H2B_LAUNCH(ctx, k_accumulate<ACC_L>, ceil_div(n_chunks, 128), 128, 0, keys_b, vals_b, off, nb_total,
               (const Affine*)d_table, buckets, partials);
    H2B_LAUNCH(ctx, k_collect, ceil_div(nb_total, 128), 128, 0, off, nb_total, ACC_L, partials, buckets, big + 1, big);
    H2B_LAUNCH(ctx, k_collect_big, 2 * ctx->sm_count, 256, 0, off, ACC_L, partials, buckets, big + 1, big);

    // hierarchical bucket reduction
    u32 pool_per_set = 0;
    {
        u32 nin = nbw;
        while (nin > 1) {
            int S = nin >= (u32)RED_S ? RED_S : (int)nin;
            nin /= S;
            pool_per_set += nin;
        }
    }
    XYZZ* ta = (XYZZ*)ctx->get(WS_REDUCE_A, (size_t)nsets * (nbw / 2 + 1) * sizeof(XYZZ));
    XYZZ* tb = (XYZZ*)ctx->get(WS_REDUCE_B, (size_t)nsets * (nbw / 2 + 1) * sizeof(XYZZ));
    XYZZ* pool = (XYZZ*)ctx->get(WS_POOL, (size_t)nsets * (pool_per_set + 1) * sizeof(XYZZ));
    XYZZ* pool2 = (XYZZ*)ctx->get(WS_POOL2, (size_t)nsets * (pool_per_set / 8 + 2) * sizeof(XYZZ));
    const XYZZ* cur = buckets;
    u32 nin = nbw, pool_off = 0;
    int dbl = 0;
    XYZZ* nxt = ta;
    while (nin > 1) {
        int S = nin >= (u32)RED_S ? RED_S : (int)nin;
        u32 nout = nin / S;
        H2B_LAUNCH(ctx, k_reduce_level, ceil_div((size_t)nout * nsets, 128), 128, 0, cur, nin, S, dbl, nsets, nxt, pool,
                   pool_per_set, pool_off);
        pool_off += nout;
        dbl += ceil_log2(S);
        nin = nout;
        cur = nxt;
        nxt = (nxt == ta) ? tb : ta;
    }
    // cur[set] is the plain sum of the set's buckets; pool[set][0..pool_per_set) the weighted parts
    const XYZZ* pcur = pool;
    u32 pn = pool_per_set, pstride = pool_per_set;
    if (pn > 2048) {
        u32 pout = (pn + 7) / 8;
        H2B_LAUNCH(ctx, k_sum_level, ceil_div((size_t)pout * nsets, 128), 128, 0, pcur, pn, pstride, 8, nsets, pool2, pout);
        pcur = pool2;
        pn = pout;
        pstride = pout;
    }
    H2B_LAUNCH(ctx, k_finalize, 1, 256, 0, cur, pcur, pn, pstride, nsets, c * q, d_out);
}

void g1_sum_run(h2b_ctx* ctx, const void* d_points, size_t m, void* d_out) {
    H2B_LAUNCH(ctx, k_g1_sum, 1, 256, 0, (const uint64_t*)d_points, (u32)m, d_out);
}
void g1_normalize_run(h2b_ctx* ctx, void* d_points, size_t m) {
    if (m == 0) return;
    H2B_LAUNCH(ctx, k_g1_normalize, ceil_div(m, 128), 128, 0, (uint64_t*)d_points, (u32)m);
}
void g1_fixed_base_mul_run(h2b_ctx* ctx, const uint64_t base_xy[8], const void* d_scalars, size_t n, void* d_out) {
    if (n == 0) return;
    Affine b;
    memcpy(&b, base_xy, sizeof(Affine));
    H2B_LAUNCH(ctx, k_fixed_base_mul, ceil_div(n, 128), 128, 0, b, (const uint64_t*)d_scalars, (u32)n, (Affine*)d_out);
}
void field_op_run(h2b_ctx* ctx, int field, int op, const void* a, const void* b, size_t n, void* out) {
    if (n == 0) return;
    if (field == 0)
        H2B_LAUNCH(ctx, k_field_op<Fq>, ceil_div(n, 128), 128, 0, op, (const uint64_t*)a, (const uint64_t*)b, (u32)n, (uint64_t*)out);
    else
        H2B_LAUNCH(ctx, k_field_op<Fr>, ceil_div(n, 128), 128, 0, op, (const uint64_t*)a, (const uint64_t*)b, (u32)n, (uint64_t*)out);
}

// ad-hoc bases: W bucket sets, no table
void msm_run_adhoc(h2b_ctx* ctx, const void* d_bases, size_t n, const void* d_scalars, void* d_out) {
    int c = msm_choose_c_adhoc(n);
    int W = (255 + c - 1) / c;
    msm_run(ctx, d_bases, n, c, W, 1, d_scalars, d_out);
}

}  // namespace h2b
