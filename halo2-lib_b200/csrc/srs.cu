// srs.cu — keygen-side SRS utilities for sm_100a:
//   g_to_lagrange   halo2-axiom 0.5.3 `poly/kzg/commitment.rs::g_to_lagrange` = best_fft over G1 with omega^-1, every
//                   point scaled by 2^-k, batch-normalised:  g_lagrange[i] = (1/n) sum_j omega^(-i j) g[j]
//   srs_setup       `ParamsKZG::setup` for a caller-supplied tau: g[i] = tau^i G, g_lagrange[i] = L_i(tau) G with
//                   L_i(tau) = (tau^n - 1)/n * omega^i / (tau - omega^i)
//   on-curve check  y^2 = x^3 + 3 or (0,0) for points read from a params file (`ParamsKZG::read`,
//                   halo2-base/src/utils/mod.rs:401-424)
// (not vendored; restated from the definitions — the outputs are unique group elements).
//
// The group FFT is radix-2 decimation in time on XYZZ points in HBM: n/2 butterflies per stage, each one scalar
// multiplication by a twiddle (double-and-add over <= 254 bits, ~4.2k Montgomery products) plus two point additions —
// integer-multiplier bound, ~ k * n/2 * 4.4k products in total (k = 19: ~2.2e10, a third of a second), run once per SRS.
#include "h2b_internal.cuh"
#include "curve.cuh"
#include "fr_domain_consts.inc"

namespace h2b {

// ---------------------------------------------------------------- compressed points (SerdeFormat::Processed)
// halo2curves `G1Affine::from_bytes` as `ParamsKZG::read` applies it to a `kzg_bn254_{k}.srs` file written in
// SerdeFormat::Processed (reference call sites halo2-base/src/utils/mod.rs:401-435): 32 bytes = x, little-endian,
// canonical; the two spare bits of the last byte carry bit 7 = point at infinity, bit 6 = parity of y (LSB of the
// canonical y).  [halo2curves-axiom 0.7.3 is not vendored: the flag positions are recalled, see DESIGN.md §2.]
// y = sqrt(x^3 + 3) = (x^3 + 3)^((p + 1) / 4) since p = 3 mod 4; a non-residue, an x >= p, or an infinity flag on a
// non-zero x is an invalid encoding.  out = (x, y) Montgomery, (0, 0) for the identity; invalid -> (0, 0) and counted.
__device__ __forceinline__ Fq fq_sqrt_candidate(const Fq& a) {
    // (p + 1) / 4, little-endian 32-bit limbs
    const u32 e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    Fq acc = Fq::one();
#pragma unroll 1
    for (int limb = 7; limb >= 0; limb--) {
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (t == limb) v = e[t];
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            acc = acc.sqr();
            if ((v >> bit) & 1) acc = acc * a;
        }
    }
    return acc;
}
__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* __restrict__ bytes, size_t n, Affine* __restrict__ out,
                                                       unsigned long long* __restrict__ invalid) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(bytes + 32 * i);
    const uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fq x;
    x.l[0] = lo.x; x.l[1] = lo.y; x.l[2] = lo.z; x.l[3] = lo.w;
    x.l[4] = hi.x; x.l[5] = hi.y; x.l[6] = hi.z; x.l[7] = hi.w;
    const bool inf = (x.l[7] >> 31) & 1, odd = (x.l[7] >> 30) & 1;
    x.l[7] &= 0x3fffffffu;
    Affine r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    bool bad = false;
    if (inf) {
        bad = !x.is_zero() || odd;
    } else {
        // canonical: x < p
        bool lt = false;
        for (int t = 7; t >= 0; t--) {
            if (x.l[t] != FqParams::MOD(t)) { lt = x.l[t] < FqParams::MOD(t); break; }
        }
        if (!lt) bad = true;
        else {
            const Fq xm = x.to_mont();
            Fq three = Fq::one();
            three = three + three + Fq::one();
            const Fq rhs = xm.sqr() * xm + three;
            Fq y = fq_sqrt_candidate(rhs);
            if (!(y.sqr() == rhs)) bad = true;
            else {
                if ((y.from_mont().l[0] & 1u) != (odd ? 1u : 0u)) y = y.neg();
                r.x = xm;
                r.y = y;
            }
        }
    }
    if (bad) atomicAdd(invalid, 1ull);
    r.store(out + i);
}

// s * p, s canonical (non-Montgomery) limbs; MSB-first double-and-add (doubling the identity is free)
static __device__ __noinline__ XYZZ xyzz_scalar_mul(const XYZZ& p, const Fr& s) {
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (int limb = 7; limb >= 0; limb--) {
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (t == limb) v = s.l[t];
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            acc = xyzz_dbl(acc);
            if ((v >> bit) & 1) xyzz_add(acc, p);
        }
    }
    return acc;
}

__device__ __forceinline__ u32 bitrev(u32 x, u32 bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

__global__ void __launch_bounds__(256) k_affine_to_xyzz_bitrev(const Affine* __restrict__ in, u32 k, XYZZ* __restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> k) return;
    XYZZ::from_affine(Affine::load(in + i)).store(out + bitrev(i, k));
}
// tw[i] = base^i (canonical limbs) for i < count, from pw[j] = base^(2^j)
__global__ void __launch_bounds__(256) k_power_table(const uint64_t* __restrict__ pw, u32 count, int canonical, uint64_t* __restrict__ tw) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr r = Fr::one();
    for (u32 j = 0; (i >> j) != 0; j++)
        if ((i >> j) & 1) r = r * Fr::load_nc(pw + 4 * (size_t)j);
    (canonical ? r.from_mont() : r).store(tw + 4 * (size_t)i);
}
// pw[j] = x^(2^j), j < 32;  extra[0] = (2^k)^-1 canonical, extra[1] = (x^(2^k) - 1) / 2^k (Montgomery)
__global__ void k_srs_consts(Fr x, u32 k, uint64_t* __restrict__ pw, uint64_t* __restrict__ extra) {
    if (threadIdx.x | blockIdx.x) return;
    Fr xn = x;
    for (u32 j = 0; j < 32; j++) {
        x.store(pw + 4 * j);
        if (j == k) xn = x;
        x = x.sqr();
    }
    Fr n = Fr::zero();
    n.l[k >> 5] = 1u << (k & 31);
    const Fr n_inv = n.to_mont().inv_bgcd();
    n_inv.from_mont().store(extra);
    ((xn - Fr::one()) * n_inv).store(extra + 4);
}
// stage s (1-based): pairs (i0, i0 + half), twiddle tw[j << (k - s)]
__global__ void __launch_bounds__(128) k_ecfft_stage(XYZZ* __restrict__ pts, const uint64_t* __restrict__ tw, u32 k, u32 s) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >> (k - 1)) return;
    const u32 half = 1u << (s - 1), j = t & (half - 1), i0 = ((t >> (s - 1)) << s) + j, i1 = i0 + half;
    XYZZ a = XYZZ::load(pts + i0), b = XYZZ::load(pts + i1);
    if (j) b = xyzz_scalar_mul(b, Fr::load_nc(tw + 4 * ((size_t)j << (k - s))));
    XYZZ lo = a;
    xyzz_add(lo, b);
    xyzz_add(a, b.neg());
    lo.store(pts + i0);
    a.store(pts + i1);
}
__global__ void __launch_bounds__(128) k_scale_to_affine(const XYZZ* __restrict__ pts, const uint64_t* __restrict__ scalar_canon, u32 n,
                                                         Affine* __restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_to_affine(xyzz_scalar_mul(XYZZ::load(pts + i), Fr::load_nc(scalar_canon))).store(out + i);
}
// den[i] = tau - omega^i
__global__ void __launch_bounds__(256) k_lagrange_den(Fr tau, const uint64_t* __restrict__ omega_pows, u32 n, uint64_t* __restrict__ den) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) (tau - Fr::load_nc(omega_pows + 4 * (size_t)i)).store(den + 4 * (size_t)i);
}
// l[i] = c * omega^i * den_inv[i]
__global__ void __launch_bounds__(256) k_lagrange_scalars(const uint64_t* __restrict__ c, const uint64_t* __restrict__ omega_pows,
                                                          const uint64_t* __restrict__ den_inv, u32 n, uint64_t* __restrict__ l) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) (Fr::load_nc(c) * Fr::load_nc(omega_pows + 4 * (size_t)i) * Fr::load_nc(den_inv + 4 * (size_t)i)).store(l + 4 * (size_t)i);
}
__global__ void __launch_bounds__(256) k_on_curve(const Affine* __restrict__ pts, size_t n, u32* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = Affine::load(pts + i);
    if (p.is_identity()) return;
    Fq three = Fq::one().dbl() + Fq::one();
    if (!(p.y.sqr() - (p.x.sqr() * p.x + three)).is_zero()) atomicAdd(bad, 1u);
}

static Fr fr_of(const uint64_t x[4]) {
    Fr r;
    memcpy(&r, x, sizeof(Fr));
    return r;
}

void g_to_lagrange_run(h2b_ctx* ctx, const void* d_g, uint32_t k, void* d_g_lagrange) {
    H2B_REQUIRE(k <= 28, "g_to_lagrange: k out of range");
    const size_t n = (size_t)1 << k;
    // workspace: XYZZ points | twiddles (n/2 canonical) | pw (32) | extra (2)
    char* w = (char*)ctx->get(WS_POOL, n * sizeof(XYZZ) + (n / 2 + 40) * 32);
    XYZZ* pts = (XYZZ*)w;
    uint64_t* tw = (uint64_t*)(w + n * sizeof(XYZZ));
    uint64_t* pw = tw + 4 * (n / 2 + 1);
    uint64_t* extra = pw + 4 * 32;
    H2B_LAUNCH(ctx, k_srs_consts, 1, 32, 0, fr_of(FR_OMEGA_INV[k]), k, pw, extra);
    if (n > 1) H2B_LAUNCH(ctx, k_power_table, ceil_div(n / 2, 256), 256, 0, pw, (u32)(n / 2), 1, tw);
    H2B_LAUNCH(ctx, k_affine_to_xyzz_bitrev, ceil_div(n, 256), 256, 0, (const Affine*)d_g, k, pts);
    for (uint32_t s = 1; s <= k; s++) H2B_LAUNCH(ctx, k_ecfft_stage, ceil_div(n / 2, 128), 128, 0, pts, tw, k, s);
    H2B_LAUNCH(ctx, k_scale_to_affine, ceil_div(n, 128), 128, 0, pts, extra, (u32)n, (Affine*)d_g_lagrange);
}

void batch_invert_run(h2b_ctx* ctx, void* d_a, size_t n);
void g1_fixed_base_mul_run(h2b_ctx* ctx, const uint64_t base_xy[8], const void* d_scalars, size_t n, void* d_out);

// g[i] = tau^i * base, g_lagrange[i] = L_i(tau) * base; either output may be null
void srs_setup_run(h2b_ctx* ctx, const uint64_t tau[4], const uint64_t base_xy[8], uint32_t k, void* d_g, void* d_g_lagrange) {
    H2B_REQUIRE(k <= 28, "srs_setup: k out of range");
    const size_t n = (size_t)1 << k;
    // workspace: scalars (n) | omega powers (n) | den (n) | pw_tau (32) | extra (2) | pw_omega (32) | extra2 (2)
    uint64_t* w = (uint64_t*)ctx->get(WS_POOL2, (3 * n + 80) * 32);
    uint64_t *sc = w, *om = w + 4 * n, *den = om + 4 * n, *pw_tau = den + 4 * n, *extra = pw_tau + 4 * 32, *pw_om = extra + 4 * 2,
             *extra2 = pw_om + 4 * 32;
    H2B_LAUNCH(ctx, k_srs_consts, 1, 32, 0, fr_of(tau), k, pw_tau, extra);
    if (d_g) {
        H2B_LAUNCH(ctx, k_power_table, ceil_div(n, 256), 256, 0, pw_tau, (u32)n, 0, sc);
        g1_fixed_base_mul_run(ctx, base_xy, sc, n, d_g);
    }
    if (d_g_lagrange) {
        H2B_LAUNCH(ctx, k_srs_consts, 1, 32, 0, fr_of(FR_OMEGA[k]), k, pw_om, extra2);
        H2B_LAUNCH(ctx, k_power_table, ceil_div(n, 256), 256, 0, pw_om, (u32)n, 0, om);
        H2B_LAUNCH(ctx, k_lagrange_den, ceil_div(n, 256), 256, 0, fr_of(tau), om, (u32)n, den);
        batch_invert_run(ctx, den, n);
        H2B_LAUNCH(ctx, k_lagrange_scalars, ceil_div(n, 256), 256, 0, extra + 4, om, den, (u32)n, sc);
        g1_fixed_base_mul_run(ctx, base_xy, sc, n, d_g_lagrange);
    }
}

// number of points that are neither (0,0) nor on y^2 = x^3 + 3 (synchronises)
size_t g1_count_off_curve_run(h2b_ctx* ctx, const void* d_points, size_t n) {
    if (n == 0) return 0;
    u32* d_bad = (u32*)ctx->get(WS_OUT, 256);
    H2B_CUDA(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    H2B_LAUNCH(ctx, k_on_curve, ceil_div(n, 256), 256, 0, (const Affine*)d_points, n, d_bad);
    u32* bounce = (u32*)ctx->get_pinned(0, 4096);
    H2B_CUDA(cudaMemcpyAsync(bounce, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    H2B_CUDA(cudaStreamSynchronize(ctx->stream));
    return bounce[0];
}

// bytes: n x 32 on the device; returns the number of invalid encodings (synchronises the stream)
size_t g1_decompress_run(h2b_ctx* ctx, const void* d_bytes, size_t n, void* d_out_xy) {
    if (n == 0) return 0;
    unsigned long long* d_cnt = (unsigned long long*)ctx->get(WS_MISC, 8);
    H2B_CUDA(cudaMemsetAsync(d_cnt, 0, 8, ctx->stream));
    H2B_LAUNCH(ctx, k_g1_decompress, ceil_div(n, 128), 128, 0, (const uint8_t*)d_bytes, n, (Affine*)d_out_xy, d_cnt);
    unsigned long long* bounce = (unsigned long long*)ctx->get_pinned(2, 4096);
    H2B_CUDA(cudaMemcpyAsync(bounce, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    H2B_CUDA(cudaStreamSynchronize(ctx->stream));
    return (size_t)bounce[0];
}

}  // namespace h2b
