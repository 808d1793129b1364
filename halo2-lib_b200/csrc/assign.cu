// assign.cu — column-wise witness assignment for sm_100a.
//
// Replaces the per-cell loop of `assign_witnesses`
// (halo2-base/src/gates/flex_gate/threads/single_phase.rs:273-312; each cell goes through
// raw_assign_advice, utils/halo2.rs:20-27, into the prover's WitnessCollection) and the lookup-column copy
// of LookupAnyManager::assign_raw (halo2-base/src/virtual_region/lookups.rs:130-155).
//
// Closed form of the walk: with V the concatenation of ctx.advice over threads,
// b_0..b_{m-1} the pinned break points and s_0 = 0, s_{c+1} = s_c + b_c, column c holds V[s_c + r] for
// 0 <= r <= b_c (the cell at the break is duplicated into row 0 of column c+1); the walk stops breaking as
// soon as fewer than b_c + 1 cells remain.  The kernels are pure 32-byte gathers: 64 B of traffic per cell.
#include "h2b_internal.cuh"
#include "field.cuh"

namespace h2b {

struct ColSpan {
    uint64_t start;  // s_c
    uint64_t len;    // cells in column c
};

__global__ void __launch_bounds__(256) k_assign_columns(const uint4* __restrict__ vcol, const ColSpan* __restrict__ spans,
                                                        u32 rows_log, u32 ncols, uint4* __restrict__ cols) {
    // one thread per 16-byte half cell: a warp moves 512 contiguous bytes
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = ((size_t)ncols << rows_log) * 2;
    if (g >= total) return;
    size_t cell = g >> 1;
    u32 half = (u32)(g & 1);
    u32 c = (u32)(cell >> rows_log);
    size_t r = cell & (((size_t)1 << rows_log) - 1);
    ColSpan sp = spans[c];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < sp.len) v = __ldg(vcol + 2 * (sp.start + r) + half);
    cols[g] = v;
}

struct ColSpans64 {
    ColSpan s[64];
};
// same gather with the spans passed by value (no staging copy, no host synchronisation): up to 64 columns
__global__ void __launch_bounds__(256) k_assign_columns_v(const uint4* __restrict__ vcol, ColSpans64 spans, u32 rows_log, u32 ncols,
                                                          uint4* __restrict__ cols) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = ((size_t)ncols << rows_log) * 2;
    if (g >= total) return;
    size_t cell = g >> 1;
    u32 half = (u32)(g & 1);
    u32 c = (u32)(cell >> rows_log);
    size_t r = cell & (((size_t)1 << rows_log) - 1);
    const ColSpan sp = spans.s[c];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < sp.len) v = __ldg(vcol + 2 * (sp.start + r) + half);
    cols[g] = v;
}

// `Assigned<Fr>` staging records (72 bytes: tag, numerator, denominator; halo2-base/src/lib.rs:157-188 re-exports the
// prover crate's Assigned::{Zero, Trivial(F), Rational(F, F)}): split into a numerator and a denominator array.
// tag 0 -> (0, 1), 1 -> (num, 1), 2 -> (num, den).  stats[0] counts the Rational cells, stats[1] the invalid tags.
__global__ void __launch_bounds__(256) k_assigned_split(const uint64_t* __restrict__ recs, size_t N, uint64_t* __restrict__ num,
                                                        uint64_t* __restrict__ den, u32* __restrict__ stats) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint64_t* r = recs + 9 * i;
    const uint64_t tag = __ldg(r);
    Fr nu = Fr::zero(), de = Fr::one();
    if (tag == 1 || tag == 2) {
        uint64_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = __ldg(r + 1 + j);
#pragma unroll
        for (int j = 0; j < 4; j++) { nu.l[2 * j] = (u32)w[j]; nu.l[2 * j + 1] = (u32)(w[j] >> 32); }
    }
    if (tag == 2) {
        uint64_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = __ldg(r + 5 + j);
#pragma unroll
        for (int j = 0; j < 4; j++) { de.l[2 * j] = (u32)w[j]; de.l[2 * j + 1] = (u32)(w[j] >> 32); }
        atomicAdd(stats, 1u);
    }
    if (tag > 2) atomicAdd(stats + 1, 1u);
    nu.store(num + 4 * i);
    de.store(den + 4 * i);
}

__global__ void __launch_bounds__(256) k_assign_lookups(const uint4* __restrict__ vals, size_t N, u32 rows_log, u32 L,
                                                        uint4* __restrict__ cols) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = ((size_t)L << rows_log) * 2;
    if (g >= total) return;
    size_t cell = g >> 1;
    u32 half = (u32)(g & 1);
    u32 c = (u32)(cell >> rows_log);
    size_t r = cell & (((size_t)1 << rows_log) - 1);
    size_t j = r * L + c;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < N) v = __ldg(vals + 2 * j + half);
    cols[g] = v;
}

// Assigned::Rational(num, den) -> num * den^-1 (den = 0 -> 0), what batch_invert_assigned produces
__global__ void __launch_bounds__(128) k_eval_rational(const uint64_t* __restrict__ num, const uint64_t* __restrict__ den,
                                                       u32 n, uint64_t* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr d = Fr::load_nc(den + 4 * (size_t)i).inv();
    (Fr::load_nc(num + 4 * (size_t)i) * d).store(out + 4 * (size_t)i);
}

void assign_columns_run(h2b_ctx* ctx, const void* d_vcol, size_t N, const uint64_t* break_points, size_t nbp,
                        uint32_t k, size_t ncols, void* d_cols) {
    H2B_REQUIRE(k <= 28, "assign: k out of range");
    const size_t rows = (size_t)1 << k;
    if (ncols == 0) {
        // single_phase.rs:279-286: "Trying to assign threads in a phase with no columns"
        if (N != 0) throw StatusError{H2B_ERR_LAYOUT, "assign_witnesses: cells present but the phase has no advice columns"};
        return;
    }
    std::vector<ColSpan> spans(ncols, ColSpan{0, 0});
    size_t s = 0, c = 0, bpi = 0;
    size_t rem = N;
    while (rem > 0) {
        if (c >= ncols)
            throw StatusError{H2B_ERR_LAYOUT, "assign_witnesses: break points walk past the last advice column (single_phase.rs:304)"};
        // The walk compares row_offset with the break point only for cells it assigns in its main step; the
        // duplicate written at row 0 of a new column is not compared, so in columns c > 0 a break point of 0
        // can never fire (and, never being consumed, disables all later ones).
        const bool can_break = bpi < nbp && rem > break_points[bpi] && (c == 0 || break_points[bpi] >= 1);
        if (can_break) {
            size_t b = (size_t)break_points[bpi++];
            if (b + 1 > rows) throw StatusError{H2B_ERR_LAYOUT, "assign_witnesses: break point beyond the 2^k rows of a column"};
            spans[c] = ColSpan{s, b + 1};
            s += b;
            rem -= b;  // the duplicated cell + the rest
            c++;
            if (c >= ncols)
                throw StatusError{H2B_ERR_LAYOUT, "assign_witnesses: break points walk past the last advice column (single_phase.rs:304)"};
        } else {
            if (rem > rows) throw StatusError{H2B_ERR_LAYOUT, "assign_witnesses: column overflows 2^k rows"};
            spans[c] = ColSpan{s, rem};
            rem = 0;
        }
    }
    size_t total = ncols * rows * 2;
    if (ncols <= 64) {  // the usual case: spans travel as a kernel argument, the call stays asynchronous
        ColSpans64 sv;
        memset(&sv, 0, sizeof(sv));
        for (size_t i = 0; i < ncols; i++) sv.s[i] = spans[i];
        H2B_LAUNCH(ctx, k_assign_columns_v, ceil_div(total, 256), 256, 0, (const uint4*)d_vcol, sv, k, (u32)ncols, (uint4*)d_cols);
        return;
    }
    ColSpan* d_spans = (ColSpan*)ctx->get(WS_MISC, ncols * sizeof(ColSpan));
    ColSpan* h_spans = (ColSpan*)ctx->get_pinned(1, ncols * sizeof(ColSpan));
    H2B_CUDA(cudaStreamSynchronize(ctx->stream));  // > 64 columns: the pinned staging block may still be in flight
    memcpy(h_spans, spans.data(), ncols * sizeof(ColSpan));
    H2B_CUDA(cudaMemcpyAsync(d_spans, h_spans, ncols * sizeof(ColSpan), cudaMemcpyHostToDevice, ctx->stream));
    H2B_LAUNCH(ctx, k_assign_columns, ceil_div(total, 256), 256, 0, (const uint4*)d_vcol, d_spans, k, (u32)ncols, (uint4*)d_cols);
}

// d_recs: N staging records of 72 bytes.  d_values (N x 32 B) receives what the prover's `batch_invert_assigned` yields:
// Zero -> 0, Trivial(x) -> x, Rational(n, d) -> n / d (d = 0 -> 0).  `stats` (device, 2 x u32, zeroed here): see above.
// invert: 0 = never (the caller knows there is no Rational cell), 1 = always (asynchronous).
void assigned_flatten_run(h2b_ctx* ctx, const void* d_recs, size_t N, void* d_values, u32* d_stats, int invert) {
    if (N == 0) return;
    uint64_t* den = (uint64_t*)ctx->get(WS_PROD, N * 32);
    H2B_CUDA(cudaMemsetAsync(d_stats, 0, 8, ctx->stream));
    H2B_LAUNCH(ctx, k_assigned_split, ceil_div(N, 256), 256, 0, (const uint64_t*)d_recs, N, (uint64_t*)d_values, den, d_stats);
    if (invert) {
        batch_invert_run(ctx, den, N);
        fr_mul_elementwise_run(ctx, d_values, den, N, d_values);
    }
}

void assign_lookups_run(h2b_ctx* ctx, const void* d_vals, size_t N, uint32_t k, size_t L, void* d_cols) {
    H2B_REQUIRE(k <= 28, "assign: k out of range");
    const size_t rows = (size_t)1 << k;
    if (L == 0) {
        if (N != 0) throw StatusError{H2B_ERR_LAYOUT, "assign_lookups: values present but no lookup advice columns (builder.rs:366)"};
        return;
    }
    if ((N + L - 1) / L > rows) throw StatusError{H2B_ERR_LAYOUT, "assign_lookups: range lookups would be assigned to unusable rows (builder.rs:368-372)"};
    size_t total = L * rows * 2;
    H2B_LAUNCH(ctx, k_assign_lookups, ceil_div(total, 256), 256, 0, (const uint4*)d_vals, N, k, (u32)L, (uint4*)d_cols);
}

void eval_rational_run(h2b_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_out) {
    if (n == 0) return;
    H2B_LAUNCH(ctx, k_eval_rational, ceil_div(n, 128), 128, 0, (const uint64_t*)d_num, (const uint64_t*)d_den, (u32)n, (uint64_t*)d_out);
}

}  // namespace h2b
