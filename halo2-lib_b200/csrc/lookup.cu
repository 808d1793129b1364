// lookup.cu — the lookup argument's permuted columns for sm_100a:
// halo2-axiom 0.5.3 `plonk/lookup/prover.rs::permute_expression_pair` (not vendored; restated from the upstream
// algorithm).  Given the compressed input column A and table column S over the usable rows u = n - (blinding + 1):
//     A' = A sorted by Fr's `Ord` (integer order of the canonical value);
//     S'[row] = A'[row]                         where A' starts a new run (first occurrence of the value),
//     S'[row] = a left-over table value         elsewhere: the sorted table minus the first instance of every distinct
//                                               input value, in ascending order.
// WHICH repeated row receives which left-over value is not determined by the protocol, and the forks differ:
//   * the sorted-table walk of PSE halo2 >= 2023 and the forks derived from it — which is what halo2-axiom 0.5.3 is
//     recalled to carry (not vendored, cannot be confirmed here: DESIGN.md §2) — fills the unfilled rows FRONT TO BACK
//     (two cursors over the sorted table and the distinct inputs): the default here;
//   * zcash halo2 iterates a BTreeMap of left-over counts and pops the repeated rows from the BACK
//     (`repeated_input_rows.pop()`): h2b_ctx_set_option("lookup.leftover_order", 1).
// Either column satisfies the argument; only proof BYTES depend on the choice.
// An input value that is not in the table is `Error::ConstraintSystemFailure` (H2B_ERR_UNSATISFIED here).
//
// Sorting 254-bit keys: LSD radix sort of a row permutation over the four 64-bit limbs of the canonical values with
// cub::DeviceRadixSort (stable), skipping every limb that is constant over the column — range-check columns carry
// < 2^lookup_bits values, so one pass is the common case.  Everything else is flag / scan / scatter work, HBM-bound on
// 32-byte records.
#include <algorithm>
#include <cub/cub.cuh>

#include "h2b_internal.cuh"
#include "field.cuh"

namespace h2b {

struct Key256 {
    uint64_t l[4];
};
__device__ __forceinline__ int key_cmp(const Key256& a, const Key256& b) {
#pragma unroll
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}
__device__ __forceinline__ Key256 key_load(const uint64_t* p, size_t i) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p + 4 * i);
    ulonglong2 a = q[0], b = q[1];
    return Key256{{a.x, a.y, b.x, b.y}};
}

// canonical (non-Montgomery) values + which limbs vary over the column (bit i of *vary set when limb i is not constant)
__global__ void __launch_bounds__(256) k_canon(const uint64_t* __restrict__ src, u32 n, uint64_t* __restrict__ canon, u32* __restrict__ idx,
                                               u32* __restrict__ vary) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 mine = 0;
    if (i < n) {
        Fr v = Fr::load_nc(src + 4 * (size_t)i).from_mont();
        v.store(canon + 4 * (size_t)i);
        idx[i] = i;
        Fr first = Fr::load_nc(src).from_mont();
#pragma unroll
        for (int l = 0; l < 4; l++)
            if (v.l[2 * l] != first.l[2 * l] || v.l[2 * l + 1] != first.l[2 * l + 1]) mine |= 1u << l;
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicOr(vary, mine);
}
__global__ void __launch_bounds__(256) k_gather_limb(const uint64_t* __restrict__ canon, const u32* __restrict__ idx, u32 n, int limb,
                                                     uint64_t* __restrict__ keys) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = canon[4 * (size_t)idx[i] + limb];
}
__global__ void __launch_bounds__(256) k_gather_rows(const uint64_t* __restrict__ src, const uint64_t* __restrict__ canon,
                                                     const u32* __restrict__ idx, u32 n, uint64_t* __restrict__ out, uint64_t* __restrict__ out_canon) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t j = idx[i];
    Fr::load_nc(src + 4 * j).store(out + 4 * (size_t)i);
    Fr::load_nc(canon + 4 * j).store(out_canon + 4 * (size_t)i);
}

// is `key` present in the sorted column `col` (n canonical keys)?
__device__ __forceinline__ bool sorted_contains(const uint64_t* col, u32 n, const Key256& key) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (key_cmp(key_load(col, mid), key) < 0) lo = mid + 1;
        else hi = mid;
    }
    return lo < n && key_cmp(key_load(col, lo), key) == 0;
}

// rep[i] = 1 where A'[i] repeats A'[i-1]; left[j] = 1 where table value j is NOT consumed by a first occurrence;
// *missing |= 1 when a first occurrence is absent from the table
__global__ void __launch_bounds__(256) k_lookup_flags(const uint64_t* __restrict__ a_canon, const uint64_t* __restrict__ t_canon, u32 n,
                                                      u32* __restrict__ rep, u32* __restrict__ left, u32* __restrict__ missing) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key256 a = key_load(a_canon, i), t = key_load(t_canon, i);
    const bool a_first = i == 0 || key_cmp(key_load(a_canon, i - 1), a) != 0;
    const bool t_first = i == 0 || key_cmp(key_load(t_canon, i - 1), t) != 0;
    rep[i] = a_first ? 0u : 1u;
    if (a_first && !sorted_contains(t_canon, n, a)) atomicOr(missing, 1u);
    left[i] = (t_first && sorted_contains(a_canon, n, t)) ? 0u : 1u;
}
// rep_rows[rank] = row for the repeated rows, ascending
__global__ void __launch_bounds__(256) k_lookup_rep_rows(const u32* __restrict__ rep, const u32* __restrict__ rep_pos, u32 n, u32* __restrict__ rep_rows) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && rep[i]) rep_rows[rep_pos[i]] = i;
}
// S' : first occurrences copy A'; left-over number q (ascending) goes to the repeated row of rank q (front to back), or of
// rank R-1-q when `backward` (the zcash order)
__global__ void __launch_bounds__(256) k_lookup_fill(const uint64_t* __restrict__ a_sorted, const uint64_t* __restrict__ t_sorted,
                                                     const u32* __restrict__ rep, const u32* __restrict__ rep_pos, const u32* __restrict__ left,
                                                     const u32* __restrict__ left_pos, const u32* __restrict__ rep_rows, u32 n,
                                                     uint64_t* __restrict__ s_out, u32* __restrict__ missing, int backward) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 n_rep = rep_pos[n - 1] + rep[n - 1], n_left = left_pos[n - 1] + left[n - 1];
    if (n_rep != n_left) {  // cannot happen when every input value is in the table
        if (i == 0) atomicOr(missing, 2u);
        return;
    }
    if (!rep[i]) Fr::load_nc(a_sorted + 4 * (size_t)i).store(s_out + 4 * (size_t)i);
    if (left[i]) Fr::load_nc(t_sorted + 4 * (size_t)i).store(s_out + 4 * (size_t)rep_rows[backward ? n_rep - 1 - left_pos[i] : left_pos[i]]);
}

// sorts one column: out = src sorted by canonical value, out_canon = the canonical values in that order
static void sort_column(h2b_ctx* ctx, const uint64_t* d_src, u32 n, uint64_t* d_out, uint64_t* d_out_canon, char* scratch, size_t cub_bytes,
                        void* d_cub) {
    // scratch: canon (32 n) | keys_a (8 n) | keys_b (8 n) | idx_a (4 n) | idx_b (4 n) | vary (4)
    uint64_t* canon = (uint64_t*)scratch;
    uint64_t* keys_a = canon + 4 * (size_t)n;
    uint64_t* keys_b = keys_a + n;
    u32* idx_a = (u32*)(keys_b + n);
    u32* idx_b = idx_a + n;
    u32* vary = idx_b + n;
    H2B_CUDA(cudaMemsetAsync(vary, 0, 4, ctx->stream));
    H2B_LAUNCH(ctx, k_canon, ceil_div(n, 256), 256, 0, d_src, n, canon, idx_a, vary);
    u32* bounce = (u32*)ctx->get_pinned(0, 4096);
    H2B_CUDA(cudaMemcpyAsync(bounce, vary, 4, cudaMemcpyDeviceToHost, ctx->stream));
    H2B_CUDA(cudaStreamSynchronize(ctx->stream));
    const u32 v = bounce[0];
    for (int limb = 0; limb < 4; limb++) {
        if (!((v >> limb) & 1)) continue;
        H2B_LAUNCH(ctx, k_gather_limb, ceil_div(n, 256), 256, 0, canon, idx_a, n, limb, keys_a);
        size_t bytes = cub_bytes;
        H2B_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, bytes, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 64, ctx->stream));
        ctx->launches += 3;  // CUB: histogram + onesweep passes (approximate; counted so that gpu_launches is not under-reported)
        std::swap(idx_a, idx_b);
    }
    H2B_LAUNCH(ctx, k_gather_rows, ceil_div(n, 256), 256, 0, d_src, canon, idx_a, n, d_out, d_out_canon);
}

// returns true when some input value is missing from the table
bool permute_expression_pair_run(h2b_ctx* ctx, const void* d_input, const void* d_table, uint32_t k, uint32_t blinding_factors,
                                 void* d_permuted_input, void* d_permuted_table) {
    H2B_REQUIRE(k <= 28, "permute_expression_pair: k out of range");
    const size_t rows = (size_t)1 << k;
    H2B_REQUIRE((size_t)blinding_factors + 1 < rows, "permute_expression_pair: no usable rows");
    const u32 n = (u32)(rows - (blinding_factors + 1));
    size_t cub_sort = 0, cub_scan = 0;
    H2B_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (uint64_t*)nullptr, (uint64_t*)nullptr, (u32*)nullptr, (u32*)nullptr, (int)n, 0, 64,
                                             ctx->stream));
    H2B_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, cub_scan, (u32*)nullptr, (u32*)nullptr, (int)n, ctx->stream));
    const size_t cub_bytes = (std::max(cub_sort, cub_scan) + 255) & ~(size_t)255;
    const size_t sort_scratch = ((size_t)n * (32 + 8 + 8 + 4 + 4) + 4 + 255) & ~(size_t)255;
    // workspace: cub | sort scratch | a_canon | t_sorted | t_canon | rep, rep_pos, left, left_pos, rep_rows | missing
    const size_t total = cub_bytes + sort_scratch + 3 * (size_t)n * 32 + 5 * (size_t)n * 4 + 256;
    char* w = (char*)ctx->get(WS_SORT_TMP, total);
    void* d_cub = w;
    char* scratch = w + cub_bytes;
    uint64_t* a_canon = (uint64_t*)(scratch + sort_scratch);
    uint64_t* t_sorted = a_canon + 4 * (size_t)n;
    uint64_t* t_canon = t_sorted + 4 * (size_t)n;
    u32* rep = (u32*)(t_canon + 4 * (size_t)n);
    u32 *rep_pos = rep + n, *left = rep_pos + n, *left_pos = left + n, *rep_rows = left_pos + n, *missing = rep_rows + n;
    H2B_CUDA(cudaMemsetAsync(missing, 0, 4, ctx->stream));
    sort_column(ctx, (const uint64_t*)d_input, n, (uint64_t*)d_permuted_input, a_canon, scratch, cub_bytes, d_cub);
    sort_column(ctx, (const uint64_t*)d_table, n, t_sorted, t_canon, scratch, cub_bytes, d_cub);
    H2B_LAUNCH(ctx, k_lookup_flags, ceil_div(n, 256), 256, 0, a_canon, t_canon, n, rep, left, missing);
    size_t bytes = cub_bytes;
    H2B_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, bytes, rep, rep_pos, (int)n, ctx->stream));
    bytes = cub_bytes;
    H2B_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, bytes, left, left_pos, (int)n, ctx->stream));
    ctx->launches += 4;
    H2B_LAUNCH(ctx, k_lookup_rep_rows, ceil_div(n, 256), 256, 0, rep, rep_pos, n, rep_rows);
    H2B_LAUNCH(ctx, k_lookup_fill, ceil_div(n, 256), 256, 0, (const uint64_t*)d_permuted_input, t_sorted, rep, rep_pos, left, left_pos, rep_rows, n,
               (uint64_t*)d_permuted_table, missing, ctx->opt_lookup_backward);
    u32* bounce = (u32*)ctx->get_pinned(0, 4096);
    H2B_CUDA(cudaMemcpyAsync(bounce, missing, 4, cudaMemcpyDeviceToHost, ctx->stream));
    H2B_CUDA(cudaStreamSynchronize(ctx->stream));
    return bounce[0] != 0;
}

}  // namespace h2b
