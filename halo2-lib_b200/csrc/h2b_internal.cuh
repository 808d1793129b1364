// h2b_internal.cuh — context, workspace and launch helpers shared by the translation units of libh2b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <array>
#include <algorithm>

#include "../../include/h2b200.h"

namespace h2b {

struct StatusError {
    int code;
    std::string msg;
};

#define H2B_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw h2b::StatusError{_e == cudaErrorMemoryAllocation ? H2B_ERR_OOM : H2B_ERR_CUDA,    \
                                   std::string(#expr) + ": " + cudaGetErrorString(_e)};             \
    } while (0)
#define H2B_REQUIRE(cond, msg)                                                   \
    do {                                                                         \
        if (!(cond)) throw h2b::StatusError{H2B_ERR_ARG, std::string(msg)};      \
    } while (0)

// grow-only device workspace slots (freed with the context)
enum WsSlot {
    WS_SCALARS = 0,   // staged scalars of host-pointer MSM calls (two of them for double buffering)
    WS_SCALARS2,
    WS_KEYS_A,
    WS_KEYS_B,
    WS_VALS_A,
    WS_VALS_B,
    WS_SORT_TMP,
    WS_OFFSETS,
    WS_BUCKETS,
    WS_PARTIALS,
    WS_BIGLIST,
    WS_REDUCE_A,
    WS_REDUCE_B,
    WS_POOL,
    WS_POOL2,
    WS_OUT,
    WS_BASES,         // ad-hoc bases staging
    WS_NTT_A,
    WS_NTT_B,
    WS_NTT_C,
    WS_NTT_D,
    WS_NTT_E,         // small (2^k) buffers of the fused lagrange -> coeff -> extended batch
    WS_NTT_F,
    WS_NTT_G,
    WS_ASSIGN_IN,
    WS_ASSIGN_OUT,
    WS_MISC,
    WS_MISC2,
    WS_RED_A,         // batch-affine group sums (ping)
    WS_RED_B,         // batch-affine group sums (pong)
    WS_PROD,          // product-column scratch (numerators, denominators, power tables)
    WS_COUNT
};

struct NttPlan;

}  // namespace h2b

struct h2b_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t copy_stream2 = nullptr;  // device-to-host leg of the pipelined batch NTT entry points
    cudaStream_t side_stream = nullptr;   // h2b_ctx_side_begin / _end / _join: work that runs beside the main stream
    cudaStream_t side_saved = nullptr;    // the main stream while the side stream is current
    int side_saved_lane = 0;              // the side queue works in the workspace set of lane 1 (its own scratch buffers)
    cudaEvent_t side_ev[2] = {nullptr, nullptr};
    cudaEvent_t pipe_ev[3][3] = {};       // [buffer][uploaded, computed, downloaded]
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int sm_count = 148;
    mutable std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    bool ntt_attr_set = false;
    int opt_affine_levels = -1;  // h2b_ctx_set_option("msm.affine_levels"): -1 = default
    int opt_affine_k = -1;       // "msm.affine_k"
    int opt_affine_pt = -1;      // "msm.affine_per_thread_inverse": 1 = every thread inverts (safegcd), 0 = one inversion per tile
    int opt_tail_priority = -1;  // "msm.tail_priority": bucket reductions of lane MSMs on high-priority streams (-1 = default on)
    int opt_ntt_ctas = 0;        // "ntt.max_ctas_per_sm": 0 = as many as fit, 1 or 2 = background transform (see ntt_run)
    int opt_msm_group = 0;       // "msm.batch_group": MSMs of a batch call that share one sort / accumulate / reduce pipeline (0 = by size)
    int opt_lookup_backward = 0; // "lookup.leftover_order": 0 = front to back (PSE / axiom walk), 1 = zcash (pop from the back)
    void* peer = nullptr;  // PeerState (peer.cu): NVLink mailboxes of the multi-GPU all-reduce
    std::vector<h2b_ctx*> members;  // device group (h2b_ctx_create_multi): members[0] == this, the others are private
    bool reduce_counter_zeroed = false;
    void* reduce_counter_ptr = nullptr;
    struct Buf {
        void* p = nullptr;
        size_t cap = 0;
    };
    // MSM lanes: independent (stream, workspace) sets so that the latency-bound tail of one MSM (bucket
    // reduction: a few CTAs) overlaps the throughput-bound phases of the next MSM of the same batch
    static constexpr int NLANES = 3;
    cudaStream_t lane_stream[NLANES] = {nullptr, nullptr, nullptr};
    cudaStream_t lane_tail[NLANES] = {nullptr, nullptr, nullptr};     // high priority: the bucket reduction of the lane's MSM
    cudaEvent_t lane_acc[NLANES] = {nullptr, nullptr, nullptr};       // accumulation of the lane's MSM enqueued
    cudaEvent_t lane_tail_done[NLANES] = {nullptr, nullptr, nullptr};
    bool in_lane = false;                                             // the current stream is lane_stream[cur_lane]
    cudaEvent_t lane_done[NLANES] = {nullptr, nullptr, nullptr};
    cudaEvent_t lane_ready[NLANES] = {nullptr, nullptr, nullptr};     // staging buffer filled (host batch API)
    cudaEvent_t lane_consumed[NLANES] = {nullptr, nullptr, nullptr};  // staging buffer read by k_digits
    cudaEvent_t fork_ev = nullptr;
    int cur_lane = 0;  // workspace set used by get()
    Buf ws[NLANES][h2b::WS_COUNT];
    Buf pinned[3];
    std::map<std::array<uint64_t, 5>, h2b::NttPlan*> ntt_plans;

    // optional per-kernel device timing (h2b_profile_*): event pairs around launches whose name matches
    std::string prof_filter;  // empty = off, "*" = every kernel, else substring of the kernel name
    struct ProfRec {
        const char* name;
        cudaEvent_t a, b;
        cudaStream_t stream;
    };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    bool prof_match(const char* name) const {
        return !prof_filter.empty() && (prof_filter == "*" || std::string(name).find(prof_filter) != std::string::npos);
    }
    cudaEvent_t prof_event();

    // grow-only; growing synchronises the device first (a kernel may still read the old block)
    void* get(int slot, size_t bytes);
    void* get_pinned(int slot, size_t bytes);
};

struct h2b_srs {
    uint32_t k = 0;
    size_t begin = 0, count = 0;
    int c = 0, W = 0;
    void* table[2] = {nullptr, nullptr};  // [basis] -> W x count affine points: table[w*count + i] = 2^(c*w) * P_i
    std::vector<h2b_srs*> parts;          // device group: one shard handle per member device (tables above unused)
};

namespace h2b {

#define H2B_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                         \
        const bool _prof = (ctx)->prof_match(#kernel);                           \
        cudaEvent_t _ea = nullptr, _eb = nullptr;                                \
        if (_prof) {                                                             \
            _ea = (ctx)->prof_event();                                           \
            _eb = (ctx)->prof_event();                                           \
            H2B_CUDA(cudaEventRecord(_ea, (ctx)->stream));                       \
        }                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
        (ctx)->launches++;                                                       \
        H2B_CUDA(cudaGetLastError());                                            \
        if (_prof) {                                                             \
            H2B_CUDA(cudaEventRecord(_eb, (ctx)->stream));                       \
            (ctx)->prof_recs.push_back({#kernel, _ea, _eb, (ctx)->stream});                     \
        }                                                                        \
    } while (0)

static inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
static inline int ceil_log2(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

// ---- msm.cu
int msm_choose_c_fixed(size_t n);
void msm_build_table(h2b_ctx* ctx, const void* d_bases, size_t count, int c, int W, void* d_table);
// table mode: q = W (one bucket set), table = W x n affine; ad-hoc mode: q = 1, table = n affine
void msm_run(h2b_ctx* ctx, const void* d_table, size_t n, int c, int W, int q, const void* d_scalars, void* d_out,
             cudaEvent_t after_digits = nullptr);
// m MSMs of one size through ONE sort / accumulate / bucket-reduction pipeline (m <= 16; tabulated bases unless m == 1)
void msm_run_group(h2b_ctx* ctx, const void* const* d_tables, size_t n, int c, int W, int q, const void* const* d_scalars, size_t m,
                   void* d_out, cudaEvent_t after_digits = nullptr);
size_t msm_group_size(const h2b_ctx* ctx, size_t n, size_t m, int W);  // MSMs per pipeline for a batch of m
// m MSMs over tabulated bases: cut into groups (msm_run_group) that are dealt to the context's lanes; joins on ctx->stream
void msm_run_batch(h2b_ctx* ctx, const void* const* d_tables, size_t n, int c, int W, const void* const* d_scalars, size_t m, void* d_out);
void g1_sum_run(h2b_ctx* ctx, const void* d_points, size_t m, void* d_out);
void g1_normalize_run(h2b_ctx* ctx, void* d_points, size_t m);
void g1_fixed_base_mul_run(h2b_ctx* ctx, const uint64_t base_xy[8], const void* d_scalars, size_t n, void* d_out);
void field_op_run(h2b_ctx* ctx, int field, int op, const void* a, const void* b, size_t n, void* out);
// ---- ntt.cu
// in-place forward transform with arbitrary root; flags: see ntt.cu
void ntt_run(h2b_ctx* ctx, const void* d_src, size_t n_src, void* d_dst, uint32_t log_n, const uint64_t omega[4],
             int inverse_scale, int coset_mode);
void ntt_free_plans(h2b_ctx* ctx);
void domain_omega(uint32_t k, uint64_t out[4], bool inverse);
// ---- assign.cu
void assign_columns_run(h2b_ctx* ctx, const void* d_vcol, size_t N, const uint64_t* break_points, size_t nbp,
                        uint32_t k, size_t ncols, void* d_cols);
void assigned_flatten_run(h2b_ctx* ctx, const void* d_recs, size_t N, void* d_values, uint32_t* d_stats, int invert);
void assign_lookups_run(h2b_ctx* ctx, const void* d_vals, size_t N, uint32_t k, size_t L, void* d_cols);
void eval_rational_run(h2b_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_out);
// ---- peer.cu
void peer_create(h2b_ctx* ctx, int rank, int nranks, uint8_t* handle_out);
void peer_connect(h2b_ctx* ctx, const uint8_t* handles);
void peer_allreduce(h2b_ctx* ctx, void* d_points, size_t m);
void peer_destroy(h2b_ctx* ctx);
bool peer_connected(const h2b_ctx* ctx);
void peer_connect_local(const std::vector<h2b_ctx*>& members);  // in-process group: direct peer mappings
// ---- quotient.cu
void flex_gate_fold_run(h2b_ctx* ctx, const void* d_q_ext, const void* d_a_ext, const uint64_t y[4], uint32_t k, uint32_t ext_k,
                        void* d_acc);
void divide_by_vanishing_run(h2b_ctx* ctx, void* d_values, uint32_t k, uint32_t ext_k);
void quotient_graph_run(h2b_ctx* ctx, const h2b_graph* g, uint32_t k, uint32_t ext_k, void* d_values);
void lookup_fold_run(h2b_ctx* ctx, const h2b_graph* g, const void* d_z, const void* d_pin, const void* d_ptab, const void* d_l0,
                     const void* d_l_last, const void* d_l_active, uint32_t k, uint32_t ext_k, void* d_values);
void permutation_fold_run(h2b_ctx* ctx, const void* const* d_z, size_t n_sets, const void* const* d_columns, const void* const* d_sigma,
                          size_t n_cols, size_t chunk_len, const void* d_l0, const void* d_l_last, const void* d_l_active,
                          const uint64_t beta[4], const uint64_t gamma[4], const uint64_t y[4], uint32_t blinding_factors, uint32_t k,
                          uint32_t ext_k, void* d_values);
// ---- lookup.cu (returns true when an input value is missing from the table)
uint32_t* permute_expression_pair_enqueue(h2b_ctx* ctx, const void* d_input, const void* d_table, uint32_t k, uint32_t blinding_factors,
                                          void* d_permuted_input, void* d_permuted_table);  // enqueue only; returns the device verdict word
bool permute_expression_pair_run(h2b_ctx* ctx, const void* d_input, const void* d_table, uint32_t k, uint32_t blinding_factors,
                                 void* d_permuted_input, void* d_permuted_table);
// ---- srs.cu
void g_to_lagrange_run(h2b_ctx* ctx, const void* d_g, uint32_t k, void* d_g_lagrange);
void srs_setup_run(h2b_ctx* ctx, const uint64_t tau[4], const uint64_t base_xy[8], uint32_t k, void* d_g, void* d_g_lagrange);
size_t g1_count_off_curve_run(h2b_ctx* ctx, const void* d_points, size_t n);
size_t g1_decompress_run(h2b_ctx* ctx, const void* d_bytes, size_t n, void* d_out_xy);
// ---- poly.cu
void eval_polynomial_run(h2b_ctx* ctx, const void* d_a, size_t n, const uint64_t x[4], void* d_out);
void eval_polynomial_batch_run(h2b_ctx* ctx, const void* const* d_polys, const uint64_t* xs, size_t m, size_t n, void* d_out);
void kate_division_run(h2b_ctx* ctx, const void* d_a, size_t n, const uint64_t z[4], void* d_q);
void poly_lincomb_run(h2b_ctx* ctx, const void* const* d_polys, const uint64_t* scalars, size_t m, size_t n, void* d_out);
// ---- scan.cu
void batch_invert_run(h2b_ctx* ctx, void* d_a, size_t n);
// d_start != nullptr: the seed is read from device memory (the previous permutation set's closing value) instead of `start`
void grand_product_run(h2b_ctx* ctx, const void* d_f, const uint64_t start[4], size_t n, void* d_z, const void* d_start = nullptr);
void eval_rational_batched_run(h2b_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_out);
// ---- prover.cu
void permutation_product_run(h2b_ctx* ctx, const void* const* d_columns, const void* const* d_sigma, size_t n_cols, size_t first_col,
                             const uint64_t beta[4], const uint64_t gamma[4], uint32_t k, uint32_t blinding_factors,
                             const void* d_start, void* d_z);
void lookup_product_run(h2b_ctx* ctx, const void* d_in, const void* d_tab, const void* d_pin, const void* d_ptab, const uint64_t beta[4],
                        const uint64_t gamma[4], uint32_t k, uint32_t blinding_factors, void* d_z);
void fr_mul_elementwise_run(h2b_ctx* ctx, const void* d_a, const void* d_b, size_t n, void* d_out);

}  // namespace h2b
