// quotient.cu — first slice of the quotient evaluation h(X) on the extended domain:
// the custom-gate term of halo2-base's single vertical gate
//     q * (a + b * c - out)      with a, b, c, out = the SAME advice column at rotations 0, 1, 2, 3
// (halo2-base/src/gates/flex_gate/mod.rs:80-91), folded into the running combination the prover keeps per
// extended-domain row:  acc[i] <- acc[i] * y + q[i] * (a[i] + a[i + s] * a[i + 2s] - a[i + 3s]),
// s = 2^(extended_k - k) (a rotation by one row of the 2^k domain is a shift by s rows of the extended coset
// domain; indices wrap).  Pointwise: 3 products per row; HBM traffic 32 B x (q, acc in, acc out, a + 3 rotated reads
// that hit L2).
#include "h2b_internal.cuh"
#include "field.cuh"

namespace h2b {

__global__ void __launch_bounds__(256) k_flex_gate_fold(const uint64_t* __restrict__ q, const uint64_t* __restrict__ a, Fr y,
                                                        u32 ext_k, u32 shift, uint64_t* __restrict__ acc) {
    const size_t n = (size_t)1 << ext_k, mask = n - 1;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = (size_t)1 << shift;
    Fr a0 = Fr::load_nc(a + 4 * i), a1 = Fr::load_nc(a + 4 * ((i + s) & mask)), a2 = Fr::load_nc(a + 4 * ((i + 2 * s) & mask)),
       a3 = Fr::load_nc(a + 4 * ((i + 3 * s) & mask));
    Fr gate = Fr::load_nc(q + 4 * i) * (a0 + a1 * a2 - a3);
    (Fr::load(acc + 4 * i) * y + gate).store(acc + 4 * i);
}

void flex_gate_fold_run(h2b_ctx* ctx, const void* d_q_ext, const void* d_a_ext, const uint64_t y[4], uint32_t k, uint32_t ext_k,
                        void* d_acc) {
    H2B_REQUIRE(ext_k >= k && ext_k <= 28, "flex_gate: extended_k out of range");
    Fr yy;
    memcpy(&yy, y, sizeof(Fr));
    H2B_LAUNCH(ctx, k_flex_gate_fold, ceil_div((size_t)1 << ext_k, 256), 256, 0, (const uint64_t*)d_q_ext, (const uint64_t*)d_a_ext, yy,
               ext_k, ext_k - k, (uint64_t*)d_acc);
}

}  // namespace h2b
