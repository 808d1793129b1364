// h2b200.hpp — C++ host-side mirror of the prover interfaces halo2-lib reaches through `halo2_base::halo2_proofs`
// (halo2-base/src/lib.rs:25-28) for the create_proof hot path, implemented on top of the C ABI in h2b200.h.
//
// The reference's host code is Rust (halo2-axiom 0.5.3 / halo2curves-axiom 0.7.3, not vendored; no Rust toolchain in
// this image), so this header restates the same operator surface in C++ with the same names, argument meaning
// and failure behaviour; the Rust binding a maintainer would add is shown in INTEGRATION.md.
//
//   halo2curves::msm::best_multiexp(coeffs, bases) -> G1            h2b::best_multiexp
//   halo2_proofs::arithmetic::best_fft(a, omega, log_n)             h2b::best_fft
//   halo2_proofs::poly::kzg::commitment::ParamsKZG::{commit, commit_lagrange}      h2b::ParamsKZG
//   halo2_proofs::poly::EvaluationDomain::{new, lagrange_to_coeff, coeff_to_lagrange, coeff_to_extended,
//                                           extended_to_coeff}                      h2b::EvaluationDomain
//   halo2_base::gates::flex_gate::threads::single_phase::assign_witnesses           h2b::assign_witnesses
//       (halo2-base/src/gates/flex_gate/threads/single_phase.rs:273-312)
//   halo2_base::virtual_region::lookups::LookupAnyManager::assign_raw               h2b::assign_lookups
//       (halo2-base/src/virtual_region/lookups.rs:130-155)
//
// Where the Rust code panics (`expect("prover should not fail")`, halo2-base/src/utils/testing.rs:48; index out of
// bounds in assign_witnesses) these wrappers throw h2b::Error; nothing is computed on the CPU.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "h2b200.h"

namespace h2b {

using Fr = std::array<uint64_t, 4>;        // [u64;4] LE Montgomery (halo2-base/src/utils/mod.rs:332-377)
using Fq = std::array<uint64_t, 4>;
struct G1Affine { Fq x, y; };              // identity = (0,0)
struct G1 { Fq x, y, z; };                 // Jacobian, identity z = 0
static_assert(sizeof(G1Affine) == 64 && sizeof(G1) == 96, "layout must match the C ABI");

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("h2b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = h2b_ctx_create(device, &ctx_);
        if (rc != H2B_OK) throw Error(rc, h2b_last_error(nullptr));
    }
    ~Context() { h2b_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    h2b_ctx* raw() const { return ctx_; }
    void check(int rc) const {
        if (rc != H2B_OK) throw Error(rc, h2b_last_error(ctx_));
    }
    void set_stream(void* cuda_stream) { check(h2b_ctx_set_stream(ctx_, cuda_stream)); }
    void synchronize() { check(h2b_ctx_synchronize(ctx_)); }
    G1 sum(const std::vector<G1>& pts) const {  // combine per-GPU partial commitments
        G1 out;
        check(h2b_g1_sum(ctx_, reinterpret_cast<const uint64_t*>(pts.data()), pts.size(), out.x.data()));
        return out;
    }
    void batch_normalize(std::vector<G1>& pts) const {
        check(h2b_g1_normalize(ctx_, reinterpret_cast<uint64_t*>(pts.data()), pts.size()));
    }

private:
    h2b_ctx* ctx_ = nullptr;
};

// best_multiexp(coeffs, bases): ad-hoc bases.  Rust asserts coeffs.len() == bases.len().
inline G1 best_multiexp(const Context& ctx, const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    if (coeffs.size() != bases.size()) throw Error(H2B_ERR_ARG, "best_multiexp: coeffs.len() != bases.len()");
    G1 out;
    ctx.check(h2b_msm_g1_bases(ctx.raw(), reinterpret_cast<const uint64_t*>(bases.data()),
                               reinterpret_cast<const uint64_t*>(coeffs.data()), coeffs.size(), out.x.data()));
    return out;
}

// best_fft(a, omega, log_n): in place.
inline void best_fft(const Context& ctx, std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != (size_t(1) << log_n)) throw Error(H2B_ERR_ARG, "best_fft: a.len() != 1 << log_n");
    ctx.check(h2b_ntt_fr(ctx.raw(), reinterpret_cast<uint64_t*>(a.data()), log_n, omega.data(), 0));
}

// The base arrays of ParamsKZG<Bn256> resident on one GPU (this rank's shard [begin, begin + count)).
class ParamsKZG {
public:
    ParamsKZG(const Context& ctx, uint32_t k, const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange,
              size_t begin = 0, size_t count = 0)
        : ctx_(ctx), k_(k), count_(count ? count : (size_t(1) << k) - begin) {
        const size_t n = size_t(1) << k;
        if ((!g.empty() && g.size() != n) || (!g_lagrange.empty() && g_lagrange.size() != n))
            throw Error(H2B_ERR_ARG, "ParamsKZG: base arrays must hold 2^k points");
        ctx.check(h2b_srs_upload(ctx.raw(), g.empty() ? nullptr : reinterpret_cast<const uint64_t*>(g.data()),
                                 g_lagrange.empty() ? nullptr : reinterpret_cast<const uint64_t*>(g_lagrange.data()), k,
                                 begin, count_, &srs_));
    }
    ~ParamsKZG() { h2b_srs_destroy(ctx_.raw(), srs_); }
    ParamsKZG(const ParamsKZG&) = delete;
    ParamsKZG& operator=(const ParamsKZG&) = delete;
    uint32_t k() const { return k_; }
    // ParamsKZG::commit(poly in coefficient form) / commit_lagrange(poly in Lagrange form)
    G1 commit(const std::vector<Fr>& poly) const { return commit_(H2B_BASIS_MONOMIAL, poly); }
    G1 commit_lagrange(const std::vector<Fr>& poly) const { return commit_(H2B_BASIS_LAGRANGE, poly); }
    // all independent commitments of one prover phase at once (basis per column: 0 monomial, 1 lagrange)
    std::vector<G1> commit_many(const std::vector<int>& basis, const std::vector<const std::vector<Fr>*>& polys) const {
        if (basis.size() != polys.size()) throw Error(H2B_ERR_ARG, "commit_many: basis.len() != polys.len()");
        std::vector<const uint64_t*> ptrs;
        for (auto* p : polys) {
            if (!p || p->size() != count_) throw Error(H2B_ERR_ARG, "commit_many: polynomial length != shard size");
            ptrs.push_back(reinterpret_cast<const uint64_t*>(p->data()));
        }
        std::vector<G1> out(polys.size());
        ctx_.check(h2b_msm_g1_batch(ctx_.raw(), srs_, basis.data(), ptrs.data(), polys.size(), count_,
                                    reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }

private:
    G1 commit_(int basis, const std::vector<Fr>& poly) const {
        G1 out;
        ctx_.check(h2b_msm_g1(ctx_.raw(), srs_, basis, reinterpret_cast<const uint64_t*>(poly.data()), poly.size(), out.x.data()));
        return out;
    }
    const Context& ctx_;
    uint32_t k_;
    size_t count_;
    h2b_srs* srs_ = nullptr;
};

// EvaluationDomain::new(j, k): j = cs.degree(); quotient_poly_degree = j - 1; extended_k = least e >= k with
// 2^e >= n * (j - 1).
class EvaluationDomain {
public:
    EvaluationDomain(const Context& ctx, uint32_t j, uint32_t k) : ctx_(ctx), k_(k), quotient_poly_degree_(j - 1) {
        extended_k_ = k;
        while ((uint64_t(1) << extended_k_) < (uint64_t(1) << k) * quotient_poly_degree_) extended_k_++;
        ctx.check(h2b_domain_omega(k, omega_.data()));
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    const Fr& get_omega() const { return omega_; }
    void lagrange_to_coeff(std::vector<Fr>& a) const {
        expect(a.size() == n(), "lagrange_to_coeff: a.len() != n");
        ctx_.check(h2b_lagrange_to_coeff(ctx_.raw(), reinterpret_cast<uint64_t*>(a.data()), k_));
    }
    void coeff_to_lagrange(std::vector<Fr>& a) const {
        expect(a.size() == n(), "coeff_to_lagrange: a.len() != n");
        ctx_.check(h2b_coeff_to_lagrange(ctx_.raw(), reinterpret_cast<uint64_t*>(a.data()), k_));
    }
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        expect(a.size() == n(), "coeff_to_extended: a.len() != n");
        std::vector<Fr> out(size_t(1) << extended_k_);
        ctx_.check(h2b_coeff_to_extended(ctx_.raw(), reinterpret_cast<const uint64_t*>(a.data()), a.size(), extended_k_,
                                         reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }
    std::vector<Fr> extended_to_coeff(std::vector<Fr> a) const {
        expect(a.size() == (size_t(1) << extended_k_), "extended_to_coeff: a.len() != extended_len");
        ctx_.check(h2b_extended_to_coeff(ctx_.raw(), reinterpret_cast<uint64_t*>(a.data()), extended_k_));
        a.resize(n() * quotient_poly_degree_);  // `a.values.truncate(n * quotient_poly_degree)`
        return a;
    }

private:
    size_t n() const { return size_t(1) << k_; }
    static void expect(bool ok, const char* msg) {
        if (!ok) throw Error(H2B_ERR_ARG, msg);
    }
    const Context& ctx_;
    uint32_t k_, extended_k_ = 0;
    uint64_t quotient_poly_degree_;
    Fr omega_{};
};

// assign_witnesses(threads, basic_gates, region, break_points): `threads[i]` = ctx.advice of the i-th Context
// (Trivial payloads); returns basic_gates.len() columns of 2^k rows.  Throws h2b::Error(H2B_ERR_LAYOUT) where Rust panics.
inline std::vector<std::vector<Fr>> assign_witnesses(const Context& ctx, const std::vector<std::vector<Fr>>& threads,
                                                     const std::vector<uint64_t>& break_points, uint32_t k, size_t num_columns) {
    std::vector<Fr> vcol;
    for (auto& t : threads) vcol.insert(vcol.end(), t.begin(), t.end());
    std::vector<Fr> flat(num_columns << k);
    ctx.check(h2b_assign_columns(ctx.raw(), reinterpret_cast<const uint64_t*>(vcol.data()), vcol.size(), break_points.data(),
                                 break_points.size(), k, num_columns, reinterpret_cast<uint64_t*>(flat.data())));
    std::vector<std::vector<Fr>> cols(num_columns);
    for (size_t c = 0; c < num_columns; c++) cols[c].assign(flat.begin() + (c << k), flat.begin() + ((c + 1) << k));
    return cols;
}

// LookupAnyManager::assign_raw: value j -> lookup column j % L, row j / L
inline std::vector<std::vector<Fr>> assign_lookups(const Context& ctx, const std::vector<Fr>& values, uint32_t k, size_t L) {
    std::vector<Fr> flat(L << k);
    ctx.check(h2b_assign_lookups(ctx.raw(), reinterpret_cast<const uint64_t*>(values.data()), values.size(), k, L,
                                 reinterpret_cast<uint64_t*>(flat.data())));
    std::vector<std::vector<Fr>> cols(L);
    for (size_t c = 0; c < L; c++) cols[c].assign(flat.begin() + (c << k), flat.begin() + ((c + 1) << k));
    return cols;
}

// arithmetic::eval_polynomial(poly, point)
inline Fr eval_polynomial(const Context& ctx, const std::vector<Fr>& poly, const Fr& point) {
    Fr out{};
    ctx.check(h2b_eval_polynomial(ctx.raw(), reinterpret_cast<const uint64_t*>(poly.data()), poly.size(),
                                  reinterpret_cast<const uint64_t*>(&point), reinterpret_cast<uint64_t*>(&out)));
    return out;
}

// arithmetic::kate_division(a, b): quotient of a(X) by (X - b), remainder dropped
inline std::vector<Fr> kate_division(const Context& ctx, const std::vector<Fr>& a, const Fr& b) {
    std::vector<Fr> q(a.empty() ? 0 : a.size() - 1);
    ctx.check(h2b_kate_division(ctx.raw(), reinterpret_cast<const uint64_t*>(a.data()), a.size(), reinterpret_cast<const uint64_t*>(&b),
                                reinterpret_cast<uint64_t*>(q.data())));
    return q;
}

}  // namespace h2b
