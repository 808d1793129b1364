// h2b200_prover.hpp — C++ host side of the RESIDENT prover: the data flow of halo2-axiom 0.5.3 `create_proof`
// (sole halo2-lib call site: halo2-base/src/utils/testing.rs:40-48) for the constraint system halo2-base builds, with
// every column kept in HBM behind `h2b_poly` handles between the phases.  It is the compiled twin of
// halo2-lib_b200/prover.py (same phases, same transcript, same order of commitments, evaluations and opening sets: the two
// produce identical bytes for the same inputs — tests/test_gpu_prover.py::test_cpp_prover_matches_python) and the shape a
// Rust `create_proof` over include/h2b200.h would take.
//
// Circuit shape (halo2-base `BaseCircuitParams`): A gate-advice columns a0..a{A-1} with selectors q{j} and the vertical
// gate q (a0 + a1 a2 - a3) (flex_gate/mod.rs:80-91); L lookup-advice columns l0..l{L-1} looked up in `table` as they are
// (range/mod.rs:131-150), or with L = 0 the selector lookup q_lookup * a0 (range/mod.rs:92-94), or no lookup; one constants
// column c; equality on [c, a0.., l0..].  Degree 5 / 4 / 3, permutation sets of degree - 2 columns, degree - 1 pieces of h.
//
// The host does what the Rust side does: the Blake2b transcript, the challenges, the blinding scalars and a handful of
// 254-bit modular operations on them (`HostFr`); no polynomial arithmetic happens here.
#pragma once
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <utility>

#include "h2b200.hpp"

namespace h2b {

// ------------------------------------------------------------------------------------------------ host-side fields (Montgomery)
// a few dozen multiplications per proof: challenges, rotations of the evaluation point, powers of v and mu (Fr); one
// inversion per commitment to bring it to affine form before it enters the transcript (Fq) — what the Rust side does on the CPU
struct FrHostParams {
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr uint64_t R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
    static constexpr uint64_t INV = 0xc2e1f593efffffffULL;  // -r^-1 mod 2^64
};
struct FqHostParams {
    static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr uint64_t R1[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
    static constexpr uint64_t INV = 0x87d20782e4866389ULL;  // -p^-1 mod 2^64
};
template <class P>
struct HostField {
    using E = std::array<uint64_t, 4>;
    static E one() { return {P::R1[0], P::R1[1], P::R1[2], P::R1[3]}; }
    static bool is_zero(const E& a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
    static bool geq_mod(const uint64_t a[4]) {
        for (int i = 3; i >= 0; i--) {
            if (a[i] != P::MOD[i]) return a[i] > P::MOD[i];
        }
        return true;
    }
    static void sub_mod(uint64_t a[4]) {
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)a[i] - P::MOD[i] - (uint64_t)borrow;
            a[i] = (uint64_t)t;
            borrow = (t >> 64) & 1;
        }
    }
    static E mul(const E& a, const E& b) {  // Montgomery product (CIOS)
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 carry = 0;
            for (int j = 0; j < 4; j++) {
                unsigned __int128 cur = (unsigned __int128)a[j] * b[i] + t[j] + (uint64_t)carry;
                t[j] = (uint64_t)cur;
                carry = cur >> 64;
            }
            unsigned __int128 cur = (unsigned __int128)t[4] + (uint64_t)carry;
            t[4] = (uint64_t)cur;
            t[5] = (uint64_t)(cur >> 64);
            const uint64_t m = t[0] * P::INV;
            carry = ((unsigned __int128)m * P::MOD[0] + t[0]) >> 64;
            for (int j = 1; j < 4; j++) {
                cur = (unsigned __int128)m * P::MOD[j] + t[j] + (uint64_t)carry;
                t[j - 1] = (uint64_t)cur;
                carry = cur >> 64;
            }
            cur = (unsigned __int128)t[4] + (uint64_t)carry;
            t[3] = (uint64_t)cur;
            t[4] = t[5] + (uint64_t)(cur >> 64);
        }
        uint64_t r[4] = {t[0], t[1], t[2], t[3]};
        if (t[4] || geq_mod(r)) sub_mod(r);
        return {r[0], r[1], r[2], r[3]};
    }
    static E add(const E& a, const E& b) {
        uint64_t r[4];
        unsigned __int128 carry = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)a[i] + b[i] + (uint64_t)carry;
            r[i] = (uint64_t)t;
            carry = t >> 64;
        }
        if (carry || geq_mod(r)) sub_mod(r);
        return {r[0], r[1], r[2], r[3]};
    }
    static E from_canonical(const uint64_t c[4]) { return mul({c[0], c[1], c[2], c[3]}, {P::R2[0], P::R2[1], P::R2[2], P::R2[3]}); }
    static E pow(E base, uint64_t e) {
        E acc = one();
        while (e) {
            if (e & 1) acc = mul(acc, base);
            base = mul(base, base);
            e >>= 1;
        }
        return acc;
    }
    static E inv(const E& a) {  // a^(modulus - 2): Fermat, ~380 products (tens of microseconds on the host)
        uint64_t e[4] = {P::MOD[0] - 2, P::MOD[1], P::MOD[2], P::MOD[3]};  // the low limbs of both moduli are >= 2
        E acc = one();
        for (int limb = 3; limb >= 0; limb--)
            for (int bit = 63; bit >= 0; bit--) {
                acc = mul(acc, acc);
                if ((e[limb] >> bit) & 1) acc = mul(acc, a);
            }
        return acc;
    }
};
using HostFq = HostField<FqHostParams>;
struct HostFr : HostField<FrHostParams> {
    // 2^28-th root of unity 7^((r - 1) >> 28), canonical (halo2curves bn256::Fr::ROOT_OF_UNITY)
    static constexpr uint64_t ROOT28[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
    // 64 little-endian bytes -> the integer mod r, Montgomery form (the transcript's challenge)
    static Fr from_wide_bytes(const uint8_t d[64]) {
        uint64_t l[4], h[4];
        std::memcpy(l, d, 32);
        std::memcpy(h, d + 32, 32);
        // value = lo + hi 2^256:  mont(lo) = mul(lo, R2),  mont(hi 2^256) = mul(mul(hi, R2), R2)
        const Fr r2 = {FrHostParams::R2[0], FrHostParams::R2[1], FrHostParams::R2[2], FrHostParams::R2[3]};
        while (geq_mod(l)) sub_mod(l);  // Montgomery multiplication wants operands < r
        while (geq_mod(h)) sub_mod(h);
        return add(mul({l[0], l[1], l[2], l[3]}, r2), mul(mul({h[0], h[1], h[2], h[3]}, r2), r2));
    }
    static Fr omega(uint32_t k) {  // generator of the 2^k domain
        Fr w = from_canonical(ROOT28);
        for (uint32_t i = k; i < 28; i++) w = mul(w, w);
        return w;
    }
};
// Jacobian (X, Y, Z) -> (X / Z^2, Y / Z^3, 1), the identity -> all zero: the form in which a commitment enters the transcript
// and the proof (the accumulation order inside an MSM is not deterministic, the Jacobian representative therefore is not either)
inline G1 g1_normalize_host(const G1& p) {
    if (HostFq::is_zero(p.z)) return G1{{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    const Fq zi = HostFq::inv(p.z), zi2 = HostFq::mul(zi, zi);
    return G1{HostFq::mul(p.x, zi2), HostFq::mul(p.y, HostFq::mul(zi2, zi)), HostFq::one()};
}

// ------------------------------------------------------------------------------------------------ Blake2b-512 (RFC 7693)
class Blake2b {
public:
    Blake2b() {
        for (int i = 0; i < 8; i++) h_[i] = IV[i];
        h_[0] ^= 0x01010000ULL ^ 64;  // digest length 64, no key
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        while (len) {
            if (fill_ == 128) {  // the buffer is only compressed when more input follows (the last block is final)
                t_ += 128;
                compress(false);
                fill_ = 0;
            }
            const size_t take = std::min(len, size_t(128) - fill_);
            std::memcpy(buf_ + fill_, p, take);
            fill_ += take;
            p += take;
            len -= take;
        }
    }
    // digest of everything absorbed so far; the state is left untouched (as hashlib's digest())
    std::array<uint8_t, 64> digest() const {
        Blake2b c = *this;
        c.t_ += c.fill_;
        std::memset(c.buf_ + c.fill_, 0, 128 - c.fill_);
        c.compress(true);
        std::array<uint8_t, 64> out;
        std::memcpy(out.data(), c.h_, 64);
        return out;
    }

private:
    static constexpr uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(bool last) {
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        std::memcpy(m, buf_, 128);
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = IV[i]; }
        v[12] ^= t_;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
    uint64_t h_[8];
    uint64_t t_ = 0;
    uint8_t buf_[128] = {};
    size_t fill_ = 0;
};

// Blake2b over what the prover writes; squeeze() yields an Fr challenge (the host side of the transcript)
class Transcript {
public:
    void absorb(const void* data, size_t bytes) { h_.update(data, bytes); }
    Fr squeeze() {
        const auto d = h_.digest();
        const uint8_t zero = 0;
        h_.update(&zero, 1);
        return HostFr::from_wide_bytes(d.data());
    }

private:
    Blake2b h_;
};

// ------------------------------------------------------------------------------------------------ device polynomials
class Poly {
public:
    Poly(const Context& ctx, size_t n) : ctx_(&ctx), n_(n) {
        ctx.check(h2b_poly_alloc(ctx.raw(), n, &h_));
        ptr_ = static_cast<char*>(h2b_poly_device_ptr(h_));
    }
    ~Poly() { if (h_) h2b_poly_free(ctx_->raw(), h_); }
    Poly(const Poly&) = delete;
    Poly& operator=(const Poly&) = delete;
    void* at(size_t elem = 0) const { return ptr_ + 32 * elem; }
    size_t len() const { return n_; }
    h2b_poly* raw() const { return h_; }
    void upload(const Fr* host, size_t n, size_t offset = 0) const { ctx_->check(h2b_poly_upload(ctx_->raw(), h_, offset, host->data(), n)); }
    void upload_async(const Fr* pinned, size_t n, size_t offset = 0) const { ctx_->check(h2b_poly_upload_async(ctx_->raw(), h_, offset, pinned->data(), n)); }
    std::vector<Fr> download(size_t offset, size_t n) const {
        std::vector<Fr> out(n);
        if (n) ctx_->check(h2b_poly_download(ctx_->raw(), h_, offset, out[0].data(), n));
        return out;
    }

private:
    const Context* ctx_;
    size_t n_;
    h2b_poly* h_ = nullptr;
    char* ptr_ = nullptr;
};
using PolyPtr = std::unique_ptr<Poly>;

// a column: n rows of a device polynomial (an advice column is a slice of the block the assignment kernels write)
struct ColRef {
    const Poly* poly = nullptr;
    size_t offset = 0;
    void* ptr(size_t row = 0) const { return poly->at(offset + row); }
};

// ------------------------------------------------------------------------------------------------ the fixed side of a circuit
class ProverCircuit {
public:
    // fixed: Lagrange values (2^k each) by name — q0..q{A-1}, [q_lookup], [table], c; sigma: one column per permutation column
    // in the order [c, a0.., l0..]
    ProverCircuit(const Context& ctx, uint32_t k, size_t A, size_t L, bool selector_lookup, const std::map<std::string, std::vector<Fr>>& fixed,
                  const std::vector<std::vector<Fr>>& sigma)
        : ctx(ctx), k(k), n(size_t(1) << k), A(A), L(L), selector_lookup(selector_lookup && L == 0) {
        degree = L ? 4 : (this->selector_lookup ? 5 : 3);
        chunk = degree - 2;
        ext_k = k + (degree == 3 ? 1 : 2);
        bf = 6;  // max(3, queries of a gate column = 4) + 2
        u = n - (bf + 1);
        for (size_t j = 0; j < A; j++) adv_names.push_back("a" + std::to_string(j));
        for (size_t t = 0; t < L; t++) adv_names.push_back("l" + std::to_string(t));
        perm_cols.push_back("c");
        perm_cols.insert(perm_cols.end(), adv_names.begin(), adv_names.end());
        n_sets = (perm_cols.size() + chunk - 1) / chunk;
        n_lookups = L ? L : (this->selector_lookup ? 1 : 0);
        for (size_t j = 0; j < A; j++) fixed_names.push_back("q" + std::to_string(j));
        if (this->selector_lookup) fixed_names.push_back("q_lookup");
        if (n_lookups) fixed_names.push_back("table");
        fixed_names.push_back("c");
        if (sigma.size() != perm_cols.size()) throw Error(H2B_ERR_ARG, "ProverCircuit: one sigma column per permutation column");
        std::vector<Fr> l0(n, Fr{}), ll(n, Fr{}), la(n, Fr{});
        l0[0] = HostFr::one();
        ll[u] = HostFr::one();
        for (size_t i = 0; i < u; i++) la[i] = HostFr::one();
        auto add = [&](const std::string& name, const std::vector<Fr>& arr) {
            if (arr.size() != n) throw Error(H2B_ERR_ARG, "ProverCircuit: column " + name + " must hold 2^k rows");
            auto lg = std::make_unique<Poly>(ctx, n), cf = std::make_unique<Poly>(ctx, n), ex = std::make_unique<Poly>(ctx, size_t(1) << ext_k);
            lg->upload(arr.data(), n);
            cf->upload(arr.data(), n);
            ctx.check(h2b_lagrange_to_coeff_dev(ctx.raw(), cf->at(), k));
            ctx.check(h2b_coeff_to_extended_dev(ctx.raw(), cf->at(), n, ext_k, ex->at()));
            lagr[name] = std::move(lg);
            coeff[name] = std::move(cf);
            ext[name] = std::move(ex);
        };
        for (auto& nm : fixed_names) {
            auto it = fixed.find(nm);
            if (it == fixed.end()) throw Error(H2B_ERR_ARG, "ProverCircuit: missing fixed column " + nm);
            add(nm, it->second);
        }
        for (size_t i = 0; i < perm_cols.size(); i++) {
            sigma_names.push_back("sigma_" + perm_cols[i]);
            add(sigma_names.back(), sigma[i]);
        }
        add("l0", l0);
        add("l_last", ll);
        add("l_active", la);
        h2b_ctx_synchronize(ctx.raw());
        // gate programs: GATES_PER_PROGRAM vertical gates each (a program holds at most 64 calculations); every program continues
        // the Horner fold in y from the previous value, so the chain of programs is the one fold evaluate_h does
        for (size_t j0 = 0; j0 < A; j0 += GATES_PER_PROGRAM) {
            GateProgram gp;
            std::vector<ValueSource> parts;
            for (size_t j = j0; j < std::min(A, j0 + GATES_PER_PROGRAM); j++) {
                const uint32_t i = uint32_t(j - j0);
                auto adv = [&](int rot) { return gp.ev.add_calculation(Calculation::Store(ValueSource::Advice(i, gp.ev.add_rotation(rot)))); };
                const ValueSource q = gp.ev.add_calculation(Calculation::Store(ValueSource::Fixed(i, gp.ev.add_rotation(0))));
                const ValueSource a0 = adv(0), a1 = adv(1), a2 = adv(2), a3 = adv(3);
                const ValueSource prod = gp.ev.add_calculation(Calculation::Mul(a1, a2));
                const ValueSource sum = gp.ev.add_calculation(Calculation::Add(a0, prod));
                const ValueSource diff = gp.ev.add_calculation(Calculation::Sub(sum, a3));
                parts.push_back(gp.ev.add_calculation(Calculation::Mul(q, diff)));
                gp.cols.push_back(j);
            }
            gp.result = gp.ev.add_calculation(Calculation::Horner(ValueSource::PreviousValue(), parts, ValueSource::Y()));
            gate_programs.push_back(std::move(gp));
        }
        // lookup program: (compressed input + beta)(compressed table + gamma); one input / table expression each, so the
        // theta-compression is the expression itself
        if (n_lookups) {
            const uint32_t r0 = lookup_ev.add_rotation(0);
            ValueSource in, tab;
            if (this->selector_lookup) {  // fixed slots [q_lookup, table], advice slot [a0]
                const ValueSource q = lookup_ev.add_calculation(Calculation::Store(ValueSource::Fixed(0, r0)));
                const ValueSource a = lookup_ev.add_calculation(Calculation::Store(ValueSource::Advice(0, r0)));
                in = lookup_ev.add_calculation(Calculation::Mul(q, a));
                tab = lookup_ev.add_calculation(Calculation::Store(ValueSource::Fixed(1, r0)));
            } else {  // fixed slot [table], advice slot [l{t}]
                in = lookup_ev.add_calculation(Calculation::Store(ValueSource::Advice(0, r0)));
                tab = lookup_ev.add_calculation(Calculation::Store(ValueSource::Fixed(0, r0)));
            }
            const ValueSource rg = lookup_ev.add_calculation(Calculation::Add(tab, ValueSource::Gamma()));
            const ValueSource lb = lookup_ev.add_calculation(Calculation::Add(in, ValueSource::Beta()));
            lookup_result = lookup_ev.add_calculation(Calculation::Mul(lb, rg));
        }
    }

    static constexpr size_t GATES_PER_PROGRAM = 5;
    struct GateProgram {
        GraphEvaluator ev;
        ValueSource result{};
        std::vector<size_t> cols;
    };
    const Context& ctx;
    uint32_t k, ext_k = 0;
    size_t n, A, L;
    bool selector_lookup;
    size_t degree = 0, chunk = 0, n_sets = 0, n_lookups = 0, u = 0;
    uint32_t bf = 0;
    std::vector<std::string> adv_names, perm_cols, fixed_names, sigma_names;
    std::map<std::string, PolyPtr> lagr, coeff, ext;
    std::vector<GateProgram> gate_programs;
    GraphEvaluator lookup_ev;
    ValueSource lookup_result{};
};

// ------------------------------------------------------------------------------------------------ one proof
struct Proof {
    std::vector<G1> commitments;
    std::vector<std::pair<std::pair<std::string, int>, Fr>> evals;  // ((column, rotation), value) in query order
    Fr theta{}, beta{}, gamma{}, y{}, x{};
    size_t h2d_bytes = 0, d2h_bytes = 0;
};

class ProverSession {
public:
    // `blind(rows)`: the caller's source of blinding scalars (Montgomery limbs), called in the order the prover blinds its columns
    using BlindSource = std::function<std::vector<Fr>(size_t)>;

    ProverSession(const Context& ctx, const ParamsKZG& params, const ProverCircuit& cs) : ctx(ctx), params(params), cs(cs) {
        const size_t n = cs.n, ne = size_t(1) << cs.ext_k;
        v = std::make_unique<Poly>(ctx, n * cs.A);
        if (cs.L) lkv = std::make_unique<Poly>(ctx, n * cs.L);
        adv_block = std::make_unique<Poly>(ctx, n * (cs.A + cs.L));
        for (size_t j = 0; j < cs.adv_names.size(); j++) lagr[cs.adv_names[j]] = ColRef{adv_block.get(), j * n};
        std::vector<std::string> names = cs.adv_names;
        for (size_t t = 0; t < cs.n_lookups; t++)
            for (const char* p : {"pa", "ps", "zl"}) names.push_back(p + std::to_string(t));
        for (size_t s = 0; s < cs.n_sets; s++) names.push_back("zp" + std::to_string(s));
        for (auto& nm : names) {
            if (!lagr.count(nm)) lagr[nm] = ColRef{own(n), 0};
            coef[nm] = own(n);
            ext[nm] = own(ne);
        }
        if (cs.selector_lookup) inp = own(n);
        rnd = own(n);
        h = own(ne);
        for (auto& t : tmp) t = own(n);
        for (auto& t : tmp_side) t = own(n);
        d_out = own(48);
        d_status = own(std::max<size_t>(1, cs.n_lookups));
        zero = own(1);  // one zero element (never written)
    }

    // witness: the virtual column of the gate cells (Montgomery), break_points as keygen pinned them, lookup_cells in
    // assign_raw order (L > 0), random_poly: the vanishing argument's random polynomial (n coefficients)
    Proof create_proof(const std::vector<Fr>& witness, const std::vector<uint64_t>& break_points, const std::vector<Fr>& lookup_cells,
                       const std::vector<Fr>& random_poly, const BlindSource& blind) {
        const uint32_t k = cs.k, ext_k = cs.ext_k, bf = cs.bf;
        const size_t n = cs.n, u = cs.u, A = cs.A, L = cs.L;
        h2b_ctx* c = ctx.raw();
        Transcript tr;
        Proof res;
        auto commit = [&](const std::vector<std::pair<int, void*>>& items, bool absorb) {
            for (size_t lo = 0; lo < items.size(); lo += 16) {
                const size_t m = std::min<size_t>(16, items.size() - lo);
                std::vector<const void*> ptrs(m);
                std::vector<int> bs(m);
                for (size_t i = 0; i < m; i++) { bs[i] = items[lo + i].first; ptrs[i] = items[lo + i].second; }
                ctx.check(h2b_msm_g1_batch_dev(c, params.raw(), bs.data(), ptrs.data(), m, n, d_out->at()));
                std::vector<G1> out(m);
                ctx.check(h2b_poly_download(c, d_out->raw(), 0, out[0].x.data(), m * 3));
                res.d2h_bytes += m * 96;
                for (auto& pt : out) pt = g1_normalize_host(pt);  // affine form: what the transcript and the proof hold
                if (absorb) tr.absorb(out.data(), m * sizeof(G1));
                res.commitments.insert(res.commitments.end(), out.begin(), out.end());
            }
        };
        auto blind_col = [&](const ColRef& col, size_t first_row) {
            const std::vector<Fr> b = blind(n - first_row);
            if (b.size() != n - first_row) throw Error(H2B_ERR_ARG, "blind source returned the wrong number of rows");
            ctx.check(h2b_poly_upload(c, col.poly->raw(), col.offset + first_row, b[0].data(), b.size()));
            res.h2d_bytes += b.size() * 32;
        };
        auto side_transforms = [&](const std::vector<std::string>& names) {
            ctx.check(h2b_ctx_side_begin(c));
            try {
                for (auto& nm : names) {
                    ctx.check(h2b_poly_copy_dev(c, coef[nm]->at(), lagr[nm].ptr(), n));
                    ctx.check(h2b_lagrange_to_coeff_dev(c, coef[nm]->at(), k));
                    ctx.check(h2b_coeff_to_extended_dev(c, coef[nm]->at(), n, ext_k, ext[nm]->at()));
                }
            } catch (...) {
                h2b_ctx_side_end(c);
                throw;
            }
            ctx.check(h2b_ctx_side_end(c));
        };
        auto lincomb = [&](const std::vector<const void*>& ptrs, const std::vector<Fr>& scalars, Poly* out) {
            bool first = true;  // h2b_poly_lincomb takes at most 32 polynomials a call
            for (size_t lo = 0; lo < ptrs.size(); lo += 31) {
                std::vector<const void*> pp(ptrs.begin() + lo, ptrs.begin() + std::min(ptrs.size(), lo + 31));
                std::vector<Fr> sc(scalars.begin() + lo, scalars.begin() + std::min(ptrs.size(), lo + 31));
                if (!first) { pp.insert(pp.begin(), out->at()); sc.insert(sc.begin(), HostFr::one()); }
                ctx.check(h2b_poly_lincomb_dev(c, pp.data(), sc[0].data(), pp.size(), n, out->at()));
                first = false;
            }
        };

        // ---- phase 0: witness up, assignment, advice commitments (the random polynomial goes up beside it)
        v->upload(witness.data(), witness.size());
        res.h2d_bytes += witness.size() * 32;
        if (L) {
            lkv->upload(lookup_cells.data(), lookup_cells.size());
            res.h2d_bytes += lookup_cells.size() * 32;
        }
        ctx.check(h2b_ctx_side_begin(c));
        rnd->upload_async(random_poly.data(), n);
        ctx.check(h2b_ctx_side_end(c));
        res.h2d_bytes += n * 32;
        ctx.check(h2b_assign_columns_dev(c, v->at(), witness.size(), break_points.empty() ? nullptr : break_points.data(), break_points.size(), k, A,
                                         adv_block->at()));
        if (L) ctx.check(h2b_assign_lookups_dev(c, lkv->at(), lookup_cells.size(), k, L, adv_block->at(A * n)));
        std::vector<std::pair<int, void*>> items;
        for (auto& nm : cs.adv_names) {
            blind_col(lagr[nm], u);
            items.push_back({H2B_BASIS_LAGRANGE, lagr[nm].ptr()});
        }
        commit(items, true);
        res.theta = tr.squeeze();
        ctx.check(h2b_ctx_side_join(c));  // the random polynomial arrived while phase 0 ran
        side_transforms(cs.adv_names);
        // ---- lookups: compressed input, permuted pair (enqueue only; the verdict words are read after this phase's commitments)
        std::vector<void*> lk_in;
        items.clear();
        std::vector<std::string> perm_names;
        for (size_t t = 0; t < cs.n_lookups; t++) {
            const std::string ts = std::to_string(t);
            if (L == 0) {
                ctx.check(h2b_fr_mul_elementwise_dev(c, cs.lagr.at("q_lookup")->at(), lagr["a0"].ptr(), n, inp->at()));
                lk_in.push_back(inp->at());
            } else {
                lk_in.push_back(lagr["l" + ts].ptr());
            }
            const ColRef &pa = lagr["pa" + ts], &ps = lagr["ps" + ts];
            ctx.check(h2b_permute_expression_pair_async_dev(c, lk_in[t], cs.lagr.at("table")->at(), k, bf, pa.ptr(), ps.ptr(),
                                                            static_cast<uint32_t*>(d_status->at(t))));
            blind_col(pa, u);
            blind_col(ps, u);
            items.push_back({H2B_BASIS_LAGRANGE, pa.ptr()});
            items.push_back({H2B_BASIS_LAGRANGE, ps.ptr()});
            perm_names.push_back("pa" + ts);
            perm_names.push_back("ps" + ts);
        }
        if (cs.n_lookups) {
            commit(items, true);
            for (auto& w : d_status->download(0, cs.n_lookups))
                if (w[0]) throw Error(H2B_ERR_UNSATISFIED, "permute_expression_pair: an input value is not in the table (ConstraintSystemFailure)");
            res.d2h_bytes += 32 * cs.n_lookups;
        }
        res.beta = tr.squeeze();
        res.gamma = tr.squeeze();
        side_transforms(perm_names);
        // ---- product columns + the vanishing argument's random polynomial
        auto col_lagr = [&](const std::string& nm) -> void* { return nm == "c" ? cs.lagr.at("c")->at() : lagr[nm].ptr(); };
        for (size_t s = 0; s < cs.n_sets; s++) {
            std::vector<const void*> cols, sig;
            for (size_t i = s * cs.chunk; i < std::min(cs.perm_cols.size(), (s + 1) * cs.chunk); i++) {
                cols.push_back(col_lagr(cs.perm_cols[i]));
                sig.push_back(cs.lagr.at("sigma_" + cs.perm_cols[i])->at());
            }
            const void* start = s == 0 ? nullptr : lagr["zp" + std::to_string(s - 1)].ptr(u);  // chained through the previous set's closing value
            ctx.check(h2b_permutation_product_dev(c, cols.data(), sig.data(), cols.size(), s * cs.chunk, res.beta.data(), res.gamma.data(), k, bf, start,
                                                  lagr["zp" + std::to_string(s)].ptr()));
        }
        for (size_t t = 0; t < cs.n_lookups; t++) {
            const std::string ts = std::to_string(t);
            ctx.check(h2b_lookup_product_dev(c, lk_in[t], cs.lagr.at("table")->at(), lagr["pa" + ts].ptr(), lagr["ps" + ts].ptr(), res.beta.data(),
                                             res.gamma.data(), k, bf, lagr["zl" + ts].ptr()));
        }
        std::vector<std::string> prod_names;
        for (size_t s = 0; s < cs.n_sets; s++) prod_names.push_back("zp" + std::to_string(s));
        for (size_t t = 0; t < cs.n_lookups; t++) prod_names.push_back("zl" + std::to_string(t));
        items.clear();
        for (auto& nm : prod_names) {
            blind_col(lagr[nm], u + 1);
            items.push_back({H2B_BASIS_LAGRANGE, lagr[nm].ptr()});
        }
        side_transforms(prod_names);  // beside the commitments below
        items.push_back({H2B_BASIS_MONOMIAL, rnd->at()});
        commit(items, true);
        res.y = tr.squeeze();
        ctx.check(h2b_ctx_side_join(c));  // every column is now in coefficient and extended form
        // ---- quotient: gate, permutation and lookup terms folded with y on the extended coset
        Challenges ch;
        ch.beta = res.beta; ch.gamma = res.gamma; ch.theta = res.theta; ch.y = res.y;
        ctx.check(h2b_poly_zero(c, h->raw()));
        for (auto& gp : cs.gate_programs) {
            std::vector<const void*> fx, ad;
            for (size_t j : gp.cols) {
                fx.push_back(cs.ext.at("q" + std::to_string(j))->at());
                ad.push_back(ext["a" + std::to_string(j)]->at());
            }
            const h2b_graph g = bind(gp.ev, gp.result, fx, ad, ch);
            ctx.check(h2b_quotient_graph_dev(c, &g, k, ext_k, h->at()));
        }
        {
            std::vector<const void*> tz, tc, ts;
            for (size_t s = 0; s < cs.n_sets; s++) tz.push_back(ext["zp" + std::to_string(s)]->at());
            for (auto& nm : cs.perm_cols) {
                tc.push_back(nm == "c" ? cs.ext.at("c")->at() : ext[nm]->at());
                ts.push_back(cs.ext.at("sigma_" + nm)->at());
            }
            ctx.check(h2b_permutation_fold_dev(c, tz.data(), cs.n_sets, tc.data(), ts.data(), tc.size(), cs.chunk, cs.ext.at("l0")->at(),
                                               cs.ext.at("l_last")->at(), cs.ext.at("l_active")->at(), res.beta.data(), res.gamma.data(), res.y.data(), bf, k,
                                               ext_k, h->at()));
        }
        for (size_t t = 0; t < cs.n_lookups; t++) {
            const std::string ts = std::to_string(t);
            std::vector<const void*> fx, ad;
            if (L == 0) {
                fx = {cs.ext.at("q_lookup")->at(), cs.ext.at("table")->at()};
                ad = {ext["a0"]->at()};
            } else {
                fx = {cs.ext.at("table")->at()};
                ad = {ext["l" + ts]->at()};
            }
            const h2b_graph g = bind(cs.lookup_ev, cs.lookup_result, fx, ad, ch);
            ctx.check(h2b_lookup_fold_dev(c, &g, ext["zl" + ts]->at(), ext["pa" + ts]->at(), ext["ps" + ts]->at(), cs.ext.at("l0")->at(),
                                          cs.ext.at("l_last")->at(), cs.ext.at("l_active")->at(), k, ext_k, h->at()));
        }
        ctx.check(h2b_divide_by_vanishing_poly_dev(c, h->at(), k, ext_k));
        ctx.check(h2b_extended_to_coeff_dev(c, h->at(), ext_k));
        const size_t pieces = cs.degree - 1;
        items.clear();
        for (size_t j = 0; j < pieces; j++) items.push_back({H2B_BASIS_MONOMIAL, h->at(j * n)});
        commit(items, true);
        res.x = tr.squeeze();
        // ---- evaluations at x and its rotations
        const Fr w = HostFr::omega(k);
        auto rot = [&](int r) { return HostFr::mul(res.x, HostFr::pow(w, uint64_t(((r % (long long)n) + (long long)n) % (long long)n))); };
        const int last = -int(bf + 1);
        struct Query { std::string name; const void* ptr; int rot; };
        std::vector<Query> queries;
        for (size_t j = 0; j < A; j++)
            for (int r : {0, 1, 2, 3}) queries.push_back({"a" + std::to_string(j), coef["a" + std::to_string(j)]->at(), r});
        for (size_t t = 0; t < L; t++) queries.push_back({"l" + std::to_string(t), coef["l" + std::to_string(t)]->at(), 0});
        for (auto& nm : cs.fixed_names) queries.push_back({nm, cs.coeff.at(nm)->at(), 0});
        for (auto& nm : cs.sigma_names) queries.push_back({nm, cs.coeff.at(nm)->at(), 0});
        for (size_t s = 0; s < cs.n_sets; s++) {  // every set at x and omega x; all but the last one also at omega^last x
            const std::string nm = "zp" + std::to_string(s);
            queries.push_back({nm, coef[nm]->at(), 0});
            queries.push_back({nm, coef[nm]->at(), 1});
            if (s + 1 < cs.n_sets) queries.push_back({nm, coef[nm]->at(), last});
        }
        for (size_t t = 0; t < cs.n_lookups; t++) {
            const std::string ts = std::to_string(t);
            queries.push_back({"pa" + ts, coef["pa" + ts]->at(), 0});
            queries.push_back({"pa" + ts, coef["pa" + ts]->at(), -1});
            queries.push_back({"ps" + ts, coef["ps" + ts]->at(), 0});
            queries.push_back({"zl" + ts, coef["zl" + ts]->at(), 0});
            queries.push_back({"zl" + ts, coef["zl" + ts]->at(), 1});
        }
        for (size_t j = 0; j < pieces; j++) queries.push_back({"h" + std::to_string(j), h->at(j * n), 0});
        queries.push_back({"rnd", rnd->at(), 0});
        {
            const size_t m = queries.size();
            std::vector<const void*> polys(m);
            std::vector<Fr> xs(m), out(m);
            for (size_t i = 0; i < m; i++) { polys[i] = queries[i].ptr; xs[i] = rot(queries[i].rot); }
            ctx.check(h2b_eval_polynomial_batch_dev(c, polys.data(), xs[0].data(), m, n, out[0].data()));
            res.d2h_bytes += m * 32;
            tr.absorb(out.data(), m * sizeof(Fr));
            for (size_t i = 0; i < m; i++) res.evals.push_back({{queries[i].name, queries[i].rot}, out[i]});
        }
        // ---- SHPLONK-shaped opening: per rotation set sum_i v^i p_i, divided by (X - point) for every point of the set
        const Fr v_ch = tr.squeeze(), mu = tr.squeeze();
        std::vector<std::pair<const void*, std::vector<int>>> by_poly;  // first-appearance order
        for (auto& q : queries) {
            auto it = std::find_if(by_poly.begin(), by_poly.end(), [&](auto& e) { return e.first == q.ptr; });
            if (it == by_poly.end()) by_poly.push_back({q.ptr, {q.rot}});
            else it->second.push_back(q.rot);
        }
        std::vector<std::pair<std::vector<int>, std::vector<const void*>>> sets;
        for (auto& e : by_poly) {
            auto it = std::find_if(sets.begin(), sets.end(), [&](auto& s) { return s.first == e.second; });
            if (it == sets.end()) sets.push_back({e.second, {e.first}});
            else it->second.push_back(e.first);
        }
        std::stable_sort(sets.begin(), sets.end(), [](auto& a, auto& b) {
            if (a.first.size() != b.first.size()) return a.first.size() < b.first.size();
            return a.first < b.first;
        });
        auto run_sets = [&](const std::vector<size_t>& which, std::array<Poly*, 3> bufs) {
            Poly *f = bufs[0], *qd = bufs[1], *acc = bufs[2];
            bool first = true;
            for (size_t si : which) {
                auto& [rots, plist] = sets[si];
                std::vector<Fr> sc;
                for (size_t i = 0; i < plist.size(); i++) sc.push_back(HostFr::pow(v_ch, i));
                lincomb(plist, sc, f);
                Poly *src = f, *dst = qd;
                for (int r : rots) {  // successive divisions by (X - point): the quotient by the set's vanishing polynomial
                    const Fr z = rot(r);
                    ctx.check(h2b_kate_division_dev(c, src->at(), n, z.data(), dst->at()));
                    // kate_division writes the n - 1 quotient coefficients; the buffer is reused as an n-coefficient
                    // polynomial (next division, linear combination), so its top coefficient is cleared
                    ctx.check(h2b_poly_copy_dev(c, dst->at(n - 1), zero->at(), 1));
                    std::swap(src, dst);
                }
                const Fr mu_s = HostFr::pow(mu, si);
                if (first) lincomb({src->at()}, {mu_s}, acc);
                else lincomb({acc->at(), src->at()}, {HostFr::one(), mu_s}, acc);
                first = false;
            }
            return !first;
        };
        std::vector<size_t> side_sets, main_sets;  // the rotation sets are independent: every other one on the side queue
        for (size_t i = 0; i < sets.size(); i++) (i % 2 ? main_sets : side_sets).push_back(i);
        ctx.check(h2b_ctx_side_begin(c));
        try {
            run_sets(side_sets, {tmp_side[0], tmp_side[1], tmp_side[2]});
        } catch (...) {
            h2b_ctx_side_end(c);
            throw;
        }
        ctx.check(h2b_ctx_side_end(c));
        const bool have_main = run_sets(main_sets, {tmp[0], tmp[1], tmp[2]});
        ctx.check(h2b_ctx_side_join(c));
        if (have_main) lincomb({tmp[2]->at(), tmp_side[2]->at()}, {HostFr::one(), HostFr::one()}, tmp[2]);
        else ctx.check(h2b_poly_copy_dev(c, tmp[2]->at(), tmp_side[2]->at(), n));
        commit({{H2B_BASIS_MONOMIAL, tmp[2]->at()}}, true);
        const Fr u_ch = tr.squeeze();
        // final quotient: W' = L / (X - u) (the remainder is dropped by kate_division)
        ctx.check(h2b_kate_division_dev(c, tmp[2]->at(), n, u_ch.data(), tmp[3]->at()));
        ctx.check(h2b_poly_copy_dev(c, tmp[3]->at(n - 1), zero->at(), 1));
        commit({{H2B_BASIS_MONOMIAL, tmp[3]->at()}}, false);
        return res;
    }

private:
    Poly* own(size_t m) {
        owned.push_back(std::make_unique<Poly>(ctx, m));
        return owned.back().get();
    }
    // the arrays an h2b_graph points to live in `hold` until the next bind()
    h2b_graph bind(const GraphEvaluator& ev, ValueSource result, const std::vector<const void*>& fixed, const std::vector<const void*>& advice,
                   const Challenges& ch) {
        hold_prog = ev.program();
        hold_fixed = fixed;
        hold_advice = advice;
        h2b_graph g{};
        g.program = hold_prog.data();
        g.program_words = hold_prog.size();
        g.n_calculations = uint32_t(ev.calculations.size());
        g.result = result.word();
        g.constants = reinterpret_cast<const uint64_t*>(ev.constants.data());
        g.n_constants = ev.constants.size();
        g.rotations = ev.rotations.data();
        g.n_rotations = ev.rotations.size();
        g.fixed = hold_fixed.data();
        g.n_fixed = hold_fixed.size();
        g.advice = hold_advice.data();
        g.n_advice = hold_advice.size();
        std::copy(ch.beta.begin(), ch.beta.end(), g.beta);
        std::copy(ch.gamma.begin(), ch.gamma.end(), g.gamma);
        std::copy(ch.theta.begin(), ch.theta.end(), g.theta);
        std::copy(ch.y.begin(), ch.y.end(), g.y);
        return g;
    }

    const Context& ctx;
    const ParamsKZG& params;
    const ProverCircuit& cs;
    std::vector<PolyPtr> owned;
    PolyPtr v, lkv, adv_block;
    std::map<std::string, ColRef> lagr;
    std::map<std::string, Poly*> coef, ext;
    Poly *inp = nullptr, *rnd = nullptr, *h = nullptr, *d_out = nullptr, *d_status = nullptr, *zero = nullptr;
    std::array<Poly*, 4> tmp{};
    std::array<Poly*, 3> tmp_side{};
    std::vector<uint32_t> hold_prog;
    std::vector<const void*> hold_fixed, hold_advice;
};

}  // namespace h2b
