// Exercises include/h2b200.hpp (the C++ mirror of the reference's prover interface) end to end on the GPU:
// two independent MSM paths agree (fixed-base table vs ad-hoc), NTT round trips, assignment layout, error mapping.
// Built and run by tests/test_cpp_mirror.py.  Exit code 0 = all checks passed.
#include <cstdio>
#include <cstring>
#include <random>
#include "../../include/h2b200.hpp"

using namespace h2b;
static int fails = 0;
#define CHECK(cond)                                                  \
    do {                                                             \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); fails++; } \
    } while (0)

static Fr small_mont(const Context& ctx, uint64_t v) {  // v -> Montgomery via the library's own to_mont hook
    Fr in{v, 0, 0, 0}, out{};
    ctx.check(h2b_test_field_op(ctx.raw(), 1, 5, in.data(), nullptr, 1, out.data()));
    return out;
}

int main() {
    Context ctx(0);
    const uint32_t k = 9;
    const size_t n = size_t(1) << k;
    // bases: (3 + 5 i) * G built by the library's fixed-base multiplication
    const uint64_t gxy[8] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL,
                             0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL};
    std::vector<Fr> idx(n);
    for (size_t i = 0; i < n; i++) idx[i] = small_mont(ctx, 3 + 5 * i);
    std::vector<G1Affine> g(n);
    ctx.check(h2b_g1_fixed_base_mul(ctx.raw(), gxy, reinterpret_cast<const uint64_t*>(idx.data()), n, reinterpret_cast<uint64_t*>(g.data())));
    std::vector<G1Affine> gl(g.rbegin(), g.rend());

    std::mt19937_64 rng(7);
    std::vector<Fr> s(n);
    for (auto& e : s) { e = {rng(), rng(), rng(), rng() & ((1ULL << 60) - 1)}; }
    s[0] = {0, 0, 0, 0};

    ParamsKZG params(ctx, k, g, gl);
    std::vector<G1> pts = {params.commit(s), best_multiexp(ctx, s, g), params.commit_lagrange(s), best_multiexp(ctx, s, gl)};
    auto many = params.commit_many({0, 1}, {&s, &s});
    pts.push_back(many[0]);
    pts.push_back(many[1]);
    ctx.batch_normalize(pts);
    CHECK(memcmp(&pts[0], &pts[1], sizeof(G1)) == 0);
    CHECK(memcmp(&pts[2], &pts[3], sizeof(G1)) == 0);
    CHECK(memcmp(&pts[0], &pts[4], sizeof(G1)) == 0);
    CHECK(memcmp(&pts[2], &pts[5], sizeof(G1)) == 0);
    CHECK(memcmp(&pts[0], &pts[2], sizeof(G1)) != 0);
    {   // a phase of 7 commitments over both bases: grouped pipelines (any group size) == one commitment at a time
        std::vector<std::vector<Fr>> colsv(7, s);
        for (size_t j = 0; j < colsv.size(); j++)
            for (size_t i = j; i < n; i += 3 + j) colsv[j][i] = small_mont(ctx, 1 + (i * (j + 1)) % 5);  // witness-like stretches
        std::fill(colsv[3].begin(), colsv[3].end(), Fr{0, 0, 0, 0});                                      // an all-zero column
        std::vector<int> basis = {0, 1, 1, 0, 1, 0, 1};
        std::vector<const std::vector<Fr>*> refs;
        for (auto& cv : colsv) refs.push_back(&cv);
        std::vector<G1> single;
        for (size_t j = 0; j < colsv.size(); j++) single.push_back(basis[j] ? params.commit_lagrange(colsv[j]) : params.commit(colsv[j]));
        ctx.batch_normalize(single);
        for (int64_t grp : {1, 3, 16, 0}) {
            ctx.set_option("msm.batch_group", grp);
            auto got = params.commit_many(basis, refs);
            ctx.batch_normalize(got);
            CHECK(memcmp(got.data(), single.data(), single.size() * sizeof(G1)) == 0);
        }
        bool rejected = false;
        try { ctx.set_option("msm.batch_group", 99); } catch (const Error& e) { rejected = (e.code == H2B_ERR_ARG); }
        CHECK(rejected);
    }

    EvaluationDomain dom(ctx, 5, k);
    CHECK(dom.extended_k() == k + 2);
    std::vector<Fr> a = s;
    dom.lagrange_to_coeff(a);
    std::vector<Fr> b = a;
    dom.coeff_to_lagrange(b);
    CHECK(b == s);
    std::vector<Fr> c = a;
    best_fft(ctx, c, dom.get_omega(), k);
    CHECK(c == s);
    auto ext = dom.coeff_to_extended(a);
    auto back = dom.extended_to_coeff(ext);
    CHECK(back.size() == 4 * n);
    CHECK(std::vector<Fr>(back.begin(), back.begin() + n) == a);
    for (size_t i = n; i < back.size(); i++) CHECK(back[i] == (Fr{0, 0, 0, 0}));

    // assignment: two threads, one break point at row 300 -> column 0 rows 0..300, column 1 starts with the duplicate
    std::vector<std::vector<Fr>> threads = {std::vector<Fr>(s.begin(), s.begin() + 200), std::vector<Fr>(s.begin() + 200, s.begin() + 450)};
    auto cols = assign_witnesses(ctx, threads, {300}, k, 2);
    CHECK(cols[0][300] == s[300] && cols[1][0] == s[300] && cols[1][1] == s[301] && cols[1][149] == s[449]);
    CHECK(cols[0][301] == (Fr{0, 0, 0, 0}) && cols[1][150] == (Fr{0, 0, 0, 0}));
    bool threw = false;
    try { assign_witnesses(ctx, threads, {300}, k, 1); } catch (const Error& e) { threw = (e.code == H2B_ERR_LAYOUT); }
    CHECK(threw);
    threw = false;
    try { best_multiexp(ctx, s, std::vector<G1Affine>(g.begin(), g.begin() + 3)); } catch (const Error& e) { threw = true; }
    CHECK(threw);
    auto lk = assign_lookups(ctx, std::vector<Fr>(s.begin(), s.begin() + 10), k, 3);
    CHECK(lk[0][0] == s[0] && lk[1][0] == s[1] && lk[2][0] == s[2] && lk[0][1] == s[3] && lk[0][3] == s[9]);

    // opening arithmetic: dividing by X (b = 0) shifts the coefficients, and then q(0) = a_1, a(0) = a_0
    Fr zero{0, 0, 0, 0};
    auto qd = kate_division(ctx, a, zero);
    CHECK(qd.size() == a.size() - 1 && qd[0] == a[1] && qd.back() == a.back());
    CHECK(eval_polynomial(ctx, a, zero) == a[0]);
    CHECK(eval_polynomial(ctx, qd, zero) == a[1]);

    // ---- the next rows through the mirror: batch inversion, the lookup permutation, a GraphEvaluator program
    std::vector<Fr> inv = a;
    inv[5] = zero;
    batch_invert(ctx, inv);
    std::vector<Fr> prod(n);
    ctx.check(h2b_test_field_op(ctx.raw(), 1, 0, reinterpret_cast<const uint64_t*>(a.data()), reinterpret_cast<const uint64_t*>(inv.data()), n,
                                reinterpret_cast<uint64_t*>(prod.data())));
    const Fr one = small_mont(ctx, 1);
    CHECK(prod[0] == one && prod[n - 1] == one && inv[5] == zero);
    auto zcol = grand_product(ctx, a, one);
    CHECK(zcol[0] == one && zcol[1] == a[0]);
    {   // table [4,9,9,2,7], inputs [9,2,9,9,4] over 5 usable rows of 2^3 (blinding_factors = 2): tests/golden/next_rows.json
        auto col = [&](std::initializer_list<uint64_t> v) {
            std::vector<Fr> c;
            for (auto x : v) c.push_back(small_mont(ctx, x));
            c.resize(8, zero);
            return c;
        };
        auto pr = permute_expression_pair(ctx, col({9, 2, 9, 9, 4}), col({4, 9, 9, 2, 7}), 3, 2);
        // left-over table values (7, 9) fill the repeated rows front to back by default, from the back with the zcash option
        CHECK(pr.first == col({2, 4, 9, 9, 9}) && pr.second == col({2, 4, 9, 7, 9}));
        ctx.check(h2b_ctx_set_option(ctx.raw(), "lookup.leftover_order", 1));
        pr = permute_expression_pair(ctx, col({9, 2, 9, 9, 4}), col({4, 9, 9, 2, 7}), 3, 2);
        CHECK(pr.first == col({2, 4, 9, 9, 9}) && pr.second == col({2, 4, 9, 9, 7}));
        ctx.check(h2b_ctx_set_option(ctx.raw(), "lookup.leftover_order", 0));
        threw = false;
        try { permute_expression_pair(ctx, col({9, 2, 9, 9, 5}), col({4, 9, 9, 2, 7}), 3, 2); } catch (const Error& e) { threw = e.code == H2B_ERR_UNSATISFIED; }
        CHECK(threw);
    }
    {   // halo2-base's gate q * (a + b*c - out) as a GraphEvaluator program == the dedicated kernel
        const uint32_t gk = 6, gek = 8;
        const size_t ge = size_t(1) << gek;
        std::vector<Fr> q(s.begin(), s.begin() + ge), adv(a.begin(), a.begin() + ge), acc1(b.begin(), b.begin() + ge), acc2 = acc1;
        GraphEvaluator ev;
        ValueSource qs = ev.add_calculation(Calculation::Store(ValueSource::Fixed(0, ev.add_rotation(0))));
        ValueSource a0 = ev.add_calculation(Calculation::Store(ValueSource::Advice(0, ev.add_rotation(0))));
        ValueSource a1 = ev.add_calculation(Calculation::Store(ValueSource::Advice(0, ev.add_rotation(1))));
        ValueSource a2 = ev.add_calculation(Calculation::Store(ValueSource::Advice(0, ev.add_rotation(2))));
        ValueSource a3 = ev.add_calculation(Calculation::Store(ValueSource::Advice(0, ev.add_rotation(3))));
        ValueSource bc = ev.add_calculation(Calculation::Mul(a1, a2));
        ValueSource sum = ev.add_calculation(Calculation::Add(a0, bc));
        ValueSource dif = ev.add_calculation(Calculation::Sub(sum, a3));
        ValueSource gate = ev.add_calculation(Calculation::Mul(qs, dif));
        ValueSource res = ev.add_calculation(Calculation::Horner(ValueSource::PreviousValue(), {gate}, ValueSource::Y()));
        Challenges ch;
        ch.y = s[7];
        quotient_graph(ctx, ev, res, {&q}, {&adv}, {}, ch, gk, gek, acc1);
        ctx.check(h2b_flex_gate_fold(ctx.raw(), reinterpret_cast<const uint64_t*>(q.data()), reinterpret_cast<const uint64_t*>(adv.data()),
                                     ch.y.data(), gk, gek, reinterpret_cast<uint64_t*>(acc2.data())));
        CHECK(acc1 == acc2);
        CHECK(ev.add_calculation(Calculation::Mul(a1, a2)) == bc);  // identical calculations are shared
        divide_by_vanishing_poly(ctx, acc1, gk, gek);
        CHECK(!(acc1 == acc2));
    }
    {   // g_to_lagrange of [G, G, G, G] at k = 2: only the constant Lagrange combination survives: out[0] = G, rest = identity
        G1Affine G;
        memcpy(&G, gxy, sizeof(G));
        auto gl4 = g_to_lagrange(ctx, std::vector<G1Affine>(4, G), 2);
        G1Affine idp{};
        CHECK(memcmp(&gl4[0], &G, sizeof(G)) == 0 && memcmp(&gl4[1], &idp, sizeof(G)) == 0 && memcmp(&gl4[3], &idp, sizeof(G)) == 0);
    }

    {   // one process, all GPUs of the box (h2b_ctx_create_multi): the sharded commitments equal the single-device ones and
        // the batched transforms dealt over the devices equal the single-device transforms
        int ndev = 0;
        for (int d = 0; d < 16; d++) {  // probe: a context on device d exists?
            h2b_ctx* probe = nullptr;
            if (h2b_ctx_create(d, &probe) != H2B_OK) break;
            h2b_ctx_destroy(probe);
            ndev++;
        }
        if (ndev >= 2) {
            std::vector<int> devs;
            for (int d = 0; d < (ndev > 4 ? 4 : ndev); d++) devs.push_back(d);
            Context grp(devs);
            CHECK(grp.device_count() == (int)devs.size());
            ParamsKZG gp(grp, k, g, gl);
            std::vector<G1> both = {params.commit(s), gp.commit(s), params.commit_lagrange(s), gp.commit_lagrange(s)};
            auto gm = gp.commit_many({0, 1, 0, 1, 0}, {&s, &s, &s, &s, &s});
            both.push_back(gm[0]);
            both.push_back(gm[3]);
            ctx.batch_normalize(both);
            CHECK(memcmp(&both[0], &both[1], sizeof(G1)) == 0);
            CHECK(memcmp(&both[2], &both[3], sizeof(G1)) == 0);
            CHECK(memcmp(&both[0], &both[4], sizeof(G1)) == 0);
            CHECK(memcmp(&both[2], &both[5], sizeof(G1)) == 0);
            // five columns through lagrange_to_coeff_batch on the group vs one at a time on device 0
            std::vector<std::vector<Fr>> cols(5, s);
            for (size_t j = 0; j < cols.size(); j++) cols[j][1] = small_mont(ctx, 100 + j);
            std::vector<std::vector<Fr>> want = cols;
            for (auto& c : want) dom.lagrange_to_coeff(c);
            std::vector<uint64_t*> ptrs;
            for (auto& c : cols) ptrs.push_back(reinterpret_cast<uint64_t*>(c.data()));
            grp.check(h2b_lagrange_to_coeff_batch(grp.raw(), ptrs.data(), ptrs.size(), k));
            for (size_t j = 0; j < cols.size(); j++) CHECK(cols[j] == want[j]);
            printf("host mirror: device group of %zu GPUs checked\n", devs.size());
        } else {
            printf("host mirror: single GPU, device-group checks skipped\n");
        }
    }

    printf(fails ? "host mirror: %d FAILED\n" : "host mirror: all checks passed\n", fails);
    return fails ? 1 : 0;
}
