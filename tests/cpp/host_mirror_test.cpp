// GPU parity test of the C++ host mirror (include/stark_rings.hpp) against the C oracle
// (oracle/libsr_oracle.so, the checker).  Built by tests/cpp/Makefile, run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "stark_rings.hpp"

extern "C" {
void sro_crt(int ring, uint64_t* buf, size_t n, int threads);
void sro_icrt(int ring, uint64_t* buf, size_t n, int threads);
void sro_ntt_mul(int ring, uint64_t* a, const uint64_t* b, size_t n, int threads);
void sro_addsub(int ring, int op, uint64_t* a, const uint64_t* b, size_t n);
void sro_sum(int ring, const uint64_t* in, size_t n, uint64_t* out);
void sro_ring_mul(int ring, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, int threads);
int sro_matvec(int ring, const uint64_t* const* rows, size_t kappa, size_t ncols, const uint64_t* v, size_t vlen,
               uint64_t* out, int threads);
int sro_sparse_matvec(int ring, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col_idx,
                      const uint64_t* vals, const uint64_t* v, size_t vlen, uint64_t* out);
int sro_matmat(int ring, const uint64_t* const* a_rows, size_t a_nrows, size_t a_ncols, const uint64_t* const* m_rows,
               size_t m_nrows, size_t m_ncols, uint64_t* const* out_rows);
void sro_scale(int ring, uint64_t* a, size_t n, const uint64_t* r);
void sro_serialize(int ring, const uint64_t* in, size_t n, unsigned char* out);
}
using namespace stark_rings;

static uint64_t rng_state = 0x5EED;
static uint64_t next64() {  // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class C>
static std::vector<uint64_t> rand_raw(size_t n) {
    std::vector<uint64_t> v(n * C::LIMBS);
    for (size_t i = 0; i < v.size(); i++) {
        uint64_t x = next64();
        if (C::ring == SR_GOLDILOCKS) { if (x >= 0xFFFFFFFF00000001ull) x -= 0xFFFFFFFF00000001ull; }
        else if (C::ring == SR_BABYBEAR) x %= 2013265921ull;
        else if (i % 4 == 3) x &= (1ull << 59) - 1;
        v[i] = x;
    }
    return v;
}
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s line %d: %s\n", name, __LINE__, #cond); return 1; } } while (0)

template <class C>
static int run(const char* name) {
    const size_t n = 777;
    auto a = rand_raw<C>(n), b = rand_raw<C>(n);
    // crt / icrt round trip and oracle parity
    auto want = a;
    sro_crt(C::ring, want.data(), n, 4);
    RqNTT<C> ntt = CRT<C>::elementwise_crt(RqPoly<C>(a));
    CHECK(ntt.limbs == want);
    RqPoly<C> back = ICRT<C>::elementwise_icrt(RqNTT<C>(ntt));
    CHECK(back.limbs == a);
    // slot-wise product
    auto want_nm = a;
    sro_ntt_mul(C::ring, want_nm.data(), b.data(), n, 4);
    CHECK((RqNTT<C>(a) * RqNTT<C>(b)).limbs == want_nm);
    // Add / Sub / Neg / Sum (ntt_form.rs:588-626, 640-654), both forms
    {
        auto wa = a, ws = a, wn = a;
        sro_addsub(C::ring, 0, wa.data(), b.data(), n);
        sro_addsub(C::ring, 1, ws.data(), b.data(), n);
        sro_addsub(C::ring, 2, wn.data(), wn.data(), n);
        std::vector<uint64_t> wsum(C::LIMBS);
        sro_sum(C::ring, a.data(), n, wsum.data());
        CHECK((RqNTT<C>(a) + RqNTT<C>(b)).limbs == wa);
        CHECK((RqPoly<C>(a) - RqPoly<C>(b)).limbs == ws);
        CHECK((-RqNTT<C>(a)).limbs == wn);
        CHECK(RqNTT<C>(a).sum().limbs == wsum);
        CHECK(RqNTT<C>::dimension() == C::D);
    }
    // fused ring product
    std::vector<uint64_t> want_rm(a.size());
    sro_ring_mul(C::ring, a.data(), b.data(), want_rm.data(), n, 4);
    CHECK((RqPoly<C>(a) * RqPoly<C>(b)).limbs == want_rm);
    // single element through CyclotomicConfig, and the length panic
    std::vector<uint64_t> one(a.begin(), a.begin() + C::LIMBS);
    C::crt_in_place(one.data(), one.size());
    CHECK(std::memcmp(one.data(), want.data(), C::LIMBS * 8) == 0);
    bool panicked = false;
    try { C::crt_in_place(a.data(), 2 * C::LIMBS); } catch (const LengthPanic&) { panicked = true; }
    CHECK(panicked);
    // mat-vec + DifferentLengths
    const size_t kappa = 3, m = 50;
    std::vector<RqNTT<C>> rows;
    std::vector<const uint64_t*> ptrs;
    for (size_t i = 0; i < kappa; i++) rows.emplace_back(rand_raw<C>(m));
    for (auto& r : rows) ptrs.push_back(r.limbs.data());
    RqNTT<C> v(rand_raw<C>(m));
    std::vector<uint64_t> want_y(kappa * C::LIMBS);
    CHECK(sro_matvec(C::ring, ptrs.data(), kappa, m, v.limbs.data(), m, want_y.data(), 4) == 0);
    Matrix<C> A(rows);
    CHECK(A.try_mul_vec(v).limbs == want_y);
    RqNTT<C> shortv(rand_raw<C>(m - 1));
    CHECK(!A.checked_mul_vec(shortv).has_value());
    bool err = false;
    try { A.try_mul_vec(shortv); } catch (const DifferentLengths& e) { err = (e.lhs == m && e.rhs == m - 1); }
    CHECK(err);
    // sparse mat-vec (CSR image of coeffs), mat-mat, scaling
    {
        const size_t nr = 9, nc = 7;
        std::vector<std::vector<std::pair<std::vector<uint64_t>, size_t>>> coeffs(nr);
        for (size_t i = 0; i < nr; i++)
            for (size_t e = 0; e < i % 4; e++) coeffs[i].push_back({rand_raw<C>(1), (size_t)(next64() % nc)});
        SparseMatrix<C> S(nr, nc, coeffs);
        RqNTT<C> x(rand_raw<C>(nc));
        std::vector<uint64_t> want_s(nr * C::LIMBS);
        CHECK(sro_sparse_matvec(C::ring, nr, nc, S.row_ptr.data(), S.col_idx.data(), S.vals.limbs.data(), x.limbs.data(),
                                nc, want_s.data()) == 0);
        CHECK(S.try_mul_vec(x).limbs == want_s);
        CHECK(!S.checked_mul_vec(shortv).has_value());
        RqNTT<C> r(rand_raw<C>(1));
        auto want_vals = S.vals.limbs;
        sro_scale(C::ring, want_vals.data(), S.vals.len(), r.limbs.data());
        S *= r;
        CHECK(S.vals.limbs == want_vals);
        // (kappa x m) * (m x 4)
        std::vector<RqNTT<C>> mrows;
        std::vector<const uint64_t*> mp;
        for (size_t k = 0; k < m; k++) mrows.emplace_back(rand_raw<C>(4));
        for (auto& rr : mrows) mp.push_back(rr.limbs.data());
        Matrix<C> M(mrows);
        std::vector<std::vector<uint64_t>> want_p(kappa, std::vector<uint64_t>(4 * C::LIMBS));
        std::vector<uint64_t*> wp;
        for (auto& w : want_p) wp.push_back(w.data());
        CHECK(sro_matmat(C::ring, ptrs.data(), kappa, m, mp.data(), m, 4, wp.data()) == 0);
        Matrix<C> P = A.try_mul_mat(M);
        for (size_t i = 0; i < kappa; i++) CHECK(P.vals[i].limbs == want_p[i]);
        CHECK(!M.checked_mul_mat(M).has_value());
    }
    // canonical serialization round trip against the oracle, InvalidData on 0xff.. bytes
    {
        RqPoly<C> pa(a);
        auto bytes = pa.serialize();
        std::vector<unsigned char> want_b(bytes.size());
        sro_serialize(C::ring, a.data(), n, want_b.data());
        CHECK(bytes == want_b);
        CHECK(RqPoly<C>::deserialize(bytes).limbs == a);
        auto bad = bytes;
        for (size_t i = 0; i < 32; i++) bad[i] = 0xff;
        bool invalid = false;
        try { RqPoly<C>::deserialize(bad); } catch (const InvalidData&) { invalid = true; }
        CHECK(invalid);
    }
    std::printf("ok %s\n", name);
    return 0;
}

int main() {
    int rc = 0;
    rc |= run<GoldilocksRingConfig>("goldilocks");
    rc |= run<BabyBearRingConfig>("babybear");
    rc |= run<StarkRingConfig>("stark_prime");
    std::printf(rc ? "FAILED\n" : "ALL OK\n");
    return rc;
}
