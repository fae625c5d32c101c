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
void sro_ring_mul(int ring, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, int threads);
int sro_matvec(int ring, const uint64_t* const* rows, size_t kappa, size_t ncols, const uint64_t* v, size_t vlen,
               uint64_t* out, int threads);
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
