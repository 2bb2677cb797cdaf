// Test harness (built and run by tests/test_cabi_cpu.py::test_poseidon_ifma_dense_layer_edges): the AVX-512 IFMA dense layer of the
// product's host Poseidon (latticefold_b200/csrc/poseidon_ifma.cpp) against a plain 128-bit `%` evaluation of
// out[i] = sum_j M[i][j] st[j] mod p on random inputs and on rows / lanes at the edges of every limb and carry decision
// (0, 1, 2^26, 2^32, 2^52, p-1, p, p+1, 2^64-1, the offset constant of the fold, ...), including non-canonical state lanes.
// Prints "lanes <n> bad <m>" (or "no-ifma" on hosts without the instructions).
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
typedef uint64_t u64; typedef unsigned __int128 u128;
namespace lf { struct PoseidonIfmaMatrix { alignas(64) u64 limb[24][2][3][8]; }; bool poseidon_ifma_supported(); void poseidon_ifma_prepare(const u64*, PoseidonIfmaMatrix*); void poseidon_ifma_dense(const PoseidonIfmaMatrix*, u64*); }
static const u64 P = 0xFFFFFFFF00000001ULL;
static long bad = 0, total = 0;
static void check(const u64* m, const u64* st) {
    static lf::PoseidonIfmaMatrix M; lf::poseidon_ifma_prepare(m, &M);
    u64 out[24]; memcpy(out, st, sizeof out); lf::poseidon_ifma_dense(&M, out);
    for (int i = 0; i < 24; ++i) { u128 acc = 0; for (int j = 0; j < 24; ++j) acc = (acc + ((u128)m[i * 24 + j] * (st[j] % P)) % P) % P;
        ++total; if (out[i] != (u64)acc) { if (++bad < 10) printf("MISMATCH lane %d got %llx want %llx\n", i, (unsigned long long)out[i], (unsigned long long)(u64)acc); } }
}
int main() {
    if (!lf::poseidon_ifma_supported()) { printf("no-ifma\n"); return 0; }
    std::mt19937_64 rng(7); std::vector<u64> m(576), st(24);
    for (int it = 0; it < 100000; ++it) { for (auto& x : m) x = rng() % P; for (auto& x : st) x = rng(); if (it & 1) for (auto& x : st) x |= 0xFFFFFFFF00000000ull; if (it & 2) for (auto& x : m) x = P - 1 - (x & 0xFFFF); check(m.data(), st.data()); }
    const u64 cst = (1ULL << 40) + (16ULL << 32) - 16;
    std::vector<u64> em = {1, 2, P - 1, P - 2, 1ULL << 32, (1ULL << 32) - 1, (1ULL << 26) - 1, 1ULL << 26, 1ULL << 52, (1ULL << 52) - 1, P - (1ULL << 32), (P - 1) / 2};
    std::vector<u64> es = {0, 1, P - 1, P, P + 1, ~0ULL, ~0ULL - 1, 0xFFFFFFFF00000000ULL, (1ULL << 32) - 1, 1ULL << 32, 1ULL << 26, 1ULL << 52, cst, cst - 1, cst + 1, P - cst, P - cst - 1, P - cst + 1, 1ULL << 40, (1ULL << 63)};
    for (u64 d = 0; d < 4; ++d) { es.push_back(P - 2 - d); es.push_back(P + 2 + d); es.push_back((1ULL << 36) - 16 + d); es.push_back(P - (1ULL << 36) + d); }
    // single- and two-term rows: row i uses lanes (2k, 2k+1)
    std::vector<std::array<u64, 4>> combos;
    for (u64 a : em) for (u64 s : es) combos.push_back({a, s, 0, 0});
    for (u64 a : em) for (u64 s : es) for (u64 b : em) for (u64 t : es) combos.push_back({a, s, b, t});
    for (size_t c = 0; c < combos.size(); c += 12) {
        std::fill(m.begin(), m.end(), 0); std::fill(st.begin(), st.end(), 0);
        for (int k = 0; k < 12 && c + k < combos.size(); ++k) { auto& q = combos[c + k]; st[2 * k] = q[1]; st[2 * k + 1] = q[3]; m[(2 * k) * 24 + 2 * k] = q[0]; m[(2 * k) * 24 + 2 * k + 1] = q[2]; m[(2 * k + 1) * 24 + 2 * k] = q[2]; m[(2 * k + 1) * 24 + 2 * k + 1] = q[0]; }
        check(m.data(), st.data());
    }
    // full rows of extreme values
    for (u64 a : em) for (u64 s : es) { for (auto& x : m) x = a; for (auto& x : st) x = s; check(m.data(), st.data()); }
    printf("lanes %ld bad %ld\n", total, bad);
}
