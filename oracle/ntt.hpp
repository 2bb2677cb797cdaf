// TEST INFRASTRUCTURE ONLY.  CPU restatement of the batched negacyclic NTT over Z_p[X]/(X^N + 1) (BASELINE.json configs[4],
// SURVEY.md 8 row C5(b)).  The reference tree holds no transform of this shape (its NTT form is the CRT of the cyclotomic rings,
// restated in ring.hpp), so there is no reference KAT to pin against: "parity unpinned" with respect to the reference; the
// definition is pinned instead by the O(N^2) evaluation below and by the schoolbook product mod X^N + 1 (tests/test_ntt_oracle.py).
//
//   forward  A[k] = sum_j a[j] psi^(j(2k+1))      inverse  a[j] = N^-1 sum_k A[k] psi^(-j(2k+1))      natural order
//   psi_N = rho^(2^A / 2N);  rho = r0^u, r0 = g^((p-1)/2^A);  Goldilocks: g = 7, A = 32, u = smallest odd exponent with
//   rho^(2^27) = 64;  BabyBear: g = 31, A = 27, u = 1.
#pragma once
#include <cstdint>
#include <vector>
#include <stdexcept>

namespace lfo { namespace nttx {
typedef uint64_t u64; typedef unsigned __int128 u128;

struct Fp {
    u64 p, gen; int adicity; bool pow2_rule;
    u64 mul(u64 a, u64 b) const { return (u64)((u128)a * b % p); }
    u64 add(u64 a, u64 b) const { u128 s = (u128)a + b; return (u64)(s >= p ? s - p : s); }
    u64 sub(u64 a, u64 b) const { return a >= b ? a - b : a + (p - b); }
    u64 pow(u64 a, u64 e) const { u64 r = 1; while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; } return r; }
    u64 inv(u64 a) const { return pow(a, p - 2); }
};
inline Fp field(int id) {
    if (id == 0) return Fp{0xFFFFFFFF00000001ULL, 7, 32, true};
    if (id == 1) return Fp{2013265921ULL, 31, 27, false};
    throw std::runtime_error("unknown field");
}
inline u64 root(const Fp& F, int log_n) {
    u64 r0 = F.pow(F.gen, (F.p - 1) >> F.adicity), rho = r0;
    if (F.pow2_rule) { bool ok = false; for (u64 u = 1; u < 64 && !ok; u += 2) { rho = F.pow(r0, u); ok = F.pow(rho, (u64)1 << 27) == 64; } if (!ok) throw std::runtime_error("no root"); }
    return F.pow(rho, (u64)1 << (F.adicity - 1 - log_n));
}
// the definition, O(N^2)
inline void naive(const Fp& F, int log_n, const u64* in, u64* out, bool inverse) {
    const u64 n = (u64)1 << log_n; u64 psi = root(F, log_n); if (inverse) psi = F.inv(psi);
    std::vector<u64> pw(2 * n); pw[0] = 1; for (u64 i = 1; i < 2 * n; ++i) pw[i] = F.mul(pw[i - 1], psi);
    const u64 ninv = F.inv(n % F.p);
    for (u64 o = 0; o < n; ++o) { u64 acc = 0;
        for (u64 i = 0; i < n; ++i) { u64 e = inverse ? (o * (2 * i + 1)) % (2 * n) : (i * (2 * o + 1)) % (2 * n); acc = F.add(acc, F.mul(in[i], pw[e])); }
        out[o] = inverse ? F.mul(acc, ninv) : acc; }
}
// textbook O(N log N): twist by psi^j, bit-reverse, iterative radix-2 Cooley-Tukey (in place on out)
inline void fast(const Fp& F, int log_n, const u64* in, u64* out, bool inverse, const std::vector<u64>& pw /* psi^i, i < 2N */) {
    const u64 n = (u64)1 << log_n;
    auto w = [&](u64 e) { e %= 2 * n; return inverse ? pw[(2 * n - e) % (2 * n)] : pw[e]; };
    std::vector<u64> a(n);
    for (u64 j = 0; j < n; ++j) a[j] = inverse ? in[j] : F.mul(in[j], pw[j]);
    for (u64 i = 0; i < n; ++i) { u64 r = 0; for (int b = 0; b < log_n; ++b) if (i >> b & 1) r |= (u64)1 << (log_n - 1 - b); if (r > i) std::swap(a[i], a[r]); }
    for (u64 len = 1; len < n; len <<= 1)
        for (u64 s = 0; s < n; s += 2 * len)
            for (u64 i = 0; i < len; ++i) { u64 tw = w(2 * (n / (2 * len)) * i), u = a[s + i], v = F.mul(a[s + i + len], tw); a[s + i] = F.add(u, v); a[s + i + len] = F.sub(u, v); }
    if (inverse) { const u64 ninv = F.inv(n % F.p); for (u64 j = 0; j < n; ++j) a[j] = F.mul(F.mul(a[j], ninv), w(j)); }
    for (u64 j = 0; j < n; ++j) out[j] = a[j];
}
inline std::vector<u64> powers(const Fp& F, int log_n) { const u64 n = (u64)1 << log_n; std::vector<u64> pw(2 * n); u64 psi = root(F, log_n); pw[0] = 1; for (u64 i = 1; i < 2 * n; ++i) pw[i] = F.mul(pw[i - 1], psi); return pw; }
// schoolbook product mod X^N + 1
inline void schoolbook(const Fp& F, int log_n, const u64* a, const u64* b, u64* out) {
    const u64 n = (u64)1 << log_n; for (u64 k = 0; k < n; ++k) out[k] = 0;
    for (u64 i = 0; i < n; ++i) for (u64 j = 0; j < n; ++j) { u64 t = F.mul(a[i], b[j]); if (i + j < n) out[i + j] = F.add(out[i + j], t); else out[i + j - n] = F.sub(out[i + j - n], t); }
}
}}  // namespace lfo::nttx
