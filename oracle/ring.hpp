// TEST INFRASTRUCTURE ONLY -- CPU oracle for the LatticeFold prover hot path.
// Nothing under oracle/ is linked into, imported by, or executed from the product library
// (latticefold_b200/csrc).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / CPU baseline.
//
// ring.hpp: base-field, slot-field (Fq^tau) and cyclotomic-ring arithmetic, CRT/ICRT, balanced
// decomposition.  The reference keeps all of this in the un-vendored dependency
//   stark-rings @ 886a89f1febb45822b0ff453e39b640f755382b6   (reference Cargo.lock:1096-1098)
// so this file restates the *published algorithm* (plain modular arithmetic on Z_p[X]/Phi_m) and is
// anchored on the reference's own call sites / KATs:
//   ring shapes                       crates/cyclotomic-rings/src/rings/{goldilocks,babybear,frog}.rs:9-20
//   X^24 = X^12 - 1, flatten order    crates/cyclotomic-rings/src/rotation.rs:174-776 (RotSum KAT, reproduced)
//   R::from(u128) broadcast           crates/latticefold/src/commitment/commitment_scheme.rs:142-160
// PARITY UNPINNED (no golden value in the reference tree fixes these; see DESIGN.md "Conventions"):
//   * slot-field non-residue nu and the order/basis of the CRT slots,
//   * the balanced-digit tie-breaking convention.
// They live in exactly one table (RingParams) on both the oracle and the product side.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>
#include <stdexcept>
#include <algorithm>

namespace lfo {

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef __int128 i128;

enum RingId { RING_GOLDILOCKS = 0, RING_BABYBEAR = 1, RING_FROG = 2 };

static const u64 P_GOLDILOCKS = 0xFFFFFFFF00000001ULL;  // 2^64 - 2^32 + 1   (rings/goldilocks.rs:86 pins p-12)
static const u64 P_BABYBEAR = 2013265921ULL;            // 15*2^27 + 1       (rings/babybear.rs KAT pins p-12)
static const u64 P_FROG = 15912092521325583641ULL;      //                   (rings/frog.rs KAT pins p-5)

// ---------------------------------------------------------------- base field
struct Fp {
    u64 p;
    inline u64 add(u64 a, u64 b) const { u128 s = (u128)a + b; return (u64)(s >= p ? s - p : s); }
    inline u64 sub(u64 a, u64 b) const { return a >= b ? a - b : (u64)((u128)a + p - b); }
    inline u64 neg(u64 a) const { return a ? p - a : 0; }
    inline u64 mul(u64 a, u64 b) const {
        if (p == P_GOLDILOCKS) {  // 2^64 = 2^32 - 1, 2^96 = -1 (mod p): the usual Goldilocks folding, exact
            u128 x = (u128)a * b; u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
            u64 t0 = lo - hh; if (lo < hh) t0 -= 0xFFFFFFFFULL;
            u64 t1 = hl * 0xFFFFFFFFULL, r = t0 + t1; if (r < t1) r += 0xFFFFFFFFULL;
            return r >= p ? r - p : r;
        }
        if (p < (1ULL << 32)) return (a * b) % p;
        return (u64)(((u128)a * b) % p);
    }
    u64 pow(u64 a, u64 e) const {
        u64 r = 1 % p;
        while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; }
        return r;
    }
    u64 inv(u64 a) const { return pow(a, p - 2); }
    // signed integer -> field
    u64 from_i128(i128 x) const { i128 m = x % (i128)p; if (m < 0) m += p; return (u64)m; }
    // field -> signed representative in (-p/2, p/2]   (balanced_decomposition, stark-rings; UNPINNED)
    i128 to_signed(u64 a) const { return a <= (p - 1) / 2 ? (i128)a : (i128)a - (i128)p; }
};

// ---------------------------------------------------------------- ring descriptor
// Family covered: Phi_m with m = tau*g, p = 1 mod g, ord_m(p) = tau; slot field Fq[Y]/(Y^tau - nu), nu a
// primitive g-th root of unity in Fq, so Y is a primitive m-th root of unity in the slot field and
//   Phi_m(X) = prod_{k in (Z/g)^*} (X^tau - nu^k),   slot_k(a) = a(Y^k).
// Goldilocks: m=72, g=24, tau=3, d=24 (X^24 - X^12 + 1);  BabyBear: m=216, g=24, tau=9, d=72 (X^72 - X^36 + 1);
// Frog: m=32, g=8, tau=4, d=16 (X^16 + 1).
struct RingParams {
    int id;
    Fp F;
    int d, S, tau, g;
    u64 nu;
    std::vector<int> k;       // slot exponents, ascending units of Z/g
    bool trinomial;           // true: X^d = X^{d/2} - 1 ; false: X^d = -1
    std::vector<u64> icrt;    // d x d matrix: coeff = icrt * ntt_limbs
    // short-challenge set (rings/*.rs)
    int cs_bytes;
    int E() const { return d; }  // limbs per element in either form
};

inline u64 find_nu(const Fp& F, int g, int ringid) {
    if (ringid == RING_GOLDILOCKS) return 1ULL << 40;  // 2 has order 192 => 2^40 has order 24; shifts-only twiddles
    // smallest generator-derived primitive g-th root: h^((p-1)/g) for the smallest h that gives exact order g
    for (u64 h = 2;; ++h) {
        u64 w = F.pow(h, (F.p - 1) / g);
        bool ok = true;
        for (int q : {2, 3}) if (g % q == 0 && F.pow(w, g / q) == 1) ok = false;
        if (ok) return w;
    }
}

// ---------------------------------------------------------------- slot field ops (tau limbs, Y^tau = nu)
inline void sf_mul(const RingParams& R, u64* out, const u64* a, const u64* b) {
    const int t = R.tau; const Fp& F = R.F;
    u64 lo[16] = {0}, hi[16] = {0};
    for (int i = 0; i < t; ++i) {
        if (!a[i]) continue;
        for (int j = 0; j < t; ++j) {
            u64 pr = F.mul(a[i], b[j]);
            if (i + j < t) lo[i + j] = F.add(lo[i + j], pr); else hi[i + j - t] = F.add(hi[i + j - t], pr);
        }
    }
    for (int i = 0; i < t; ++i) out[i] = F.add(lo[i], F.mul(hi[i], R.nu));
}
inline void sf_add(const RingParams& R, u64* out, const u64* a, const u64* b) { for (int i = 0; i < R.tau; ++i) out[i] = R.F.add(a[i], b[i]); }
inline void sf_sub(const RingParams& R, u64* out, const u64* a, const u64* b) { for (int i = 0; i < R.tau; ++i) out[i] = R.F.sub(a[i], b[i]); }
inline bool sf_is_zero(const RingParams& R, const u64* a) { for (int i = 0; i < R.tau; ++i) if (a[i]) return false; return true; }
inline void sf_pow(const RingParams& R, u64* out, const u64* a, u128 e) {
    u64 r[16] = {0}, b[16], t[16]; r[0] = 1; memcpy(b, a, 8 * R.tau);
    while (e) { if (e & 1) { sf_mul(R, t, r, b); memcpy(r, t, 8 * R.tau); } sf_mul(R, t, b, b); memcpy(b, t, 8 * R.tau); e >>= 1; }
    memcpy(out, r, 8 * R.tau);
}
// inverse through the norm to the base field: a^{-1} = a^{(q^tau - 1)/(q-1) - 1} * N(a)^{-1}
inline void sf_inv(const RingParams& R, u64* out, const u64* a) {
    // e = 1 + q + ... + q^{tau-1}; compute a^{e-1} by Frobenius-free square-and-multiply on big exponent
    // tau*64 bits do not fit u128 for BabyBear (9*31=279) -> do it as product of a^{q^i}, i=1..tau-1
    const int t = R.tau;
    u64 fr[16], acc[16] = {0}, tmp[16]; acc[0] = 1; memcpy(fr, a, 8 * t);
    for (int i = 1; i < t; ++i) { sf_pow(R, tmp, fr, R.F.p); memcpy(fr, tmp, 8 * t); sf_mul(R, tmp, acc, fr); memcpy(acc, tmp, 8 * t); }
    sf_mul(R, tmp, acc, a);  // norm, lies in the base field
    for (int i = 1; i < t; ++i) if (tmp[i]) throw std::runtime_error("sf_inv: norm not in base field");
    u64 ninv = R.F.inv(tmp[0]);
    for (int i = 0; i < t; ++i) out[i] = R.F.mul(acc[i], ninv);
}

// ---------------------------------------------------------------- ring elements: E = d limbs
// NTT form: limb index = slot*tau + l  (absorbed slot-major by the transcript, transcript/poseidon.rs:40-47)
// coefficient form: limb index = power of X
inline void ntt_mul(const RingParams& R, u64* out, const u64* a, const u64* b) {
    u64 t[16];
    for (int s = 0; s < R.S; ++s) { sf_mul(R, t, a + s * R.tau, b + s * R.tau); memcpy(out + s * R.tau, t, 8 * R.tau); }
}
inline void el_add(const RingParams& R, u64* out, const u64* a, const u64* b) { for (int i = 0; i < R.d; ++i) out[i] = R.F.add(a[i], b[i]); }
inline void el_sub(const RingParams& R, u64* out, const u64* a, const u64* b) { for (int i = 0; i < R.d; ++i) out[i] = R.F.sub(a[i], b[i]); }
inline bool el_is_zero(const RingParams& R, const u64* a) { for (int i = 0; i < R.d; ++i) if (a[i]) return false; return true; }
inline void ntt_from_u64(const RingParams& R, u64* out, u64 x) {  // R::from(u128): same integer in every slot
    memset(out, 0, 8 * R.d); u64 v = x % R.F.p; for (int s = 0; s < R.S; ++s) out[s * R.tau] = v;
}
inline void ntt_from_sf(const RingParams& R, u64* out, const u64* sf) {  // R::from(slot field elem): broadcast
    for (int s = 0; s < R.S; ++s) memcpy(out + s * R.tau, sf, 8 * R.tau);
}

// multiply a coefficient-form polynomial by X (one rotation step; cf. Cyclotomic::into_rot_iter, rotation.rs:60)
inline void coeff_mul_x(const RingParams& R, u64* a) {
    const int d = R.d; u64 top = a[d - 1];
    for (int i = d - 1; i > 0; --i) a[i] = a[i - 1];
    a[0] = R.F.neg(top);                                   // X^d = ... - 1
    if (R.trinomial) a[d / 2] = R.F.add(a[d / 2], top);    // X^d = X^{d/2} - 1
}
// schoolbook product mod Phi -- the CRT-independent definition of ring multiplication
inline void coeff_mul(const RingParams& R, u64* out, const u64* a, const u64* b) {
    const int d = R.d; const Fp& F = R.F;
    std::vector<u64> w(2 * d, 0);
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) w[i + j] = F.add(w[i + j], F.mul(a[i], b[j]));
    for (int i = 2 * d - 2; i >= d; --i) {
        u64 c = w[i]; if (!c) continue; w[i] = 0;
        w[i - d] = F.sub(w[i - d], c);
        if (R.trinomial) w[i - d / 2] = F.add(w[i - d / 2], c);
    }
    memcpy(out, w.data(), 8 * d);
}

// CRT: slot_s(a) = a(Y^{k_s}) by Horner in the slot field.  Y^k = nu^{k div tau} * Y^{k mod tau}.
inline void crt(const RingParams& R, u64* out, const u64* a) {
    const int t = R.tau;
    for (int s = 0; s < R.S; ++s) {
        u64 yk[16] = {0}; yk[R.k[s] % t] = R.F.pow(R.nu, R.k[s] / t);
        u64 acc[16] = {0}, tmp[16];
        for (int j = R.d - 1; j >= 0; --j) { sf_mul(R, tmp, acc, yk); memcpy(acc, tmp, 8 * t); acc[0] = R.F.add(acc[0], a[j]); }
        memcpy(out + s * t, acc, 8 * t);
    }
}
inline void icrt(const RingParams& R, u64* out, const u64* a) {
    const int d = R.d; u64 tmp[128];
    for (int i = 0; i < d; ++i) { u64 acc = 0; for (int j = 0; j < d; ++j) if (R.icrt[i * d + j] && a[j]) acc = R.F.add(acc, R.F.mul(R.icrt[i * d + j], a[j])); tmp[i] = acc; }
    memcpy(out, tmp, 8 * d);
}

inline std::vector<u64> mat_inverse(const Fp& F, std::vector<u64> M, int n) {
    std::vector<u64> I(n * n, 0); for (int i = 0; i < n; ++i) I[i * n + i] = 1;
    for (int c = 0; c < n; ++c) {
        int piv = -1; for (int r = c; r < n; ++r) if (M[r * n + c]) { piv = r; break; }
        if (piv < 0) throw std::runtime_error("CRT matrix singular");
        if (piv != c) for (int j = 0; j < n; ++j) { std::swap(M[piv * n + j], M[c * n + j]); std::swap(I[piv * n + j], I[c * n + j]); }
        u64 iv = F.inv(M[c * n + c]);
        for (int j = 0; j < n; ++j) { M[c * n + j] = F.mul(M[c * n + j], iv); I[c * n + j] = F.mul(I[c * n + j], iv); }
        for (int r = 0; r < n; ++r) if (r != c && M[r * n + c]) {
            u64 f = M[r * n + c];
            for (int j = 0; j < n; ++j) { M[r * n + j] = F.sub(M[r * n + j], F.mul(f, M[c * n + j])); I[r * n + j] = F.sub(I[r * n + j], F.mul(f, I[c * n + j])); }
        }
    }
    return I;
}

inline RingParams make_ring(int id) {
    RingParams R; R.id = id;
    switch (id) {
        case RING_GOLDILOCKS: R.F.p = P_GOLDILOCKS; R.tau = 3; R.g = 24; R.trinomial = true; R.cs_bytes = 18; break;
        case RING_BABYBEAR:   R.F.p = P_BABYBEAR;   R.tau = 9; R.g = 24; R.trinomial = true; R.cs_bytes = 18; break;
        case RING_FROG:       R.F.p = P_FROG;       R.tau = 4; R.g = 8;  R.trinomial = false; R.cs_bytes = 16; break;
        default: throw std::runtime_error("unknown ring id");
    }
    for (int k = 1; k < R.g; ++k) if (std::__gcd(k, R.g) == 1) R.k.push_back(k);
    R.S = (int)R.k.size(); R.d = R.S * R.tau;
    R.nu = find_nu(R.F, R.g, id);
    // ICRT matrix = inverse of the CRT matrix (columns = CRT of the monomials)
    std::vector<u64> M(R.d * R.d, 0), e(R.d), o(R.d);
    for (int j = 0; j < R.d; ++j) { std::fill(e.begin(), e.end(), 0); e[j] = 1; crt(R, o.data(), e.data()); for (int i = 0; i < R.d; ++i) M[i * R.d + j] = o[i]; }
    R.icrt = mat_inverse(R.F, M, R.d);
    return R;
}

// ---------------------------------------------------------------- balanced decomposition (UNPINNED convention)
// stark-rings balanced_decomposition: signed representative in (-p/2, p/2]; digits least-significant first;
// rem = curr % b (truncated); |rem| <= b/2 is kept, otherwise rem -/+ b with a carry of +/-1; padded with zeros.
inline void decompose_balanced(const Fp& F, u64 v, u128 b, int len, i128* digits) {
    i128 curr = F.to_signed(v); i128 bb = (i128)b, half = bb / 2; int n = 0;
    for (;;) {
        i128 rem = curr % bb; i128 q = curr / bb;
        i128 arem = rem < 0 ? -rem : rem;
        if (arem <= half) { if (n < len) digits[n] = rem; curr = q; }
        else { i128 dg = rem < 0 ? rem + bb : rem - bb; if (n < len) digits[n] = dg; curr = q + (rem < 0 ? -1 : 1); }
        ++n;
        if (curr == 0) break;
    }
    if (n > len) throw std::runtime_error("decompose_balanced: value does not fit in the requested number of digits");
    for (; n < len; ++n) digits[n] = 0;
}
// one coefficient-form element -> len digit elements (digit l of every coefficient)
inline void decompose_elem(const RingParams& R, const u64* a, u128 b, int len, u64* out /* len x d */) {
    std::vector<i128> dg(len);
    for (int c = 0; c < R.d; ++c) {
        decompose_balanced(R.F, a[c], b, len, dg.data());
        for (int l = 0; l < len; ++l) out[(size_t)l * R.d + c] = R.F.from_i128(dg[l]);
    }
}
// recompose(chunk, base) = sum_i chunk[i] * base^i   (works limb-wise in either form: base is an integer scalar)
inline void recompose_elems(const RingParams& R, const u64* chunk, int len, u128 base, u64* out) {
    u64 bmod = (u64)(base % R.F.p), pw = 1; std::vector<u64> acc(R.d, 0);
    for (int i = 0; i < len; ++i) { for (int c = 0; c < R.d; ++c) acc[c] = R.F.add(acc[c], R.F.mul(chunk[(size_t)i * R.d + c], pw)); pw = R.F.mul(pw, bmod); }
    memcpy(out, acc.data(), 8 * R.d);
}

// ---------------------------------------------------------------- short challenge sets (rings/*.rs)
// goldilocks.rs:32-68 / babybear.rs:32-68: 18 bytes -> 24 six-bit values - 32 (BabyBear: remaining 48 coeffs zero)
// frog.rs:32-56: 16 bytes -> byte - 128
inline void short_challenge_from_bytes(const RingParams& R, const uint8_t* bs, u64* coeffs) {
    memset(coeffs, 0, 8 * R.d);
    if (R.id == RING_FROG) { for (int i = 0; i < 16; ++i) coeffs[i] = R.F.from_i128((i128)bs[i] - 128); return; }
    for (int i = 0; i < 6; ++i) {
        int x0 = (bs[3 * i] & 0x3F) - 32;
        int x1 = (((bs[3 * i] & 0xC0) >> 6) | ((bs[3 * i + 1] & 0x0F) << 2)) - 32;
        int x2 = (((bs[3 * i + 1] & 0xF0) >> 4) | ((bs[3 * i + 2] & 0x03) << 4)) - 32;
        int x3 = ((bs[3 * i + 2] & 0xFC) >> 2) - 32;
        int xs[4] = {x0, x1, x2, x3};
        for (int j = 0; j < 4; ++j) coeffs[4 * i + j] = R.F.from_i128(xs[j]);
    }
}

}  // namespace lfo
