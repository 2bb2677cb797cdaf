// Host-side ring tables and small-vector ring arithmetic of the product library (transcript-adjacent work that the
// reference also keeps on the CPU: x_s, y_0 Horner, RotSum, cm_0/u_0/x_0).  Independent of oracle/.
#pragma once
#include "field.cuh"
#include <vector>
#include <array>
#include <stdexcept>
#include <cstring>
#include <string>

namespace lf {

struct LfException : std::runtime_error { int code; LfException(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

// CRT / ICRT as D x D matrices over Fq with NNZ = D/TAU * ... non-zeros per row (8 for every supported ring):
//   slot_s(X^j) = Y^{k_s j} = nu^{(k_s j) div TAU} * Y^{(k_s j) mod TAU},  k_s the s-th unit of Z/G ascending.
// The choice of nu and of the slot order is the convention table DESIGN.md calls "unpinned"; it lives only here.
template <class Rg> struct RingTables {
    typedef typename Rg::F F;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU, NNZ = Rg::S;
    u64 nu;
    int k[Rg::S];
    u64 crt[D][D], icrt[D][D];
    // sparse rows (device constant-memory image): every row has exactly NNZ non-zeros
    int crt_idx[D][NNZ]; u64 crt_val[D][NNZ];
    int icrt_idx[D][NNZ]; u64 icrt_val[D][NNZ];

    RingTables() {
        nu = F::NU;
        int c = 0;
        for (int x = 1; x < Rg::G; ++x) { int a = x, b = Rg::G; while (b) { int t = a % b; a = b; b = t; } if (a == 1) k[c++] = x; }
        if (c != S) throw std::logic_error("slot count");
        std::memset(crt, 0, sizeof crt);
        for (int s = 0; s < S; ++s) for (int j = 0; j < D; ++j) { int e = k[s] * j; crt[s * TAU + e % TAU][j] = F::pow(nu, (u64)((e / TAU) % Rg::G)); }
        // Gauss-Jordan inverse
        static u64 M[D][D], I[D][D];
        std::memcpy(M, crt, sizeof M); std::memset(I, 0, sizeof I); for (int i = 0; i < D; ++i) I[i][i] = 1;
        for (int col = 0; col < D; ++col) {
            int piv = -1; for (int r = col; r < D; ++r) if (M[r][col]) { piv = r; break; }
            if (piv < 0) throw std::logic_error("CRT matrix singular");
            if (piv != col) for (int j = 0; j < D; ++j) { std::swap(M[piv][j], M[col][j]); std::swap(I[piv][j], I[col][j]); }
            u64 iv = F::inv(M[col][col]);
            for (int j = 0; j < D; ++j) { M[col][j] = F::mul(M[col][j], iv); I[col][j] = F::mul(I[col][j], iv); }
            for (int r = 0; r < D; ++r) if (r != col && M[r][col]) { u64 f = M[r][col];
                for (int j = 0; j < D; ++j) { M[r][j] = F::sub(M[r][j], F::mul(f, M[col][j])); I[r][j] = F::sub(I[r][j], F::mul(f, I[col][j])); } }
        }
        std::memcpy(icrt, I, sizeof icrt);
        sparsify(crt, crt_idx, crt_val); sparsify(icrt, icrt_idx, icrt_val);
    }
    static void sparsify(const u64 (*m)[D], int (*idx)[NNZ], u64 (*val)[NNZ]) {
        for (int r = 0; r < D; ++r) { int c = 0;
            for (int j = 0; j < D; ++j) if (m[r][j]) { if (c == NNZ) throw std::logic_error("CRT row denser than expected"); idx[r][c] = j; val[r][c] = m[r][j]; ++c; }
            for (; c < NNZ; ++c) { idx[r][c] = 0; val[r][c] = 0; } }
    }
};

template <class Rg> struct HostRing {
    typedef typename Rg::F F; typedef SlotField<Rg> SF;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU;
    typedef std::array<u64, Rg::D> El;
    const RingTables<Rg>& T;
    explicit HostRing(const RingTables<Rg>& t) : T(t) {}

    static El zero() { El e; e.fill(0); return e; }
    static El from_u64(u64 x) { El e = zero(); for (int s = 0; s < S; ++s) e[s * TAU] = x % F::P; return e; }   // R::from(u128): same integer in every slot
    static El from_sf(const u64* sf) { El e; for (int s = 0; s < S; ++s) for (int l = 0; l < TAU; ++l) e[s * TAU + l] = sf[l]; return e; }
    static El load(const u64* p) { El e; std::memcpy(e.data(), p, 8 * D); return e; }
    static El add(const El& a, const El& b) { El c; for (int i = 0; i < D; ++i) c[i] = F::add(a[i], b[i]); return c; }
    static El sub(const El& a, const El& b) { El c; for (int i = 0; i < D; ++i) c[i] = F::sub(a[i], b[i]); return c; }
    static El mul(const El& a, const El& b) { El c; for (int s = 0; s < S; ++s) SF::mul(&c[s * TAU], &a[s * TAU], &b[s * TAU]); return c; }   // NTT form
    static El scale(const El& a, u64 k) { El c; for (int i = 0; i < D; ++i) c[i] = F::mul(a[i], k); return c; }
    static bool is_zero(const El& a) { for (u64 v : a) if (v) return false; return true; }
    El crt(const El& a) const { El o; for (int r = 0; r < D; ++r) { typename F::Acc x; x.clear(); for (int c = 0; c < T.NNZ; ++c) x.mac(T.crt_val[r][c], a[T.crt_idx[r][c]]); o[r] = F::reduce(x); } return o; }
    El icrt(const El& a) const { El o; for (int r = 0; r < D; ++r) { typename F::Acc x; x.clear(); for (int c = 0; c < T.NNZ; ++c) x.mac(T.icrt_val[r][c], a[T.icrt_idx[r][c]]); o[r] = F::reduce(x); } return o; }
    // multiply a coefficient-form element by X (Cyclotomic::into_rot_iter step, cyclotomic-rings/src/rotation.rs:60)
    static void mul_x(El& a) { u64 top = a[D - 1]; for (int i = D - 1; i > 0; --i) a[i] = a[i - 1]; a[0] = F::neg(top); if (Rg::TRINOMIAL) a[D / 2] = F::add(a[D / 2], top); }
};

// balanced base-b digits of a signed value, least significant first (stark-rings balanced_decomposition convention:
// truncated remainder; |rem| <= b/2 kept, otherwise wrapped with a carry).  Returns false if the value needs more
// than `len` digits.
LF_HD bool balanced_digits(int64_t v, int64_t b, int len, int64_t* out) {
    int64_t half = b / 2; int n = 0;
    for (;;) {
        int64_t rem = v % b, q = v / b, ar = rem < 0 ? -rem : rem;
        int64_t dg;
        if (ar <= half) { dg = rem; v = q; } else if (rem < 0) { dg = rem + b; v = q - 1; } else { dg = rem - b; v = q + 1; }
        if (n < len) out[n] = dg;
        ++n;
        if (v == 0) break;
    }
    if (n > len) return false;
    for (; n < len; ++n) out[n] = 0;
    return true;
}

}  // namespace lf
