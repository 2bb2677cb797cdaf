// Fiat-Shamir transcript of the product's host side: Poseidon duplex sponge (width 24 = rate 20 + capacity 4,
// 8 full + 22 partial rounds, x^7) driven exactly as the reference drives arkworks' PoseidonSponge:
//   crates/latticefold/src/transcript/poseidon.rs:29-75  (absorb ring element = its D base limbs; get_challenge squeezes
//   TAU limbs and absorbs them back; short challenges from squeeze_bytes)
//   crates/latticefold/src/transcript.rs:13-51           (absorb_field_element = broadcast to a ring element)
// Sponge algorithm: ark-crypto-primitives 0.4.0 PoseidonSponge (duplex with Absorbing/Squeezing cursor).
// Runs on the CPU by design (sequential); the MDS product uses lazily reduced 192-bit sums (x86-64 add/adc chains).
#pragma once
#include "ring_host.hpp"
#include <type_traits>
#include <cstdlib>
#include "poseidon_w24_tables.inc"

namespace lf {

// AVX-512 IFMA dense layer for the Goldilocks field (poseidon_ifma.cpp, compiled by g++; chosen at run time)
struct PoseidonIfmaMatrix { alignas(64) u64 limb[POSEIDON_W24_WIDTH][2][3][8]; };
bool poseidon_ifma_supported();
void poseidon_ifma_prepare(const u64* m, PoseidonIfmaMatrix* out);
void poseidon_ifma_dense(const PoseidonIfmaMatrix* M, u64* st);
// LF_POSEIDON_SCALAR=1 in the environment keeps the scalar dense layer (tests compare the two)
inline bool poseidon_use_ifma() { static const bool on = poseidon_ifma_supported() && !(std::getenv("LF_POSEIDON_SCALAR") && std::getenv("LF_POSEIDON_SCALAR")[0] == '1'); return on; }

constexpr u64 poseidon_inv64(u64 p) { u64 x = 1; for (int i = 0; i < 6; ++i) x *= 2 - p * x; return x; }      // p^-1 mod 2^64 (Newton, p odd)

template <class Rg> class Transcript {
    typedef typename Rg::F F;
    static constexpr int W = POSEIDON_W24_WIDTH, RATE = POSEIDON_W24_RATE, CAP = POSEIDON_W24_CAP;
    static constexpr int RF = POSEIDON_W24_FULL, RP = POSEIDON_W24_PARTIAL;
    // Partial rounds use the standard sparse factorisation of the MDS matrix (Poseidon paper, appendix B): with
    // M = Ms * Md, Ms = [[m00, v^T Mh^-1], [w, I]], Md = diag(1, Mh), the block-diagonal factor commutes with the
    // single S-box and is pushed into the previous round, so a partial round costs 2W-1 multiplications instead of W^2.
    // The permutation computed is identical (checked against the dense form in tests and by the reference KATs).
    struct Tables {
        u64 ark[(RF + RP) * W];       // round constants; partial rounds hold Md_r * c_r
        u64 mds[W * W];               // dense MDS (full rounds)
        u64 pre[W * W];               // replaces MDS in the last full round before the partial rounds
        u64 sp_row0[RP][W];           // Ms_r first row
        u64 sp_col0[RP][W];           // Ms_r first column (entry 0 unused)
        u64 sp_c0[RP];                // partial rounds: the only constant that has to be added before the S-box (lane 0)
        PoseidonIfmaMatrix mds_ifma, pre_ifma;   // limb images of mds / pre (Goldilocks on IFMA hosts only)
        // IFMA hosts run the RP partial rounds in closed form (partial_rounds_closed): with s the state on entry and x_r the S-box
        // output of round r, lane i >= 1 on exit is s_i + sum_r col0_r[i] x_r and lane 0 after round r is
        // row0_r[0] x_r + sum_{j>=1} row0_r[j] s_j + sum_{r'<r} pr_t[r][r'] x_r',  pr_t[r][r'] = sum_{j>=1} row0_r[j] col0_r'[j].
        PoseidonIfmaMatrix pr_v_ifma, pr_c_ifma; // rows r < RP: row0_r without lane 0;  rows i >= 1: (col0_r[i])_r
        u64 pr_t[RP][RP];
        bool ifma = false;
        static void matmul(u64* o, const u64* a, const u64* b) {   // o = a * b (W x W)
            for (int i = 0; i < W; ++i) for (int j = 0; j < W; ++j) { u64 acc = 0; for (int k = 0; k < W; ++k) acc = F::add(acc, F::mul(a[i * W + k], b[k * W + j])); o[i * W + j] = acc; }
        }
        static void invert(u64* inv, const u64* m, int n) {         // Gauss-Jordan over Fq, n x n row-major
            std::vector<u64> a(m, m + n * n); for (int i = 0; i < n * n; ++i) inv[i] = 0; for (int i = 0; i < n; ++i) inv[i * n + i] = 1;
            for (int c = 0; c < n; ++c) {
                int piv = -1; for (int r = c; r < n; ++r) if (a[r * n + c]) { piv = r; break; }
                if (piv < 0) throw std::logic_error("Poseidon: singular MDS sub-matrix");
                if (piv != c) for (int k = 0; k < n; ++k) { std::swap(a[piv * n + k], a[c * n + k]); std::swap(inv[piv * n + k], inv[c * n + k]); }
                u64 iv = F::inv(a[c * n + c]);
                for (int k = 0; k < n; ++k) { a[c * n + k] = F::mul(a[c * n + k], iv); inv[c * n + k] = F::mul(inv[c * n + k], iv); }
                for (int r = 0; r < n; ++r) if (r != c && a[r * n + c]) { u64 f = a[r * n + c];
                    for (int k = 0; k < n; ++k) { a[r * n + k] = F::sub(a[r * n + k], F::mul(f, a[c * n + k])); inv[r * n + k] = F::sub(inv[r * n + k], F::mul(f, inv[c * n + k])); } }
            }
        }
        Tables() {
            for (int i = 0; i < (RF + RP) * W; ++i) ark[i] = POSEIDON_W24_ARK[i] % F::P;
            for (int i = 0; i < W * W; ++i) mds[i] = POSEIDON_W24_MDS[i] % F::P;
            std::vector<u64> cur(mds, mds + W * W), md(W * W), nxt(W * W), hat((W - 1) * (W - 1)), hinv((W - 1) * (W - 1));
            for (int r = RP - 1; r >= 0; --r) {
                for (int i = 1; i < W; ++i) for (int j = 1; j < W; ++j) hat[(i - 1) * (W - 1) + (j - 1)] = cur[i * W + j];
                invert(hinv.data(), hat.data(), W - 1);
                sp_row0[r][0] = cur[0];
                for (int j = 1; j < W; ++j) { u64 acc = 0; for (int k = 1; k < W; ++k) acc = F::add(acc, F::mul(cur[k], hinv[(k - 1) * (W - 1) + (j - 1)])); sp_row0[r][j] = acc; }   // v^T Mh^-1
                sp_col0[r][0] = 0; for (int i = 1; i < W; ++i) sp_col0[r][i] = cur[i * W];
                std::fill(md.begin(), md.end(), 0); md[0] = 1;
                for (int i = 1; i < W; ++i) for (int j = 1; j < W; ++j) md[i * W + j] = cur[i * W + j];
                // constants of this partial round move through Md
                u64* c = ark + (RF / 2 + r) * W; u64 nc[W];
                for (int i = 0; i < W; ++i) { u64 acc = 0; for (int k = 0; k < W; ++k) acc = F::add(acc, F::mul(md[i * W + k], c[k])); nc[i] = acc; }
                std::memcpy(c, nc, sizeof nc);
                matmul(nxt.data(), md.data(), mds); cur = nxt;   // matrix the previous round has to apply
            }
            std::memcpy(pre, cur.data(), sizeof pre);
            // Lanes 1..W-1 skip the S-box, so their constants commute with it: push them through Ms_r into the next round
            // (new_0 = row0 . k, new_i = k_i) until they land in the constants of the first full round after the partial ones.
            u64 carry[W] = {0};
            for (int r = 0; r < RP; ++r) {
                u64* c = ark + (RF / 2 + r) * W; u64 k[W];
                for (int i = 0; i < W; ++i) k[i] = F::add(c[i], carry[i]);
                sp_c0[r] = k[0];
                u64 acc = 0; for (int j = 1; j < W; ++j) acc = F::add(acc, F::mul(sp_row0[r][j], k[j]));
                carry[0] = acc; for (int i = 1; i < W; ++i) carry[i] = k[i];
            }
            u64* c = ark + (RF / 2 + RP) * W;
            for (int i = 0; i < W; ++i) c[i] = F::add(c[i], carry[i]);
            if (std::is_same<F, Goldilocks>::value && poseidon_use_ifma()) {
                static_assert(RP <= W, "closed-form partial rounds keep one S-box output per state lane");
                poseidon_ifma_prepare(mds, &mds_ifma); poseidon_ifma_prepare(pre, &pre_ifma);
                std::vector<u64> v(W * W, 0), cm(W * W, 0);
                for (int r = 0; r < RP; ++r) for (int j = 1; j < W; ++j) { v[r * W + j] = sp_row0[r][j]; cm[j * W + r] = sp_col0[r][j]; }
                poseidon_ifma_prepare(v.data(), &pr_v_ifma); poseidon_ifma_prepare(cm.data(), &pr_c_ifma);
                for (int r = 0; r < RP; ++r) for (int q = 0; q < RP; ++q) {
                    u64 acc = 0; if (q < r) for (int j = 1; j < W; ++j) acc = F::add(acc, F::mul(sp_row0[r][j], sp_col0[q][j]));
                    pr_t[r][q] = acc;
                }
                ifma = true;
            }
            if constexpr (MONT) {      // constants in Montgomery form, matrices with R^2 (see MONT above)
                for (u64& x : ark) x = to_mont(x);
                for (u64& x : sp_c0) x = to_mont(x);
                for (u64& x : mds) x = to_mont(to_mont(x));
                for (u64& x : pre) x = to_mont(to_mont(x));
                for (auto& row : sp_row0) for (u64& x : row) x = to_mont(to_mont(x));
                for (auto& col : sp_col0) for (u64& x : col) x = to_mont(x);
            }
        }
    };
    static const Tables& tables() { static const Tables t; return t; }
    u64 st_[W];
    int cursor_;        // next rate lane to absorb into / squeeze from
    bool squeezing_;
    unsigned long long permutations_ = 0;

    // Lazy ("weak") representatives inside a round: any u64 congruent to the value.  For Goldilocks that drops the canonicalising
    // subtract from every reduction; the dense layers' reduce_wide returns canonical lanes again.  Other fields keep canonical ops.
    static constexpr bool LAZY = std::is_same<F, Goldilocks>::value;
    // 64-bit generic primes (the Frog ring's q): the state, the round constants and the matrices live in Montgomery form inside the sponge
    // (REDC = two multiplies instead of a 128-by-64-bit division per reduction: 20.8 -> ~5 us per permutation); absorb / squeeze convert.
    // Matrix entries carry R^2 so that a lazily accumulated row (sum m s R^3, 192 bits) comes back to Montgomery form with two REDC steps.
    static constexpr bool MONT = !LAZY && F::P > (1ull << 32);
    // 31-bit field (BabyBear): a dot product of 24 lanes runs on 32 x 16-bit partial products in plain 64-bit sums (no carry chains, vectorisable),
    // the matrix entries split into their low and high 16 bits: sum a m = sum a m_lo + 2^16 sum a m_hi, both sums < 2^52
    static constexpr bool SMALLP = F::P < (1ull << 32);
    static inline u64 dot_small(const u64* row, const u64* v, int n) {
        u64 s0 = 0, s1 = 0;
        for (int j = 0; j < n; ++j) { const u64 a = (u32)v[j], m = row[j]; s0 += a * (m & 0xFFFFu); s1 += a * (m >> 16); }
        return (s0 % F::P + ((s1 % F::P) << 16)) % F::P;
    }
    static constexpr u64 NINV = ~poseidon_inv64(F::P) + 1;
    static inline u64 redc(u64 lo, u64 hi) {                // (hi:lo) / 2^64 mod p for hi < p; branch-free (the conditions are coin flips)
        const u64 m = lo * NINV; const u64 mph = (u64)(((u128)m * F::P) >> 64);
        const u64 t = hi + mph; const u64 t2 = t + (u64)(lo != 0);
        const u64 over = (u64)(t < hi) | (u64)(t2 < t) | (u64)(t2 >= F::P);
        return t2 - (F::P & (0 - over));
    }
    static inline u64 madd(u64 a, u64 b) { const u64 s = a + b; return s - (F::P & (0 - ((u64)(s < a) | (u64)(s >= F::P)))); }      // a + b mod p, branch-free
    static u64 mont_r1() { return (u64)((((u128)1) << 64) % F::P); }
    static u64 to_mont(u64 a) { return (u64)((u128)(a % F::P) * mont_r1() % F::P); }
    static inline u64 wred(u64 lo, u64 hi) {               // (hi:lo) mod p, weak
        if constexpr (LAZY) {
            u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFull, t1 = (hl << 32) - hl;
            u64 t0 = lo - hh; t0 -= (0 - (u64)(lo < hh)) & 0xFFFFFFFFull;
            u64 r = t0 + t1; r += (0 - (u64)(r < t1)) & 0xFFFFFFFFull;
            return r;
        } else if constexpr (MONT) return redc(lo, hi);
        else if constexpr (F::P < (1ull << 32)) return lo % F::P;      // 31-bit field: every product and multiply-add of the sponge fits one word (hi == 0); a constant-divisor remainder, no 128-bit division
        else return F::reduce128(lo, hi);
    }
    static inline u64 wmul(u64 a, u64 b) { u128 x = (u128)a * b; return wred((u64)x, (u64)(x >> 64)); }
    static inline u64 wadd(u64 a, u64 c) {                 // a weak, c canonical
        if constexpr (LAZY) { u64 s = a + c; s += (0 - (u64)(s < a)) & 0xFFFFFFFFull; return s; } else if constexpr (MONT) return madd(a, c); else return F::add(a, c);
    }
    // 192-bit sum of products in three registers, one add/adc/adc chain per product
    struct Acc3 {
        u64 c0 = 0, c1 = 0, c2 = 0;
        inline void mac(u64 a, u64 b) {
            u128 x = (u128)a * b; u64 lo = (u64)x, hi = (u64)(x >> 64);
            asm("add %3, %0\n\tadc %4, %1\n\tadc $0, %2" : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(lo), "r"(hi) : "cc");
        }
        inline u64 reduce() const {
            if constexpr (MONT) {      // two REDC steps on c2:c1:c0 (c2 < 2^6)
                const u64 m = c0 * NINV; const u64 mph = (u64)(((u128)m * F::P) >> 64);
                const u64 lo = c1 + mph; u64 k = (u64)(lo < c1); const u64 lo2 = lo + (u64)(c0 != 0); k += (u64)(lo2 < lo); const u64 hi = c2 + k;
                const u64 m2 = lo2 * NINV; const u64 mph2 = (u64)(((u128)m2 * F::P) >> 64);
                const u64 r = hi + mph2 + (u64)(lo2 != 0);      // < p + 2^7: no wrap (p < 2^64 - 2^61)
                return r - (F::P & (0 - (u64)(r >= F::P)));
            } else return F::reduce_wide((u128)c0, ((u128)c2 << 64) | c1);
        }
    };
    // x -> (x + c)^7 on all lanes, one multiplication level at a time: 24 independent products per level keep the multiplier
    // busy, where lane-by-lane x^7 chains four dependent reductions
    void sbox_layer(const u64* c) {
        u64 a[W], x2[W], x3[W];
        for (int i = 0; i < W; ++i) a[i] = wadd(st_[i], c[i]);
        for (int i = 0; i < W; ++i) x2[i] = wmul(a[i], a[i]);
        for (int i = 0; i < W; ++i) x3[i] = wmul(x2[i], a[i]);
        for (int i = 0; i < W; ++i) x2[i] = wmul(x3[i], x3[i]);
        for (int i = 0; i < W; ++i) st_[i] = wmul(x2[i], a[i]);
    }
    void dense_layer(const u64* m, const PoseidonIfmaMatrix* mi, bool ifma) {
        if (ifma) { poseidon_ifma_dense(mi, st_); return; }
        u64 nx[W];
        for (int i = 0; i < W; ++i) {
            const u64* row = m + i * W;
            if constexpr (SMALLP) nx[i] = dot_small(row, st_, W);
            else {
                Acc3 acc;
#pragma GCC unroll 24
                for (int j = 0; j < W; ++j) acc.mac(row[j], st_[j]);
                nx[i] = acc.reduce();
            }
        }
        std::memcpy(st_, nx, sizeof st_);
    }
    // all RP partial rounds with two IFMA matrix products and a short scalar chain (see Tables::pr_t): the 23 x RP lane updates lose
    // their per-round reductions, and only x -> x^7 plus two multiply-adds per round remain serial
    void partial_rounds_closed(const Tables& t) {
        u64 u[W], x[W] = {0};
        std::memcpy(u, st_, sizeof u); poseidon_ifma_dense(&t.pr_v_ifma, u);        // u_r = sum_{j>=1} row0_r[j] s_j, canonical
        u64 s0 = st_[0];
        for (int r = 0; r < RP; ++r) {
            const u64 a = wadd(s0, t.sp_c0[r]);
            const u64 a2 = wmul(a, a), a3 = wmul(a2, a), a4 = wmul(a2, a2); x[r] = wmul(a4, a3);   // depth 3: a^4 and a^3 in parallel
            Acc3 acc; acc.mac(u[r], 1);
            for (int q = 0; q < r; ++q) acc.mac(t.pr_t[r][q], x[q]);
            acc.mac(t.sp_row0[r][0], x[r]);
            s0 = acc.reduce();
        }
        poseidon_ifma_dense(&t.pr_c_ifma, x);                                       // x_i <- sum_r col0_r[i] x_r, canonical
        for (int i = 1; i < W; ++i) st_[i] = wadd(st_[i], x[i]);
        st_[0] = s0;
    }
    void permute() {
        const Tables& t = tables(); ++permutations_;
        int r = 0;
        for (; r < RF / 2; ++r) {
            sbox_layer(&t.ark[r * W]);
            if (r == RF / 2 - 1) dense_layer(t.pre, &t.pre_ifma, t.ifma); else dense_layer(t.mds, &t.mds_ifma, t.ifma);
        }
        if (t.ifma) { partial_rounds_closed(t); r += RP; }
        else for (int pr = 0; pr < RP; ++pr, ++r) {
            const u64 a = wadd(st_[0], t.sp_c0[pr]);
            const u64 a2 = wmul(a, a), a3 = wmul(a2, a), a4 = wmul(a2, a2), x0 = wmul(a4, a3);     // depth 3: a^4 and a^3 in parallel
            const u64* row = t.sp_row0[pr]; const u64* col = t.sp_col0[pr];
            Acc3 acc; u64 small_dot = 0;                   // lane 0 last: the other 23 products do not wait for the S-box
            if constexpr (SMALLP) small_dot = dot_small(row + 1, st_ + 1, W - 1);
            else {
#pragma GCC unroll 23
                for (int j = 1; j < W; ++j) acc.mac(row[j], st_[j]);
                acc.mac(row[0], x0);
            }
#pragma GCC unroll 23
            for (int i = 1; i < W; ++i) {
                if constexpr (MONT) st_[i] = madd(st_[i], wmul(col[i], x0));
                else { u128 x = (u128)col[i] * x0 + st_[i]; st_[i] = wred((u64)x, (u64)(x >> 64)); }
            }
            if constexpr (SMALLP) st_[0] = (small_dot + (row[0] * x0) % F::P) % F::P; else st_[0] = acc.reduce();
        }
        for (; r < RF + RP; ++r) {
            sbox_layer(&t.ark[r * W]);
            dense_layer(t.mds, &t.mds_ifma, t.ifma);
        }
    }
    static inline u64 to_mont_fast(u64 v) { static const u64 r2 = to_mont(mont_r1()); const u128 x = (u128)(v >= F::P ? v - F::P : v) * r2; return redc((u64)x, (u64)(x >> 64)); }
    static inline void copy_out(u64* dst, const u64* src, size_t n) { if constexpr (MONT) { for (size_t i = 0; i < n; ++i) dst[i] = redc(src[i], 0); } else std::memcpy(dst, src, 8 * n); }
public:
    Transcript() { std::memset(st_, 0, sizeof st_); cursor_ = 0; squeezing_ = false; }
    unsigned long long permutations() const { return permutations_; }

    void absorb_base(const u64* v, size_t n) {
        if (!n) return;
        if (squeezing_ || cursor_ == RATE) { permute(); cursor_ = 0; }
        squeezing_ = false;
        for (size_t i = 0; i < n; ++i) {
            if (cursor_ == RATE) { permute(); cursor_ = 0; }
            st_[CAP + cursor_] = MONT ? madd(st_[CAP + cursor_], to_mont_fast(v[i])) : F::add(st_[CAP + cursor_], v[i]); ++cursor_;
        }
    }
    void squeeze_base(u64* out, size_t n) {
        if (!n) return;
        if (!squeezing_ || cursor_ == RATE) { permute(); cursor_ = 0; }
        squeezing_ = true;
        size_t off = 0;
        for (;;) {
            size_t rem = n - off;
            if (cursor_ + rem <= (size_t)RATE) { copy_out(out + off, st_ + CAP + cursor_, rem); cursor_ += (int)rem; return; }
            size_t take = RATE - cursor_;
            copy_out(out + off, st_ + CAP + cursor_, take);
            if (rem != (size_t)RATE) permute();   // arkworks 0.4 squeeze_internal tests the remaining length before advancing
            off += take; cursor_ = 0;
        }
    }
    void squeeze_bytes(uint8_t* out, size_t n) {
        int bits = 64 - __builtin_clzll(F::P); size_t usable = (size_t)(bits - 1) / 8, ne = (n + usable - 1) / usable;
        std::vector<u64> el(ne); squeeze_base(el.data(), ne);
        size_t w = 0; for (size_t e = 0; e < ne && w < n; ++e) for (size_t b = 0; b < usable && w < n; ++b) out[w++] = (uint8_t)(el[e] >> (8 * b));
    }
    // ---- LatticeFold transcript surface
    void absorb(const u64* ring_el) { absorb_base(ring_el, Rg::D); }
    void absorb_slice(const u64* els, size_t count) { for (size_t i = 0; i < count; ++i) absorb(els + i * Rg::D); }
    void absorb_sf(const u64* sf) { auto e = HostRing<Rg>::from_sf(sf); absorb(e.data()); }
    void absorb_u64(u64 x) { auto e = HostRing<Rg>::from_u64(x); absorb(e.data()); }
    void absorb_tag(const char* tag) { u128 acc = 0; for (const char* c = tag; *c; ++c) acc = ((acc << 8) | (uint8_t)*c) % F::P; u64 sf[Rg::TAU] = {0}; sf[0] = (u64)acc; absorb_sf(sf); }
    void get_challenge(u64* sf) { squeeze_base(sf, Rg::TAU); absorb_base(sf, Rg::TAU); }
    // challenge-set decode: crates/cyclotomic-rings/src/rings/goldilocks.rs:32-68 (18 bytes -> 24 six-bit values - 32)
    void get_short_challenge(u64* coeffs) {
        uint8_t bs[32]; squeeze_bytes(bs, Rg::CS_BYTES); short_challenge_from_bytes(bs, coeffs);
    }
    static void short_challenge_from_bytes(const uint8_t* bs, u64* coeffs) {
        std::memset(coeffs, 0, 8 * Rg::D);
        if (Rg::CS_BYTES == 16) {      // frog.rs:32-56: 16 bytes -> byte - 128
            for (int i = 0; i < 16; ++i) coeffs[i] = F::from_i64((int64_t)bs[i] - 128);
            return;
        }
        for (int i = 0; i < 6; ++i) {  // goldilocks.rs:32-68 / babybear.rs:32-68: 18 bytes -> 24 six-bit values - 32
            const uint8_t b0 = bs[3 * i], b1 = bs[3 * i + 1], b2 = bs[3 * i + 2];
            int v[4] = {b0 & 0x3F, ((b0 >> 6) & 3) | ((b1 & 0x0F) << 2), ((b1 >> 4) & 0x0F) | ((b2 & 3) << 4), (b2 >> 2) & 0x3F};
            for (int j = 0; j < 4; ++j) coeffs[4 * i + j] = F::from_i64(v[j] - 32);
        }
    }
};

}  // namespace lf
