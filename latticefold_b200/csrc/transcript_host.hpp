// Fiat-Shamir transcript of the product's host side: Poseidon duplex sponge (width 24 = rate 20 + capacity 4,
// 8 full + 22 partial rounds, x^7) driven exactly as the reference drives arkworks' PoseidonSponge:
//   crates/latticefold/src/transcript/poseidon.rs:29-75  (absorb ring element = its D base limbs; get_challenge squeezes
//   TAU limbs and absorbs them back; short challenges from squeeze_bytes)
//   crates/latticefold/src/transcript.rs:13-51           (absorb_field_element = broadcast to a ring element)
// Sponge algorithm: ark-crypto-primitives 0.4.0 PoseidonSponge (duplex with Absorbing/Squeezing cursor).
// Runs on the CPU by design (sequential); the MDS product uses lazily reduced 192-bit sums.
#pragma once
#include "ring_host.hpp"
#include "poseidon_w24_tables.inc"

namespace lf {

template <class Rg> class Transcript {
    typedef typename Rg::F F;
    static constexpr int W = POSEIDON_W24_WIDTH, RATE = POSEIDON_W24_RATE, CAP = POSEIDON_W24_CAP;
    static constexpr int RF = POSEIDON_W24_FULL, RP = POSEIDON_W24_PARTIAL;
    struct Tables { u64 ark[(RF + RP) * W]; u64 mds[W * W]; Tables() { for (int i = 0; i < (RF + RP) * W; ++i) ark[i] = POSEIDON_W24_ARK[i] % F::P; for (int i = 0; i < W * W; ++i) mds[i] = POSEIDON_W24_MDS[i] % F::P; } };
    static const Tables& tables() { static const Tables t; return t; }
    u64 st_[W];
    int cursor_;        // next rate lane to absorb into / squeeze from
    bool squeezing_;
    unsigned long long permutations_ = 0;

    static u64 pow7(u64 x) { u64 x2 = F::mul(x, x), x3 = F::mul(x2, x), x6 = F::mul(x3, x3); return F::mul(x6, x); }
    void permute() {
        const Tables& t = tables(); ++permutations_;
        for (int r = 0; r < RF + RP; ++r) {
            for (int i = 0; i < W; ++i) st_[i] = F::add(st_[i], t.ark[r * W + i]);
            if (r < RF / 2 || r >= RF / 2 + RP) { for (int i = 0; i < W; ++i) st_[i] = pow7(st_[i]); } else st_[0] = pow7(st_[0]);
            u64 nx[W];
            for (int i = 0; i < W; ++i) { Acc192 a; a.clear(); const u64* row = t.mds + i * W; for (int j = 0; j < W; ++j) a.mac(row[j], st_[j]); nx[i] = F::reduce192(a); }
            std::memcpy(st_, nx, sizeof st_);
        }
    }
public:
    Transcript() { std::memset(st_, 0, sizeof st_); cursor_ = 0; squeezing_ = false; }
    unsigned long long permutations() const { return permutations_; }

    void absorb_base(const u64* v, size_t n) {
        if (!n) return;
        if (squeezing_ || cursor_ == RATE) { permute(); cursor_ = 0; }
        squeezing_ = false;
        for (size_t i = 0; i < n; ++i) {
            if (cursor_ == RATE) { permute(); cursor_ = 0; }
            st_[CAP + cursor_] = F::add(st_[CAP + cursor_], v[i]); ++cursor_;
        }
    }
    void squeeze_base(u64* out, size_t n) {
        if (!n) return;
        if (!squeezing_ || cursor_ == RATE) { permute(); cursor_ = 0; }
        squeezing_ = true;
        size_t off = 0;
        for (;;) {
            size_t rem = n - off;
            if (cursor_ + rem <= (size_t)RATE) { std::memcpy(out + off, st_ + CAP + cursor_, 8 * rem); cursor_ += (int)rem; return; }
            size_t take = RATE - cursor_;
            std::memcpy(out + off, st_ + CAP + cursor_, 8 * take);
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
        static_assert(Rg::CS_BYTES == 18, "only the 6-bit challenge sets are implemented");
        for (int i = 0; i < 6; ++i) {
            const uint8_t b0 = bs[3 * i], b1 = bs[3 * i + 1], b2 = bs[3 * i + 2];
            int v[4] = {b0 & 0x3F, ((b0 >> 6) & 3) | ((b1 & 0x0F) << 2), ((b1 >> 4) & 0x0F) | ((b2 & 3) << 4), (b2 >> 2) & 0x3F};
            for (int j = 0; j < 4; ++j) coeffs[4 * i + j] = F::from_i64(v[j] - 32);
        }
    }
};

}  // namespace lf
