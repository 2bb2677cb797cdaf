// TEST INFRASTRUCTURE ONLY (see ring.hpp header).
// Poseidon duplex sponge + LatticeFold transcript, restating
//   crates/latticefold/src/transcript/poseidon.rs:29-75   (PoseidonTranscript: absorb / get_challenge / squeeze_bytes)
//   crates/latticefold/src/transcript.rs:13-51            (absorb_field_element, absorb_slice, get_challenges)
// The sponge itself is ark-crypto-primitives 0.4.0 `PoseidonSponge` (reference Cargo.lock:60-62, not in the
// tree); its published duplex algorithm is restated below and PINNED by the reference KATs
// transcript/poseidon.rs:86-142 (tests/golden/transcript_goldilocks.json).
#pragma once
#include "ring.hpp"
#include "poseidon_w24_tables.inc"

namespace lfo {

struct PoseidonSponge {
    const RingParams* R;
    static const int W = POSEIDON_W24_WIDTH, RATE = POSEIDON_W24_RATE, CAP = POSEIDON_W24_CAP;
    u64 st[W];
    bool absorbing; int idx;       // DuplexSpongeMode::{Absorbing{next_absorb_index}, Squeezing{next_squeeze_index}}
    std::vector<u64> ark, mds;

    explicit PoseidonSponge(const RingParams& r) : R(&r) {
        memset(st, 0, sizeof st); absorbing = true; idx = 0;
        const int nr = POSEIDON_W24_FULL + POSEIDON_W24_PARTIAL;
        ark.resize(nr * W); mds.resize(W * W);
        for (int i = 0; i < nr * W; ++i) ark[i] = POSEIDON_W24_ARK[i] % r.F.p;
        for (int i = 0; i < W * W; ++i) mds[i] = POSEIDON_W24_MDS[i] % r.F.p;
    }
    void permute() {
        const Fp& F = R->F; const int full = POSEIDON_W24_FULL, part = POSEIDON_W24_PARTIAL;
        for (int r = 0; r < full + part; ++r) {
            for (int i = 0; i < W; ++i) st[i] = F.add(st[i], ark[r * W + i]);
            bool is_full = r < full / 2 || r >= full / 2 + part;
            int n_sbox = is_full ? W : 1;
            for (int i = 0; i < n_sbox; ++i) st[i] = F.pow(st[i], POSEIDON_W24_ALPHA);
            u64 ns[W];
            for (int i = 0; i < W; ++i) { u64 acc = 0; for (int j = 0; j < W; ++j) acc = F.add(acc, F.mul(mds[i * W + j], st[j])); ns[i] = acc; }
            memcpy(st, ns, sizeof st);
        }
    }
    void absorb(const u64* el, size_t n) {
        if (n == 0) return;
        int i;
        if (absorbing) { i = idx; if (i == RATE) { permute(); i = 0; } } else { permute(); i = 0; }
        size_t off = 0;
        for (;;) {
            if (i + (n - off) <= (size_t)RATE) {
                for (size_t j = off; j < n; ++j) st[CAP + i + (j - off)] = R->F.add(st[CAP + i + (j - off)], el[j]);
                absorbing = true; idx = i + (int)(n - off); return;
            }
            int take = RATE - i;
            for (int j = 0; j < take; ++j) st[CAP + i + j] = R->F.add(st[CAP + i + j], el[off + j]);
            permute(); off += take; i = 0;
        }
    }
    void squeeze(u64* out, size_t n) {
        if (n == 0) return;
        int i;
        if (absorbing) { permute(); i = 0; } else { i = idx; if (i == RATE) { permute(); i = 0; } }
        size_t off = 0;
        for (;;) {
            if (i + (n - off) <= (size_t)RATE) {
                for (size_t j = off; j < n; ++j) out[j] = st[CAP + i + (j - off)];
                absorbing = false; idx = i + (int)(n - off); return;
            }
            int take = RATE - i;
            for (int j = 0; j < take; ++j) out[off + j] = st[CAP + i + j];
            // arkworks 0.4 squeeze_internal, literally: "Unless we are done with squeezing in this call, permute",
            // tested as `output_remaining.len() != rate` BEFORE the slice is advanced.
            if (n - off != (size_t)RATE) permute();
            off += take; i = 0;
        }
    }
    // squeeze_bytes: ceil(n/usable) elements, low `usable` little-endian bytes of each; usable = (bits(p)-1)/8
    void squeeze_bytes(uint8_t* out, size_t n) {
        int bits = 64 - __builtin_clzll(R->F.p); size_t usable = (size_t)(bits - 1) / 8;
        size_t ne = (n + usable - 1) / usable; std::vector<u64> el(ne); squeeze(el.data(), ne);
        size_t w = 0;
        for (size_t e = 0; e < ne && w < n; ++e) for (size_t b = 0; b < usable && w < n; ++b) out[w++] = (uint8_t)(el[e] >> (8 * b));
    }
};

struct Transcript {
    const RingParams* R; PoseidonSponge sp;
    explicit Transcript(const RingParams& r) : R(&r), sp(r) {}
    void absorb(const u64* ring_el) { sp.absorb(ring_el, R->d); }                       // poseidon.rs:40-47
    void absorb_slice(const u64* els, size_t n) { for (size_t i = 0; i < n; ++i) absorb(els + i * R->d); }
    void absorb_sf(const u64* sf) { std::vector<u64> e(R->d); ntt_from_sf(*R, e.data(), sf); absorb(e.data()); }  // transcript.rs:20-22
    void absorb_u64(u64 x) { std::vector<u64> e(R->d); ntt_from_u64(*R, e.data(), x); absorb(e.data()); }
    void absorb_tag(const char* tag) {                                                   // from_be_bytes_mod_order(tag), nifs.rs:180
        u128 acc = 0; for (const char* c = tag; *c; ++c) acc = ((acc << 8) | (uint8_t)*c) % R->F.p;
        u64 sf[16] = {0}; sf[0] = (u64)acc; absorb_sf(sf);
    }
    void get_challenge(u64* sf) { sp.squeeze(sf, R->tau); sp.absorb(sf, R->tau); }        // poseidon.rs:49-57
    void get_short_challenge(u64* coeffs) {                                              // poseidon.rs:69-74
        uint8_t bs[32]; sp.squeeze_bytes(bs, R->cs_bytes); short_challenge_from_bytes(*R, bs, coeffs);
    }
};

}  // namespace lfo
