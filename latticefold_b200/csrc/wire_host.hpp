// Proof wire format: the byte image of `LFProof::serialize_with_mode(.., Compress::Yes)` (crates/latticefold/src/nifs.rs:28-34,
// examples/e2e.rs:126-146) for the flat u64 proof layout of this library.  Host code, no GPU.
//
// ark-serialize 0.4 rules restated (the derive macros serialise struct fields in declaration order):
//   Vec<T>            u64 little-endian length, then the elements
//   prime field Fp    ceil(bits(p) / 8) little-endian bytes of the canonical value (8 for Goldilocks / Frog, 4 for BabyBear);
//                     deserialisation rejects values >= p.  Compress::Yes and ::No coincide (no curve points anywhere)
//   extension field   its base-field coordinates in order
// and, for the un-vendored stark-rings ring type, ASSUMED (parity unpinned): an NTT-form ring element is its
// S slot-field elements in slot order with no length prefix (a fixed-size array).
//   LFProof             = LinearizationProof | DecompositionProof (acc) | DecompositionProof (new) | FoldingProof
//   LinearizationProof  = sumcheck::Proof | v: Vec<R> | u: Vec<R>                          nifs/linearization/structs.rs:14-40
//   DecompositionProof  = u_s | v_s | x_s: Vec<Vec<R>> each | y_s: Vec<Commitment{val: Vec<R>}>   nifs/decomposition/structs.rs:18-46
//   FoldingProof        = sumcheck::Proof | theta_s | eta_s: Vec<Vec<R>>                   nifs/folding/structs.rs:17-45
//   sumcheck::Proof     = Vec<ProverMsg{evaluations: Vec<R>}>                              utils/sumcheck.rs:41-42, sumcheck/prover.rs:13-17
#pragma once
#include "ring_host.hpp"
#include "../../include/lf_b200.h"
#include <vector>
#include <cstring>

namespace lf {

template <class Rg> struct Wire {
    typedef typename Rg::F F;
    static constexpr int D = Rg::D, TAU = Rg::TAU;
    static constexpr size_t FB = F::P >> 32 ? 8 : 4;                 // bytes per base-field element
    static constexpr size_t RB = (size_t)D * FB;                     // bytes per ring element
    struct Shape { size_t s, dlin, t, K, l, kappa, b; };
    static Shape shape(const lf_problem& P) { return Shape{(size_t)P.s, (size_t)P.d + 1, (size_t)P.t, (size_t)P.K, (size_t)P.l, (size_t)P.kappa, (size_t)P.b}; }
    static size_t vec_ring(size_t n) { return 8 + n * RB; }
    static size_t sumcheck_bytes(size_t rounds, size_t deg) { return 8 + rounds * vec_ring(deg + 1); }
    static size_t bytes(const lf_problem& P) {
        const Shape h = shape(P);
        const size_t lin = sumcheck_bytes(h.s, h.dlin) + vec_ring(TAU) + vec_ring(h.t);
        const size_t dec = (8 + h.K * vec_ring(h.t)) + (8 + h.K * vec_ring(TAU)) + (8 + h.K * vec_ring(h.l + 1)) + (8 + h.K * vec_ring(h.kappa));
        const size_t fold = sumcheck_bytes(h.s, 2 * h.b) + (8 + 2 * h.K * vec_ring(TAU)) + (8 + 2 * h.K * vec_ring(h.t));
        return lin + 2 * dec + fold;
    }
    // ---- writer
    struct Out { uint8_t* p; void u64le(u64 v) { for (int i = 0; i < 8; ++i) *p++ = (uint8_t)(v >> (8 * i)); }
                 void ring(const u64* e) { for (int i = 0; i < D; ++i) { u64 v = e[i]; for (size_t k = 0; k < FB; ++k) *p++ = (uint8_t)(v >> (8 * k)); } }
                 void vec(const u64* e, size_t n) { u64le(n); for (size_t i = 0; i < n; ++i) ring(e + i * D); } };
    // flat layout (prover.cuh): lin{msgs, v, u} | dec{per piece: x, y, u, v} x 2 | fold{msgs, theta, eta}
    static void serialize(const lf_problem& P, const u64* w, uint8_t* out) {
        const Shape h = shape(P); Out o{out};
        auto sumcheck = [&](size_t deg) { o.u64le(h.s); for (size_t r = 0; r < h.s; ++r) { o.vec(w, deg + 1); w += (deg + 1) * D; } };
        sumcheck(h.dlin); o.vec(w, TAU); w += (size_t)TAU * D; o.vec(w, h.t); w += h.t * D;
        for (int half = 0; half < 2; ++half) {
            const size_t per = ((h.l + 1) + h.kappa + h.t + TAU) * D; const u64* base = w;
            auto field = [&](size_t off, size_t n) { o.u64le(h.K); for (size_t k = 0; k < h.K; ++k) o.vec(base + k * per + off * D, n); };
            field((h.l + 1) + h.kappa, h.t);                 // u_s
            field((h.l + 1) + h.kappa + h.t, TAU);           // v_s
            field(0, h.l + 1);                               // x_s
            field(h.l + 1, h.kappa);                         // y_s (Commitment = its Vec)
            w += h.K * per;
        }
        sumcheck(2 * h.b);
        o.u64le(2 * h.K); for (size_t i = 0; i < 2 * h.K; ++i) { o.vec(w, TAU); w += (size_t)TAU * D; }
        o.u64le(2 * h.K); for (size_t i = 0; i < 2 * h.K; ++i) { o.vec(w, h.t); w += h.t * D; }
    }
    // ---- reader (validating: every length prefix must match the problem's shape, every field element must be canonical)
    struct In { const uint8_t* p; const uint8_t* end;
                u64 u64le() { need(8); u64 v = 0; for (int i = 0; i < 8; ++i) v |= (u64)*p++ << (8 * i); return v; }
                void need(size_t n) const { if ((size_t)(end - p) < n) throw LfException(LF_ERR_INCORRECT_LENGTH, "proof bytes: unexpected end of input"); }
                void len(u64 want, const char* what) { if (u64le() != want) throw LfException(LF_ERR_INCORRECT_LENGTH, std::string("proof bytes: wrong length prefix of ") + what); }
                void ring(u64* e) { need(RB); for (int i = 0; i < D; ++i) { u64 v = 0; for (size_t k = 0; k < FB; ++k) v |= (u64)*p++ << (8 * k);
                                      if (v >= F::P) throw LfException(LF_ERR_INVALID_ARG, "proof bytes: non-canonical field element"); e[i] = v; } }
                void vec(u64* e, size_t n, const char* what) { len(n, what); for (size_t i = 0; i < n; ++i) ring(e + i * D); } };
    static void deserialize(const lf_problem& P, const uint8_t* in, size_t n_bytes, u64* w) {
        const Shape h = shape(P); In r{in, in + n_bytes};
        auto sumcheck = [&](size_t deg, const char* what) { r.len(h.s, what); for (size_t i = 0; i < h.s; ++i) { r.vec(w, deg + 1, "a sumcheck message"); w += (deg + 1) * D; } };
        sumcheck(h.dlin, "the linearization sumcheck"); r.vec(w, TAU, "v"); w += (size_t)TAU * D; r.vec(w, h.t, "u"); w += h.t * D;
        for (int half = 0; half < 2; ++half) {
            const size_t per = ((h.l + 1) + h.kappa + h.t + TAU) * D; u64* base = w;
            auto field = [&](size_t off, size_t n, const char* what) { r.len(h.K, what); for (size_t k = 0; k < h.K; ++k) r.vec(base + k * per + off * D, n, what); };
            field((h.l + 1) + h.kappa, h.t, "u_s"); field((h.l + 1) + h.kappa + h.t, TAU, "v_s"); field(0, h.l + 1, "x_s"); field(h.l + 1, h.kappa, "y_s");
            w += h.K * per;
        }
        sumcheck(2 * h.b, "the folding sumcheck");
        r.len(2 * h.K, "theta_s"); for (size_t i = 0; i < 2 * h.K; ++i) { r.vec(w, TAU, "theta_s"); w += (size_t)TAU * D; }
        r.len(2 * h.K, "eta_s"); for (size_t i = 0; i < 2 * h.K; ++i) { r.vec(w, h.t, "eta_s"); w += h.t * D; }
        if (r.p != r.end) throw LfException(LF_ERR_INCORRECT_LENGTH, "proof bytes: trailing bytes");
    }
};

}  // namespace lf
