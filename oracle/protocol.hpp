// TEST INFRASTRUCTURE ONLY (see ring.hpp header).
// CPU restatement of the LatticeFold prover step and its verifier:
//   crates/latticefold/src/arith.rs:230-338                    Witness::{from_w_ccs, from_f, from_f_coeff, get_fhat}
//   crates/latticefold/src/arith/utils.rs:52-65                mat_vec_mul
//   crates/latticefold/src/utils/mle_helpers.rs:65-146         evaluate_mles, calculate_Mz_mles
//   crates/latticefold/src/commitment/commitment_scheme.rs     AjtaiCommitmentScheme::commit
//   crates/latticefold/src/nifs/linearization.rs:145-285       LFLinearization{Prover,Verifier}
//   crates/latticefold/src/nifs/decomposition.rs:33-293        LFDecomposition{Prover,Verifier}
//   crates/latticefold/src/nifs/folding.rs:42-370 + utils.rs   LFFolding{Prover,Verifier}
//   crates/latticefold/src/nifs.rs:48-197                      NIFSProver::prove / NIFSVerifier::verify
//   crates/cyclotomic-rings/src/rotation.rs:45-104             rot_sum / rot_lin_combination
// Loops that the reference parallelises with rayon (cfg_iter!/cfg_into_iter!) carry an OpenMP pragma on the
// same axis, so the CPU baseline uses the host cores the way `--features parallel` would.
#pragma once
#include "sumcheck.hpp"

namespace lfo {

typedef std::vector<u64> Vec;  // flat vector of ring elements: count * d limbs

struct DecompParams { u128 B; int L; u64 b; int K; };  // decomposition_parameters.rs:11-20

struct LfError : std::runtime_error { int code; LfError(int c, const std::string& m) : std::runtime_error(m), code(c) {} };
enum { ERR_WRONG_WITNESS_LEN = -1, ERR_LENGTHS_NOT_EQUAL = -2, ERR_MLE_LEN = -3, ERR_INVALID_SIZE_BOUNDS = -4,
       ERR_INCORRECT_LENGTH = -5, ERR_SUMCHECK_FAILED = -6, ERR_RECOMPOSED = -7, ERR_UNSUPPORTED = -8 };

struct SparseMatrix {  // stark-rings-linalg SparseMatrix{nrows,ncols,coeffs: Vec<Vec<(R,usize)>>} as CSR
    size_t nrows = 0, ncols = 0; std::vector<u64> row_ptr, col; Vec val;
};
struct CCS {  // arith.rs:51-74
    size_t m = 0, n = 0, l = 0, t = 0, q = 0, d = 0, s = 0;
    std::vector<SparseMatrix> M; std::vector<std::vector<int>> S; Vec c;
};
struct Witness { Vec w_ccs, f, f_coeff; std::vector<Mle> f_hat; };  // arith.rs:205-223
struct CCCS { Vec cm, x_ccs; };                                     // arith.rs:180-185
struct LCCCS { Vec r, v, cm, u, x_w, h; };                          // arith.rs:193-206

inline size_t count(const RingParams& R, const Vec& v) { return v.size() / R.d; }
inline int ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) ++l; return l; }

// ------------------------------------------------------------------ elementwise CRT / ICRT
inline Vec elementwise_crt(const RingParams& R, const Vec& a) {
    Vec o(a.size()); const long n = (long)count(R, a);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) crt(R, o.data() + i * R.d, a.data() + i * R.d);
    return o;
}
inline Vec elementwise_icrt(const RingParams& R, const Vec& a) {
    Vec o(a.size()); const long n = (long)count(R, a);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) icrt(R, o.data() + i * R.d, a.data() + i * R.d);
    return o;
}

// ------------------------------------------------------------------ decompositions (coefficient form)
// gadget_decompose(B, L): element i -> elements [i*L, (i+1)*L), digit index = power of B  (decomposition/utils.rs:25-26)
inline Vec gadget_decompose(const RingParams& R, const Vec& a, u128 B, int L) {
    const long n = (long)count(R, a); Vec o((size_t)n * L * R.d); bool bad = false;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) { try { decompose_elem(R, a.data() + i * R.d, B, L, o.data() + (size_t)i * L * R.d); } catch (...) { bad = true; } }
    if (bad) throw LfError(ERR_UNSUPPORTED, "gadget_decompose: a coefficient does not fit in L digits of base B");
    return o;
}
// decompose_to_vec(b, K).transpose(): K vectors of n elements, piece k has weight b^k  (decomposition/utils.rs:45-49)
inline std::vector<Vec> decompose_to_k_vecs(const RingParams& R, const Vec& a, u128 b, int K) {
    const long n = (long)count(R, a); std::vector<Vec> out(K, Vec(a.size())); bool bad = false;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        std::vector<u64> tmp((size_t)K * R.d);
        try { decompose_elem(R, a.data() + i * R.d, b, K, tmp.data()); } catch (...) { bad = true; continue; }
        for (int k = 0; k < K; ++k) memcpy(out[k].data() + i * R.d, tmp.data() + (size_t)k * R.d, 8 * R.d);
    }
    if (bad) throw LfError(ERR_UNSUPPORTED, "decompose_to_vec: a coefficient does not fit in K digits of base b");
    return out;
}
// gadget_recompose(B, L): w[i] = sum_l f[i*L+l] * B^l (either form; B acts as an integer scalar)
inline Vec gadget_recompose(const RingParams& R, const Vec& f, u128 B, int L) {
    const long n = (long)count(R, f) / L; Vec o((size_t)n * R.d);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) recompose_elems(R, f.data() + (size_t)i * L * R.d, L, B, o.data() + i * R.d);
    return o;
}

// ------------------------------------------------------------------ f-hat (arith.rs:273-297; KAT arith.rs:456-502)
inline std::vector<Mle> get_fhat(const RingParams& R, const Vec& f_coeff) {
    const size_t n = count(R, f_coeff); const int nv = ceil_log2(n ? n : 1);
    std::vector<Mle> fh(R.tau);
    for (int j = 0; j < R.tau; ++j) {
        Mle& m = fh[j]; m.nv = nv; m.d = R.d; m.ev.assign(n * R.d, 0);
        for (size_t i = 0; i < n; ++i) for (int k = 0; k < R.S; ++k) m.ev[i * R.d + (size_t)k * R.tau] = f_coeff[i * R.d + (size_t)j * R.S + k];
        size_t len = n; while (len > 0 && el_is_zero(R, m.at(len - 1))) --len;   // truncate_lnze
        m.ev.resize(len * R.d);
    }
    return fh;
}
inline Witness witness_from_f_coeff(const RingParams& R, const DecompParams& P, Vec f_coeff) {  // arith.rs:324-338
    Witness w; w.f = elementwise_crt(R, f_coeff); w.f_hat = get_fhat(R, f_coeff);
    w.w_ccs = gadget_recompose(R, w.f, P.B, P.L); w.f_coeff = std::move(f_coeff); return w;
}
inline Witness witness_from_f(const RingParams& R, const DecompParams& P, Vec f) {              // arith.rs:299-313
    Witness w; w.f_coeff = elementwise_icrt(R, f); w.f_hat = get_fhat(R, w.f_coeff);
    w.w_ccs = gadget_recompose(R, f, P.B, P.L); w.f = std::move(f); return w;
}
inline Witness witness_from_w_ccs(const RingParams& R, const DecompParams& P, Vec w_ccs) {      // arith.rs:230-248
    Witness w; Vec wc = elementwise_icrt(R, w_ccs);
    w.f_coeff = gadget_decompose(R, wc, P.B, P.L); w.f = elementwise_crt(R, w.f_coeff); w.f_hat = get_fhat(R, w.f_coeff);
    w.w_ccs = std::move(w_ccs); return w;
}

// ------------------------------------------------------------------ Ajtai commitment (commitment_scheme.rs:37-55)
struct Ajtai { size_t kappa = 0, n = 0; Vec A; /* kappa x n row-major, NTT form */ };
inline Vec ajtai_commit(const RingParams& R, const Ajtai& S, const Vec& f) {
    if (count(R, f) != S.n) throw LfError(ERR_WRONG_WITNESS_LEN, "WrongWitnessLength");
    Vec cm(S.kappa * R.d, 0); const int d = R.d;
    #pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < (long)S.kappa; ++i) {
        u64 acc[128] = {0}, t[128];
        const u64* row = S.A.data() + (size_t)i * S.n * d;
        for (size_t j = 0; j < S.n; ++j) { ntt_mul(R, t, row + j * d, f.data() + j * d); el_add(R, acc, acc, t); }
        memcpy(cm.data() + (size_t)i * d, acc, 8 * d);
    }
    return cm;
}

// ------------------------------------------------------------------ sparse mat-vec and Mz MLEs
inline Vec mat_vec_mul(const RingParams& R, const SparseMatrix& M, const Vec& z) {  // arith/utils.rs:52-65
    if (M.ncols != count(R, z)) throw LfError(ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
    Vec o(M.nrows * R.d, 0); const int d = R.d;
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)M.nrows; ++r) {
        u64 t[128];
        for (u64 e = M.row_ptr[r]; e < M.row_ptr[r + 1]; ++e) {
            ntt_mul(R, t, M.val.data() + e * d, z.data() + M.col[e] * d); el_add(R, o.data() + r * d, o.data() + r * d, t);
        }
    }
    return o;
}
inline std::vector<Mle> calculate_Mz_mles(const RingParams& R, const CCS& ccs, const Vec& z) {  // mle_helpers.rs:137-146
    std::vector<Mle> out;
    for (const auto& M : ccs.M) {
        Vec mz = mat_vec_mul(R, M, z);
        if (((size_t)1 << ccs.s) < count(R, mz)) throw LfError(ERR_MLE_LEN, "IncorrectLength");
        out.push_back(mle_from(R, (int)ccs.s, mz.data(), count(R, mz)));
    }
    return out;
}
// evaluate_mles (mle_helpers.rs:65-88); point = ring elements
inline Vec evaluate_mles(const RingParams& R, const std::vector<Mle>& mles, const Vec& point) {
    const int np = (int)count(R, point); Vec out(mles.size() * R.d);
    for (const auto& m : mles) if (m.nv != np) throw LfError(ERR_MLE_LEN, "IncorrectLength");
    #pragma omp parallel for schedule(dynamic, 1)
    for (long k = 0; k < (long)mles.size(); ++k) mle_evaluate(R, mles[k], point.data(), np, out.data() + k * R.d);
    return out;
}
inline Vec point_to_ring(const RingParams& R, const std::vector<std::vector<u64>>& pt) {
    Vec o(pt.size() * R.d); for (size_t i = 0; i < pt.size(); ++i) ntt_from_sf(R, o.data() + i * R.d, pt[i].data()); return o;
}
inline Vec get_z_vector(const RingParams& R, const Vec& x, const u64* h, const Vec& w) {  // arith.rs:399-421
    Vec z; z.reserve(x.size() + R.d + w.size()); z.insert(z.end(), x.begin(), x.end()); z.insert(z.end(), h, h + R.d);
    z.insert(z.end(), w.begin(), w.end()); return z;
}
inline void sanity_check(const CCS& ccs, const DecompParams& P) {  // nifs.rs:165-173
    size_t want = std::max((ccs.n - ccs.l - 1) * (size_t)P.L, ccs.m); size_t p2 = 1; while (p2 < want) p2 <<= 1;
    if (ccs.m != p2) throw LfError(ERR_INVALID_SIZE_BOUNDS, "InvalidSizeBounds");
}
inline std::vector<u64> squeeze_challenges(const RingParams& R, Transcript& T, const char* tag, int n) {  // n slot-field elems
    T.absorb_tag(tag); std::vector<u64> out((size_t)n * R.tau);
    for (int i = 0; i < n; ++i) T.get_challenge(out.data() + (size_t)i * R.tau);
    return out;
}
inline Vec sf_to_ring_vec(const RingParams& R, const std::vector<u64>& sf) {
    size_t n = sf.size() / R.tau; Vec o(n * R.d); for (size_t i = 0; i < n; ++i) ntt_from_sf(R, o.data() + i * R.d, sf.data() + i * R.tau); return o;
}

// ------------------------------------------------------------------ linearization (nifs/linearization.rs)
struct LinearizationProof { SumcheckProof sumcheck; Vec v, u; };
inline Comb lin_comb(const RingParams& R, const CCS& ccs) {  // linearization/utils.rs:90-107
    Comb C; C.kind = COMB_LIN;
    for (size_t i = 0; i < ccs.q; ++i) { C.coef.emplace_back(ccs.c.begin() + i * R.d, ccs.c.begin() + (i + 1) * R.d); C.idx.push_back(ccs.S[i]); }
    return C;
}
inline void linearization_prove(const RingParams& R, const CCCS& cm_i, const Witness& wit, Transcript& T, const CCS& ccs,
                                LCCCS& out, LinearizationProof& proof) {
    std::vector<u64> one(R.d); ntt_from_u64(R, one.data(), 1);
    Vec z = get_z_vector(R, cm_i.x_ccs, one.data(), wit.w_ccs);
    Vec beta = sf_to_ring_vec(R, squeeze_challenges(R, T, "beta_s", (int)ccs.s));  // linearization/utils.rs:113-124
    std::vector<Mle> Mz = calculate_Mz_mles(R, ccs, z);
    std::vector<Mle> g;                                                             // linearization/utils.rs:63-88
    for (size_t i = 0; i < ccs.q; ++i) { if (el_is_zero(R, ccs.c.data() + i * R.d)) continue; for (int j : ccs.S[i]) g.push_back(Mz.at(j)); }
    g.push_back(build_eq_x_r(R, beta.data(), (int)ccs.s));
    Comb C = lin_comb(R, ccs);
    for (auto& ix : C.idx) for (int j : ix) if ((size_t)j >= g.size()) throw LfError(ERR_INCORRECT_LENGTH, "comb index outside MLE list");
    std::vector<std::vector<u64>> pt;
    proof.sumcheck = prove_as_subprotocol(R, T, std::move(g), (int)ccs.s, (int)ccs.d + 1, C, pt);
    Vec r = point_to_ring(R, pt);
    proof.v = evaluate_mles(R, wit.f_hat, r); proof.u = evaluate_mles(R, Mz, r);
    T.absorb_slice(proof.v.data(), count(R, proof.v)); T.absorb_slice(proof.u.data(), count(R, proof.u));
    out.r = r; out.v = proof.v; out.cm = cm_i.cm; out.u = proof.u; out.x_w = cm_i.x_ccs; out.h = one;
}
inline void linearization_verify(const RingParams& R, const CCCS& cm_i, const LinearizationProof& proof, Transcript& T, const CCS& ccs, LCCCS& out) {
    Vec beta = sf_to_ring_vec(R, squeeze_challenges(R, T, "beta_s", (int)ccs.s));
    std::vector<u64> zero(R.d, 0), one(R.d); ntt_from_u64(R, one.data(), 1);
    SubClaim sc = verify_as_subprotocol(R, T, (int)ccs.s, (int)ccs.d + 1, zero.data(), proof.sumcheck);
    if (!sc.ok) throw LfError(ERR_SUMCHECK_FAILED, "linearization sumcheck failed");
    Vec r = point_to_ring(R, sc.point);
    std::vector<u64> e(R.d), acc(R.d, 0), term(R.d);
    eq_eval(R, r.data(), beta.data(), (int)ccs.s, e.data());
    for (size_t i = 0; i < ccs.q; ++i) {
        memcpy(term.data(), ccs.c.data() + i * R.d, 8 * R.d);
        for (int j : ccs.S[i]) ntt_mul(R, term.data(), term.data(), proof.u.data() + (size_t)j * R.d);
        el_add(R, acc.data(), acc.data(), term.data());
    }
    ntt_mul(R, acc.data(), acc.data(), e.data());
    if (acc != sc.expected) throw LfError(ERR_SUMCHECK_FAILED, "linearization evaluation claim failed");
    T.absorb_slice(proof.v.data(), count(R, proof.v)); T.absorb_slice(proof.u.data(), count(R, proof.u));
    out.r = r; out.v = proof.v; out.cm = cm_i.cm; out.u = proof.u; out.x_w = cm_i.x_ccs; out.h = one;
}

// ------------------------------------------------------------------ decomposition (nifs/decomposition.rs)
struct DecompositionProof { std::vector<Vec> u_s, v_s, x_s, y_s; };
// decompose_big_vec_into_k_vec_and_compose_back (decomposition/utils.rs:12-42)
inline std::vector<Vec> compute_x_s(const RingParams& R, const DecompParams& P, Vec x_w, const Vec& h) {
    x_w.insert(x_w.end(), h.begin(), h.end());
    Vec coeff = elementwise_icrt(R, x_w);
    Vec inB = gadget_decompose(R, coeff, P.B, P.L);
    std::vector<Vec> pieces = decompose_to_k_vecs(R, inB, P.b, P.K);  // K x (len*L)
    std::vector<Vec> out;
    for (auto& pc : pieces) out.push_back(elementwise_crt(R, gadget_recompose(R, pc, P.B, P.L)));
    return out;
}
inline void scale_u64(const RingParams& R, u64* el, u64 s) { for (int i = 0; i < R.d; ++i) el[i] = R.F.mul(el[i], s % R.F.p); }
inline void decomposition_prove(const RingParams& R, const DecompParams& P, const LCCCS& cm_i, const Witness& wit, Transcript& T,
                                const CCS& ccs, const Ajtai& scheme, std::vector<std::vector<Mle>>& mz_mles,
                                std::vector<LCCCS>& lcccs_s, std::vector<Witness>& wit_s, DecompositionProof& proof) {
    sanity_check(ccs, P);
    const int K = P.K, d = R.d;
    std::vector<Vec> f_s = decompose_to_k_vecs(R, wit.f_coeff, P.b, K);                 // decomposition.rs:162-167
    wit_s.clear(); for (auto& f : f_s) wit_s.push_back(witness_from_f_coeff(R, P, std::move(f)));
    proof.x_s = compute_x_s(R, P, cm_i.x_w, cm_i.h);
    // commit_witnesses (decomposition.rs:178-201): y_0 = cm - b*(y_1 + b*(y_2 + ...))
    proof.y_s.assign(K, Vec());
    for (int k = 1; k < K; ++k) proof.y_s[k] = ajtai_commit(R, scheme, wit_s[k].f);
    Vec bsum(scheme.kappa * d, 0);
    for (int k = K - 1; k >= 1; --k) for (size_t e = 0; e < scheme.kappa; ++e) { el_add(R, bsum.data() + e * d, bsum.data() + e * d, proof.y_s[k].data() + e * d); scale_u64(R, bsum.data() + e * d, P.b); }
    proof.y_s[0].resize(scheme.kappa * d);
    if (cm_i.cm.size() != bsum.size()) throw LfError(ERR_WRONG_WITNESS_LEN, "WrongCommitmentLength");
    for (size_t e = 0; e < scheme.kappa; ++e) el_sub(R, proof.y_s[0].data() + e * d, cm_i.cm.data() + e * d, bsum.data() + e * d);
    proof.v_s.clear(); for (int k = 0; k < K; ++k) proof.v_s.push_back(evaluate_mles(R, wit_s[k].f_hat, cm_i.r));   // :204-211
    mz_mles.clear();                                                                                                   // :229-256
    for (int k = 0; k < K; ++k) {
        Vec z = proof.x_s[k]; z.insert(z.end(), wit_s[k].w_ccs.begin(), wit_s[k].w_ccs.end());
        std::vector<Mle> ms;
        for (const auto& M : ccs.M) { Vec mz = mat_vec_mul(R, M, z); if (((size_t)1 << ccs.s) < count(R, mz)) throw LfError(ERR_MLE_LEN, "IncorrectLength"); ms.push_back(mle_from(R, (int)ccs.s, mz.data(), count(R, mz))); }
        mz_mles.push_back(std::move(ms));
    }
    proof.u_s.clear(); for (int k = 0; k < K; ++k) proof.u_s.push_back(evaluate_mles(R, mz_mles[k], cm_i.r));          // :214-228
    lcccs_s.clear();
    for (int k = 0; k < K; ++k) {
        const Vec& x = proof.x_s[k];
        T.absorb_slice(x.data(), count(R, x)); T.absorb_slice(proof.y_s[k].data(), count(R, proof.y_s[k]));
        T.absorb_slice(proof.u_s[k].data(), count(R, proof.u_s[k])); T.absorb_slice(proof.v_s[k].data(), count(R, proof.v_s[k]));
        if (x.empty()) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
        LCCCS L; L.r = cm_i.r; L.v = proof.v_s[k]; L.cm = proof.y_s[k]; L.u = proof.u_s[k];
        L.x_w.assign(x.begin(), x.end() - d); L.h.assign(x.end() - d, x.end()); lcccs_s.push_back(std::move(L));
    }
}
inline Vec recompose_vecs(const RingParams& R, const std::vector<Vec>& s, u64 b) {  // decomposition.rs:259-272
    if (s.empty()) throw LfError(ERR_RECOMPOSED, "RecomposedError");
    Vec out(s[0].size(), 0); u64 pw = 1 % R.F.p;
    for (const auto& si : s) { for (size_t i = 0; i < out.size() && i < si.size(); ++i) out[i] = R.F.add(out[i], R.F.mul(si[i], pw)); pw = R.F.mul(pw, b % R.F.p); }
    return out;
}
inline void decomposition_verify(const RingParams& R, const DecompParams& P, const LCCCS& cm_i, const DecompositionProof& proof, Transcript& T,
                                 std::vector<LCCCS>& lcccs_s) {
    const int d = R.d; lcccs_s.clear();
    for (size_t k = 0; k < proof.x_s.size() && k < proof.y_s.size() && k < proof.u_s.size() && k < proof.v_s.size(); ++k) {
        const Vec& x = proof.x_s[k];
        T.absorb_slice(x.data(), count(R, x)); T.absorb_slice(proof.y_s[k].data(), count(R, proof.y_s[k]));
        T.absorb_slice(proof.u_s[k].data(), count(R, proof.u_s[k])); T.absorb_slice(proof.v_s[k].data(), count(R, proof.v_s[k]));
        if (x.empty()) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
        LCCCS L; L.r = cm_i.r; L.v = proof.v_s[k]; L.cm = proof.y_s[k]; L.u = proof.u_s[k];
        L.x_w.assign(x.begin(), x.end() - d); L.h.assign(x.end() - d, x.end()); lcccs_s.push_back(std::move(L));
    }
    if (recompose_vecs(R, proof.y_s, P.b) != cm_i.cm) throw LfError(ERR_RECOMPOSED, "RecomposedError(y)");
    if (recompose_vecs(R, proof.v_s, P.b) != cm_i.v) throw LfError(ERR_RECOMPOSED, "RecomposedError(v)");
    if (recompose_vecs(R, proof.u_s, P.b) != cm_i.u) throw LfError(ERR_RECOMPOSED, "RecomposedError(u)");
    Vec x = recompose_vecs(R, proof.x_s, P.b);
    if (x.size() < (size_t)d) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
    Vec h(x.end() - d, x.end()); x.resize(x.size() - d);
    if (x != cm_i.x_w || h != cm_i.h) throw LfError(ERR_RECOMPOSED, "RecomposedError(x)");
}

// ------------------------------------------------------------------ RotSum (cyclotomic-rings/src/rotation.rs:45-104)
// rho: coefficient-form polynomials; theta_i: tau NTT elements -> flattened element-major/slot-minor into d slot-field
// values; acc[j] += embed(coeff_j(X^i * rho)) * flat[i]; result promoted back to tau NTT elements.
inline Vec rot_lin_combination(const RingParams& R, const std::vector<Vec>& rho_coeff, const std::vector<Vec>& theta) {
    const int d = R.d, t = R.tau; const size_t ne = count(R, theta.at(0));
    if (ne * R.S != (size_t)d) throw LfError(ERR_INCORRECT_LENGTH, "rot_sum: b.len() != dimension");
    std::vector<u64> acc((size_t)d * t, 0);  // d slot-field values
    for (size_t i = 0; i < rho_coeff.size(); ++i) {
        std::vector<u64> a(rho_coeff[i]); const u64* flat = theta[i].data();  // flat[e*S+s] = limbs at (e*d + s*tau)
        for (int bi = 0; bi < d; ++bi) {
            const u64* B = flat + (size_t)bi * t;  // element-major, slot-minor == plain memory order of NTT elements
            for (int j = 0; j < d; ++j) for (int l = 0; l < t; ++l) acc[(size_t)j * t + l] = R.F.add(acc[(size_t)j * t + l], R.F.mul(a[j], B[l]));
            coeff_mul_x(R, a.data());
        }
    }
    return acc;  // d slot values * tau limbs == tau NTT elements in memory order
}

// ------------------------------------------------------------------ folding (nifs/folding.rs, folding/utils.rs)
struct FoldingProof { SumcheckProof sumcheck; std::vector<Vec> theta_s, eta_s; };
inline void mle_add_assign(const RingParams& R, Mle& a, const Mle& b) {  // DenseMultilinearExtension += (zero() adopts the rhs shape)
    if (a.ev.empty() && a.nv == 0) { a.nv = b.nv; a.d = R.d; }
    if (a.nv != b.nv) throw LfError(ERR_MLE_LEN, "MLE += with different num_vars");
    if (b.ev.size() > a.ev.size()) a.ev.resize(b.ev.size(), 0);
    const long n = (long)b.len();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) el_add(R, a.at(i), a.at(i), b.at(i));
}
inline void mle_mul_assign(const RingParams& R, Mle& a, const u64* s) {
    const long n = (long)a.len();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) ntt_mul(R, a.at(i), a.at(i), s);
}
inline Mle horner_combine(const RingParams& R, const std::vector<std::vector<Mle>>& groups, size_t from, size_t to, const Vec& chal) {
    Mle comb; comb.d = R.d;                                  // folding.rs:208-226 / folding/utils.rs:524-546
    for (size_t i = from; i < to; ++i) {
        Mle m; m.d = R.d;
        for (size_t j = groups[i].size(); j-- > 0;) { mle_add_assign(R, m, groups[i][j]); mle_mul_assign(R, m, chal.data() + i * R.d); }
        mle_add_assign(R, comb, m);
    }
    return comb;
}
inline void compute_v0_u0_x0_cm_0(const RingParams& R, const std::vector<Vec>& rho_coeff, const Vec& rho, const std::vector<Vec>& theta_s,
                                  const std::vector<LCCCS>& cm_i_s, const std::vector<Vec>& eta_s, const CCS& ccs, Vec& v0, Vec& cm0, Vec& u0, Vec& x0) {
    const int d = R.d; u64 t[128];
    v0 = rot_lin_combination(R, rho_coeff, theta_s);
    cm0.assign(cm_i_s.empty() ? 0 : cm_i_s[0].cm.size(), 0); u0.assign(ccs.t * d, 0); x0.assign((ccs.l + 1) * d, 0);
    for (size_t i = 0; i < cm_i_s.size(); ++i) {
        const u64* r = rho.data() + i * d;
        for (size_t e = 0; e < count(R, cm0); ++e) { ntt_mul(R, t, cm_i_s[i].cm.data() + e * d, r); el_add(R, cm0.data() + e * d, cm0.data() + e * d, t); }
        for (size_t e = 0; e < ccs.t && e < count(R, eta_s[i]); ++e) { ntt_mul(R, t, eta_s[i].data() + e * d, r); el_add(R, u0.data() + e * d, u0.data() + e * d, t); }
        Vec xh = cm_i_s[i].x_w; xh.insert(xh.end(), cm_i_s[i].h.begin(), cm_i_s[i].h.end());
        for (size_t e = 0; e < ccs.l + 1 && e < count(R, xh); ++e) { ntt_mul(R, t, xh.data() + e * d, r); el_add(R, x0.data() + e * d, x0.data() + e * d, t); }
    }
}
struct FoldChallenges { Vec alpha, beta, zeta, mu; };
inline FoldChallenges squeeze_alpha_beta_zeta_mu(const RingParams& R, Transcript& T, int K, int log_m) {  // folding/utils.rs:51-96
    FoldChallenges c;
    c.alpha = sf_to_ring_vec(R, squeeze_challenges(R, T, "alpha_s", 2 * K));
    c.zeta = sf_to_ring_vec(R, squeeze_challenges(R, T, "zeta_s", 2 * K));
    c.mu = sf_to_ring_vec(R, squeeze_challenges(R, T, "mu_s", 2 * K - 1));
    Vec one(R.d); ntt_from_u64(R, one.data(), 1); c.mu.insert(c.mu.end(), one.begin(), one.end());
    c.beta = sf_to_ring_vec(R, squeeze_challenges(R, T, "beta_s", log_m));
    return c;
}
inline void get_rhos(const RingParams& R, Transcript& T, int K, std::vector<Vec>& rho_coeff, Vec& rho) {  // folding/utils.rs:116-131
    T.absorb_tag("rho_s"); rho_coeff.clear();
    for (int i = 0; i < 2 * K - 1; ++i) { Vec c(R.d); T.get_short_challenge(c.data()); rho_coeff.push_back(c); }
    Vec one(R.d, 0); one[0] = 1; rho_coeff.push_back(one);
    rho.resize((size_t)2 * K * R.d); for (int i = 0; i < 2 * K; ++i) crt(R, rho.data() + (size_t)i * R.d, rho_coeff[i].data());
}
inline void folding_prove(const RingParams& R, const DecompParams& P, const std::vector<LCCCS>& cm_i_s, std::vector<Witness>& w_s, Transcript& T,
                          const CCS& ccs, const std::vector<std::vector<Mle>>& mz_mles, LCCCS& out, Witness& w0, FoldingProof& proof) {
    sanity_check(ccs, P);
    const int K = P.K, d = R.d; const int log_m = (int)ccs.s;
    if ((int)cm_i_s.size() != 2 * K) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
    FoldChallenges ch = squeeze_alpha_beta_zeta_mu(R, T, K, log_m);
    std::vector<std::vector<Mle>> f_hat; for (auto& w : w_s) f_hat.push_back(w.f_hat);
    Mle Ms1 = horner_combine(R, mz_mles, 0, K, ch.zeta), Ms2 = horner_combine(R, mz_mles, K, 2 * K, ch.zeta);
    // create_sumcheck_polynomial (folding/utils.rs:200-259)
    if ((int)f_hat.size() != 2 * K || (int)count(R, ch.beta) != log_m) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
    std::vector<Mle> g;
    { Mle c1 = horner_combine(R, f_hat, 0, K, ch.alpha); mle_add_assign(R, c1, Ms1); g.push_back(build_eq_x_r(R, cm_i_s[0].r.data(), (int)count(R, cm_i_s[0].r))); g.push_back(std::move(c1)); }
    { Mle c2 = horner_combine(R, f_hat, K, 2 * K, ch.alpha); mle_add_assign(R, c2, Ms2); g.push_back(build_eq_x_r(R, cm_i_s[K].r.data(), (int)count(R, cm_i_s[K].r))); g.push_back(std::move(c2)); }
    g.push_back(build_eq_x_r(R, ch.beta.data(), log_m));
    for (auto& fh : f_hat) for (auto& m : fh) g.push_back(m);
    Comb C; C.kind = COMB_FOLD; C.n_mu = 2 * K; C.tau = R.tau; C.b = (int)P.b; C.mu = ch.mu;
    std::vector<std::vector<u64>> pt;
    proof.sumcheck = prove_as_subprotocol(R, T, std::move(g), log_m, 2 * (int)P.b, C, pt);
    Vec r0 = point_to_ring(R, pt);
    proof.theta_s.clear(); proof.eta_s.clear();
    for (auto& fh : f_hat) proof.theta_s.push_back(evaluate_mles(R, fh, r0));
    for (auto& mz : mz_mles) proof.eta_s.push_back(evaluate_mles(R, mz, r0));
    for (auto& th : proof.theta_s) T.absorb_slice(th.data(), count(R, th));
    for (auto& et : proof.eta_s) T.absorb_slice(et.data(), count(R, et));
    std::vector<Vec> rho_coeff; Vec rho; get_rhos(R, T, K, rho_coeff, rho);
    // compute_f_0 (folding.rs:258-268)
    const long n = (long)count(R, w_s[0].f); Vec f0((size_t)n * d, 0);
    #pragma omp parallel for schedule(static)
    for (long j = 0; j < n; ++j) { u64 t[128]; for (int i = 0; i < 2 * K; ++i) { ntt_mul(R, t, rho.data() + (size_t)i * d, w_s[i].f.data() + j * d); el_add(R, f0.data() + j * d, f0.data() + j * d, t); } }
    Vec v0, cm0, u0, x0; compute_v0_u0_x0_cm_0(R, rho_coeff, rho, proof.theta_s, cm_i_s, proof.eta_s, ccs, v0, cm0, u0, x0);
    out.r = r0; out.v = v0; out.cm = cm0; out.u = u0; out.h.assign(x0.end() - d, x0.end()); out.x_w.assign(x0.begin(), x0.end() - d);
    w0 = witness_from_f(R, P, std::move(f0));
}
inline void folding_verify(const RingParams& R, const DecompParams& P, const std::vector<LCCCS>& cm_i_s, const FoldingProof& proof, Transcript& T,
                           const CCS& ccs, LCCCS& out) {
    sanity_check(ccs, P);
    const int K = P.K, d = R.d, tau = R.tau; const int log_m = (int)ccs.s; u64 t[128], pw[128], acc[128];
    if ((int)cm_i_s.size() != 2 * K || (int)proof.theta_s.size() != 2 * K || (int)proof.eta_s.size() != 2 * K) throw LfError(ERR_INCORRECT_LENGTH, "IncorrectLength");
    FoldChallenges ch = squeeze_alpha_beta_zeta_mu(R, T, K, log_m);
    // calculate_claims (folding.rs:310-342)
    std::vector<u64> claim(d, 0);
    for (int i = 0; i < 2 * K; ++i) {
        const u64* al = ch.alpha.data() + (size_t)i * d; memcpy(pw, al, 8 * d);
        for (size_t j = 0; j < count(R, cm_i_s[i].v); ++j) { ntt_mul(R, t, pw, cm_i_s[i].v.data() + j * d); el_add(R, claim.data(), claim.data(), t); ntt_mul(R, pw, pw, al); }
        const u64* ze = ch.zeta.data() + (size_t)i * d; memcpy(pw, ze, 8 * d);
        for (size_t j = 0; j < count(R, cm_i_s[i].u); ++j) { ntt_mul(R, t, pw, cm_i_s[i].u.data() + j * d); el_add(R, claim.data(), claim.data(), t); ntt_mul(R, pw, pw, ze); }
    }
    SubClaim sc = verify_as_subprotocol(R, T, log_m, 2 * (int)P.b, claim.data(), proof.sumcheck);
    if (!sc.ok) throw LfError(ERR_SUMCHECK_FAILED, "folding sumcheck failed");
    Vec r0 = point_to_ring(R, sc.point);
    // verify_evaluation (folding.rs:271-308) + compute_sumcheck_claim_expected_value (folding/utils.rs:327-372)
    std::vector<u64> e_ast(d), should(d, 0), e_i(d), jj(d), prod(d), a(d), b(d);
    eq_eval(R, ch.beta.data(), r0.data(), log_m, e_ast.data());
    for (int i = 0; i < 2 * K; ++i) {
        eq_eval(R, cm_i_s[i].r.data(), r0.data(), log_m, e_i.data());
        const Vec& th = proof.theta_s[i]; const Vec& et = proof.eta_s[i];
        const u64* al = ch.alpha.data() + (size_t)i * d; memcpy(pw, al, 8 * d);
        for (int j = 0; j < tau && j < (int)count(R, th); ++j) { ntt_mul(R, t, pw, e_i.data()); ntt_mul(R, t, t, th.data() + (size_t)j * d); el_add(R, should.data(), should.data(), t); ntt_mul(R, pw, pw, al); }
        const u64* mu = ch.mu.data() + (size_t)i * d; memcpy(pw, mu, 8 * d); memset(acc, 0, 8 * d);
        for (int j = 0; j < tau && j < (int)count(R, th); ++j) {
            const u64* theta = th.data() + (size_t)j * d; ntt_from_u64(R, prod.data(), 1);
            for (u64 x = 1; x < P.b; ++x) { ntt_from_u64(R, jj.data(), x); el_sub(R, a.data(), theta, jj.data()); el_add(R, b.data(), theta, jj.data()); ntt_mul(R, a.data(), a.data(), b.data()); ntt_mul(R, prod.data(), prod.data(), a.data()); }
            ntt_mul(R, t, pw, theta); ntt_mul(R, t, t, prod.data()); el_add(R, acc, acc, t); ntt_mul(R, pw, pw, mu);
        }
        ntt_mul(R, acc, acc, e_ast.data()); el_add(R, should.data(), should.data(), acc);
        const u64* ze = ch.zeta.data() + (size_t)i * d; memcpy(pw, ze, 8 * d); memset(acc, 0, 8 * d);
        for (size_t j = 0; j < count(R, et); ++j) { ntt_mul(R, t, pw, et.data() + j * d); el_add(R, acc, acc, t); ntt_mul(R, pw, pw, ze); }
        ntt_mul(R, acc, acc, e_i.data()); el_add(R, should.data(), should.data(), acc);
    }
    if (should != sc.expected) throw LfError(ERR_SUMCHECK_FAILED, "folding evaluation claim failed");
    for (auto& th : proof.theta_s) T.absorb_slice(th.data(), count(R, th));
    for (auto& et : proof.eta_s) T.absorb_slice(et.data(), count(R, et));
    std::vector<Vec> rho_coeff; Vec rho; get_rhos(R, T, K, rho_coeff, rho);
    Vec v0, cm0, u0, x0; compute_v0_u0_x0_cm_0(R, rho_coeff, rho, proof.theta_s, cm_i_s, proof.eta_s, ccs, v0, cm0, u0, x0);
    out.r = r0; out.v = v0; out.cm = cm0; out.u = u0; out.h.assign(x0.end() - d, x0.end()); out.x_w.assign(x0.begin(), x0.end() - d);
}

// ------------------------------------------------------------------ NIFS (nifs.rs:48-197)
struct LFProof { LinearizationProof lin; DecompositionProof dl, dr; FoldingProof fold; };
inline void absorb_public_input(const RingParams& R, const LCCCS& acc, const CCCS& cm_i, Transcript& T) {
    T.absorb_tag("acc");
    T.absorb_slice(acc.r.data(), count(R, acc.r)); T.absorb_slice(acc.v.data(), count(R, acc.v)); T.absorb_slice(acc.cm.data(), count(R, acc.cm));
    T.absorb_slice(acc.u.data(), count(R, acc.u)); T.absorb_slice(acc.x_w.data(), count(R, acc.x_w)); T.absorb(acc.h.data());
    T.absorb_tag("cm_i");
    T.absorb_slice(cm_i.cm.data(), count(R, cm_i.cm)); T.absorb_slice(cm_i.x_ccs.data(), count(R, cm_i.x_ccs));
}
inline void nifs_prove(const RingParams& R, const DecompParams& P, const LCCCS& acc, const Witness& w_acc, const CCCS& cm_i, const Witness& w_i,
                       Transcript& T, const CCS& ccs, const Ajtai& scheme, LCCCS& out, Witness& w_out, LFProof& proof) {
    sanity_check(ccs, P);
    absorb_public_input(R, acc, cm_i, T);
    LCCCS lin; linearization_prove(R, cm_i, w_i, T, ccs, lin, proof.lin);
    std::vector<std::vector<Mle>> mz_l, mz_r; std::vector<LCCCS> lc_l, lc_r; std::vector<Witness> w_l, w_r;
    decomposition_prove(R, P, acc, w_acc, T, ccs, scheme, mz_l, lc_l, w_l, proof.dl);
    decomposition_prove(R, P, lin, w_i, T, ccs, scheme, mz_r, lc_r, w_r, proof.dr);
    for (auto& x : lc_r) lc_l.push_back(std::move(x));
    for (auto& x : w_r) w_l.push_back(std::move(x));
    for (auto& x : mz_r) mz_l.push_back(std::move(x));
    folding_prove(R, P, lc_l, w_l, T, ccs, mz_l, out, w_out, proof.fold);
}
inline void nifs_verify(const RingParams& R, const DecompParams& P, const LCCCS& acc, const CCCS& cm_i, const LFProof& proof, Transcript& T,
                        const CCS& ccs, LCCCS& out) {
    sanity_check(ccs, P);
    absorb_public_input(R, acc, cm_i, T);
    LCCCS lin; linearization_verify(R, cm_i, proof.lin, T, ccs, lin);
    std::vector<LCCCS> a, b;
    decomposition_verify(R, P, acc, proof.dl, T, a);
    decomposition_verify(R, P, lin, proof.dr, T, b);
    for (auto& x : b) a.push_back(std::move(x));
    folding_verify(R, P, a, proof.fold, T, ccs, out);
}

}  // namespace lfo
