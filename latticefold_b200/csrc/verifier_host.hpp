// NIFSVerifier::verify (crates/latticefold/src/nifs.rs:117-162) for the product library: SURVEY.md 8(f) rank 2.
// The verifier touches no witness-sized data (a few hundred ring elements and the transcript), so like the reference's it is
// host code: it shares the Poseidon transcript, the ring tables and RotSum with the prover's host half and nothing with oracle/.
//   MLSumcheck::verify_as_subprotocol      utils/sumcheck.rs:84-104, sumcheck/verifier.rs:40-123 (+ interpolate_uni_poly :125-254)
//   LFLinearizationVerifier::verify        nifs/linearization.rs:192-285
//   LFDecompositionVerifier::verify        nifs/decomposition.rs:91-156, recompose :259-293
//   LFFoldingVerifier::verify              nifs/folding.rs:133-195, calculate_claims :310-342, verify_evaluation :271-308,
//                                          compute_sumcheck_claim_expected_value folding/utils.rs:327-372
#pragma once
#include "prover.cuh"

namespace lf {

template <class Rg> struct Verifier {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef HostRing<Rg> HR; typedef typename HR::El El;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU;
    const lf_problem& in; RingTables<Rg> tab; HR H;
    explicit Verifier(const lf_problem& p) : in(p), tab(), H(tab) {}

    static size_t cnt(const HV& v) { return v.size() / D; }
    static HV take(const u64*& p, size_t elems) { HV v(p, p + elems * D); p += elems * D; return v; }
    static El bcast(const u64* sf) { return HR::from_sf(sf); }
    static void acc_add(El& a, const El& b) { a = HR::add(a, b); }
    static std::vector<u64> squeeze(Transcript<Rg>& T, const char* tag, int n) { return Prover<Rg>::squeeze(T, tag, n); }

    // p(r) from p(0..len-1): Lagrange basis over the integers 0..len-1 (denominators are base-field integers), r a slot-field element
    static El interpolate(const u64* evals, int len, const u64* r) {
        std::vector<std::array<u64, Rg::TAU>> rj(len);
        for (int j = 0; j < len; ++j) { for (int l = 0; l < TAU; ++l) rj[j][l] = r[l]; rj[j][0] = F::sub(r[0], (u64)j % F::P); }
        El out = HR::zero();
        for (int i = 0; i < len; ++i) {
            u64 num[Rg::TAU] = {0}; num[0] = 1; int64_t den = 1;
            for (int j = 0; j < len; ++j) if (j != i) { u64 t[Rg::TAU]; SF::mul(t, num, rj[j].data()); std::memcpy(num, t, sizeof num); den *= (i - j); }
            const u64 dinv = F::inv(F::from_i64(den));
            for (int l = 0; l < TAU; ++l) num[l] = F::mul(num[l], dinv);
            acc_add(out, HR::mul(HR::load(evals + (size_t)i * D), bcast(num)));
        }
        return out;
    }
    struct SubClaim { std::vector<u64> point; El expected; };
    static SubClaim sumcheck_verify(Transcript<Rg>& T, int nvars, int degree, const El& claimed, const u64* msgs) {
        SubClaim sc; const int ne = degree + 1;
        T.absorb_u64((u64)nvars); T.absorb_u64((u64)degree);
        sc.point.assign((size_t)nvars * TAU, 0);
        for (int i = 0; i < nvars; ++i) { T.absorb_slice(msgs + (size_t)i * ne * D, ne); T.get_challenge(&sc.point[(size_t)i * TAU]); T.absorb_sf(&sc.point[(size_t)i * TAU]); }
        El expected = claimed;
        for (int i = 0; i < nvars; ++i) {
            const u64* msg = msgs + (size_t)i * ne * D;
            if (HR::add(HR::load(msg), HR::load(msg + D)) != expected) throw LfException(LF_ERR_SUMCHECK_FAILED, "SumCheckFailed: p(0) + p(1) differs from the running claim");
            expected = interpolate(msg, ne, &sc.point[(size_t)i * TAU]);
        }
        sc.expected = expected; return sc;
    }
    // eq(x, y) = prod_i (x_i y_i + (1 - x_i)(1 - y_i))   (sumcheck/utils.rs:78-98)
    static El eq_eval(const u64* x, const u64* y, int n) {
        const El one = HR::from_u64(1); El acc = one;
        for (int i = 0; i < n; ++i) { El a = HR::load(x + (size_t)i * D), b = HR::load(y + (size_t)i * D);
            acc = HR::mul(acc, HR::add(HR::mul(a, b), HR::mul(HR::sub(one, a), HR::sub(one, b)))); }
        return acc;
    }
    // sum_j s[j] * b^j  (decomposition.rs:259-272)
    HV recompose(const std::vector<HV>& s) const {
        if (s.empty()) throw LfException(LF_ERR_RECOMPOSED, "RecomposedError");
        HV out(s[0].size(), 0); u64 pw = 1;
        for (const HV& si : s) { for (size_t i = 0; i < out.size() && i < si.size(); ++i) out[i] = F::add(out[i], F::mul(si[i], pw)); pw = F::mul(pw, in.b % F::P); }
        return out;
    }

    LCCCS verify_linearization(const HV& cm, const HV& x_ccs, const u64* msgs, const HV& v, const HV& u, Transcript<Rg>& T) const {
        const int s = (int)in.s;
        HV beta = Prover<Rg>::sf_to_ring(squeeze(T, "beta_s", s));
        SubClaim sc = sumcheck_verify(T, s, (int)in.d + 1, HR::zero(), msgs);
        HV r = Prover<Rg>::sf_to_ring(sc.point);
        El acc = HR::zero(); const int32_t* sp = in.S_flat;
        for (size_t i = 0; i < in.q; ++i) { El term = HR::load(in.c + i * D);
            for (int f = 0; f < in.S_len[i]; ++f) { int j = *sp++; if (j < 0 || (size_t)j >= cnt(u)) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength"); term = HR::mul(term, HR::load(&u[(size_t)j * D])); }
            acc_add(acc, term); }
        if (HR::mul(acc, eq_eval(r.data(), beta.data(), s)) != sc.expected) throw LfException(LF_ERR_SUMCHECK_FAILED, "linearization: evaluation claim does not match the sumcheck");
        T.absorb_slice(v.data(), cnt(v)); T.absorb_slice(u.data(), cnt(u));
        LCCCS o; o.r = r; o.v = v; o.cm = cm; o.u = u; o.x_w = x_ccs; El one = HR::from_u64(1); o.h.assign(one.begin(), one.end()); return o;
    }
    struct DecProof { std::vector<HV> x_s, y_s, u_s, v_s; };
    std::vector<LCCCS> verify_decomposition(const LCCCS& cm, const DecProof& pf, Transcript<Rg>& T) const {
        std::vector<LCCCS> out;
        for (int k = 0; k < in.K; ++k) {
            const HV& x = pf.x_s[k];
            T.absorb_slice(x.data(), cnt(x)); T.absorb_slice(pf.y_s[k].data(), cnt(pf.y_s[k])); T.absorb_slice(pf.u_s[k].data(), cnt(pf.u_s[k])); T.absorb_slice(pf.v_s[k].data(), cnt(pf.v_s[k]));
            if (x.empty()) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
            LCCCS L; L.r = cm.r; L.v = pf.v_s[k]; L.cm = pf.y_s[k]; L.u = pf.u_s[k]; L.x_w.assign(x.begin(), x.end() - D); L.h.assign(x.end() - D, x.end());
            out.push_back(std::move(L));
        }
        if (recompose(pf.y_s) != cm.cm) throw LfException(LF_ERR_RECOMPOSED, "RecomposedError: commitments");
        if (recompose(pf.v_s) != cm.v) throw LfException(LF_ERR_RECOMPOSED, "RecomposedError: v");
        if (recompose(pf.u_s) != cm.u) throw LfException(LF_ERR_RECOMPOSED, "RecomposedError: u");
        HV x = recompose(pf.x_s);
        if (x.size() < (size_t)D) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
        HV h(x.end() - D, x.end()); x.resize(x.size() - D);
        if (x != cm.x_w || h != cm.h) throw LfException(LF_ERR_RECOMPOSED, "RecomposedError: x");
        return out;
    }
    // Horner-free power sum: sum_j c^(j+1) * v[j]
    static El pow_sum(const u64* c_sf, const HV& v, size_t limit) {
        El acc = HR::zero(), pw = bcast(c_sf); const El c = pw;
        for (size_t j = 0; j < cnt(v) && j < limit; ++j) { acc_add(acc, HR::mul(pw, HR::load(&v[j * D]))); pw = HR::mul(pw, c); }
        return acc;
    }
    LCCCS verify_folding(const std::vector<LCCCS>& lcs, const u64* msgs, const std::vector<HV>& theta, const std::vector<HV>& eta, Transcript<Rg>& T) const {
        const int K = in.K, s = (int)in.s;
        if ((int)lcs.size() != 2 * K || (int)theta.size() != 2 * K || (int)eta.size() != 2 * K) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
        std::vector<u64> alpha = squeeze(T, "alpha_s", 2 * K), zeta = squeeze(T, "zeta_s", 2 * K), mu = squeeze(T, "mu_s", 2 * K - 1);
        { u64 one[Rg::TAU] = {0}; one[0] = 1; mu.insert(mu.end(), one, one + TAU); }
        HV beta = Prover<Rg>::sf_to_ring(squeeze(T, "beta_s", s));
        El claim = HR::zero();                                                    // calculate_claims
        for (int i = 0; i < 2 * K; ++i) { acc_add(claim, pow_sum(&alpha[(size_t)i * TAU], lcs[i].v, ~(size_t)0)); acc_add(claim, pow_sum(&zeta[(size_t)i * TAU], lcs[i].u, ~(size_t)0)); }
        SubClaim sc = sumcheck_verify(T, s, 2 * (int)in.b, claim, msgs);
        HV r0 = Prover<Rg>::sf_to_ring(sc.point);
        const El e_ast = eq_eval(beta.data(), r0.data(), s); El should = HR::zero();
        for (int i = 0; i < 2 * K; ++i) {
            const El e_i = eq_eval(lcs[i].r.data(), r0.data(), s);
            acc_add(should, HR::mul(e_i, pow_sum(&alpha[(size_t)i * TAU], theta[i], TAU)));
            El acc = HR::zero(), pw = bcast(&mu[(size_t)i * TAU]); const El m = pw;      // sum_j mu^(j+1) theta_j prod_{x=1}^{b-1} (theta_j^2 - x^2)
            for (int j = 0; j < TAU && j < (int)cnt(theta[i]); ++j) {
                const El th = HR::load(&theta[i][(size_t)j * D]); El prod = HR::from_u64(1);
                for (u64 x = 1; x < in.b; ++x) { const El xe = HR::from_u64(x); prod = HR::mul(prod, HR::mul(HR::sub(th, xe), HR::add(th, xe))); }
                acc_add(acc, HR::mul(HR::mul(pw, th), prod)); pw = HR::mul(pw, m);
            }
            acc_add(should, HR::mul(acc, e_ast));
            acc_add(should, HR::mul(e_i, pow_sum(&zeta[(size_t)i * TAU], eta[i], ~(size_t)0)));
        }
        if (should != sc.expected) throw LfException(LF_ERR_SUMCHECK_FAILED, "folding: evaluation claim does not match the sumcheck");
        for (const HV& th : theta) T.absorb_slice(th.data(), cnt(th));
        for (const HV& et : eta) T.absorb_slice(et.data(), cnt(et));
        T.absorb_tag("rho_s");                                                     // get_rhos, folding/utils.rs:116-131
        std::vector<El> rho_coeff, rho;
        for (int i = 0; i < 2 * K - 1; ++i) { El c; T.get_short_challenge(c.data()); rho_coeff.push_back(c); }
        { El one = HR::zero(); one[0] = 1; rho_coeff.push_back(one); }
        for (const El& c : rho_coeff) rho.push_back(H.crt(c));
        LCCCS o; o.r = r0; o.v = Prover<Rg>::rot_lin_combination(rho_coeff, theta);   // compute_v0_u0_x0_cm_0, folding/utils.rs:460-521
        const size_t kappa = cnt(lcs[0].cm), t = in.t;
        o.cm.assign(kappa * D, 0); o.u.assign(t * D, 0); HV x0((in.l + 1) * D, 0);
        auto axpy = [&](HV& acc, size_t e, const u64* v, const El& r) { El p = HR::mul(HR::load(v), r); for (int l = 0; l < D; ++l) acc[e * D + l] = F::add(acc[e * D + l], p[l]); };
        for (int i = 0; i < 2 * K; ++i) {
            for (size_t e = 0; e < kappa && e < cnt(lcs[i].cm); ++e) axpy(o.cm, e, &lcs[i].cm[e * D], rho[i]);
            for (size_t e = 0; e < t && e < cnt(eta[i]); ++e) axpy(o.u, e, &eta[i][e * D], rho[i]);
            HV xh = lcs[i].x_w; xh.insert(xh.end(), lcs[i].h.begin(), lcs[i].h.end());
            for (size_t e = 0; e < in.l + 1 && e < cnt(xh); ++e) axpy(x0, e, &xh[e * D], rho[i]);
        }
        o.h.assign(x0.end() - D, x0.end()); o.x_w.assign(x0.begin(), x0.end() - D);
        return o;
    }

    // LFLinearizationVerifier::verify alone (linearization.rs:192-285): lin_proof = msgs | v | u
    LCCCS verify_linearization_only(const u64* lin_proof, Transcript<Rg>& T) const {
        const size_t words = ((size_t)in.s * (in.d + 2) + TAU + in.t) * D;
        for (size_t i = 0; i < words; ++i) if (lin_proof[i] >= F::P) throw LfException(LF_ERR_INVALID_ARG, "non-canonical field element in the proof");
        HV cm_i = Prover<Rg>::load_canonical(in.cm_i_cm, in.kappa, "cm_i"), x_ccs = Prover<Rg>::load_canonical(in.cm_i_x_ccs, in.l, "x_ccs");
        const u64* p = lin_proof; const u64* msgs = p; p += (size_t)in.s * (in.d + 2) * D;
        HV v = take(p, TAU), u = take(p, in.t);
        return verify_linearization(cm_i, x_ccs, msgs, v, u, T);
    }
    // the whole step; throws LfException(LF_ERR_SUMCHECK_FAILED / LF_ERR_RECOMPOSED / LF_ERR_INCORRECT_LENGTH / ...) on rejection
    LCCCS verify(const u64* proof, Transcript<Rg>& T) const {
        { size_t want = std::max((size_t)(in.n_ccs - in.l - 1) * (size_t)in.L, (size_t)in.m), p2 = 1; while (p2 < want) p2 <<= 1;      // nifs.rs:165-173
          if (in.m != p2 || ((u64)1 << in.s) != in.m) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "InvalidSizeBounds"); }
        // Every limb that crosses the boundary must be a canonical representative (< p): the reference's arkworks deserialisation
        // rejects anything else, and a non-canonical limb would both make the proof malleable and overflow the lazily reduced host sums.
        auto canonical = [](const u64* p, size_t limbs, const char* what) { for (size_t i = 0; i < limbs; ++i) if (p[i] >= F::P) throw LfException(LF_ERR_INVALID_ARG, std::string("non-canonical field element in ") + what); };
        auto ld = [&](const u64* p, size_t n) { if (!p && n) throw LfException(LF_ERR_INVALID_ARG, "accumulator field is NULL"); canonical(p, n * D, "the public input"); return HV(p, p + n * D); };
        canonical(proof, (size_t)Prover<Rg>::proof_words_of(in), "the proof");
        for (size_t i = 0; i < in.q * (size_t)D; ++i) if (in.c[i] >= F::P) throw LfException(LF_ERR_INVALID_ARG, "non-canonical field element in the CCS constants");
        LCCCS acc; acc.r = ld(in.acc_r, in.s); acc.v = ld(in.acc_v, TAU); acc.cm = ld(in.acc_cm, in.kappa); acc.u = ld(in.acc_u, in.t); acc.x_w = ld(in.acc_x_w, in.l); acc.h = ld(in.acc_h, 1);
        HV cm_i = ld(in.cm_i_cm, in.kappa), x_ccs = ld(in.cm_i_x_ccs, in.l);
        T.absorb_tag("acc");                                                       // absorb_public_input, nifs.rs:175-197
        T.absorb_slice(acc.r.data(), cnt(acc.r)); T.absorb_slice(acc.v.data(), cnt(acc.v)); T.absorb_slice(acc.cm.data(), cnt(acc.cm));
        T.absorb_slice(acc.u.data(), cnt(acc.u)); T.absorb_slice(acc.x_w.data(), cnt(acc.x_w)); T.absorb(acc.h.data());
        T.absorb_tag("cm_i"); T.absorb_slice(cm_i.data(), cnt(cm_i)); T.absorb_slice(x_ccs.data(), cnt(x_ccs));
        const u64* p = proof; const int K = in.K;
        const u64* lin_msgs = p; p += (size_t)in.s * (in.d + 2) * D;
        HV lin_v = take(p, TAU), lin_u = take(p, in.t);
        DecProof dp[2];
        for (DecProof& d : dp) for (int k = 0; k < K; ++k) { d.x_s.push_back(take(p, in.l + 1)); d.y_s.push_back(take(p, in.kappa)); d.u_s.push_back(take(p, in.t)); d.v_s.push_back(take(p, TAU)); }
        const u64* fold_msgs = p; p += (size_t)in.s * (2 * in.b + 1) * D;
        std::vector<HV> theta, eta;
        for (int i = 0; i < 2 * K; ++i) theta.push_back(take(p, TAU));
        for (int i = 0; i < 2 * K; ++i) eta.push_back(take(p, in.t));
        LCCCS lin = verify_linearization(cm_i, x_ccs, lin_msgs, lin_v, lin_u, T);
        std::vector<LCCCS> a = verify_decomposition(acc, dp[0], T), b = verify_decomposition(lin, dp[1], T);
        a.insert(a.end(), b.begin(), b.end());
        return verify_folding(a, fold_msgs, theta, eta, T);
    }
};

}  // namespace lf
