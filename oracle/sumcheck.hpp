// TEST INFRASTRUCTURE ONLY (see ring.hpp header).
// Dense multilinear extensions, eq table, ring sumcheck prover + verifier, restating
//   crates/latticefold/src/utils/sumcheck.rs:53-104             (prove_as_subprotocol / verify_as_subprotocol)
//   crates/latticefold/src/utils/sumcheck/prover.rs:56-162      (prove_round, the Jolt-style inner loop)
//   crates/latticefold/src/utils/sumcheck/verifier.rs:40-254    (verify_round, check_and_generate_subclaim, interpolate_uni_poly)
//   crates/latticefold/src/utils/sumcheck/utils.rs:78-170       (eq_eval, build_eq_x_r)
//   crates/latticefold/src/utils/mle_helpers.rs:21-88           (evaluate_mles)
// DenseMultilinearExtension is stark-rings-poly @ 886a89f (not in the tree): evaluations may be shorter than
// 2^num_vars (missing tail = 0), fix_variables binds the LOWEST variable, evaluate(point) has point[0] on bit 0.
// Its values are fully determined by the maths (MLE evaluation is unique) and by prover.rs:112-124's indexing.
#pragma once
#include "transcript.hpp"
#include <functional>

namespace lfo {

struct Mle {
    int nv = 0; int d = 0;
    std::vector<u64> ev;  // len * d limbs
    size_t len() const { return d ? ev.size() / d : 0; }
    const u64* at(size_t i) const { return ev.data() + i * d; }
    u64* at(size_t i) { return ev.data() + i * d; }
};
inline Mle mle_from(const RingParams& R, int nv, const u64* data, size_t len) {
    if (len > ((size_t)1 << nv)) throw std::runtime_error("MLE longer than 2^num_vars");  // mle_helpers.rs:104-108
    Mle m; m.nv = nv; m.d = R.d; m.ev.assign(data, data + len * R.d); return m;
}
inline void mle_get(const RingParams& R, const Mle& m, size_t i, u64* out) {
    if (i < m.len()) memcpy(out, m.at(i), 8 * R.d); else memset(out, 0, 8 * R.d);
}
// fix the lowest variable to the ring element r: new[b] = old[2b] + r*(old[2b+1]-old[2b])
inline void mle_fix_low(const RingParams& R, Mle& m, const u64* r) {
    size_t half = (size_t)1 << (m.nv - 1); size_t nl = std::min(half, (m.len() + 1) / 2);
    std::vector<u64> out(nl * R.d);
    #pragma omp parallel for schedule(static) if (nl > 256)
    for (long i = 0; i < (long)nl; ++i) {
        u64 a[128], b[128], t[128];
        mle_get(R, m, 2 * (size_t)i, a); mle_get(R, m, 2 * (size_t)i + 1, b);
        el_sub(R, t, b, a); ntt_mul(R, t, t, r); el_add(R, out.data() + (size_t)i * R.d, a, t);
    }
    m.ev.swap(out); m.nv -= 1;
}
// evaluate at point (ring elements, point[0] on bit 0) by successive halving
inline void mle_evaluate(const RingParams& R, Mle m, const u64* point, int npoint, u64* out) {
    if (npoint != m.nv) throw std::runtime_error("MLE evaluate: point length != num_vars");  // -> MleEvaluationError::IncorrectLength
    for (int i = 0; i < npoint; ++i) mle_fix_low(R, m, point + (size_t)i * R.d);
    mle_get(R, m, 0, out);
}

// eq(x, r) for all x in {0,1}^s, r[0] on bit 0   (sumcheck/utils.rs:100-170)
inline Mle build_eq_x_r(const RingParams& R, const u64* r, int s) {
    if (s == 0) throw std::runtime_error("r length is 0");
    const int d = R.d; std::vector<u64> buf, one(d), t(d);
    ntt_from_u64(R, one.data(), 1);
    // recursion from the last variable down: start with [1-r_{s-1}, r_{s-1}]
    buf.resize(2 * d); el_sub(R, buf.data(), one.data(), r + (size_t)(s - 1) * d); memcpy(buf.data() + d, r + (size_t)(s - 1) * d, 8 * d);
    for (int v = s - 2; v >= 0; --v) {
        size_t n = buf.size() / d; std::vector<u64> res(2 * n * d);
        for (size_t i = 0; i < 2 * n; ++i) {
            const u64* bi = buf.data() + (i >> 1) * d; ntt_mul(R, t.data(), r + (size_t)v * d, bi);
            if ((i & 1) == 0) el_sub(R, res.data() + i * d, bi, t.data()); else memcpy(res.data() + i * d, t.data(), 8 * d);
        }
        buf.swap(res);
    }
    Mle m; m.nv = s; m.d = d; m.ev.swap(buf); return m;
}
// eq_eval(x, y) = prod (2 x_i y_i - x_i - y_i + 1)   (sumcheck/utils.rs:78-92)
inline void eq_eval(const RingParams& R, const u64* x, const u64* y, int n, u64* out) {
    const int d = R.d; std::vector<u64> res(d), one(d), t(d), u(d);
    ntt_from_u64(R, one.data(), 1); res = one;
    for (int i = 0; i < n; ++i) {
        ntt_mul(R, t.data(), x + (size_t)i * d, y + (size_t)i * d);
        el_add(R, u.data(), t.data(), t.data()); el_sub(R, u.data(), u.data(), x + (size_t)i * d); el_sub(R, u.data(), u.data(), y + (size_t)i * d);
        el_add(R, u.data(), u.data(), one.data()); ntt_mul(R, res.data(), res.data(), u.data());
    }
    memcpy(out, res.data(), 8 * d);
}

// ---------------------------------------------------------------- comb functions (the closures of the reference)
enum CombKind { COMB_PRODUCTS = 0, COMB_LIN = 1, COMB_FOLD = 2 };
struct Comb {
    int kind = COMB_PRODUCTS;
    // COMB_PRODUCTS (sumcheck/utils.rs:60-73 rand_poly_comb_fn) and COMB_LIN (linearization/utils.rs:90-107):
    std::vector<std::vector<u64>> coef;       // ring element per term
    std::vector<std::vector<int>> idx;        // MLE indices per term
    // COMB_FOLD (folding/utils.rs:273-325)
    int n_mu = 0, tau = 0, b = 0; std::vector<u64> mu;  // n_mu ring elements
};
inline void comb_eval(const RingParams& R, const Comb& C, const u64* vals /* M x d */, int M, u64* out) {
    const int d = R.d; u64 res[128], term[128], t[128]; memset(res, 0, 8 * d);
    if (C.kind == COMB_PRODUCTS || C.kind == COMB_LIN) {
        for (size_t i = 0; i < C.coef.size(); ++i) {
            if (C.kind == COMB_LIN && el_is_zero(R, C.coef[i].data())) continue;
            memcpy(term, C.coef[i].data(), 8 * d); bool skip = false;
            for (int j : C.idx[i]) {
                if (C.kind == COMB_LIN && el_is_zero(R, vals + (size_t)j * d)) { skip = true; break; }
                ntt_mul(R, term, term, vals + (size_t)j * d);
            }
            if (!skip) el_add(R, res, res, term);
        }
        if (C.kind == COMB_LIN) ntt_mul(R, res, res, vals + (size_t)(M - 1) * d);  // eq() is the last MLE
        memcpy(out, res, 8 * d); return;
    }
    // FOLD: v0*v1 + v2*v3 + sum_k mu-Horner over d of v4 * f * prod_{j=1}^{b-1}(f^2 - j^2)
    u64 inter[128], ev[128], f2[128], mult[128], jj[128];
    ntt_mul(R, res, vals, vals + d);
    ntt_mul(R, t, vals + 2 * (size_t)d, vals + 3 * (size_t)d); el_add(R, res, res, t);
    for (int k = 0; k < C.n_mu; ++k) {
        const u64* mu = C.mu.data() + (size_t)k * d;
        memset(inter, 0, 8 * d);
        for (int dd = C.tau - 1; dd >= 0; --dd) {
            const u64* f = vals + (size_t)(5 + k * C.tau + dd) * d;
            if (el_is_zero(R, f)) { if (!el_is_zero(R, inter)) ntt_mul(R, inter, inter, mu); continue; }
            memcpy(ev, vals + 4 * (size_t)d, 8 * d);
            ntt_mul(R, f2, f, f);
            for (int j = 1; j < C.b; ++j) {
                ntt_from_u64(R, jj, (u64)j * j); el_sub(R, mult, f2, jj);
                if (el_is_zero(R, mult)) { memset(ev, 0, 8 * d); break; }
                ntt_mul(R, ev, ev, mult);
            }
            ntt_mul(R, ev, ev, f);
            el_add(R, inter, inter, ev);
            ntt_mul(R, inter, inter, mu);
        }
        el_add(R, res, res, inter);
    }
    memcpy(out, res, 8 * d);
}

// ---------------------------------------------------------------- prover
struct SumcheckProof { int nvars = 0, degree = 0; std::vector<u64> msgs; /* nvars x (degree+1) x d */ };
struct ProverState { std::vector<std::vector<u64>> randomness; /* slot-field elems */ std::vector<Mle> mles; int nv, deg, round; };

// one round, prover.rs:56-162.  prev: slot-field challenge or nullptr
inline void prove_round(const RingParams& R, ProverState& st, const u64* prev, const Comb& C, u64* evals_out /* (deg+1) x d */) {
    const int d = R.d;
    if (prev) {
        if (st.round == 0) throw std::runtime_error("first round should be prover first.");
        st.randomness.emplace_back(prev, prev + R.tau);
        std::vector<u64> r(d); ntt_from_sf(R, r.data(), prev);
        for (auto& m : st.mles) mle_fix_low(R, m, r.data());
    } else if (st.round > 0) throw std::runtime_error("verifier message is empty");
    st.round += 1;
    if (st.round > st.nv) throw std::runtime_error("Prover is not active");
    const int M = (int)st.mles.size(), deg = st.deg; const size_t nb = (size_t)1 << (st.nv - st.round);
    std::vector<u64> evals((deg + 1) * d, 0);
    #pragma omp parallel if (nb > 16)
    {   // rayon fold over b (prover.rs:111-143) then reduce (prover.rs:145-161)
        std::vector<u64> le((deg + 1) * d, 0), v0(M * d), v1(M * d), steps(M * d), vals(M * d), lev(d);
        #pragma omp for schedule(static) nowait
        for (long b = 0; b < (long)nb; ++b) {
            for (int k = 0; k < M; ++k) { mle_get(R, st.mles[k], 2 * (size_t)b, v0.data() + (size_t)k * d); mle_get(R, st.mles[k], 2 * (size_t)b + 1, v1.data() + (size_t)k * d); }
            comb_eval(R, C, v0.data(), M, lev.data()); el_add(R, le.data(), le.data(), lev.data());
            comb_eval(R, C, v1.data(), M, lev.data()); el_add(R, le.data() + d, le.data() + d, lev.data());
            for (int k = 0; k < M; ++k) { el_sub(R, steps.data() + (size_t)k * d, v1.data() + (size_t)k * d, v0.data() + (size_t)k * d); }
            vals = v1;
            for (int e = 2; e <= deg; ++e) {
                for (int k = 0; k < M; ++k) el_add(R, vals.data() + (size_t)k * d, vals.data() + (size_t)k * d, steps.data() + (size_t)k * d);
                comb_eval(R, C, vals.data(), M, lev.data()); el_add(R, le.data() + (size_t)e * d, le.data() + (size_t)e * d, lev.data());
            }
        }
        #pragma omp critical
        for (int e = 0; e <= deg; ++e) el_add(R, evals.data() + (size_t)e * d, evals.data() + (size_t)e * d, le.data() + (size_t)e * d);
    }
    memcpy(evals_out, evals.data(), 8 * (size_t)(deg + 1) * d);
}

// sumcheck.rs:53-80.  Returns the proof; `point` gets nvars slot-field challenges; final state keeps the MLEs
// folded by all but the LAST challenge (sumcheck.rs:75-77).
inline SumcheckProof prove_as_subprotocol(const RingParams& R, Transcript& T, std::vector<Mle> mles, int nvars, int degree,
                                          const Comb& C, std::vector<std::vector<u64>>& point, ProverState* final_state = nullptr) {
    if (nvars == 0) throw std::runtime_error("Attempt to prove a constant.");
    T.absorb_u64((u64)nvars); T.absorb_u64((u64)degree);
    ProverState st; st.mles = std::move(mles); st.nv = nvars; st.deg = degree; st.round = 0;
    SumcheckProof pf; pf.nvars = nvars; pf.degree = degree; pf.msgs.resize((size_t)nvars * (degree + 1) * R.d);
    std::vector<u64> r(R.tau); bool have = false;
    for (int i = 0; i < nvars; ++i) {
        u64* msg = pf.msgs.data() + (size_t)i * (degree + 1) * R.d;
        prove_round(R, st, have ? r.data() : nullptr, C, msg);
        T.absorb_slice(msg, degree + 1);
        T.get_challenge(r.data()); have = true;
        T.absorb_sf(r.data());
    }
    st.randomness.emplace_back(r.begin(), r.end());
    point = st.randomness;
    if (final_state) *final_state = std::move(st);
    return pf;
}

// ---------------------------------------------------------------- verifier
// interpolate the degree<=len-1 polynomial through (i, p_i[i]) and evaluate at the slot-field point x (broadcast).
// verifier.rs:139-254 computes the same Lagrange sum with a particular operation order; field arithmetic is exact,
// so any order gives the identical element.
inline void interpolate_uni_poly(const RingParams& R, const u64* p_i, int len, const u64* x_sf, u64* out) {
    const int d = R.d, t = R.tau; std::vector<u64> res(d, 0), term(d), w(d);
    for (int i = 0; i < len; ++i) {
        // L_i(x) = prod_{j != i} (x - j)/(i - j)   in the slot field
        u64 num[16] = {0}, tmp[16], f[16]; num[0] = 1; u64 den = 1;
        for (int j = 0; j < len; ++j) if (j != i) {
            memcpy(f, x_sf, 8 * t); f[0] = R.F.sub(f[0], (u64)j); sf_mul(R, tmp, num, f); memcpy(num, tmp, 8 * t);
            den = R.F.mul(den, R.F.from_i128((i128)i - j));
        }
        u64 di = R.F.inv(den); for (int l = 0; l < t; ++l) num[l] = R.F.mul(num[l], di);
        ntt_from_sf(R, w.data(), num); ntt_mul(R, term.data(), p_i + (size_t)i * d, w.data()); el_add(R, res.data(), res.data(), term.data());
    }
    memcpy(out, res.data(), 8 * d);
}
struct SubClaim { std::vector<std::vector<u64>> point; std::vector<u64> expected; bool ok = false; };
// sumcheck.rs:84-104 + verifier.rs:92-123
inline SubClaim verify_as_subprotocol(const RingParams& R, Transcript& T, int nvars, int degree, const u64* claimed_sum, const SumcheckProof& pf) {
    const int d = R.d; SubClaim sc;
    T.absorb_u64((u64)nvars); T.absorb_u64((u64)degree);
    if (pf.nvars != nvars || pf.degree != degree) return sc;
    std::vector<u64> r(R.tau);
    for (int i = 0; i < nvars; ++i) {
        const u64* msg = pf.msgs.data() + (size_t)i * (degree + 1) * d;
        T.absorb_slice(msg, degree + 1);
        T.get_challenge(r.data()); sc.point.emplace_back(r.begin(), r.end());
        T.absorb_sf(r.data());
    }
    std::vector<u64> expected(claimed_sum, claimed_sum + d), s(d);
    for (int i = 0; i < nvars; ++i) {
        const u64* msg = pf.msgs.data() + (size_t)i * (degree + 1) * d;
        el_add(R, s.data(), msg, msg + d);
        if (memcmp(s.data(), expected.data(), 8 * d) != 0) return sc;  // SumCheckFailed
        interpolate_uni_poly(R, msg, degree + 1, sc.point[i].data(), expected.data());
    }
    sc.expected = expected; sc.ok = true; return sc;
}

}  // namespace lfo
