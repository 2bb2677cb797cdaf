// TEST INFRASTRUCTURE ONLY (see ring.hpp header).
// LatticeFold+ consumers of the commitment / sumcheck kernels, restated on the coefficient-form ring
// R = Z_q[X]/(X^d + 1) (`frog_ring::RqPoly`, BaseRing = Fq) exactly as the reference runs them -- dense ring-valued MLEs and
// a ring-valued sumcheck -- so that the product's base-field redesign is checked against the reference's own data flow:
//   crates/latticefold-plus/src/transcript.rs:16-56   PoseidonTranscript<R: OverField> (absorb = coefficients, get_challenge = 1 Fq)
//   crates/latticefold-plus/src/setchk.rs:39-262      ev, In::set_check   (monomial set check, Construction 4.2)
//   crates/latticefold-plus/src/setchk.rs:264-356     Out::verify, absorb_evaluations
//   crates/latticefold-plus/src/utils.rs:12-43        split
//   crates/latticefold-plus/src/utils.rs:49-86        tensor_product, tensor
//   crates/latticefold-plus/src/rgchk.rs:75-187       Rg::range_check
//   crates/latticefold-plus/src/rgchk.rs:190-246      Dcom::verify
//   crates/latticefold-plus/src/rgchk.rs:259-336      RgInstance::from_f (double commitment)
//   crates/latticefold-plus/src/cm.rs:57-601          Cm::prove, sumchecker, CmProof::verify, ComX, calculate_t_z
//   crates/latticefold-plus/src/r1cs.rs:72-165        ComR1CS::linearize, ComR1CSProof::verify
//   crates/latticefold-plus/src/mlin.rs:41-106        Mlin::mlin
//   crates/latticefold-plus/src/decomp.rs:32-127      Decomp::decompose, DecompProof::verify
//   crates/latticefold/src/utils/sumcheck.rs:53-104, sumcheck/prover.rs:56-162, sumcheck/verifier.rs:92-254 (generic over R: OverField)
// Third-party pieces restated from their published definition (stark-rings @ 886a89f, not in the tree):
//   exp(a) = sgn(a) X^a = X^(a mod d) for |a| < d/2, psi = sum_{0<i<d/2} i (X^{-i} + X^i), ct = constant coefficient
//   (LatticeFold+ paper, section 4.1: ct(psi * exp(a)) = a), Matrix::gadget_decompose / decompose_to_vec = the balanced
//   digits of ring.hpp, DenseMultilinearExtension as in sumcheck.hpp.
// PARITY UNPINNED in addition to ring.hpp's list: exp(0) (taken as X^0 = 1; 0 and X^{d/2} also satisfy every check of the
// reference) and the element order of Matrix::gadget_decompose inside `split` (taken row-major, the l digits of an entry
// adjacent).  The reference's tests for this path are prove -> verify acceptance / rejection tests (setchk.rs:358-495,
// rgchk.rs:344-433); tests/test_oracle_plus.py runs the same cases.
#pragma once
#include "sumcheck.hpp"

namespace lfo { namespace plus {

// ---------------------------------------------------------------- coefficient-form ring Z_q[X]/(X^d + 1)
inline void r_mul(const RingParams& R, u64* out, const u64* a, const u64* b) {
    const int d = R.d; const Fp& F = R.F;
    if (R.trinomial) { coeff_mul(R, out, a, b); return; }
    // operands with a single non-zero coefficient (constants, monomials) are a scaled rotation: same element, fewer products
    int na = 0, ia = 0, nb = 0, ib = 0;
    for (int i = 0; i < d; ++i) { if (a[i]) { ++na; ia = i; } if (b[i]) { ++nb; ib = i; } }
    u64 w[64];
    if (na == 0 || nb == 0) { memset(out, 0, 8 * d); return; }
    if (nb == 1 || na == 1) {
        const u64* x = nb == 1 ? a : b; const int sh = nb == 1 ? ib : ia; const u64 c = nb == 1 ? b[ib] : a[ia];
        for (int i = 0; i < d; ++i) { u64 v = x[i] ? F.mul(x[i], c) : 0; int k = i + sh; if (k >= d) { k -= d; v = F.neg(v); } w[k] = v; }
        memcpy(out, w, 8 * d); return;
    }
    memset(w, 0, sizeof w);
    for (int i = 0; i < d; ++i) { if (!a[i]) continue;
        for (int j = 0; j < d; ++j) { u64 pr = F.mul(a[i], b[j]); int k = i + j; if (k >= d) w[k - d] = F.sub(w[k - d], pr); else w[k] = F.add(w[k], pr); } }
    memcpy(out, w, 8 * d);
}
inline void r_const(const RingParams& R, u64* out, u64 c) { memset(out, 0, 8 * R.d); out[0] = c % R.F.p; }     // R::from(BaseRing)
inline void r_scale(const RingParams& R, u64* out, const u64* a, u64 c) { for (int i = 0; i < R.d; ++i) out[i] = a[i] ? R.F.mul(a[i], c) : 0; }
// ev(r, x) = sum_i r_i x^i   (setchk.rs:46-57)
inline u64 ev(const RingParams& R, const u64* r, u64 x) { u64 acc = 0, e = 1; for (int i = 0; i < R.d; ++i) { acc = R.F.add(acc, R.F.mul(r[i], e)); e = R.F.mul(e, x); } return acc; }
// exp, psi, ct (stark-rings; see the header)
inline void r_exp(const RingParams& R, u64* out, u64 c) {
    const i128 a = R.F.to_signed(c); const int d = R.d;
    if (a <= -(d / 2) || a >= d / 2) throw std::runtime_error("exp: exponent outside (-d/2, d/2)");
    memset(out, 0, 8 * d); out[(int)((a + d) % d)] = 1;
}
inline void r_psi(const RingParams& R, u64* out) {
    const int d = R.d; memset(out, 0, 8 * d);
    for (int i = 1; i < d / 2; ++i) { out[i] = R.F.add(out[i], (u64)i); out[d - i] = R.F.sub(out[d - i], (u64)i); }    // X^{-i} = -X^{d-i}
}
inline u64 r_ct(const u64* a) { return a[0]; }

// ---------------------------------------------------------------- transcript (latticefold-plus/src/transcript.rs)
struct PlusTranscript {
    const RingParams* R; PoseidonSponge sp;
    explicit PlusTranscript(const RingParams& r) : R(&r), sp(r) {}
    void absorb(const u64* el) { sp.absorb(el, R->d); }                                               // transcript.rs:37-44
    void absorb_slice(const u64* els, size_t n) { for (size_t i = 0; i < n; ++i) absorb(els + i * R->d); }
    void absorb_field(u64 c) { std::vector<u64> e(R->d); r_const(*R, e.data(), c); absorb(e.data()); }   // Transcript::absorb_field_element
    u64 get_challenge() { u64 c; sp.squeeze(&c, 1); sp.absorb(&c, 1); return c; }                      // transcript.rs:46-55 (extension degree 1)
    std::vector<u64> get_challenges(int n) { std::vector<u64> v(n); for (int i = 0; i < n; ++i) v[i] = get_challenge(); return v; }
};

// ---------------------------------------------------------------- dense MLEs over the coefficient ring
struct PMle { int nv = 0; std::vector<u64> ev; };      // ev: len x d, missing tail = 0
inline void pm_get(const RingParams& R, const PMle& m, size_t i, u64* out) { if ((i + 1) * R.d <= m.ev.size()) memcpy(out, m.ev.data() + i * R.d, 8 * R.d); else memset(out, 0, 8 * R.d); }
inline void pm_fix_low(const RingParams& R, PMle& m, u64 r) {      // fix_variables(&[R::from(r)])
    const int d = R.d; const size_t len = m.ev.size() / d, half = (size_t)1 << (m.nv - 1), nl = std::min(half, (len + 1) / 2);
    std::vector<u64> out(nl * d);
    #pragma omp parallel for schedule(static) if (nl > 1024)
    for (long i = 0; i < (long)nl; ++i) { u64 a[64], b[64];
        pm_get(R, m, 2 * (size_t)i, a); pm_get(R, m, 2 * (size_t)i + 1, b);
        for (int c = 0; c < d; ++c) out[(size_t)i * d + c] = R.F.add(a[c], R.F.mul(R.F.sub(b[c], a[c]), r)); }
    m.ev.swap(out); m.nv -= 1;
}
inline void pm_evaluate(const RingParams& R, PMle m, const u64* point, int np, u64* out) {
    if (np != m.nv) throw std::runtime_error("MLE evaluate: point length != num_vars");
    for (int i = 0; i < np; ++i) pm_fix_low(R, m, point[i]);
    pm_get(R, m, 0, out);
}
// eq(x, c) table with base-field point, c[0] on bit 0 (sumcheck/utils.rs:100-170), entries as constants of R
inline std::vector<u64> eq_table_base(const RingParams& R, const u64* c, int s) {
    std::vector<u64> t(1, 1);
    for (int v = s - 1; v >= 0; --v) { std::vector<u64> n2(t.size() * 2);
        for (size_t i = 0; i < t.size(); ++i) { u64 hi = R.F.mul(t[i], c[v]); n2[2 * i] = R.F.sub(t[i], hi); n2[2 * i + 1] = hi; }
        t.swap(n2); }
    return t;
}
inline u64 eq_eval_base(const RingParams& R, const u64* x, const u64* y, int n) {     // sumcheck/utils.rs:78-92
    u64 res = 1; for (int i = 0; i < n; ++i) { u64 xy = R.F.mul(x[i], y[i]); u64 t = R.F.add(R.F.sub(R.F.sub(R.F.add(xy, xy), x[i]), y[i]), 1); res = R.F.mul(res, t); } return res;
}

// ---------------------------------------------------------------- sumcheck over R (generic MLSumcheck with R = RqPoly)
typedef std::function<void(const u64* vals, u64* out)> CombFn;
struct PProof { int nvars = 0, degree = 0; std::vector<u64> msgs; };      // nvars x (degree+1) x d
inline void p_prove_round(const RingParams& R, std::vector<PMle>& mles, int nv, int deg, int round /* 1-based, after increment */, const CombFn& comb, u64* evals_out) {
    const int d = R.d, M = (int)mles.size(); const size_t nb = (size_t)1 << (nv - round);
    std::vector<u64> evals((size_t)(deg + 1) * d, 0);
    #pragma omp parallel if (nb > 64)
    {
        std::vector<u64> le((size_t)(deg + 1) * d, 0), v0((size_t)M * d), v1((size_t)M * d), st((size_t)M * d), vals((size_t)M * d), lev(d);
        #pragma omp for schedule(static) nowait
        for (long b = 0; b < (long)nb; ++b) {
            for (int k = 0; k < M; ++k) { pm_get(R, mles[k], 2 * (size_t)b, v0.data() + (size_t)k * d); pm_get(R, mles[k], 2 * (size_t)b + 1, v1.data() + (size_t)k * d); }
            comb(v0.data(), lev.data()); el_add(R, le.data(), le.data(), lev.data());
            comb(v1.data(), lev.data()); el_add(R, le.data() + d, le.data() + d, lev.data());
            for (size_t i = 0; i < (size_t)M * d; ++i) st[i] = R.F.sub(v1[i], v0[i]);
            vals = v1;
            for (int e = 2; e <= deg; ++e) {
                for (size_t i = 0; i < (size_t)M * d; ++i) vals[i] = R.F.add(vals[i], st[i]);
                comb(vals.data(), lev.data()); el_add(R, le.data() + (size_t)e * d, le.data() + (size_t)e * d, lev.data());
            }
        }
        #pragma omp critical
        for (int e = 0; e <= deg; ++e) el_add(R, evals.data() + (size_t)e * d, evals.data() + (size_t)e * d, le.data() + (size_t)e * d);
    }
    memcpy(evals_out, evals.data(), 8 * (size_t)(deg + 1) * d);
}
inline PProof p_prove(const RingParams& R, PlusTranscript& T, std::vector<PMle> mles, int nvars, int degree, const CombFn& comb, std::vector<u64>& point) {
    if (nvars == 0) throw std::runtime_error("Attempt to prove a constant.");
    T.absorb_field((u64)nvars); T.absorb_field((u64)degree);
    PProof pf; pf.nvars = nvars; pf.degree = degree; pf.msgs.resize((size_t)nvars * (degree + 1) * R.d); point.clear();
    u64 r = 0;
    for (int i = 0; i < nvars; ++i) {
        if (i > 0) { for (auto& m : mles) pm_fix_low(R, m, r); }
        u64* msg = pf.msgs.data() + (size_t)i * (degree + 1) * R.d;
        p_prove_round(R, mles, nvars, degree, i + 1, comb, msg);
        T.absorb_slice(msg, degree + 1);
        r = T.get_challenge(); T.absorb_field(r); point.push_back(r);
    }
    return pf;
}
// Lagrange interpolation through (i, p_i) evaluated at the base-field point x (verifier.rs:139-254 computes the same value)
inline void p_interpolate(const RingParams& R, const u64* p_i, int len, u64 x, u64* out) {
    const int d = R.d; std::vector<u64> res(d, 0);
    for (int i = 0; i < len; ++i) { u64 num = 1, den = 1;
        for (int j = 0; j < len; ++j) if (j != i) { num = R.F.mul(num, R.F.sub(x, (u64)j)); den = R.F.mul(den, R.F.from_i128((i128)i - j)); }
        const u64 w = R.F.mul(num, R.F.inv(den));
        for (int c = 0; c < d; ++c) res[c] = R.F.add(res[c], R.F.mul(p_i[(size_t)i * d + c], w)); }
    memcpy(out, res.data(), 8 * d);
}
struct PSubClaim { std::vector<u64> point, expected; bool ok = false; };
inline PSubClaim p_verify(const RingParams& R, PlusTranscript& T, int nvars, int degree, const u64* claimed, const PProof& pf) {
    const int d = R.d; PSubClaim sc;
    T.absorb_field((u64)nvars); T.absorb_field((u64)degree);
    if (pf.nvars != nvars || pf.degree != degree || pf.msgs.size() != (size_t)nvars * (degree + 1) * d) return sc;
    for (int i = 0; i < nvars; ++i) { T.absorb_slice(pf.msgs.data() + (size_t)i * (degree + 1) * d, degree + 1); u64 r = T.get_challenge(); sc.point.push_back(r); T.absorb_field(r); }
    std::vector<u64> expected(claimed, claimed + d), s(d);
    for (int i = 0; i < nvars; ++i) { const u64* msg = pf.msgs.data() + (size_t)i * (degree + 1) * d;
        el_add(R, s.data(), msg, msg + d); if (memcmp(s.data(), expected.data(), 8 * d) != 0) return sc;
        p_interpolate(R, msg, degree + 1, sc.point[i], expected.data()); }
    sc.expected = expected; sc.ok = true; return sc;
}

// ---------------------------------------------------------------- sparse matrices of ring elements (stark-rings-linalg SparseMatrix, CSR)
struct SparseR { size_t nrows = 0, ncols = 0; std::vector<u64> row_ptr, col, val; };      // val: nnz x d
inline std::vector<u64> sp_column_dense(const RingParams& R, const SparseR& M, size_t j) {      // row j of M^T scattered into a dense vector
    std::vector<u64> v(M.nrows * R.d, 0);
    for (size_t r = 0; r < M.nrows; ++r) for (u64 e = M.row_ptr[r]; e < M.row_ptr[r + 1]; ++e) if (M.col[e] == j) memcpy(v.data() + r * R.d, M.val.data() + e * R.d, 8 * R.d);
    return v;
}
inline std::vector<u64> sp_mul_vec(const RingParams& R, const SparseR& M, const std::vector<u64>& x) {      // try_mul_vec
    const int d = R.d; if (x.size() != M.ncols * d) throw std::runtime_error("sparse mat-vec: length mismatch");
    std::vector<u64> y(M.nrows * d, 0);
    #pragma omp parallel for schedule(static) if (M.nrows > 256)
    for (long r = 0; r < (long)M.nrows; ++r) { u64 t[64];
        for (u64 e = M.row_ptr[r]; e < M.row_ptr[r + 1]; ++e) { r_mul(R, t, M.val.data() + e * d, x.data() + M.col[e] * d); el_add(R, y.data() + (size_t)r * d, y.data() + (size_t)r * d, t); } }
    return y;
}

// ---------------------------------------------------------------- monomial set check (setchk.rs)
struct MonSet { bool matrix = true; SparseR M; std::vector<u64> v; };      // v: n x d
struct SetOut { int nvars = 0, n_mat = 0, ncols = 0, n_vec = 0, n_M = 0; std::vector<u64> r; PProof pf; std::vector<u64> e /* (1+n_M) x n_mat x ncols x d */, b /* n_vec x d */; };

inline void absorb_evaluations(const RingParams& R, const SetOut& o, PlusTranscript& T) {      // setchk.rs:346-356
    T.absorb_slice(o.e.data(), o.e.size() / R.d); T.absorb_slice(o.b.data(), o.b.size() / R.d);
}
inline SetOut set_check(const RingParams& R, int nvars, const std::vector<MonSet>& sets, const std::vector<SparseR>& M, PlusTranscript& T) {
    const int d = R.d; const Fp& F = R.F;
    std::vector<const SparseR*> Ms; std::vector<const std::vector<u64>*> ms;
    for (auto& s : sets) { if (s.matrix) Ms.push_back(&s.M); else ms.push_back(&s.v); }
    if (Ms.empty()) throw std::runtime_error("set_check needs at least one matrix set");      // setchk.rs:85 (Ms[0])
    const size_t ncols = Ms[0]->ncols, nrows = Ms[0]->nrows, N = (size_t)1 << nvars;
    if (nrows > N) throw std::runtime_error("set larger than 2^nvars");
    std::vector<PMle> mles; std::vector<u64> alphas;
    auto push_const_mle = [&](const std::vector<u64>& vals) { PMle m; m.nv = nvars; m.ev.assign(vals.size() * d, 0); for (size_t i = 0; i < vals.size(); ++i) m.ev[i * d] = vals[i]; mles.push_back(std::move(m)); };
    for (const SparseR* Mp : Ms) {
        if (Mp->ncols != ncols || Mp->nrows != nrows) throw std::runtime_error("matrix sets of different shapes");
        std::vector<u64> c = T.get_challenges(nvars); const u64 beta = T.get_challenge();
        for (size_t j = 0; j < ncols; ++j) {
            std::vector<u64> mj(nrows, 0), mpj(nrows);
            for (size_t r = 0; r < nrows; ++r) for (u64 e = Mp->row_ptr[r]; e < Mp->row_ptr[r + 1]; ++e) if (Mp->col[e] == j) mj[r] = ev(R, Mp->val.data() + e * d, beta);
            for (size_t r = 0; r < nrows; ++r) mpj[r] = F.mul(mj[r], mj[r]);
            push_const_mle(mj); push_const_mle(mpj);
        }
        push_const_mle(eq_table_base(R, c.data(), nvars));
        alphas.push_back(T.get_challenge());
    }
    for (const std::vector<u64>* m : ms) {
        if (m->size() != nrows * d) throw std::runtime_error("vector set of a different length");
        std::vector<u64> c = T.get_challenges(nvars); const u64 beta = T.get_challenge();
        std::vector<u64> mj(nrows), mpj(nrows);
        for (size_t r = 0; r < nrows; ++r) { mj[r] = ev(R, m->data() + r * d, beta); mpj[r] = F.mul(mj[r], mj[r]); }
        push_const_mle(mj); push_const_mle(mpj); push_const_mle(eq_table_base(R, c.data(), nvars));
        alphas.push_back(T.get_challenge());
    }
    const bool have_rc = Ms.size() > 1; const u64 rc = have_rc ? T.get_challenge() : 0;
    const size_t nM = Ms.size(), nv_ = ms.size();
    CombFn comb = [&](const u64* vals, u64* out) {      // setchk.rs:157-189, literally (including the early return when there is no rc)
        u64 lc[64], res[64], t[64], u[64], cst[64]; memset(lc, 0, 8 * d);
        for (size_t i = 0; i < nM; ++i) {
            const size_t s = i * (2 * ncols + 1); memset(res, 0, 8 * d);
            for (size_t j = 0; j < ncols; ++j) {
                r_mul(R, t, vals + (s + 2 * j) * d, vals + (s + 2 * j) * d); el_sub(R, t, t, vals + (s + 2 * j + 1) * d);
                r_scale(R, u, t, F.pow(alphas[i], j)); el_add(R, res, res, u);
            }
            r_mul(R, res, res, vals + (s + 2 * ncols) * d);
            if (!have_rc) { memcpy(out, res, 8 * d); return; }
            r_scale(R, t, res, F.pow(rc, i)); el_add(R, lc, lc, t);
        }
        for (size_t i = 0; i < nv_; ++i) {
            const size_t s = nM * (2 * ncols + 1) + 3 * i, ai = nM + i;
            r_mul(R, t, vals + s * d, vals + s * d); el_sub(R, t, t, vals + (s + 1) * d);
            r_const(R, cst, alphas[ai]); r_mul(R, res, t, cst); r_mul(R, res, res, vals + (s + 2) * d);
            if (!have_rc) { memcpy(out, res, 8 * d); return; }
            r_scale(R, t, res, F.pow(rc, ai)); el_add(R, lc, lc, t);
        }
        memcpy(out, lc, 8 * d);
    };
    SetOut o; o.nvars = nvars; o.n_mat = (int)nM; o.ncols = (int)ncols; o.n_vec = (int)nv_; o.n_M = (int)M.size();
    o.pf = p_prove(R, T, std::move(mles), nvars, 3, comb, o.r);
    // Step 3: evaluations of the columns (and of M_i * column) at r
    o.e.assign((size_t)(1 + M.size()) * nM * ncols * d, 0);
    for (size_t mi = 0; mi <= M.size(); ++mi) for (size_t i = 0; i < nM; ++i) for (size_t j = 0; j < ncols; ++j) {
        std::vector<u64> col = sp_column_dense(R, *Ms[i], j);
        PMle m; m.nv = nvars;
        if (mi == 0) m.ev = std::move(col); else m.ev = sp_mul_vec(R, M[mi - 1], col);
        pm_evaluate(R, std::move(m), o.r.data(), nvars, o.e.data() + ((mi * nM + i) * ncols + j) * d);
    }
    o.b.assign(nv_ * d, 0);
    for (size_t i = 0; i < nv_; ++i) { PMle m; m.nv = nvars; m.ev = *ms[i]; pm_evaluate(R, std::move(m), o.r.data(), nvars, o.b.data() + i * d); }
    absorb_evaluations(R, o, T);
    return o;
}
// setchk.rs:264-344.  Returns true iff the reference returns Ok(())
inline bool set_check_verify(const RingParams& R, const SetOut& o, PlusTranscript& T) {
    const int d = R.d; const Fp& F = R.F; const int nclaims = o.n_mat + o.n_vec, nv = o.nvars;
    struct Cba { std::vector<u64> c; u64 beta, alpha; }; std::vector<Cba> cba(nclaims);
    for (auto& x : cba) { x.c = T.get_challenges(nv); x.beta = T.get_challenge(); x.alpha = T.get_challenge(); }
    const bool have_rc = o.n_mat > 1; const u64 rc = have_rc ? T.get_challenge() : 1;
    std::vector<u64> zero(d, 0);
    PSubClaim sc = p_verify(R, T, nv, 3, zero.data(), o.pf);
    if (!sc.ok) return false;
    absorb_evaluations(R, o, T);
    u64 ver = 0;      // every term is a constant of R: the ring products of setchk.rs:299-336 stay in Fq
    for (int i = 0; i < o.n_mat; ++i) {
        const u64 eq = eq_eval_base(R, cba[i].c.data(), sc.point.data(), nv); u64 esum = 0;
        for (int j = 0; j < o.ncols; ++j) { const u64* ej = o.e.data() + ((size_t)i * o.ncols + j) * d;
            const u64 e1 = ev(R, ej, cba[i].beta), e2 = ev(R, ej, F.mul(cba[i].beta, cba[i].beta));
            esum = F.add(esum, F.mul(F.sub(F.mul(e1, e1), e2), F.pow(cba[i].alpha, j))); }
        ver = F.add(ver, F.mul(F.mul(eq, esum), F.pow(rc, i)));
    }
    for (int i = 0; i < o.n_vec; ++i) { const Cba& x = cba[o.n_mat + i];
        const u64 eq = eq_eval_base(R, x.c.data(), sc.point.data(), nv); const u64* bi = o.b.data() + (size_t)i * d;
        const u64 e1 = ev(R, bi, x.beta), e2 = ev(R, bi, F.mul(x.beta, x.beta));
        ver = F.add(ver, F.mul(F.mul(F.mul(eq, x.alpha), F.sub(F.mul(e1, e1), e2)), F.pow(rc, o.n_mat + i))); }
    if (sc.expected[0] != ver) return false;
    for (int c = 1; c < d; ++c) if (sc.expected[c]) return false;
    return true;
}

// ---------------------------------------------------------------- utils.rs
// split: gadget-decompose the double commitment, flatten to coefficients, pad to n (utils.rs:12-43)
inline std::vector<u64> split(const RingParams& R, const std::vector<u64>& com /* rows x cols x d */, size_t n, u128 b, int k) {
    const int d = R.d; const size_t ne = com.size() / d; std::vector<u64> tau; tau.reserve(n); std::vector<u64> dig((size_t)k * d);
    for (size_t e = 0; e < ne; ++e) { decompose_elem(R, com.data() + e * d, b, k, dig.data()); tau.insert(tau.end(), dig.begin(), dig.end()); }
    if (tau.size() >= n) throw std::runtime_error("small n unsupported, must be >= tau unpadded");      // utils.rs:34-40 (the `<` test)
    tau.resize(n, 0); return tau;
}
inline std::vector<u64> tensor(const RingParams& R, const u64* r, int n) {      // utils.rs:74-86: sequential (1 - r_i, r_i) products
    std::vector<u64> res(1, 1);
    for (int i = 0; i < n; ++i) { std::vector<u64> n2; n2.reserve(res.size() * 2); for (u64 a : res) { n2.push_back(R.F.mul(a, R.F.sub(1, r[i]))); n2.push_back(R.F.mul(a, r[i])); } res.swap(n2); }
    return res;
}

// ---------------------------------------------------------------- range check (rgchk.rs)
struct DecompParameters { u128 b; int k, l; };
struct FComs { std::vector<u64> cm_f, C_Mf, cm_mtau; };      // kappa x d each
struct RgInstance { size_t n = 0, kappa = 0; int k = 0;
    std::vector<std::vector<u64>> M_f;      // k matrices, n x d x d (dense, row-major)
    std::vector<u64> tau;                   // n base-field values
    std::vector<u64> m_tau, f;              // n x d
    std::vector<std::vector<u64>> comM_f;   // k matrices, kappa x d x d
    FComs fcoms; };
inline std::vector<u64> mat_mul_vec(const RingParams& R, const std::vector<u64>& A, size_t kappa, size_t n, const std::vector<u64>& x) {      // A.try_mul_vec
    const int d = R.d; std::vector<u64> y(kappa * d, 0);
    for (size_t r = 0; r < kappa; ++r) {
        #pragma omp parallel
        { std::vector<u64> acc(d, 0); u64 t[64];
          #pragma omp for schedule(static) nowait
          for (long i = 0; i < (long)n; ++i) { r_mul(R, t, A.data() + (r * n + i) * d, x.data() + (size_t)i * d); el_add(R, acc.data(), acc.data(), t); }
          #pragma omp critical
          el_add(R, y.data() + r * d, y.data() + r * d, acc.data()); }
    }
    return y;
}
inline RgInstance rg_from_f(const RingParams& R, const std::vector<u64>& f, const std::vector<u64>& A, size_t kappa, const DecompParameters& dp) {      // rgchk.rs:259-336
    const int d = R.d; const size_t n = f.size() / d; RgInstance I; I.n = n; I.kappa = kappa; I.k = dp.k; I.f = f;
    I.M_f.assign(dp.k, std::vector<u64>(n * d * d, 0));
    std::vector<i128> dg(dp.k); u64 e[64];
    for (size_t i = 0; i < n; ++i) for (int c = 0; c < d; ++c) {
        decompose_balanced(R.F, f[i * d + c], dp.b, dp.k, dg.data());
        for (int kk = 0; kk < dp.k; ++kk) { r_exp(R, e, R.F.from_i128(dg[kk])); memcpy(I.M_f[kk].data() + (i * d + c) * d, e, 8 * d); }
    }
    // comM_f[kk] = A * M_f[kk]  (kappa x d), com = hconcat
    std::vector<u64> com(kappa * (size_t)d * dp.k * d, 0);
    for (int kk = 0; kk < dp.k; ++kk) {
        std::vector<u64> cm(kappa * d * d, 0);
        for (size_t r = 0; r < kappa; ++r)
            #pragma omp parallel for schedule(static)
            for (int c = 0; c < d; ++c) { u64 acc[64] = {0}, t[64];
                for (size_t i = 0; i < n; ++i) { r_mul(R, t, A.data() + (r * n + i) * d, I.M_f[kk].data() + (i * d + c) * d); el_add(R, acc, acc, t); }
                memcpy(cm.data() + (r * d + c) * d, acc, 8 * d); }
        for (size_t r = 0; r < kappa; ++r) for (int c = 0; c < d; ++c) memcpy(com.data() + ((r * dp.k + kk) * d + c) * d, cm.data() + (r * d + c) * d, 8 * d);
        I.comM_f.push_back(std::move(cm));
    }
    I.tau = split(R, com, n, (u128)(d / 2), dp.l);
    I.m_tau.assign(n * d, 0); for (size_t i = 0; i < n; ++i) r_exp(R, I.m_tau.data() + i * d, I.tau[i]);
    std::vector<u64> tau_r(n * d, 0); for (size_t i = 0; i < n; ++i) tau_r[i * d] = I.tau[i];
    I.fcoms.cm_f = mat_mul_vec(R, A, kappa, n, f); I.fcoms.C_Mf = mat_mul_vec(R, A, kappa, n, tau_r); I.fcoms.cm_mtau = mat_mul_vec(R, A, kappa, n, I.m_tau);
    return I;
}
struct DcomEvals { std::vector<u64> v /* d */, a /* 1+n_M */, b /* (1+n_M) x d */, c /* (1+n_M) x d */; };
struct Dcom { std::vector<DcomEvals> evals; std::vector<FComs> fcoms; SetOut out; DecompParameters dp; };
inline SparseR sparse_from_dense(const RingParams& R, const std::vector<u64>& m, size_t nrows, size_t ncols) {      // SparseMatrix::from_dense: non-zero entries
    SparseR S; S.nrows = nrows; S.ncols = ncols; S.row_ptr.push_back(0); const int d = R.d;
    for (size_t r = 0; r < nrows; ++r) { for (size_t c = 0; c < ncols; ++c) { const u64* e = m.data() + (r * ncols + c) * d; if (!el_is_zero(R, e)) { S.col.push_back(c); S.val.insert(S.val.end(), e, e + d); } } S.row_ptr.push_back(S.col.size()); }
    return S;
}
inline Dcom range_check(const RingParams& R, int nvars, const std::vector<RgInstance>& inst, const DecompParameters& dp, const std::vector<SparseR>& M, PlusTranscript& T) {      // rgchk.rs:75-187
    const int d = R.d; std::vector<MonSet> sets;
    for (auto& I : inst) for (auto& m : I.M_f) { MonSet s; s.matrix = true; s.M = sparse_from_dense(R, m, I.n, d); sets.push_back(std::move(s)); }
    for (auto& I : inst) { MonSet s; s.matrix = false; s.v = I.m_tau; sets.push_back(std::move(s)); }
    Dcom D; D.dp = dp; D.out = set_check(R, nvars, sets, M, T);
    const std::vector<u64>& r = D.out.r;
    for (size_t l = 0; l < inst.size(); ++l) { const RgInstance& I = inst[l]; DcomEvals E; const size_t n = I.n;
        E.v.resize(d); { PMle m; m.nv = nvars; m.ev = I.f; std::vector<u64> o(d); pm_evaluate(R, std::move(m), r.data(), nvars, o.data()); E.v = o; }      // coefficient-wise MLEs of cf(f)
        std::vector<u64> tau_r(n * d, 0); for (size_t i = 0; i < n; ++i) tau_r[i * d] = I.tau[i];
        auto eval = [&](const std::vector<u64>& tbl) { PMle m; m.nv = nvars; m.ev = tbl; std::vector<u64> o(d); pm_evaluate(R, std::move(m), r.data(), nvars, o.data()); return o; };
        E.a.push_back(eval(tau_r)[0]);
        E.b.insert(E.b.end(), D.out.b.begin() + l * d, D.out.b.begin() + (l + 1) * d);
        { auto c0 = eval(I.f); E.c.insert(E.c.end(), c0.begin(), c0.end()); }
        for (auto& m : M) {
            E.a.push_back(r_ct(eval(sp_mul_vec(R, m, tau_r)).data()));
            { auto x = eval(sp_mul_vec(R, m, I.m_tau)); E.b.insert(E.b.end(), x.begin(), x.end()); }
            { auto x = eval(sp_mul_vec(R, m, I.f)); E.c.insert(E.c.end(), x.begin(), x.end()); }
        }
        D.evals.push_back(std::move(E)); D.fcoms.push_back(I.fcoms);
    }
    for (auto& E : D.evals) { for (u64 a : E.a) T.absorb_field(a); T.absorb_slice(E.c.data(), E.c.size() / d); }      // rgchk.rs:338-343
    return D;
}
inline bool range_check_verify(const RingParams& R, const Dcom& D, PlusTranscript& T) {      // rgchk.rs:190-246
    const int d = R.d; if (!set_check_verify(R, D.out, T)) return false;
    for (auto& E : D.evals) { for (u64 a : E.a) T.absorb_field(a); T.absorb_slice(E.c.data(), E.c.size() / d); }
    u64 psi[64], t[64]; r_psi(R, psi);
    const int nE = 1 + D.out.n_M, k = D.dp.k; const size_t ncols = D.out.ncols;
    for (size_t l = 0; l < D.evals.size(); ++l) { const DcomEvals& E = D.evals[l];
        for (size_t i = 0; i < E.a.size(); ++i) { r_mul(R, t, psi, E.b.data() + i * d); if (r_ct(t) != E.a[i]) return false; }
        for (int ni = 0; ni < nE; ++ni) {
            std::vector<u64> ucomb(ncols * d, 0);
            for (int i = 0; i < k; ++i) { const u64 dpw = R.F.pow((u64)(d / 2), i); const u64* ui = D.out.e.data() + (((size_t)ni * D.out.n_mat + (size_t)k * l + i) * ncols) * d;
                for (size_t j = 0; j < ncols * d; ++j) ucomb[j] = R.F.add(ucomb[j], R.F.mul(ui[j], dpw)); }
            for (size_t j = 0; j < ncols; ++j) { r_mul(R, t, psi, ucomb.data() + j * d); const u64 got = r_ct(t);
                const u64 want = ni == 0 ? E.v[j] : E.c[(size_t)ni * d + j]; if (got != want) return false; }
        }
    }
    return true;
}


// ---------------------------------------------------------------- flat u64 images (the layout the product C ABI returns, include/lf_b200.h)
// SetOut: [nvars, n_mat, ncols, n_vec, n_M] r[nvars] msgs[nvars x 4 x d] e[(1+n_M) x n_mat x ncols x d] b[n_vec x d]
inline std::vector<u64> set_out_words(const RingParams& R, const SetOut& o) {
    std::vector<u64> w = {(u64)o.nvars, (u64)o.n_mat, (u64)o.ncols, (u64)o.n_vec, (u64)o.n_M};
    w.insert(w.end(), o.r.begin(), o.r.end()); w.insert(w.end(), o.pf.msgs.begin(), o.pf.msgs.end()); w.insert(w.end(), o.e.begin(), o.e.end()); w.insert(w.end(), o.b.begin(), o.b.end());
    return w;
}
inline size_t set_out_parse(const RingParams& R, const u64* w, size_t len, SetOut& o) {
    const size_t d = R.d; if (len < 5) throw std::runtime_error("set-check image too short");
    o.nvars = (int)w[0]; o.n_mat = (int)w[1]; o.ncols = (int)w[2]; o.n_vec = (int)w[3]; o.n_M = (int)w[4];
    if (w[0] > 40 || w[1] > 4096 || w[2] > 4096 || w[3] > 4096 || w[4] > 64) throw std::runtime_error("set-check image: implausible header");
    const size_t nr = o.nvars, nm = (size_t)o.nvars * 4 * d, ne = (size_t)(1 + o.n_M) * o.n_mat * o.ncols * d, nb = (size_t)o.n_vec * d;
    if (len < 5 + nr + nm + ne + nb) throw std::runtime_error("set-check image truncated");
    const u64* p = w + 5; o.r.assign(p, p + nr); p += nr; o.pf.nvars = o.nvars; o.pf.degree = 3; o.pf.msgs.assign(p, p + nm); p += nm; o.e.assign(p, p + ne); p += ne; o.b.assign(p, p + nb); p += nb;
    return (size_t)(p - w);
}
// Dcom: [L, k, l, kappa, b] SetOut, then per instance v[d] a[1+n_M] b[(1+n_M) x d] c[(1+n_M) x d] cm_f[kappa x d] C_Mf[kappa x d] cm_mtau[kappa x d]
inline std::vector<u64> dcom_words(const RingParams& R, const Dcom& D, size_t kappa) {
    std::vector<u64> w = {(u64)D.evals.size(), (u64)D.dp.k, (u64)D.dp.l, (u64)kappa, (u64)D.dp.b};
    auto so = set_out_words(R, D.out); w.insert(w.end(), so.begin(), so.end());
    for (size_t l = 0; l < D.evals.size(); ++l) { const DcomEvals& E = D.evals[l]; const FComs& C = D.fcoms[l];
        for (auto* v : {&E.v, &E.a, &E.b, &E.c, &C.cm_f, &C.C_Mf, &C.cm_mtau}) w.insert(w.end(), v->begin(), v->end()); }
    return w;
}
inline void dcom_parse(const RingParams& R, const u64* w, size_t len, Dcom& D, size_t& kappa) {
    const size_t d = R.d; if (len < 5) throw std::runtime_error("range-check image too short");
    const size_t L = w[0]; D.dp.k = (int)w[1]; D.dp.l = (int)w[2]; kappa = w[3]; D.dp.b = w[4];
    if (L > 64 || w[1] > 64 || kappa > 4096) throw std::runtime_error("range-check image: implausible header");
    size_t off = 5 + set_out_parse(R, w + 5, len - 5, D.out); const size_t nE = 1 + D.out.n_M;
    const size_t per = d + nE + 2 * nE * d + 3 * kappa * d; if (len < off + L * per) throw std::runtime_error("range-check image truncated");
    for (size_t l = 0; l < L; ++l) { DcomEvals E; FComs C; const u64* p = w + off + l * per;
        E.v.assign(p, p + d); p += d; E.a.assign(p, p + nE); p += nE; E.b.assign(p, p + nE * d); p += nE * d; E.c.assign(p, p + nE * d); p += nE * d;
        C.cm_f.assign(p, p + kappa * d); p += kappa * d; C.C_Mf.assign(p, p + kappa * d); p += kappa * d; C.cm_mtau.assign(p, p + kappa * d);
        D.evals.push_back(std::move(E)); D.fcoms.push_back(std::move(C)); }
}

// ---------------------------------------------------------------- commitment transformation (cm.rs)
// utils.rs:88-103 short_challenge(lambda = 128): u = 2^(lambda / d) = 256, coefficient = (byte mod u) - u/2 from squeeze_bytes(d)
inline void short_challenge128(const RingParams& R, PlusTranscript& T, u64* out) {
    const int d = R.d; const unsigned u = 1u << (128 / d); std::vector<uint8_t> bs(d); T.sp.squeeze_bytes(bs.data(), d);
    for (int i = 0; i < d; ++i) out[i] = R.F.from_i128((i128)(bs[i] % u) - (i128)(u / 2));
}
inline int ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) ++l; return l; }      // ark_std::log2
// t(z) = tensor(c) (x) s' (x) (1, d', .., d'^(l-1)) (x) (1, X, .., X^(d-1))   (cm.rs:590-601), ring elements
inline std::vector<u64> calculate_t_z(const RingParams& R, const std::vector<u64>& c, const std::vector<u64>& s_prime_flat /* kd x d */, int l) {
    const int d = R.d; const size_t kd = s_prime_flat.size() / d; std::vector<u64> tc = tensor(R, c.data(), (int)c.size());
    std::vector<u64> out; out.reserve(tc.size() * kd * l * d * d); u64 e[64], xb[64], t[64];
    for (u64 tci : tc) for (size_t j = 0; j < kd; ++j) for (int a = 0; a < l; ++a) for (int b = 0; b < d; ++b) {
        r_scale(R, e, s_prime_flat.data() + j * d, tci);                              // tensor(c)_i * s'_j
        r_scale(R, e, e, R.F.pow((u64)(d / 2), a));                                   // * d'^a
        memset(xb, 0, 8 * d); xb[b] = 1; r_mul(R, t, e, xb);                          // * X^b
        out.insert(out.end(), t, t + d);
    }
    return out;
}
struct CmProof { Dcom dcom; std::vector<u64> comh /* L x kappa x d */; PProof pf[2]; std::vector<u64> evals[2] /* L x (1+n_M) x 4 x d */; size_t kappa = 0; };
struct ComX { std::vector<u64> cm_g /* L x kappa x d */, ro /* nvars x 2 */, vo /* L x (1+n_M) x 2 x d */; };
struct Com { std::vector<u64> g /* L x n x d */; ComX x; };

inline ComX cm_x(const RingParams& R, const CmProof& P, const u64* s /* 3 x d */, const std::vector<u64>& ro_a, const std::vector<u64>& ro_b) {      // cm.rs:537-575
    const int d = R.d; const size_t L = P.dcom.fcoms.size(), kappa = P.kappa, nE = 1 + P.dcom.out.n_M; ComX X; u64 t[64], acc[64];
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < kappa; ++i) { const FComs& C = P.dcom.fcoms[l];
        r_mul(R, acc, s, C.C_Mf.data() + i * d); r_mul(R, t, s + d, C.cm_mtau.data() + i * d); el_add(R, acc, acc, t);
        r_mul(R, t, s + 2 * d, C.cm_f.data() + i * d); el_add(R, acc, acc, t); el_add(R, acc, acc, P.comh.data() + (l * kappa + i) * d);
        X.cm_g.insert(X.cm_g.end(), acc, acc + d); }
    for (size_t i = 0; i < ro_a.size(); ++i) { X.ro.push_back(ro_a[i]); X.ro.push_back(ro_b[i]); }
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < nE; ++i) for (int z = 0; z < 2; ++z) { const u64* e = P.evals[z].data() + ((l * nE + i) * 4) * d;
        r_mul(R, acc, s, e); r_mul(R, t, s + d, e + d); el_add(R, acc, acc, t); r_mul(R, t, s + 2 * d, e + 2 * d); el_add(R, acc, acc, t); el_add(R, acc, acc, e + 3 * d);
        X.vo.insert(X.vo.end(), acc, acc + d); }
    return X;
}
// Cm::sumchecker (cm.rs:205-342)
inline void cm_sumchecker(const RingParams& R, int nvars, const std::vector<RgInstance>& inst, const Dcom& dcom, const std::vector<std::vector<u64>>& h,
                          const std::vector<u64>& t0, const std::vector<u64>& t1, const std::vector<SparseR>& M, PlusTranscript& T, PProof& pf, std::vector<u64>& evals, std::vector<u64>& ro) {
    const int d = R.d; const Fp& F = R.F; const size_t L = inst.size(), Mlen = M.size();
    const u64 rc = T.get_challenge();
    std::vector<PMle> mles;
    auto push = [&](std::vector<u64> ev) { PMle m; m.nv = nvars; m.ev = std::move(ev); mles.push_back(std::move(m)); };
    { std::vector<u64> eq = eq_table_base(R, dcom.out.r.data(), nvars), e(eq.size() * d, 0); for (size_t i = 0; i < eq.size(); ++i) e[i * d] = eq[i]; push(std::move(e)); }
    for (size_t i = 0; i < L; ++i) { const RgInstance& I = inst[i];
        std::vector<u64> rtau(I.n * d, 0); for (size_t x = 0; x < I.n; ++x) rtau[x * d] = I.tau[x];
        push(rtau); push(I.m_tau); push(I.f); push(h[i]);
        for (auto& m : M) { push(sp_mul_vec(R, m, rtau)); push(sp_mul_vec(R, m, I.m_tau)); push(sp_mul_vec(R, m, I.f)); push(sp_mul_vec(R, m, h[i])); } }
    push(t0); push(t1);
    std::vector<u64> rcps; u64 rcp = 1;
    for (size_t i = 0; i < L * (4 + 4 * Mlen); ++i) { rcps.push_back(rcp); rcp = F.mul(rcp, rc); }
    rcps.push_back(rcp); rcp = F.mul(rcp, rc); rcps.push_back(rcp);
    const size_t nvals = mles.size();
    CombFn comb = [&](const u64* vals, u64* out) {      // cm.rs:285-307
        u64 tot[64], in[64], t[64]; memset(tot, 0, 8 * d);
        for (size_t l = 0; l < L; ++l) { const size_t l_idx = 1 + l * (4 + 4 * Mlen); memset(in, 0, 8 * d);
            for (size_t q = 0; q < 4 + 4 * Mlen; ++q) { r_scale(R, t, vals + (l_idx + q) * d, rcps[l_idx + q - 1]); el_add(R, in, in, t); }
            r_mul(R, t, vals, in); el_add(R, tot, tot, t);
            r_mul(R, t, vals + l_idx * d, vals + (nvals - 2) * d); r_scale(R, t, t, rcps[nvals - 3]); el_add(R, tot, tot, t);
            r_mul(R, t, vals + l_idx * d, vals + (nvals - 1) * d); r_scale(R, t, t, rcps[nvals - 2]); el_add(R, tot, tot, t); }
        memcpy(out, tot, 8 * d);
    };
    pf = p_prove(R, T, mles, nvars, 2, comb, ro);
    evals.assign(L * (1 + Mlen) * 4 * d, 0);
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i <= Mlen; ++i) for (int q = 0; q < 4; ++q)
        pm_evaluate(R, mles[1 + l * (4 + 4 * Mlen) + 4 * i + q], ro.data(), nvars, evals.data() + ((l * (1 + Mlen) + i) * 4 + q) * d);
    T.absorb_slice(evals.data(), evals.size() / d);      // absorb_evaluations, cm.rs:581-588
}
// Cm::prove (cm.rs:57-203)
inline void cm_prove(const RingParams& R, int nvars, const std::vector<RgInstance>& inst, const DecompParameters& dp, const std::vector<SparseR>& M, PlusTranscript& T, Com& com, CmProof& P) {
    const int d = R.d, k = dp.k; const size_t L = inst.size(), n = inst[0].tau.size(), kappa = inst[0].kappa, N = (size_t)1 << nvars;
    P.kappa = kappa; P.dcom = range_check(R, nvars, inst, dp, M, T);
    std::vector<u64> s(3 * d), sp((size_t)k * d * d);
    for (int i = 0; i < 3; ++i) short_challenge128(R, T, s.data() + i * d);
    for (int i = 0; i < k * d; ++i) short_challenge128(R, T, sp.data() + (size_t)i * d);
    std::vector<std::vector<u64>> h(L, std::vector<u64>(N * d, 0)); u64 t[64];
    for (size_t l = 0; l < L; ++l)
        #pragma omp parallel for schedule(static) private(t)
        for (long x = 0; x < (long)n; ++x) for (int kk = 0; kk < k; ++kk) for (int c = 0; c < d; ++c) {
            r_mul(R, t, inst[l].M_f[kk].data() + ((size_t)x * d + c) * d, sp.data() + ((size_t)kk * d + c) * d); el_add(R, h[l].data() + (size_t)x * d, h[l].data() + (size_t)x * d, t); }
    P.comh.assign(L * kappa * d, 0);
    for (size_t l = 0; l < L; ++l) for (int kk = 0; kk < k; ++kk) for (size_t r = 0; r < kappa; ++r) for (int c = 0; c < d; ++c) {
        r_mul(R, t, inst[l].comM_f[kk].data() + (r * d + c) * d, sp.data() + ((size_t)kk * d + c) * d); el_add(R, P.comh.data() + (l * kappa + r) * d, P.comh.data() + (l * kappa + r) * d, t); }
    T.absorb_slice(P.comh.data(), L * kappa);
    const int log_kappa = ceil_log2(kappa);
    std::vector<u64> c0 = T.get_challenges(log_kappa), c1 = T.get_challenges(log_kappa);
    std::vector<u64> t0 = calculate_t_z(R, c0, sp, dp.l), t1 = calculate_t_z(R, c1, sp, dp.l);
    if (t0.size() > n * d) throw std::runtime_error("t0 too large!");
    t0.resize(n * d, 0); t1.resize(n * d, 0);
    std::vector<u64> ro_a, ro_b;
    cm_sumchecker(R, nvars, inst, P.dcom, h, t0, t1, M, T, P.pf[0], P.evals[0], ro_a);
    cm_sumchecker(R, nvars, inst, P.dcom, h, t0, t1, M, T, P.pf[1], P.evals[1], ro_b);
    com.g.assign(L * n * d, 0);
    for (size_t l = 0; l < L; ++l)
        #pragma omp parallel for schedule(static)
        for (long x = 0; x < (long)n; ++x) { u64 a[64], u[64], cst[64]; const RgInstance& I = inst[l];
            r_const(R, cst, I.tau[x]); r_mul(R, a, s.data(), cst); r_mul(R, u, s.data() + d, I.m_tau.data() + (size_t)x * d); el_add(R, a, a, u);
            r_mul(R, u, s.data() + 2 * d, I.f.data() + (size_t)x * d); el_add(R, a, a, u); el_add(R, a, a, h[l].data() + (size_t)x * d);
            memcpy(com.g.data() + (l * n + x) * d, a, 8 * d); }
    com.x = cm_x(R, P, s.data(), ro_a, ro_b);
}
// CmProof::verify (cm.rs:349-535): returns false where the reference returns Err / panics on a failed check
inline bool cm_verify(const RingParams& R, const CmProof& P, const std::vector<SparseR>& M, PlusTranscript& T, ComX& out) {
    const int d = R.d, k = P.dcom.dp.k, nvars = P.dcom.out.nvars; const Fp& F = R.F; const size_t L = P.dcom.evals.size(), kappa = P.kappa, Mlen = M.size(), nE = 1 + Mlen;
    if ((size_t)P.dcom.out.n_M != Mlen) return false;
    if (!range_check_verify(R, P.dcom, T)) return false;
    std::vector<u64> s(3 * d), sp((size_t)k * d * d);
    for (int i = 0; i < 3; ++i) short_challenge128(R, T, s.data() + i * d);
    for (int i = 0; i < k * d; ++i) short_challenge128(R, T, sp.data() + (size_t)i * d);
    T.absorb_slice(P.comh.data(), L * kappa);
    const int log_kappa = ceil_log2(kappa);
    std::vector<u64> c0 = T.get_challenges(log_kappa), c1 = T.get_challenges(log_kappa);
    u64 t[64];
    std::vector<u64> u(L * nE * d, 0);      // u[l][ni] = sum over the k sets of instance l and their d columns of e * s'
    for (size_t l = 0; l < L; ++l) for (size_t ni = 0; ni < nE; ++ni) for (int q = 0; q < k * d; ++q) {
        const u64* e = P.dcom.out.e.data() + ((ni * P.dcom.out.n_mat + l * k) * (size_t)P.dcom.out.ncols + q) * d;
        r_mul(R, t, e, sp.data() + (size_t)q * d); el_add(R, u.data() + (l * nE + ni) * d, u.data() + (l * nE + ni) * d, t); }
    std::vector<u64> tc0 = tensor(R, c0.data(), log_kappa), tc1 = tensor(R, c1.data(), log_kappa), tcch0(L * d, 0), tcch1(L * d, 0);
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < std::min(tc0.size(), kappa); ++i) {
        r_scale(R, t, P.comh.data() + (l * kappa + i) * d, tc0[i]); el_add(R, tcch0.data() + l * d, tcch0.data() + l * d, t);
        r_scale(R, t, P.comh.data() + (l * kappa + i) * d, tc1[i]); el_add(R, tcch1.data() + l * d, tcch1.data() + l * d, t); }
    std::vector<u64> ro[2];
    for (int z = 0; z < 2; ++z) {
        const u64 rc = T.get_challenge(); const size_t z_idx = L * (4 + 4 * Mlen);
        std::vector<u64> claimed(d, 0);
        auto add_scaled = [&](std::vector<u64>& acc, const u64* e, u64 c) { r_scale(R, t, e, c); el_add(R, acc.data(), acc.data(), t); };
        for (size_t l = 0; l < L; ++l) { const DcomEvals& E = P.dcom.evals[l]; const size_t l_idx = l * (4 + 4 * Mlen); u64 cst[64];
            for (size_t i = 0; i < nE; ++i) { const size_t idx = l_idx + 4 * i;
                r_const(R, cst, E.a[i]); add_scaled(claimed, cst, F.pow(rc, idx)); add_scaled(claimed, E.b.data() + i * d, F.pow(rc, idx + 1));
                add_scaled(claimed, E.c.data() + i * d, F.pow(rc, idx + 2)); add_scaled(claimed, u.data() + (l * nE + i) * d, F.pow(rc, idx + 3)); }
            add_scaled(claimed, tcch0.data() + l * d, F.pow(rc, z_idx)); add_scaled(claimed, tcch1.data() + l * d, F.pow(rc, z_idx + 1)); }
        PSubClaim sc = p_verify(R, T, nvars, 2, claimed.data(), P.pf[z]);
        if (!sc.ok) return false;
        ro[z] = sc.point;
        u64 t0_ro[64], t1_ro[64];
        { PMle m; m.nv = nvars; m.ev = calculate_t_z(R, c0, sp, P.dcom.dp.l); if (m.ev.size() > ((size_t)d << nvars)) return false; pm_evaluate(R, std::move(m), ro[z].data(), nvars, t0_ro); }
        { PMle m; m.nv = nvars; m.ev = calculate_t_z(R, c1, sp, P.dcom.dp.l); pm_evaluate(R, std::move(m), ro[z].data(), nvars, t1_ro); }
        if (P.evals[z].size() != L * nE * 4 * d) return false;
        T.absorb_slice(P.evals[z].data(), P.evals[z].size() / d);
        const u64 eq = eq_eval_base(R, P.dcom.out.r.data(), ro[z].data(), nvars);
        std::vector<u64> ev(d, 0);
        for (size_t l = 0; l < L; ++l) { const size_t l_idx = l * (4 + 4 * Mlen); std::vector<u64> in(d, 0); const u64* el = P.evals[z].data() + (l * nE * 4) * d;
            for (size_t q = 0; q < 4 * nE; ++q) add_scaled(in, el + q * d, F.pow(rc, l_idx + q));
            add_scaled(ev, in.data(), eq);
            u64 pr[64]; r_mul(R, pr, t0_ro, el); add_scaled(ev, pr, F.pow(rc, z_idx)); r_mul(R, pr, t1_ro, el); add_scaled(ev, pr, F.pow(rc, z_idx + 1)); }
        if (memcmp(ev.data(), sc.expected.data(), 8 * d) != 0) return false;      // assert_eq!(expected_eval, eval)
    }
    out = cm_x(R, P, s.data(), ro[0], ro[1]);
    return true;
}
// CmProof image: the Dcom image, comh[L x kappa x d], two sumcheck proofs [nvars x 3 x d], two evaluation blocks [L x (1+n_M) x 4 x d]
inline std::vector<u64> cm_proof_words(const RingParams& R, const CmProof& P) {
    std::vector<u64> w = dcom_words(R, P.dcom, P.kappa);
    for (auto* v : {&P.comh, &P.pf[0].msgs, &P.pf[1].msgs, &P.evals[0], &P.evals[1]}) w.insert(w.end(), v->begin(), v->end());
    return w;
}
inline void cm_proof_parse(const RingParams& R, const u64* w, size_t len, CmProof& P) {
    const size_t d = R.d; dcom_parse(R, w, len, P.dcom, P.kappa);
    const size_t L = P.dcom.evals.size(), nE = 1 + P.dcom.out.n_M, nv = P.dcom.out.nvars, off = dcom_words(R, P.dcom, P.kappa).size();
    const size_t n1 = L * P.kappa * d, n2 = nv * 3 * d, n3 = L * nE * 4 * d;
    if (len < off + n1 + 2 * n2 + 2 * n3) throw std::runtime_error("cm proof image truncated");
    const u64* p = w + off; P.comh.assign(p, p + n1); p += n1;
    for (int z = 0; z < 2; ++z) { P.pf[z].nvars = (int)nv; P.pf[z].degree = 2; P.pf[z].msgs.assign(p, p + n2); p += n2; }
    for (int z = 0; z < 2; ++z) { P.evals[z].assign(p, p + n3); p += n3; }
}
// ComX image: cm_g[L x kappa x d] ro[nvars x 2] vo[L x (1+n_M) x 2 x d]
inline std::vector<u64> comx_words(const ComX& X) { std::vector<u64> w = X.cm_g; w.insert(w.end(), X.ro.begin(), X.ro.end()); w.insert(w.end(), X.vo.begin(), X.vo.end()); return w; }

// ---------------------------------------------------------------- r1cs.rs: linearization of a committed R1CS
struct R1csLinProof { int nvars = 0; PProof pf; std::vector<u64> r /* nvars */, v4 /* v, va, vb, vc: 4 x d */; };
// ComR1CS::linearize (r1cs.rs:72-134): sumcheck of eq(r, x) (ga(x) gb(x) - gc(x)) with g* = A|B|C f as ring-valued MLEs (real ring products)
inline R1csLinProof r1cs_linearize(const RingParams& R, const SparseR Mabc[3], const std::vector<u64>& f, PlusTranscript& T) {
    const int d = R.d; const size_t n = f.size() / d; const int nvars = ceil_log2(n);
    std::vector<u64> g[3]; for (int i = 0; i < 3; ++i) g[i] = sp_mul_vec(R, Mabc[i], f);
    std::vector<u64> r = T.get_challenges(nvars);
    std::vector<PMle> mles; auto push = [&](std::vector<u64> ev) { PMle m; m.nv = nvars; m.ev = std::move(ev); mles.push_back(std::move(m)); };
    { std::vector<u64> eq = eq_table_base(R, r.data(), nvars), e(eq.size() * d, 0); for (size_t i = 0; i < eq.size(); ++i) e[i * d] = eq[i]; push(std::move(e)); }
    for (int i = 0; i < 3; ++i) push(g[i]);
    CombFn comb = [&](const u64* vals, u64* out) { u64 t[64]; r_mul(R, t, vals + d, vals + 2 * d); el_sub(R, t, t, vals + 3 * d); r_mul(R, out, vals, t); };      // r1cs.rs:92
    R1csLinProof P; P.nvars = nvars;
    P.pf = p_prove(R, T, mles, nvars, 3, comb, P.r);
    P.v4.assign(4 * d, 0);
    { PMle m; m.nv = nvars; m.ev = f; pm_evaluate(R, std::move(m), P.r.data(), nvars, P.v4.data()); }
    for (int i = 0; i < 3; ++i) pm_evaluate(R, mles[1 + i], P.r.data(), nvars, P.v4.data() + (size_t)(1 + i) * d);
    T.absorb_slice(P.v4.data(), 4);
    return P;
}
// ComR1CSProof::verify (r1cs.rs:136-162)
inline bool r1cs_linearize_verify(const RingParams& R, const R1csLinProof& P, PlusTranscript& T) {
    const int d = R.d; std::vector<u64> r = T.get_challenges(P.nvars), zero(d, 0);
    PSubClaim sc = p_verify(R, T, P.nvars, 3, zero.data(), P.pf);
    if (!sc.ok) return false;
    T.absorb_slice(P.v4.data(), 4);
    const u64 e = eq_eval_base(R, r.data(), sc.point.data(), P.nvars);
    u64 t[64]; r_mul(R, t, P.v4.data() + d, P.v4.data() + 2 * d); el_sub(R, t, t, P.v4.data() + 3 * d); r_scale(R, t, t, e);
    return memcmp(t, sc.expected.data(), 8 * d) == 0;
}
// image: [nvars] r[nvars] messages[nvars x 4 x d] v|va|vb|vc [4 x d]
inline std::vector<u64> r1cs_lin_words(const R1csLinProof& P) { std::vector<u64> w = {(u64)P.nvars}; for (auto* v : {&P.r, &P.pf.msgs, &P.v4}) w.insert(w.end(), v->begin(), v->end()); return w; }
inline void r1cs_lin_parse(const RingParams& R, const u64* w, size_t len, R1csLinProof& P) {
    const size_t d = R.d; if (len < 1 || w[0] > 40) throw std::runtime_error("linearization image: bad header");
    P.nvars = (int)w[0]; const size_t nv = P.nvars; if (len < 1 + nv + nv * 4 * d + 4 * d) throw std::runtime_error("linearization image truncated");
    const u64* p = w + 1; P.r.assign(p, p + nv); p += nv; P.pf.nvars = P.nvars; P.pf.degree = 3; P.pf.msgs.assign(p, p + nv * 4 * d); p += nv * 4 * d; P.v4.assign(p, p + 4 * d);
}

// ---------------------------------------------------------------- mlin.rs / decomp.rs
struct LinB2 { std::vector<u64> g /* n x d */, cm_g /* kappa x d */, ro /* nvars x 2 */, vo /* (1+n_M) x 2 x d */; };
// Mlin::mlin (mlin.rs:41-106): from_f on every f, Cm::prove, then the sums over the instances
inline LinB2 mlin(const RingParams& R, const std::vector<std::vector<u64>>& fs, const std::vector<u64>& A, size_t kappa, const DecompParameters& dp, const std::vector<SparseR>& M, PlusTranscript& T, CmProof& P) {
    const int d = R.d; const size_t n = fs[0].size() / d, L = fs.size(), nE = 1 + M.size();
    std::vector<RgInstance> inst; for (auto& f : fs) inst.push_back(rg_from_f(R, f, A, kappa, dp));
    Com com; cm_prove(R, ceil_log2(n), inst, dp, M, T, com, P);
    LinB2 o; o.cm_g.assign(kappa * d, 0); o.vo.assign(nE * 2 * d, 0); o.g.assign(n * d, 0); o.ro = com.x.ro;
    for (size_t l = 0; l < L; ++l) {
        for (size_t i = 0; i < kappa * d; ++i) o.cm_g[i] = R.F.add(o.cm_g[i], com.x.cm_g[l * kappa * d + i]);
        for (size_t i = 0; i < nE * 2 * d; ++i) o.vo[i] = R.F.add(o.vo[i], com.x.vo[l * nE * 2 * d + i]);
        for (size_t i = 0; i < n * d; ++i) o.g[i] = R.F.add(o.g[i], com.g[l * n * d + i]); }
    return o;
}
struct DecompProof { std::vector<u64> C[2] /* kappa x d */, v[2] /* (1+n_M) x 2 x d */; };
// Decomp::decompose (decomp.rs:32-99): F = decompose_to_vec(f, B, 2); v_i over [F_i, M_j F_i] at the two points; C_i = A F_i.  r: nvars x 2 (constants of R)
inline DecompProof decompose(const RingParams& R, const std::vector<u64>& f, const std::vector<u64>& r, const std::vector<SparseR>& M, const std::vector<u64>& A, size_t kappa, u128 B, std::vector<u64> F[2]) {
    const int d = R.d; const size_t n = f.size() / d; const int nvars = ceil_log2(n); DecompProof P; i128 dg[2];
    F[0].assign(n * d, 0); F[1].assign(n * d, 0);
    for (size_t i = 0; i < n * d; ++i) { decompose_balanced(R.F, f[i], B, 2, dg); F[0][i] = R.F.from_i128(dg[0]); F[1][i] = R.F.from_i128(dg[1]); }
    std::vector<u64> ra(nvars), rb(nvars); for (int i = 0; i < nvars; ++i) { ra[i] = r[2 * i]; rb[i] = r[2 * i + 1]; }
    for (int z = 0; z < 2; ++z) {
        auto two = [&](const std::vector<u64>& tbl) { u64 o[64]; PMle m; m.nv = nvars; m.ev = tbl; pm_evaluate(R, m, ra.data(), nvars, o); P.v[z].insert(P.v[z].end(), o, o + d); pm_evaluate(R, std::move(m), rb.data(), nvars, o); P.v[z].insert(P.v[z].end(), o, o + d); };
        two(F[z]); for (auto& m : M) two(sp_mul_vec(R, m, F[z]));
        P.C[z] = mat_mul_vec(R, A, kappa, n, F[z]);
    }
    return P;
}
// DecompProof::verify (decomp.rs:102-126): recompose([C0, C1], B) = cm_f and recompose of the evaluation pairs = v
inline bool decompose_verify(const RingParams& R, const DecompProof& P, const std::vector<u64>& cm_f, const std::vector<u64>& v, u128 B) {
    const u64 Bm = (u64)(B % R.F.p);
    if (P.C[0].size() != cm_f.size() || P.v[0].size() != v.size()) return false;
    for (size_t i = 0; i < cm_f.size(); ++i) if (R.F.add(P.C[0][i], R.F.mul(P.C[1][i], Bm)) != cm_f[i]) return false;
    for (size_t i = 0; i < v.size(); ++i) if (R.F.add(P.v[0][i], R.F.mul(P.v[1][i], Bm)) != v[i]) return false;
    return true;
}

} }  // namespace lfo::plus
