// TEST INFRASTRUCTURE ONLY (see ring.hpp header).  C entry points over the CPU oracle so that tests/, smoke()
// and bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.  Never linked into the product.
//
// All ring elements cross this boundary as d canonical u64 limbs (NTT form: slot-major; coefficient form:
// power of X), vectors as contiguous arrays of such elements -- the same host format the product C-ABI uses.
#include "protocol.hpp"
#include "ntt.hpp"
#include "lfplus.hpp"
#include <map>
#include <memory>
#include <chrono>
#include <omp.h>

using namespace lfo;

namespace {
const RingParams& ring(int id) {
    static std::map<int, std::unique_ptr<RingParams>> cache;
    #pragma omp critical(lfo_ring_cache)
    { if (!cache.count(id)) cache[id].reset(new RingParams(make_ring(id))); }
    return *cache[id];
}
thread_local std::string g_err;
template <class F> int guard(F&& f) {
    try { f(); return 0; }
    catch (const LfError& e) { g_err = e.what(); return e.code; }
    catch (const std::exception& e) { g_err = e.what(); return -100; }
}
Vec vec_of(const u64* p, size_t n_elems, int d) { return p ? Vec(p, p + n_elems * d) : Vec(); }
}  // namespace

extern "C" {

const char* lfo_last_error() { return g_err.c_str(); }
int lfo_num_threads() { return omp_get_max_threads(); }
void lfo_set_num_threads(int n) { omp_set_num_threads(n); }

// out[0..7] = p, d, S, tau, g, nu, trinomial, cs_bytes
int lfo_ring_info(int id, u64* out) {
    return guard([&] { const RingParams& R = ring(id); out[0] = R.F.p; out[1] = R.d; out[2] = R.S; out[3] = R.tau; out[4] = R.g; out[5] = R.nu; out[6] = R.trinomial; out[7] = R.cs_bytes; });
}
int lfo_crt(int id, const u64* in, u64* out, size_t n) {
    return guard([&] { const RingParams& R = ring(id); Vec o = elementwise_crt(R, vec_of(in, n, R.d)); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_icrt(int id, const u64* in, u64* out, size_t n) {
    return guard([&] { const RingParams& R = ring(id); Vec o = elementwise_icrt(R, vec_of(in, n, R.d)); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_coeff_mul(int id, const u64* a, const u64* b, u64* out) { return guard([&] { coeff_mul(ring(id), out, a, b); }); }
int lfo_ntt_mul(int id, const u64* a, const u64* b, u64* out, size_t n) {
    return guard([&] { const RingParams& R = ring(id); for (size_t i = 0; i < n; ++i) ntt_mul(R, out + i * R.d, a + i * R.d, b + i * R.d); });
}
int lfo_gadget_decompose(int id, const u64* in, size_t n, u64 B_lo, u64 B_hi, int L, u64* out) {
    return guard([&] { const RingParams& R = ring(id); Vec o = gadget_decompose(R, vec_of(in, n, R.d), ((u128)B_hi << 64) | B_lo, L); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_gadget_recompose(int id, const u64* in, size_t n_in, u64 B_lo, u64 B_hi, int L, u64* out) {
    return guard([&] { const RingParams& R = ring(id); Vec o = gadget_recompose(R, vec_of(in, n_in, R.d), ((u128)B_hi << 64) | B_lo, L); memcpy(out, o.data(), 8 * o.size()); });
}
// out: K x n x d (piece-major)
int lfo_decompose_to_vec(int id, const u64* in, size_t n, u64 b, int K, u64* out) {
    return guard([&] { const RingParams& R = ring(id); auto ps = decompose_to_k_vecs(R, vec_of(in, n, R.d), b, K);
                       for (int k = 0; k < K; ++k) memcpy(out + (size_t)k * n * R.d, ps[k].data(), 8 * n * R.d); });
}
// out: tau x n x d (zero padded), lens[tau] = effective (truncated) lengths
int lfo_fhat(int id, const u64* f_coeff, size_t n, u64* out, u64* lens) {
    return guard([&] { const RingParams& R = ring(id); auto fh = get_fhat(R, vec_of(f_coeff, n, R.d));
                       memset(out, 0, 8 * (size_t)R.tau * n * R.d);
                       for (int j = 0; j < R.tau; ++j) { memcpy(out + (size_t)j * n * R.d, fh[j].ev.data(), 8 * fh[j].ev.size()); lens[j] = fh[j].len(); } });
}
int lfo_commit(int id, const u64* A, size_t kappa, size_t n, const u64* f, size_t nf, u64* out) {
    return guard([&] { const RingParams& R = ring(id); Ajtai S; S.kappa = kappa; S.n = n; S.A = vec_of(A, kappa * n, R.d);
                       Vec cm = ajtai_commit(R, S, vec_of(f, nf, R.d)); memcpy(out, cm.data(), 8 * cm.size()); });
}
int lfo_spmv(int id, size_t nrows, size_t ncols, const u64* row_ptr, const u64* col, const u64* val, const u64* z, size_t nz, u64* out) {
    return guard([&] { const RingParams& R = ring(id); SparseMatrix M; M.nrows = nrows; M.ncols = ncols; M.row_ptr.assign(row_ptr, row_ptr + nrows + 1);
                       M.col.assign(col, col + row_ptr[nrows]); M.val = vec_of(val, row_ptr[nrows], R.d);
                       Vec o = mat_vec_mul(R, M, vec_of(z, nz, R.d)); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_eq_table(int id, const u64* r, int s, u64* out) {
    return guard([&] { const RingParams& R = ring(id); Mle m = build_eq_x_r(R, r, s); memcpy(out, m.ev.data(), 8 * m.ev.size()); });
}
int lfo_eq_eval(int id, const u64* x, const u64* y, int n, u64* out) { return guard([&] { eq_eval(ring(id), x, y, n, out); }); }
// mles: count x len x d ; all with num_vars nv; point: np ring elements
int lfo_evaluate_mles(int id, const u64* mles, int count, size_t len, int nv, const u64* point, int np, u64* out) {
    return guard([&] { const RingParams& R = ring(id); std::vector<Mle> ms;
                       for (int k = 0; k < count; ++k) ms.push_back(mle_from(R, nv, mles + (size_t)k * len * R.d, len));
                       Vec o = evaluate_mles(R, ms, vec_of(point, np, R.d)); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_rot_lin_combination(int id, const u64* rho_coeff, const u64* theta, int count, u64* out) {
    return guard([&] { const RingParams& R = ring(id); std::vector<Vec> rho, th;
                       for (int i = 0; i < count; ++i) { rho.push_back(vec_of(rho_coeff + (size_t)i * R.d, 1, R.d)); th.push_back(vec_of(theta + (size_t)i * R.tau * R.d, R.tau, R.d)); }
                       Vec o = rot_lin_combination(R, rho, th); memcpy(out, o.data(), 8 * o.size()); });
}
int lfo_short_challenge_from_bytes(int id, const uint8_t* bs, u64* coeffs) { return guard([&] { short_challenge_from_bytes(ring(id), bs, coeffs); }); }

// ------------------------------------------------------------------ transcript handles
void* lfo_tr_new(int id) { try { return new Transcript(ring(id)); } catch (...) { return nullptr; } }
void* lfo_tr_clone(void* h) { return new Transcript(*(Transcript*)h); }
void lfo_tr_free(void* h) { delete (Transcript*)h; }
void lfo_tr_absorb(void* h, const u64* els, size_t n) { ((Transcript*)h)->absorb_slice(els, n); }
void lfo_tr_absorb_base(void* h, const u64* limbs, size_t n) { ((Transcript*)h)->sp.absorb(limbs, n); }   // raw sponge absorb (KATs)
void lfo_tr_absorb_tag(void* h, const char* tag) { ((Transcript*)h)->absorb_tag(tag); }
void lfo_tr_absorb_u64(void* h, u64 x) { ((Transcript*)h)->absorb_u64(x); }
void lfo_tr_squeeze_base(void* h, u64* out, size_t n) { ((Transcript*)h)->sp.squeeze(out, n); }
void lfo_tr_get_challenge(void* h, u64* sf) { ((Transcript*)h)->get_challenge(sf); }
void lfo_tr_get_short_challenge(void* h, u64* coeffs) { ((Transcript*)h)->get_short_challenge(coeffs); }
void lfo_tr_state(void* h, u64* out24) { memcpy(out24, ((Transcript*)h)->sp.st, sizeof(((Transcript*)h)->sp.st)); }

// ------------------------------------------------------------------ stand-alone sumcheck
// comb_kind: 0 = PRODUCTS (sum_i coef_i * prod_{j in idx_i} v_j), 1 = LIN, 2 = FOLD.
// For 0/1: nterms, coef (nterms x d), idx_flat + idx_len[nterms].  For 2: n_mu, b, mu (n_mu x d).
// mles: M x len x d, effective lengths lens[M] (<= len).  Outputs: msgs nv x (deg+1) x d; point nv x tau.
int lfo_sumcheck_prove(int id, void* tr, const u64* mles, int M, size_t len, const u64* lens, int nv, int degree, int comb_kind,
                       int nterms, const u64* coef, const int* idx_flat, const int* idx_len, int n_mu, int b, const u64* mu,
                       u64* msgs, u64* point, u64* final_vals /* M x d after applying the last challenge, may be null */) {
    return guard([&] {
        const RingParams& R = ring(id); Comb C; C.kind = comb_kind;
        if (comb_kind != COMB_FOLD) { int o = 0; for (int i = 0; i < nterms; ++i) { C.coef.push_back(vec_of(coef + (size_t)i * R.d, 1, R.d)); C.idx.emplace_back(idx_flat + o, idx_flat + o + idx_len[i]); o += idx_len[i]; } }
        else { C.n_mu = n_mu; C.tau = R.tau; C.b = b; C.mu = vec_of(mu, n_mu, R.d); }
        std::vector<Mle> ms; for (int k = 0; k < M; ++k) ms.push_back(mle_from(R, nv, mles + (size_t)k * len * R.d, lens ? lens[k] : len));
        std::vector<std::vector<u64>> pt; ProverState st;
        SumcheckProof pf = prove_as_subprotocol(R, *(Transcript*)tr, std::move(ms), nv, degree, C, pt, &st);
        memcpy(msgs, pf.msgs.data(), 8 * pf.msgs.size());
        for (int i = 0; i < nv; ++i) memcpy(point + (size_t)i * R.tau, pt[i].data(), 8 * R.tau);
        if (final_vals) { Vec r(R.d); ntt_from_sf(R, r.data(), pt[nv - 1].data());
                          for (int k = 0; k < M; ++k) { Mle m = st.mles[k]; mle_fix_low(R, m, r.data()); mle_get(R, m, 0, final_vals + (size_t)k * R.d); } }
    });
}
// claimed_sum: d limbs.  returns 0 and fills expected (d limbs) + point on accept, ERR_SUMCHECK_FAILED otherwise
int lfo_sumcheck_verify(int id, void* tr, int nv, int degree, const u64* claimed_sum, const u64* msgs, u64* expected, u64* point) {
    return guard([&] {
        const RingParams& R = ring(id); SumcheckProof pf; pf.nvars = nv; pf.degree = degree; pf.msgs.assign(msgs, msgs + (size_t)nv * (degree + 1) * R.d);
        SubClaim sc = verify_as_subprotocol(R, *(Transcript*)tr, nv, degree, claimed_sum, pf);
        if (!sc.ok) throw LfError(ERR_SUMCHECK_FAILED, "SumCheckFailed");
        memcpy(expected, sc.expected.data(), 8 * R.d);
        for (int i = 0; i < nv; ++i) memcpy(point + (size_t)i * R.tau, sc.point[i].data(), 8 * R.tau);
    });
}

// ------------------------------------------------------------------ the full prover step on flat inputs
struct lfo_csr { u64 nrows, ncols; const u64* row_ptr; const u64* col; const u64* val; };
struct lfo_problem {
    int ring; int L, K; u64 B_lo, B_hi, b;
    u64 kappa, n; const u64* A;                              // Ajtai matrix kappa x n x d (NTT form)
    u64 m, n_ccs, l, t, q, d, s;                             // CCS shape (arith.rs:51-74)
    const lfo_csr* M; const int* S_flat; const int* S_len; const u64* c;   // t matrices; q multisets; q ring elements
    const u64 *acc_r, *acc_v, *acc_cm, *acc_u, *acc_x_w, *acc_h;           // running LCCCS
    const u64* w_acc_f;                                      // accumulator witness, NTT form, n x d  (Witness::from_f)
    const u64 *cm_i_cm, *cm_i_x_ccs;                         // incoming CCCS
    const u64* w_i_f;                                        // incoming witness f (NTT form, n x d)
};
static void load_problem(const lfo_problem& P, const RingParams& R, DecompParams& dp, CCS& ccs, Ajtai& sch, LCCCS& acc, CCCS& cmi) {
    dp.B = ((u128)P.B_hi << 64) | P.B_lo; dp.L = P.L; dp.b = P.b; dp.K = P.K;
    ccs.m = P.m; ccs.n = P.n_ccs; ccs.l = P.l; ccs.t = P.t; ccs.q = P.q; ccs.d = P.d; ccs.s = P.s;
    int o = 0;
    for (u64 i = 0; i < P.q; ++i) { ccs.S.emplace_back(P.S_flat + o, P.S_flat + o + P.S_len[i]); o += P.S_len[i]; }
    ccs.c = vec_of(P.c, P.q, R.d);
    for (u64 j = 0; j < P.t; ++j) { SparseMatrix M; M.nrows = P.M[j].nrows; M.ncols = P.M[j].ncols; M.row_ptr.assign(P.M[j].row_ptr, P.M[j].row_ptr + M.nrows + 1);
                                    u64 nnz = M.row_ptr[M.nrows]; M.col.assign(P.M[j].col, P.M[j].col + nnz); M.val = vec_of(P.M[j].val, nnz, R.d); ccs.M.push_back(std::move(M)); }
    if (P.A) { sch.kappa = P.kappa; sch.n = P.n; sch.A = vec_of(P.A, P.kappa * P.n, R.d); }
    if (P.acc_r) { acc.r = vec_of(P.acc_r, P.s, R.d); acc.v = vec_of(P.acc_v, R.tau, R.d); acc.cm = vec_of(P.acc_cm, P.kappa, R.d); acc.u = vec_of(P.acc_u, P.t, R.d);
                   acc.x_w = vec_of(P.acc_x_w, P.l, R.d); acc.h = vec_of(P.acc_h, 1, R.d); }
    cmi.cm = vec_of(P.cm_i_cm, P.kappa, R.d); cmi.x_ccs = vec_of(P.cm_i_x_ccs, P.l, R.d);
}
static void put(u64*& p, const Vec& v) { memcpy(p, v.data(), 8 * v.size()); p += v.size(); }
static void get(const u64*& p, Vec& v, size_t n) { v.assign(p, p + n); p += n; }
static size_t lcccs_words(const lfo_problem& P, const RingParams& R) { return (P.s + R.tau + P.kappa + P.t + P.l + 1) * R.d; }
static void put_lcccs(u64*& p, const LCCCS& L) { put(p, L.r); put(p, L.v); put(p, L.cm); put(p, L.u); put(p, L.x_w); put(p, L.h); }

// number of u64 words of the serialized proof for this problem shape
u64 lfo_proof_words(const lfo_problem* P) {
    const RingParams& R = ring(P->ring); u64 d = R.d, tau = R.tau;
    u64 lin = P->s * (P->d + 2) * d + tau * d + P->t * d;
    u64 dec = (u64)P->K * ((P->l + 1) + P->kappa + P->t + tau) * d;
    u64 fold = P->s * (2 * P->b + 1) * d + 2 * (u64)P->K * (tau + P->t) * d;
    return lin + 2 * dec + fold;
}
u64 lfo_lcccs_words(const lfo_problem* P) { return lcccs_words(*P, ring(P->ring)); }
static void put_proof(u64*& p, const LFProof& pf) {
    put(p, pf.lin.sumcheck.msgs); put(p, pf.lin.v); put(p, pf.lin.u);
    for (const DecompositionProof* dp : {&pf.dl, &pf.dr}) for (size_t k = 0; k < dp->x_s.size(); ++k) { put(p, dp->x_s[k]); put(p, dp->y_s[k]); put(p, dp->u_s[k]); put(p, dp->v_s[k]); }
    put(p, pf.fold.sumcheck.msgs); for (auto& v : pf.fold.theta_s) put(p, v); for (auto& v : pf.fold.eta_s) put(p, v);
}
static void get_proof(const u64*& p, LFProof& pf, const lfo_problem& P, const RingParams& R) {
    size_t d = R.d, tau = R.tau;
    pf.lin.sumcheck.nvars = (int)P.s; pf.lin.sumcheck.degree = (int)P.d + 1; get(p, pf.lin.sumcheck.msgs, P.s * (P.d + 2) * d); get(p, pf.lin.v, tau * d); get(p, pf.lin.u, P.t * d);
    for (DecompositionProof* dp : {&pf.dl, &pf.dr}) { dp->x_s.resize(P.K); dp->y_s.resize(P.K); dp->u_s.resize(P.K); dp->v_s.resize(P.K);
        for (int k = 0; k < P.K; ++k) { get(p, dp->x_s[k], (P.l + 1) * d); get(p, dp->y_s[k], P.kappa * d); get(p, dp->u_s[k], P.t * d); get(p, dp->v_s[k], tau * d); } }
    pf.fold.sumcheck.nvars = (int)P.s; pf.fold.sumcheck.degree = 2 * (int)P.b; get(p, pf.fold.sumcheck.msgs, P.s * (2 * P.b + 1) * d);
    pf.fold.theta_s.resize(2 * P.K); pf.fold.eta_s.resize(2 * P.K);
    for (auto& v : pf.fold.theta_s) get(p, v, tau * d); for (auto& v : pf.fold.eta_s) get(p, v, P.t * d);
}

// Witness::commit (arith.rs:357-362) for a witness given by f
// LFLinearizationProver::prove on (cm_i, w_i): fills out_lcccs (lcccs_words) and lin proof part (msgs, v, u).
int lfo_linearize(const lfo_problem* P, void* tr, u64* out_lcccs, u64* out_proof /* s*(d+2)*dd + tau*dd + t*dd */) {
    return guard([&] {
        const RingParams& R = ring(P->ring); DecompParams dp; CCS ccs; Ajtai sch; LCCCS acc; CCCS cmi; load_problem(*P, R, dp, ccs, sch, acc, cmi);
        Witness wi = witness_from_f(R, dp, vec_of(P->w_i_f, P->n, R.d));
        LCCCS out; LinearizationProof pf; linearization_prove(R, cmi, wi, *(Transcript*)tr, ccs, out, pf);
        u64* p = out_lcccs; put_lcccs(p, out);
        if (out_proof) { p = out_proof; put(p, pf.sumcheck.msgs); put(p, pf.v); put(p, pf.u); }
    });
}
// LFLinearizationVerifier::verify (nifs/linearization.rs:192-285) on the linearization part of a proof (msgs, v, u): returns 0 when accepted
int lfo_linearization_verify(const lfo_problem* P, void* tr, const u64* lin_proof, u64* out_lcccs) {
    return guard([&] {
        const RingParams& R = ring(P->ring); DecompParams dp; CCS ccs; Ajtai sch; LCCCS acc; CCCS cmi; load_problem(*P, R, dp, ccs, sch, acc, cmi);
        LinearizationProof pf; const u64* p = lin_proof; const size_t d = R.d;
        pf.sumcheck.nvars = (int)P->s; pf.sumcheck.degree = (int)P->d + 1; get(p, pf.sumcheck.msgs, P->s * (P->d + 2) * d); get(p, pf.v, R.tau * d); get(p, pf.u, P->t * d);
        LCCCS out; linearization_verify(R, cmi, pf, *(Transcript*)tr, ccs, out);
        if (out_lcccs) { u64* q = out_lcccs; put_lcccs(q, out); }
    });
}
// NIFSProver::prove (nifs.rs:48-103).  out_proof: lfo_proof_words; out_lcccs: lfo_lcccs_words; out_f: n x d (folded witness, NTT form)
// timing_ms (optional, 4 doubles): linearization, decomposition x2, folding, total
int lfo_nifs_prove(const lfo_problem* P, void* tr, u64* out_proof, u64* out_lcccs, u64* out_f, double* timing_ms) {
    return guard([&] {
        const RingParams& R = ring(P->ring); DecompParams dp; CCS ccs; Ajtai sch; LCCCS acc; CCCS cmi; load_problem(*P, R, dp, ccs, sch, acc, cmi);
        Witness wa = witness_from_f(R, dp, vec_of(P->w_acc_f, P->n, R.d)), wi = witness_from_f(R, dp, vec_of(P->w_i_f, P->n, R.d));
        LCCCS out; Witness wo; LFProof pf;
        auto t0 = std::chrono::steady_clock::now();
        nifs_prove(R, dp, acc, wa, cmi, wi, *(Transcript*)tr, ccs, sch, out, wo, pf);
        auto t1 = std::chrono::steady_clock::now();
        if (timing_ms) timing_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        u64* p = out_proof; put_proof(p, pf);
        p = out_lcccs; put_lcccs(p, out);
        if (out_f) memcpy(out_f, wo.f.data(), 8 * wo.f.size());
    });
}
// NIFSVerifier::verify (nifs.rs:117-162): returns 0 when the proof is accepted and fills out_lcccs
int lfo_nifs_verify(const lfo_problem* P, void* tr, const u64* proof, u64* out_lcccs) {
    return guard([&] {
        const RingParams& R = ring(P->ring); DecompParams dp; CCS ccs; Ajtai sch; LCCCS acc; CCCS cmi; load_problem(*P, R, dp, ccs, sch, acc, cmi);
        LFProof pf; const u64* p = proof; get_proof(p, pf, *P, R);
        LCCCS out; nifs_verify(R, dp, acc, cmi, pf, *(Transcript*)tr, ccs, out);
        if (out_lcccs) { u64* q = out_lcccs; put_lcccs(q, out); }
    });
}


// ---- negacyclic NTT (ntt.hpp): field 0 = Goldilocks, 1 = BabyBear; elements are u64 on this side for both fields
u64 lfo_ntt_root(int field, int log_n) { u64 r = 0; guard([&] { r = nttx::root(nttx::field(field), log_n); }); return r; }
int lfo_ntt_naive(int field, int log_n, const u64* in, u64* out, int inverse) {
    return guard([&] { nttx::naive(nttx::field(field), log_n, in, out, inverse != 0); });
}
int lfo_ntt_fast(int field, int log_n, const u64* in, u64* out, size_t batch, int inverse) {
    return guard([&] { auto F = nttx::field(field); auto pw = nttx::powers(F, log_n); const size_t n = (size_t)1 << log_n;
        #pragma omp parallel for schedule(static)
        for (long long b = 0; b < (long long)batch; ++b) nttx::fast(F, log_n, in + b * n, out + b * n, inverse != 0, pw); });
}
int lfo_ntt_schoolbook(int field, int log_n, const u64* a, const u64* b, u64* out) {
    return guard([&] { nttx::schoolbook(nttx::field(field), log_n, a, b, out); });
}

// ---- LatticeFold+ set check / range check (lfplus.hpp).  Same struct layout as the product's lf_csr / lf_plus_set.
struct lfo_plus_set { int32_t kind; int32_t pad; lfo_csr m; const u64* v; u64 n; };      // kind 0: matrix (m), 1: vector (v, n elements)
namespace {
plus::SparseR sparse_of(const RingParams& R, const lfo_csr& m) {
    plus::SparseR S; S.nrows = m.nrows; S.ncols = m.ncols; S.row_ptr.assign(m.row_ptr, m.row_ptr + m.nrows + 1); const u64 nnz = S.row_ptr.back();
    S.col.assign(m.col, m.col + nnz); S.val.assign(m.val, m.val + nnz * R.d); return S;
}
plus::PlusTranscript plus_transcript(const RingParams& R, const u64* seed, size_t n_seed) { plus::PlusTranscript T(R); if (n_seed) T.sp.absorb(seed, n_seed); return T; }
}
// returns the number of words written (or needed, when cap is too small: nothing is written), < 0 on error
long lfo_plus_set_check(int id, int nvars, const lfo_plus_set* sets, int n_sets, const lfo_csr* M, int n_M, const u64* seed, size_t n_seed, u64* out, size_t cap) {
    long n = -1; int rc = guard([&] { const RingParams& R = ring(id);
        std::vector<plus::MonSet> S; for (int i = 0; i < n_sets; ++i) { plus::MonSet s; s.matrix = sets[i].kind == 0; if (s.matrix) s.M = sparse_of(R, sets[i].m); else s.v.assign(sets[i].v, sets[i].v + sets[i].n * R.d); S.push_back(std::move(s)); }
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        auto T = plus_transcript(R, seed, n_seed);
        auto w = plus::set_out_words(R, plus::set_check(R, nvars, S, Ms, T)); n = (long)w.size(); if (w.size() <= cap) memcpy(out, w.data(), 8 * w.size()); });
    return rc ? rc : n;
}
int lfo_plus_set_check_verify(int id, const u64* words, size_t len, const u64* seed, size_t n_seed) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); plus::SetOut o; plus::set_out_parse(R, words, len, o); auto T = plus_transcript(R, seed, n_seed); ok = plus::set_check_verify(R, o, T) ? 1 : 0; });
    return rc ? rc : ok;
}
// RgInstance::from_f: tau[n], fcoms[3 x kappa x d], comM[k x kappa x d x d] (any output pointer may be NULL)
int lfo_plus_rg_from_f(int id, const u64* f, size_t n, const u64* A, size_t kappa, u64 b, int k, int l, u64* tau, u64* fcoms, u64* comM) {
    return guard([&] { const RingParams& R = ring(id); const size_t d = R.d; plus::DecompParameters dp{b, k, l};
        auto I = plus::rg_from_f(R, Vec(f, f + n * d), Vec(A, A + kappa * n * d), kappa, dp);
        if (tau) memcpy(tau, I.tau.data(), 8 * n);
        if (fcoms) { memcpy(fcoms, I.fcoms.cm_f.data(), 8 * kappa * d); memcpy(fcoms + kappa * d, I.fcoms.C_Mf.data(), 8 * kappa * d); memcpy(fcoms + 2 * kappa * d, I.fcoms.cm_mtau.data(), 8 * kappa * d); }
        if (comM) for (int kk = 0; kk < k; ++kk) memcpy(comM + (size_t)kk * kappa * d * d, I.comM_f[kk].data(), 8 * kappa * d * d); });
}
// Rg{nvars, L instances from_f(f_l, A)}.range_check(M, transcript): f = L x n x d
long lfo_plus_range_check(int id, int nvars, int L, const u64* f, size_t n, const u64* A, size_t kappa, u64 b, int k, int l, const lfo_csr* M, int n_M,
                          const u64* seed, size_t n_seed, u64* out, size_t cap) {
    long nw = -1; int rc = guard([&] { const RingParams& R = ring(id); const size_t d = R.d; plus::DecompParameters dp{b, k, l};
        std::vector<plus::RgInstance> inst; for (int i = 0; i < L; ++i) inst.push_back(plus::rg_from_f(R, Vec(f + (size_t)i * n * d, f + (size_t)(i + 1) * n * d), Vec(A, A + kappa * n * d), kappa, dp));
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        auto T = plus_transcript(R, seed, n_seed);
        auto w = plus::dcom_words(R, plus::range_check(R, nvars, inst, dp, Ms, T), kappa); nw = (long)w.size(); if (w.size() <= cap) memcpy(out, w.data(), 8 * w.size()); });
    return rc ? rc : nw;
}
int lfo_plus_range_check_verify(int id, const u64* words, size_t len, const u64* seed, size_t n_seed) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); plus::Dcom D; size_t kappa = 0; plus::dcom_parse(R, words, len, D, kappa); auto T = plus_transcript(R, seed, n_seed); ok = plus::range_check_verify(R, D, T) ? 1 : 0; });
    return rc ? rc : ok;
}
int lfo_plus_tensor(int id, const u64* r, int n, u64* out) { return guard([&] { auto t = plus::tensor(ring(id), r, n); memcpy(out, t.data(), 8 * t.size()); }); }
int lfo_plus_ring_mul(int id, const u64* a, const u64* b, u64* out) { return guard([&] { plus::r_mul(ring(id), out, a, b); }); }


// Cm{rg}.prove(M, transcript) (cm.rs:57-203) on instances built by from_f: returns the CmProof image; comx (ComX image) and g (L x n x d) optional
long lfo_plus_cm_prove(int id, int nvars, int L, const u64* f, size_t n, const u64* A, size_t kappa, u64 b, int k, int l, const lfo_csr* M, int n_M,
                       const u64* seed, size_t n_seed, u64* out, size_t cap, u64* comx, size_t comx_cap, u64* g) {
    long nw = -1; int rc = guard([&] { const RingParams& R = ring(id); const size_t d = R.d; plus::DecompParameters dp{b, k, l};
        std::vector<plus::RgInstance> inst; for (int i = 0; i < L; ++i) inst.push_back(plus::rg_from_f(R, Vec(f + (size_t)i * n * d, f + (size_t)(i + 1) * n * d), Vec(A, A + kappa * n * d), kappa, dp));
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        auto T = plus_transcript(R, seed, n_seed); plus::Com com; plus::CmProof P;
        plus::cm_prove(R, nvars, inst, dp, Ms, T, com, P);
        auto w = plus::cm_proof_words(R, P); nw = (long)w.size(); if (w.size() <= cap) memcpy(out, w.data(), 8 * w.size());
        if (comx) { auto x = plus::comx_words(com.x); if (x.size() > comx_cap) throw std::runtime_error("ComX buffer too small"); memcpy(comx, x.data(), 8 * x.size()); }
        if (g) memcpy(g, com.g.data(), 8 * com.g.size()); });
    return rc ? rc : nw;
}
// CmProof::verify(M, transcript): 1 accept (comx filled when not NULL), 0 reject
int lfo_plus_cm_verify(int id, const u64* words, size_t len, const lfo_csr* M, int n_M, const u64* seed, size_t n_seed, u64* comx, size_t comx_cap) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); plus::CmProof P; plus::cm_proof_parse(R, words, len, P);
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        auto T = plus_transcript(R, seed, n_seed); plus::ComX X; ok = plus::cm_verify(R, P, Ms, T, X) ? 1 : 0;
        if (ok && comx) { auto x = plus::comx_words(X); if (x.size() > comx_cap) throw std::runtime_error("ComX buffer too small"); memcpy(comx, x.data(), 8 * x.size()); } });
    return rc ? rc : ok;
}


// Matrix::try_mul_vec on the coefficient ring: out[kappa x d] = A[kappa x n] * x[n]
int lfo_plus_mat_vec(int id, const u64* A, size_t kappa, size_t n, const u64* x, u64* out) {
    return guard([&] { const RingParams& R = ring(id); auto y = plus::mat_mul_vec(R, Vec(A, A + kappa * n * R.d), kappa, n, Vec(x, x + n * R.d)); memcpy(out, y.data(), 8 * y.size()); });
}


// Mlin::mlin: linb2x = cm_g[kappa x d] ro[nvars x 2] vo[(1+n_M) x 2 x d]; g[n x d]; returns the CmProof image length
long lfo_plus_mlin(int id, int L, const u64* f, size_t n, const u64* A, size_t kappa, u64 b, int k, int l, const lfo_csr* M, int n_M, const u64* seed, size_t n_seed,
                   u64* proof, size_t cap, u64* linb2x, u64* g) {
    long nw = -1; int rc = guard([&] { const RingParams& R = ring(id); const size_t d = R.d; plus::DecompParameters dp{b, k, l};
        std::vector<Vec> fs; for (int i = 0; i < L; ++i) fs.emplace_back(f + (size_t)i * n * d, f + (size_t)(i + 1) * n * d);
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        auto T = plus_transcript(R, seed, n_seed); plus::CmProof P;
        plus::LinB2 o = plus::mlin(R, fs, Vec(A, A + kappa * n * d), kappa, dp, Ms, T, P);
        auto w = plus::cm_proof_words(R, P); nw = (long)w.size(); if (w.size() <= cap) memcpy(proof, w.data(), 8 * w.size());
        if (linb2x) { u64* p = linb2x; for (auto* v : {&o.cm_g, &o.ro, &o.vo}) { memcpy(p, v->data(), 8 * v->size()); p += v->size(); } }
        if (g) memcpy(g, o.g.data(), 8 * o.g.size()); });
    return rc ? rc : nw;
}
// Decomp::decompose: proof = C0 C1 [kappa x d each] v0 v1 [(1+n_M) x 2 x d each]; F = 2 x n x d (optional)
int lfo_plus_decompose(int id, const u64* f, size_t n, const u64* r_pairs, const lfo_csr* M, int n_M, const u64* A, size_t kappa, u64 B, u64* proof, u64* F) {
    return guard([&] { const RingParams& R = ring(id); const size_t d = R.d; int nv = plus::ceil_log2(n);
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        Vec Fs[2]; plus::DecompProof P = plus::decompose(R, Vec(f, f + n * d), Vec(r_pairs, r_pairs + 2 * nv), Ms, Vec(A, A + kappa * n * d), kappa, B, Fs);
        u64* p = proof; for (auto* v : {&P.C[0], &P.C[1], &P.v[0], &P.v[1]}) { memcpy(p, v->data(), 8 * v->size()); p += v->size(); }
        if (F) { memcpy(F, Fs[0].data(), 8 * n * d); memcpy(F + n * d, Fs[1].data(), 8 * n * d); } });
}
int lfo_plus_decompose_verify(int id, const u64* proof, size_t kappa, int n_M, const u64* cm_f, const u64* v, u64 B) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); const size_t d = R.d, nc = kappa * d, nv = (size_t)(1 + n_M) * 2 * d; plus::DecompProof P;
        P.C[0].assign(proof, proof + nc); P.C[1].assign(proof + nc, proof + 2 * nc); P.v[0].assign(proof + 2 * nc, proof + 2 * nc + nv); P.v[1].assign(proof + 2 * nc + nv, proof + 2 * nc + 2 * nv);
        ok = plus::decompose_verify(R, P, Vec(cm_f, cm_f + nc), Vec(v, v + nv), B) ? 1 : 0; });
    return rc ? rc : ok;
}


// ComR1CS::linearize on (A, B, C, f): image [nvars] r msgs v|va|vb|vc (the LinB it returns is f with r = (ro, ro), v = the four evaluations twice)
long lfo_plus_r1cs_linearize(int id, const lfo_csr* abc, const u64* f, size_t n, void* tr, u64* out, size_t cap) {
    long nw = -1; int rc = guard([&] { const RingParams& R = ring(id); plus::SparseR M[3]; for (int i = 0; i < 3; ++i) M[i] = sparse_of(R, abc[i]);
        auto w = plus::r1cs_lin_words(plus::r1cs_linearize(R, M, Vec(f, f + n * R.d), *(plus::PlusTranscript*)tr)); nw = (long)w.size(); if (w.size() <= cap) memcpy(out, w.data(), 8 * w.size()); });
    return rc ? rc : nw;
}
int lfo_plus_r1cs_linearize_verify(int id, const u64* words, size_t len, void* tr) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); plus::R1csLinProof P; plus::r1cs_lin_parse(R, words, len, P); ok = plus::r1cs_linearize_verify(R, P, *(plus::PlusTranscript*)tr) ? 1 : 0; });
    return rc ? rc : ok;
}
// stateful LatticeFold+ transcript for multi-protocol flows (plus.rs): create / absorb base elements / free; the other lfo_plus_*_t entry points take it
void* lfo_plus_tr_new(int id, const u64* seed, size_t n_seed) { void* p = nullptr; guard([&] { p = new plus::PlusTranscript(plus_transcript(ring(id), seed, n_seed)); }); return p; }
void lfo_plus_tr_free(void* t) { delete (plus::PlusTranscript*)t; }
u64 lfo_plus_tr_challenge(void* t) { return ((plus::PlusTranscript*)t)->get_challenge(); }
long lfo_plus_mlin_t(int id, int L, const u64* f, size_t n, const u64* A, size_t kappa, u64 b, int k, int l, const lfo_csr* M, int n_M, void* tr, u64* proof, size_t cap, u64* linb2x, u64* g) {
    long nw = -1; int rc = guard([&] { const RingParams& R = ring(id); const size_t d = R.d; plus::DecompParameters dp{b, k, l};
        std::vector<Vec> fs; for (int i = 0; i < L; ++i) fs.emplace_back(f + (size_t)i * n * d, f + (size_t)(i + 1) * n * d);
        std::vector<plus::SparseR> Ms; for (int i = 0; i < n_M; ++i) Ms.push_back(sparse_of(R, M[i]));
        plus::CmProof P; plus::LinB2 o = plus::mlin(R, fs, Vec(A, A + kappa * n * d), kappa, dp, Ms, *(plus::PlusTranscript*)tr, P);
        auto w = plus::cm_proof_words(R, P); nw = (long)w.size(); if (w.size() <= cap) memcpy(proof, w.data(), 8 * w.size());
        if (linb2x) { u64* p = linb2x; for (auto* v : {&o.cm_g, &o.ro, &o.vo}) { memcpy(p, v->data(), 8 * v->size()); p += v->size(); } }
        if (g) memcpy(g, o.g.data(), 8 * o.g.size()); });
    return rc ? rc : nw;
}
int lfo_plus_cm_verify_t(int id, const u64* words, size_t len, int n_M, void* tr) {
    int ok = 0; int rc = guard([&] { const RingParams& R = ring(id); plus::CmProof P; plus::cm_proof_parse(R, words, len, P);
        std::vector<plus::SparseR> Ms(n_M); plus::ComX X; ok = plus::cm_verify(R, P, Ms, *(plus::PlusTranscript*)tr, X) ? 1 : 0; });
    return rc ? rc : ok;
}

}  // extern "C"
