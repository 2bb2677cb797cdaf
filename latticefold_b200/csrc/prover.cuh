// One LatticeFold folding step on the device: the host side of NIFSProver::prove
//   crates/latticefold/src/nifs.rs:48-103   = linearization (nifs/linearization.rs:145-189)
//                                            + 2 x decomposition (nifs/decomposition.rs:33-88)
//                                            + folding (nifs/folding.rs:42-130)
// written above the kernels.  What the reference does on small vectors between transcript calls (x_s, y_0 Horner,
// RotSum, cm_0/u_0/x_0) stays on the CPU here as well; everything witness-sized runs on the GPU and stays in HBM.
// Algebraic regrouping relative to the reference (results are identical field elements because the arithmetic is exact):
//   * f-hat MLEs are never materialised: MLE j slot k of a witness IS coefficient plane j*S+k (arith.rs:282-291);
//     decomposed pieces are held as int8 balanced digits;
//   * theta (folding.rs:236-246) is read off the sumcheck's final fold instead of a second evaluate_mles pass;
//   * eq(r,.) tables are built once per point and shared by decomposition and folding.
#pragma once
#include "sumcheck.cuh"

struct lf_witness { lf_words *f = nullptr, *f_coeff = nullptr, *w_ccs = nullptr; size_t n = 0, pitch = 0, W = 0, w_pitch = 0; };

struct lf_prover {
    lf_ctx* ctx = nullptr;
    lf_ctx* aux = nullptr;         // second stream + scratch: the accumulator's decomposition runs beside the linearization
    cudaEvent_t acc_ready = nullptr;   // host-buffer entry point: recorded when the accumulator witness is on the device, so that its
                                       // decomposition starts while the incoming witness is still being copied
    int ring = 0, L = 0, K = 0; uint64_t B = 0, b = 0;
    size_t kappa = 0, n = 0;
    size_t m = 0, n_ccs = 0, l = 0, t = 0, q = 0, d = 0, s = 0;
    std::vector<std::vector<int>> S; lf::HV c;
    lf_ajtai* A = nullptr; bool own_A = true;
    std::vector<lf_sparse*> M;
    double timings[5] = {0, 0, 0, 0, 0};
    bool detail = false; std::vector<std::pair<std::string, double>> marks; std::chrono::steady_clock::time_point last_mark;
};

namespace lf {

struct LCCCS { HV r, v, cm, u, x_w, h; };

template <class Rg> struct Prover {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef HostRing<Rg> HR; typedef typename HR::El El; typedef typename Rg::W W;
    typedef PtrListT<W> PL;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU;
    static W* wp(lf_words* p) { return reinterpret_cast<W*>(p); }
    static const W* wp(const lf_words* p) { return reinterpret_cast<const W*>(p); }
    static lf_words* ow(W* p) { return reinterpret_cast<lf_words*>(p); }
    lf_prover* P; Engine<Rg> E; HR H;
    explicit Prover(lf_prover* p) : P(p), E(p->ctx), H(E.tab()) {}
    // diagnostic phase marks (LF_TIMING_DETAIL=1): synchronise and record the time since the previous mark
    void mark(const char* name) { if (!P->detail) return; E.sync(); auto now = std::chrono::steady_clock::now(); P->marks.push_back({name, std::chrono::duration<double, std::milli>(now - P->last_mark).count()}); P->last_mark = now; }

    static HV sf_to_ring(const std::vector<u64>& sf) { size_t n = sf.size() / TAU; HV o(n * D); for (size_t i = 0; i < n; ++i) { El e = HR::from_sf(&sf[i * TAU]); std::memcpy(&o[i * D], e.data(), 8 * D); } return o; }
    static std::vector<u64> squeeze(Transcript<Rg>& T, const char* tag, int n) { T.absorb_tag(tag); std::vector<u64> o((size_t)n * TAU); for (int i = 0; i < n; ++i) T.get_challenge(&o[(size_t)i * TAU]); return o; }
    static size_t cnt(const HV& v) { return v.size() / D; }
    // sharding: rank g owns witness elements / hypercube indices [g * n/G, (g+1) * n/G)
    int world() const { return P->ctx->world; }
    int rank() const { return P->ctx->rank; }
    size_t nl() const { return P->n / world(); }        // local witness length
    size_t ml() const { return P->m / world(); }        // local hypercube slab
    size_t Wl() const { return nl() / P->L; }

    void sanity_check() const {   // nifs.rs:165-173
        size_t want = std::max((P->n_ccs - P->l - 1) * (size_t)P->L, P->m), p2 = 1; while (p2 < want) p2 <<= 1;
        if (P->m != p2 || ((size_t)1 << P->s) != P->m) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "InvalidSizeBounds");
        if (P->n > P->m) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "witness longer than 2^s");
        const size_t G = (size_t)world();
        if (G > 1 && ((G & (G - 1)) || P->n != P->m || P->n % (G * P->L) || ml() < 2))
            throw LfException(LF_ERR_UNSUPPORTED, "sharding needs a power-of-two rank count, n == 2^s, ranks | W and at least two hypercube entries per rank");
    }

    // ------------------------------------------------------------------ witnesses (arith.rs:299-313, Witness::from_f)
    lf_witness* witness_from_f_device(W* f_dev /* takes ownership, pitch = pitch_of(n) */) {
        lf_witness* w = new lf_witness; w->n = nl(); w->pitch = pitch_of(nl()); w->f = ow(f_dev);
        if (P->n % (P->L * (size_t)world())) throw LfException(LF_ERR_INCORRECT_LENGTH, "witness length is not a multiple of L (times the rank count)");
        w->W = Wl(); w->w_pitch = pitch_of(w->W);
        w->f_coeff = ow(E.template dalloc<W>(w->pitch * D)); E.crt(wp(w->f), w->pitch, wp(w->f_coeff), w->pitch, w->n, true);
        w->w_ccs = ow(E.template dalloc<W>(w->w_pitch * D)); E.gadget_recompose(wp(w->f), w->pitch, wp(w->w_ccs), w->w_pitch, w->W, P->B, P->L);
        return w;
    }
    lf_witness* upload_witness(const u64* f_host) {
        W* f = E.template dalloc<W>(pitch_of(nl()) * D); E.upload_planes(f_host, nl(), f, pitch_of(nl()));   // this rank's slice
        return witness_from_f_device(f);
    }
    void free_witness(lf_witness* w) { if (!w) return; E.dfree(w->f); E.dfree(w->f_coeff); E.dfree(w->w_ccs); delete w; }

    // ------------------------------------------------------------------ shared device helpers
    struct DevVec { W* p = nullptr; size_t n = 0, pitch = 0; };
    DevVec eq_table(const HV& r) {      // this rank's slab of eq(., r)
        DevVec v; v.n = ((size_t)1 << cnt(r)) / world(); v.pitch = pitch_of(v.n); v.p = E.template dalloc<W>(v.pitch * D);
        E.eq_table(r.data(), (int)cnt(r), v.p, v.pitch, (size_t)rank() * v.n, v.n); return v;
    }
    // Mz tables for a list of z = head_k || tail_k; out: [count * t] rows of pitch mz_pitch, effective length eff[j]
    struct MzSet { W* p = nullptr; size_t pitch = 0, stride = 0; int rows = 0; size_t* d_len = nullptr; std::vector<size_t> len; };
    // upload = false: the caller copies the length table itself, on the stream that will read it (upload_mz_len)
    MzSet alloc_mz(int count, bool upload = true) {
        MzSet z; z.rows = count * (int)P->t; size_t mx = 1; for (auto* M : P->M) mx = std::max(mx, M->eff_rows);
        z.pitch = pitch_of(mx); z.stride = z.pitch * D; z.p = E.template dalloc<W>((size_t)z.rows * z.stride);
        z.len.resize(z.rows); for (int i = 0; i < z.rows; ++i) z.len[i] = P->M[i % P->t]->eff_rows;
        z.d_len = E.template dalloc<size_t>(z.rows);
        if (upload) upload_mz_len(z);
        return z;
    }
    void upload_mz_len(MzSet& z) { E.h2d(z.d_len, z.len.data(), z.rows * sizeof(size_t)); }
    void free_mz(MzSet& z) { E.dfree(z.p); E.dfree(z.d_len); z.p = nullptr; }
    // z = head || tail where the tail (w_ccs) is, when sharded, the all-gathered concatenation of the ranks' slabs:
    // chunk r (tail_chunk elements) at tail + r * tail_chunk_stride.  Rows of M are this rank's slab.
    void compute_mz(MzSet& z, int k, const HV& head, const W* tail, size_t tail_pitch, size_t tail_len, size_t tail_chunk = ~(size_t)0, size_t tail_chunk_stride = 0) {   // mat_vec_mul x t (arith/utils.rs:52-65)
        const size_t hl = cnt(head), hp = pitch_of(hl);
        W* d_head = E.template dalloc<W>(hp * D); E.upload_small(head.data(), hl, d_head, hp);
        for (size_t j = 0; j < P->t; ++j) {
            if (P->M[j]->ncols != hl + tail_len) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
            E.spmv(P->M[j], d_head, hl, hp, tail, tail_pitch, z.p + ((size_t)k * P->t + j) * z.stride, z.pitch, P->M[j]->eff_rows, tail_chunk, tail_chunk_stride);
        }
        E.dfree(d_head);
    }
    // the K pieces of one decomposition against every CCS matrix: t launches (blockIdx.z = piece) instead of K * t
    void compute_mz_batch(MzSet& z, int k0, int K, const std::vector<HV>& heads, const W* tail, size_t tail_pitch, size_t tail_piece_stride, size_t tail_len, size_t tail_chunk, size_t tail_chunk_stride) {
        const size_t hl = cnt(heads[0]), hp = pitch_of(hl);
        W* d_heads = E.template dalloc<W>((size_t)K * hp * D);
        for (int k = 0; k < K; ++k) { if (cnt(heads[k]) != hl) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength"); E.upload_small(heads[k].data(), hl, d_heads + (size_t)k * hp * D, hp); }
        for (size_t j = 0; j < P->t; ++j) {
            if (P->M[j]->ncols != hl + tail_len) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
            E.spmv(P->M[j], d_heads, hl, hp, tail, tail_pitch, z.p + ((size_t)k0 * P->t + j) * z.stride, z.pitch, P->M[j]->eff_rows, tail_chunk, tail_chunk_stride,
                   K, hp * D, tail_piece_stride, P->t * z.stride);
        }
        E.dfree(d_heads);
    }
    // all-gather `count` consecutive per-piece w_ccs slabs (each wc_stride words) of every rank: returns [world][count][D][pitch]
    W* gather_wccs(const W* local, size_t words) {      // words of type W; the collective moves them as u64 lanes (pitches are multiples of 32 words)
        if (world() == 1) return const_cast<W*>(local);
        W* all = E.template dalloc<W>(words * world());
        LF_CUDA(cudaMemcpyAsync(all + (size_t)rank() * words, local, words * sizeof(W), cudaMemcpyDeviceToDevice, E.st()));
        E.collective(1, reinterpret_cast<u64*>(all), words * sizeof(W) / 8);
        return all;
    }
    // evaluate every Mz row at the point whose eq table is given -> rows x D limbs (host)
    const u64* eval_mz_async(const MzSet& z, int row0, int rows, const DevVec& eq) {      // pinned result, valid after the next sync / event
        PL Y; Y.p[0] = eq.p; Y.len[0] = eq.n;
        u64* d_out = E.template dalloc<u64>((size_t)rows * D);
        E.dot(z.p + (size_t)row0 * z.stride, z.stride, z.pitch, rows, z.d_len + row0, Y, eq.pitch, 1, z.pitch, d_out, "k_dot_eval");
        const u64* land = E.d2h_async(d_out, (size_t)rows * D); E.dfree(d_out); return land;
    }
    HV eval_mz(const MzSet& z, int row0, int rows, const DevVec& eq) {
        const u64* land = eval_mz_async(z, row0, rows, eq); E.sync(); return HV(land, land + (size_t)rows * D);
    }

    // local z vectors: z = head || tail with the l+1 head elements (x || 1, or x_s[k]) in front of this rank's slice of w_ccs; the head
    // entries are real on rank 0 and zero elsewhere, so that sums over ranks count them once
    size_t hc() const { return P->l + 1; }
    size_t zcols() const { return hc() + Wl(); }
    // (M_j z_k)(r) for every matrix j and `count` local z vectors, transposed (k_csc_eq): pinned result [t][count][D], valid after the next sync / event
    const u64* eval_z_async(const W* z, size_t z_pitch, size_t z_stride, int count, const HV& point) {
        EvalPrep e = eval_prepare(point); const u64* land = eval_with(e, z, z_pitch, z_stride, count); E.dfree(e.v); return land;
    }
    // head entries of `count` consecutive local z vectors (rank 0: the given elements; other ranks: zero), one strided copy
    void write_heads(W* z, size_t z_pitch, int count, const std::vector<HV>& heads) {
        const size_t h = hc();
        if (rank() != 0) { LF_CUDA(cudaMemset2DAsync(z, z_pitch * sizeof(W), 0, h * sizeof(W), (size_t)count * D, E.st())); return; }
        W* stage = (W*)E.arena_alloc((size_t)count * D * h * sizeof(W));
        for (int k = 0; k < count; ++k) { if (cnt(heads[k]) != h) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
            for (int l = 0; l < D; ++l) for (size_t e = 0; e < h; ++e) stage[((size_t)k * D + l) * h + e] = (W)heads[k][e * D + l]; }
        LF_CUDA(cudaMemcpy2DAsync(z, z_pitch * sizeof(W), stage, h * sizeof(W), h * sizeof(W), (size_t)count * D, cudaMemcpyHostToDevice, E.st()));
    }

    // ------------------------------------------------------------------ linearization (linearization.rs:145-189)
    struct LinOut { LCCCS lc; HV msgs; DevVec eq_r; };
    // pre_tail: the all-gathered w_ccs slabs when the caller has already queued that collective (the sharded step issues it before
    // the accumulator's decomposition so that NCCL's cross-stream ordering does not park the linearization behind the auxiliary stream)
    LinOut linearize(const HV& cm_i_cm, const HV& x_ccs, const lf_witness* w, Transcript<Rg>& T, W* pre_tail = nullptr) {
        LinOut o; const int s = (int)P->s; const size_t m = ml();    // tables hold this rank's slab
        HV head = x_ccs; { El one = HR::from_u64(1); head.insert(head.end(), one.begin(), one.end()); }      // z = x || 1 || w  (arith.rs:399-409)
        HV beta = sf_to_ring(squeeze(T, "beta_s", s));                                                       // linearization/utils.rs:113-124
        MzSet mz = alloc_mz(1);
        { const size_t words = w->w_pitch * D; W* tail = pre_tail ? pre_tail : gather_wccs(wp(w->w_ccs), words);
          compute_mz(mz, 0, head, tail, w->w_pitch, w->W * world(), world() == 1 ? ~(size_t)0 : w->W, words);
          if (tail != wp(w->w_ccs)) E.dfree(tail); }
        // sumcheck list: for each term with c_i != 0, the Mz named by S_i; eq(beta,.) last (linearization/utils.rs:63-88)
        std::vector<int> list; for (size_t i = 0; i < P->q; ++i) { bool z = true; for (int l = 0; l < D; ++l) z = z && P->c[i * D + l] == 0; if (z) continue; for (int j : P->S[i]) list.push_back(j); }
        const int Mn = (int)list.size() + 1;
        lf_sumcheck sc; sc.ctx = P->ctx; sc.nv = s; sc.deg = (int)P->d + 1; sc.kind = LF_COMB_LIN; sc.len = m; sc.sharded = world() > 1;
        SumcheckDriver<Rg> drv(P->ctx, &sc);
        SumcheckDriver<Rg>::alloc_group(E, sc.dense, Mn, m);
        W* dense = wp(sc.dense.cur); constexpr size_t WB = sizeof(W);
        LF_CUDA(cudaMemsetAsync(dense, 0, (size_t)Mn * sc.dense.stride * WB, E.st()));
        for (int k = 0; k + 1 < Mn; ++k) { const int j = list[k];
            LF_CUDA(cudaMemcpy2DAsync(dense + (size_t)k * sc.dense.stride, sc.dense.pitch * WB, mz.p + (size_t)j * mz.stride, mz.pitch * WB, mz.len[j] * WB, D, cudaMemcpyDeviceToDevice, E.st())); }
        E.eq_table(beta.data(), s, dense + (size_t)(Mn - 1) * sc.dense.stride, sc.dense.pitch, (size_t)rank() * m, m);
        // LIN comb (linearization/utils.rs:90-107): vals[] is indexed by the CCS matrix index j.  The list position of
        // matrix j coincides with j for R1CS and the degree-3 CCS; reproduce the reference literally and refuse anything else.
        drv.set_terms(Mn, sc.deg, true, P->S);
        sc.d_coef = E.template dalloc<u64>(P->q * D);
        E.h2d(sc.d_coef, P->c.data(), P->q * D * 8);
        std::vector<u64> point; o.msgs = run_sumcheck(drv, T, point);
        drv.free_all();
        o.lc.r = sf_to_ring(point);
        o.eq_r = eq_table(o.lc.r);
        // v = f-hat(r) (linearization.rs:126-131), u = Mz(r) (:133-139)
        u64* d_v = E.small_dev((size_t)TAU * D);
        E.template coeff_eval<W>(wp(w->f_coeff), w->pitch, 0, 1, o.eq_r.p, o.eq_r.pitch, w->n, d_v);
        o.lc.v.resize((size_t)TAU * D); E.download_words(d_v, o.lc.v.size(), o.lc.v.data());
        o.lc.u = eval_mz(mz, 0, (int)P->t, o.eq_r);
        free_mz(mz);
        T.absorb_slice(o.lc.v.data(), cnt(o.lc.v)); T.absorb_slice(o.lc.u.data(), cnt(o.lc.u));
        o.lc.cm = cm_i_cm; o.lc.x_w = x_ccs; { El one = HR::from_u64(1); o.lc.h.assign(one.begin(), one.end()); }
        return o;
    }
    // MLSumcheck::prove_as_subprotocol (sumcheck.rs:53-80); returns nv x (deg+1) x D message limbs, point = nv x TAU
    HV run_sumcheck(SumcheckDriver<Rg>& drv, Transcript<Rg>& T, std::vector<u64>& point, HV* final_vals = nullptr) {
        lf_sumcheck* sc = drv.sc; const int ne = sc->deg + 1;
        if (sc->nv == 0) throw LfException(LF_ERR_SUMCHECK_MISUSE, "Attempt to prove a constant.");
        T.absorb_u64((u64)sc->nv); T.absorb_u64((u64)sc->deg);
        HV msgs((size_t)sc->nv * ne * D); point.assign((size_t)sc->nv * TAU, 0);
        for (int i = 0; i < sc->nv; ++i) {
            if (i > 0) drv.apply_challenge(&point[(size_t)(i - 1) * TAU]);
            u64* msg = &msgs[(size_t)i * ne * D];
            drv.evaluate(msg);
            auto t0 = std::chrono::steady_clock::now();
            T.absorb_slice(msg, ne); T.get_challenge(&point[(size_t)i * TAU]); T.absorb_sf(&point[(size_t)i * TAU]);
            P->timings[3] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        if (final_vals) { drv.apply_challenge(&point[(size_t)(sc->nv - 1) * TAU]); final_vals->resize((size_t)(sc->dense.count + sc->n_f) * D); drv.final_values(final_vals->data()); }
        return msgs;
    }

    // ------------------------------------------------------------------ decomposition (decomposition.rs:33-88)
    struct StepBuffers {    // witness-sized state shared by the two decompositions and the folding
        int8_t* dig = nullptr; size_t dig_pitch = 0, dig_stride = 0;     // [2K][D][pitch]
        W* pieces = nullptr; size_t pc_pitch = 0, pc_stride = 0;         // NTT form of every piece, [2K][D][pitch]
        W* zl = nullptr; size_t zl_pitch = 0, zl_stride = 0;           // local z of every piece (head || gadget_recompose of the piece), [2K][D][pitch]
    };
    struct DecOut { std::vector<HV> x_s, y_s, u_s, v_s; std::vector<LCCCS> lc; };
    // decompose_big_vec_into_k_vec_and_compose_back (decomposition/utils.rs:12-42) on l+1 elements: host
    std::vector<HV> compute_x_s(const LCCCS& cm) {
        HV xs = cm.x_w; xs.insert(xs.end(), cm.h.begin(), cm.h.end());
        const size_t ne = cnt(xs); const int L = P->L, K = P->K;
        std::vector<HV> out(K, HV(ne * D, 0));
        std::vector<int64_t> dB(L), db(K);
        for (size_t e = 0; e < ne; ++e) {
            El co = H.icrt(HR::load(&xs[e * D]));
            // piece k, chunk digit l, coefficient c  ->  recompose over l with base B
            std::vector<El> acc(K, HR::zero());
            for (int c = 0; c < D; ++c) {
                if (!balanced_digits(F::to_signed(co[c]), (int64_t)P->B, L, dB.data())) throw LfException(LF_ERR_DOES_NOT_FIT, "x_s: coefficient does not fit L digits of base B");
                u64 pw = 1;
                for (int l = 0; l < L; ++l) {
                    if (!balanced_digits(dB[l], (int64_t)P->b, K, db.data())) throw LfException(LF_ERR_DOES_NOT_FIT, "x_s: digit does not fit K digits of base b");
                    for (int k = 0; k < K; ++k) acc[k][c] = F::add(acc[k][c], F::mul(F::from_i64(db[k]), pw));
                    pw = F::mul(pw, P->B % F::P);
                }
            }
            for (int k = 0; k < K; ++k) { El nt = H.crt(acc[k]); std::memcpy(&out[k][e * D], nt.data(), 8 * D); }
        }
        return out;
    }
    // The device half of a decomposition is enqueued without any host synchronisation (results land in the pinned arena
    // behind an event); the host half (y_0 Horner, transcript absorbs) runs later, overlapping whatever the GPU does next.
    // Results arrive in groups of pieces, each behind its own event: the commitments of all pieces first (one pass over the matrix),
    // then CRT / recompose / v_s / u_s group by group, so that the host hashes group g (Fiat-Shamir absorbs x_s, y_s, u_s, v_s piece by
    // piece) while the device works on group g+1 -- only the last group's absorb is exposed.
    static constexpr int DEC_GROUPS = 4;
    struct DecPending { LCCCS cm; std::vector<HV> x_s; const u64* y_pin = nullptr; bool y_early = false; int ngroups = 0; int g0[DEC_GROUPS + 1] = {0};
                        const u64 *v_pin[DEC_GROUPS] = {nullptr}, *u_pin[DEC_GROUPS] = {nullptr}; cudaEvent_t ev[DEC_GROUPS] = {nullptr}; };
    struct EvalPrep { W* v = nullptr; size_t vp = 0, vs = 0; };
    EvalPrep eval_prepare(const HV& point) {      // v_j = M_j^T eq(., point) on this rank's columns
        const size_t t = P->t;
        for (auto* M : P->M) if (M->t_ncols != zcols()) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
        typename Engine<Rg>::EqHalves q = E.eq_halves(point.data(), (int)cnt(point));
        EvalPrep e; e.vp = pitch_of(zcols()); e.vs = e.vp * D; e.v = E.template dalloc<W>(t * e.vs);
        for (size_t j = 0; j < t; ++j) E.csc_eq(P->M[j], q, e.v + j * e.vs, e.vp);
        E.free_halves(q); return e;
    }
    const u64* eval_with(const EvalPrep& e, const W* z, size_t z_pitch, size_t z_stride, int count) {      // pinned [t][count][D]
        if (count > MAX_LIST) throw LfException(LF_ERR_UNSUPPORTED, "more than MAX_LIST z vectors in one evaluation");
        const size_t t = P->t;
        PL Y; for (int k = 0; k < count; ++k) { Y.p[k] = z + (size_t)k * z_stride; Y.len[k] = zcols(); }
        u64* d_out = E.template dalloc<u64>(t * count * D);
        E.dot(e.v, e.vs, e.vp, (int)t, nullptr, Y, z_pitch, count, zcols(), d_out, "k_dot_eval");
        const u64* land = E.d2h_async(d_out, t * count * D); E.dfree(d_out); return land;
    }
    // Part A needs only the witness and the instance's public part (x_w, h): digit split, x_s, the K-1 commitments, CRT and
    // gadget recomposition of every piece.  Part B needs the evaluation point r: v_s and u_s, group by group.  The incoming witness's
    // part A is therefore queued BEFORE the linearization sumcheck has produced its r and runs beside it on the auxiliary stream.
    DecPending decompose_enqueue_a(const HV& x_w, const HV& h, const lf_witness* w, StepBuffers& sb, int half) {
        DecPending o; o.cm.x_w = x_w; o.cm.h = h; const int K = P->K; const size_t n = nl(), kappa = P->kappa;
        int8_t* dig = sb.dig + (size_t)half * K * sb.dig_stride;
        W* pieces = sb.pieces + (size_t)half * K * sb.pc_stride;
        W* zl = sb.zl + (size_t)half * K * sb.zl_stride;
        // decompose_witness: f_coeff.decompose_to_vec(b, K).transpose() (decomposition.rs:162-167)
        E.digit_split(wp(w->f_coeff), w->pitch, dig, sb.dig_pitch, n, P->b, K);
        o.x_s = compute_x_s(o.cm);
        write_heads(zl, sb.zl_pitch, K, o.x_s);
        mark("dec.split_x_s");
        // commit_witnesses (decomposition.rs:178-201): K-1 commits in one pass over A.  Digit pieces: integer GEMM on the tensor cores
        // straight from the int8 digits (commit_mma.cuh); otherwise (other rings) lazily reduced dot products on the pieces' NTT forms
        const bool mma = K > 1 && E.can_commit_digits(P->A, K - 1, sb.dig_pitch); o.y_early = mma;
        if (mma) { u64* d_y = E.template dalloc<u64>(kappa * (K - 1) * D);
            E.commit_digits(P->A, dig + sb.dig_stride, sb.dig_pitch, sb.dig_stride, K - 1, d_y);
            o.y_pin = E.d2h_async(d_y, kappa * (K - 1) * D); E.dfree(d_y); }
        mark("dec.commit");
        // CRT and recompose of every piece (arith.rs:324-338), a few pieces at a time: the NTT forms a CRT launch writes (50 MB per
        // piece at C2) are still in L2 when the recomposition reads them
        for (int k0 = 0; k0 < K; k0 += 4) { const int c = std::min(4, K - k0);
            E.crt_digits(dig + (size_t)k0 * sb.dig_stride, sb.dig_pitch, pieces + (size_t)k0 * sb.pc_stride, sb.pc_pitch, n, c, sb.dig_stride, sb.pc_stride);
            E.gadget_recompose(pieces + (size_t)k0 * sb.pc_stride, sb.pc_pitch, zl + (size_t)k0 * sb.zl_stride + hc(), sb.zl_pitch, w->W, P->B, P->L, c, sb.pc_stride, sb.zl_stride); }
        if (!mma && K > 1) {      // the dot-product commit needs every piece's NTT form
            PL Y; for (int k = 1; k < K; ++k) { Y.p[k - 1] = pieces + (size_t)k * sb.pc_stride; Y.len[k - 1] = n; }
            u64* d_y = E.template dalloc<u64>(kappa * (K - 1) * D);
            E.dot(wp(P->A->p), P->A->pitch * D, P->A->pitch, (int)kappa, nullptr, Y, sb.pc_pitch, K - 1, n, d_y, "k_dot_commit");
            o.y_pin = E.d2h_async(d_y, kappa * (K - 1) * D); E.dfree(d_y);
        }
        mark("dec.crt_recompose");
        return o;
    }
    // Results of part B arrive in groups of pieces, each behind its own event, so that the host hashes group g while the device works on g+1
    void decompose_enqueue_b(DecPending& o, const LCCCS& cm, const DevVec& eq_r, StepBuffers& sb, int half) {
        const int K = P->K; const size_t n = nl();
        if (cnt(cm.cm) != P->kappa) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongCommitmentLength");
        o.cm = cm;
        int8_t* dig = sb.dig + (size_t)half * K * sb.dig_stride;
        W* zl = sb.zl + (size_t)half * K * sb.zl_stride;
        EvalPrep ep = eval_prepare(cm.r);
        o.ngroups = std::min(DEC_GROUPS, K); if (std::getenv("LF_DEC_ONE_GROUP")) o.ngroups = 1;
        for (int g = 0; g <= o.ngroups; ++g) o.g0[g] = (int)((size_t)K * g / o.ngroups);
        for (int g = 0; g < o.ngroups; ++g) {
            const int k0 = o.g0[g], c = o.g0[g + 1] - k0;
            // compute_v_s (decomposition.rs:204-211): f-hat of piece k evaluated at r, straight from the digits
            { u64* d_v = E.template dalloc<u64>((size_t)c * TAU * D);
              E.template coeff_eval<int8_t>(dig + (size_t)k0 * sb.dig_stride, sb.dig_pitch, sb.dig_stride, c, eq_r.p, eq_r.pitch, n, d_v);
              o.v_pin[g] = E.d2h_async(d_v, (size_t)c * TAU * D); E.dfree(d_v); }
            // compute_mz_mles / compute_u_s (decomposition.rs:214-256): u_s[k][j] = (M_j z_k)(r), z_k = x_s[k] || w_ccs_k, evaluated through
            // the transposed matrices on this rank's columns -- no Mz table per piece, no gather
            o.u_pin[g] = eval_with(ep, zl + (size_t)k0 * sb.zl_stride, sb.zl_pitch, sb.zl_stride, c);
            LF_CUDA(cudaEventCreateWithFlags(&o.ev[g], cudaEventDisableTiming)); LF_CUDA(cudaEventRecord(o.ev[g], E.st()));
        }
        E.dfree(ep.v);
        mark("dec.groups");
    }
    DecPending decompose_enqueue(const LCCCS& cm, const lf_witness* w, const DevVec& eq_r, StepBuffers& sb, int half) {
        DecPending o = decompose_enqueue_a(cm.x_w, cm.h, w, sb, half); decompose_enqueue_b(o, cm, eq_r, sb, half); return o;
    }
    DecOut decompose_finish(DecPending& pd, Transcript<Rg>& T) {
        DecOut o; const int K = P->K; const size_t kappa = P->kappa, t = P->t; const LCCCS& cm = pd.cm;
        o.x_s = std::move(pd.x_s);
        o.y_s.assign(K, HV(kappa * D, 0)); o.v_s.resize(K); o.u_s.resize(K);
        // the commitments are queued in part A, before any group: they have landed when the first group's event has
        const bool y_early = true;
        auto take_y = [&] {
            for (int k = 1; k < K; ++k) for (size_t i = 0; i < kappa; ++i) std::memcpy(&o.y_s[k][i * D], pd.y_pin + (i * (K - 1) + (k - 1)) * D, 8 * D);
            HV bsum(kappa * D, 0); const u64 bm = P->b % F::P;      // y_0 = cm - b (y_1 + b (y_2 + ...))
            for (int k = K - 1; k >= 1; --k) for (size_t i = 0; i < kappa * D; ++i) bsum[i] = F::mul(F::add(bsum[i], o.y_s[k][i]), bm);
            for (size_t i = 0; i < kappa * D; ++i) o.y_s[0][i] = F::sub(cm.cm[i], bsum[i]);
        };
        if (!y_early) { LF_CUDA(cudaEventSynchronize(pd.ev[pd.ngroups - 1])); take_y(); }
        for (int g = 0; g < pd.ngroups; ++g) {
            LF_CUDA(cudaEventSynchronize(pd.ev[g])); cudaEventDestroy(pd.ev[g]); pd.ev[g] = nullptr;
            if (g == 0 && y_early) take_y();
            const int k0 = pd.g0[g], c = pd.g0[g + 1] - k0;
            auto t0 = std::chrono::steady_clock::now();
            for (int k = k0; k < k0 + c; ++k) {
                o.v_s[k].assign(pd.v_pin[g] + (size_t)(k - k0) * TAU * D, pd.v_pin[g] + (size_t)(k - k0 + 1) * TAU * D);
                o.u_s[k].resize(t * D); for (size_t j = 0; j < t; ++j) std::memcpy(&o.u_s[k][j * D], pd.u_pin[g] + (j * c + (k - k0)) * D, 8 * D);      // [t][c] -> per piece
                const HV& x = o.x_s[k];
                T.absorb_slice(x.data(), cnt(x)); T.absorb_slice(o.y_s[k].data(), kappa); T.absorb_slice(o.u_s[k].data(), cnt(o.u_s[k])); T.absorb_slice(o.v_s[k].data(), cnt(o.v_s[k]));
                if (x.empty()) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
                LCCCS L; L.r = cm.r; L.v = o.v_s[k]; L.cm = o.y_s[k]; L.u = o.u_s[k]; L.x_w.assign(x.begin(), x.end() - D); L.h.assign(x.end() - D, x.end());
                o.lc.push_back(std::move(L));
            }
            P->timings[3] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        return o;
    }

    // ------------------------------------------------------------------ folding (folding.rs:42-130)
    struct FoldOut { HV msgs; std::vector<HV> theta, eta; LCCCS lc; W* f0 = nullptr; };
    // rot_lin_combination (cyclotomic-rings/src/rotation.rs:45-104), host: d^2 base-by-slot-field products per term
    static HV rot_lin_combination(const std::vector<El>& rho_coeff, const std::vector<HV>& theta) {
        HV acc((size_t)D * TAU, 0);
        for (size_t i = 0; i < rho_coeff.size(); ++i) {
            if (theta[i].size() != (size_t)TAU * D) throw LfException(LF_ERR_INCORRECT_LENGTH, "rot_sum: b.len() != dimension");
            El a = rho_coeff[i];
            for (int bi = 0; bi < D; ++bi) {          // flatten_to_coeffs: element-major, slot-minor == memory order
                const u64* Bv = &theta[i][(size_t)bi * TAU];
                for (int j = 0; j < D; ++j) for (int l = 0; l < TAU; ++l) acc[(size_t)j * TAU + l] = F::add(acc[(size_t)j * TAU + l], F::mul(a[j], Bv[l]));
                HR::mul_x(a);
            }
        }
        return acc;
    }
    FoldOut fold(const std::vector<LCCCS>& lcs, StepBuffers& sb, const DevVec& eq_acc, const DevVec& eq_new, Transcript<Rg>& T) {
        FoldOut o; const int K = P->K, s = (int)P->s; const size_t n = nl(), m = ml(), t = P->t;
        if ((int)lcs.size() != 2 * K) throw LfException(LF_ERR_INCORRECT_LENGTH, "IncorrectLength");
        // squeeze_alpha_beta_zeta_mu (folding/utils.rs:51-96)
        // alpha and zeta first: the G tables below need only these two, so their kernels run while the host squeezes mu and beta
        std::vector<u64> alpha = squeeze(T, "alpha_s", 2 * K), zeta = squeeze(T, "zeta_s", 2 * K);
        // dense tables [eq(r_acc), G_acc, eq(r_new), G_new, eq(beta)]  (create_sumcheck_polynomial, folding/utils.rs:200-259)
        lf_sumcheck sc; sc.ctx = P->ctx; sc.nv = s; sc.deg = 2 * (int)P->b; sc.kind = LF_COMB_FOLD; sc.len = m; sc.sharded = world() > 1;
        SumcheckDriver<Rg> drv(P->ctx, &sc);
        SumcheckDriver<Rg>::alloc_group(E, sc.dense, 5, m);
        constexpr size_t WB = sizeof(W);
        auto tbl = [&](int i) { return wp(sc.dense.cur) + (size_t)i * sc.dense.stride; };
        LF_CUDA(cudaMemcpy2DAsync(tbl(0), sc.dense.pitch * WB, eq_acc.p, eq_acc.pitch * WB, m * WB, D, cudaMemcpyDeviceToDevice, E.st()));
        LF_CUDA(cudaMemcpy2DAsync(tbl(2), sc.dense.pitch * WB, eq_new.p, eq_new.pitch * WB, m * WB, D, cudaMemcpyDeviceToDevice, E.st()));
        for (int half = 0; half < 2; ++half) {
            W* G = tbl(1 + 2 * half);
            LF_CUDA(cudaMemsetAsync(G, 0, sc.dense.stride * WB, E.st()));
            // sum_i Horner_{alpha_i}(f-hat_i[tau-1..0]) = sum_i sum_d alpha_i^{d+1} f-hat_{i,d}   (folding/utils.rs:524-546)
            std::vector<u64> wts((size_t)K * TAU * TAU);
            for (int i = 0; i < K; ++i) { const u64* a = &alpha[(size_t)(half * K + i) * TAU]; u64 pw[TAU]; std::memcpy(pw, a, 8 * TAU);
                for (int dd = 0; dd < TAU; ++dd) { std::memcpy(&wts[((size_t)i * TAU + dd) * TAU], pw, 8 * TAU); SF::mul(pw, pw, a); } }
            u64* d_w = E.template dalloc<u64>(wts.size());
            E.h2d(d_w, wts.data(), wts.size() * 8);
            E.launch("k_digit_lincomb", [&] { k_digit_lincomb<Rg><<<dim3(Engine<Rg>::blocks_for(n, 128), S), 128, 0, E.st()>>>(sb.dig + (size_t)half * K * sb.dig_stride, sb.dig_pitch, sb.dig_stride, K, d_w, G, sc.dense.pitch, n, 0); });
            E.dfree(d_w);
            // + sum_i Horner_{zeta_i}(Mz_i[t-1..0]) = sum_j M_j (sum_i zeta_i^(j+1) z_i)   (calculate_challenged_mz_mle, folding.rs:208-226;
            // slot-wise scalars commute with the sparse product): the z vectors are combined first, on every rank's own columns, and
            // only the t combined vectors are gathered for the row owners' sparse products
            { const size_t zs = sb.zl_stride, zp = sb.zl_pitch; W* Z = E.template dalloc<W>(t * zs);
              std::vector<HV> zhead(t, HV(hc() * D, 0));
              for (size_t j = 0; j < t; ++j) {
                  HV coef((size_t)K * D); PL pl;
                  for (int i = 0; i < K; ++i) { const u64* z = &zeta[(size_t)(half * K + i) * TAU]; u64 pw[TAU]; std::memcpy(pw, z, 8 * TAU);
                      for (size_t e = 0; e < j; ++e) SF::mul(pw, pw, z);
                      El ce = HR::from_sf(pw); std::memcpy(&coef[(size_t)i * D], ce.data(), 8 * D);
                      pl.p[i] = sb.zl + (size_t)(half * K + i) * zs; pl.len[i] = zcols();
                      const LCCCS& L = lcs[half * K + i]; HV xh = L.x_w; xh.insert(xh.end(), L.h.begin(), L.h.end());
                      for (size_t e = 0; e < hc() && e < cnt(xh); ++e) { El pr = HR::mul(HR::load(&xh[e * D]), ce); for (int l = 0; l < D; ++l) zhead[j][e * D + l] = F::add(zhead[j][e * D + l], pr[l]); } }
                  E.lincomb(pl, zp, K, coef.data(), Z + j * zs, zp, zcols(), false);
              }
              W* all = gather_wccs(Z, t * zs);
              const size_t hp = pitch_of(hc()); W* d_head = E.template dalloc<W>(t * hp * D);
              for (size_t j = 0; j < t; ++j) {
                  if (P->M[j]->eff_rows > sc.dense.pitch) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength");
                  E.upload_small(zhead[j].data(), hc(), d_head + j * hp * D, hp);
                  E.spmv(P->M[j], d_head + j * hp * D, hc(), hp, all + j * zs + hc(), zp, G, sc.dense.pitch, P->M[j]->eff_rows,
                         world() == 1 ? ~(size_t)0 : Wl(), t * zs, 1, 0, 0, 0, true);
              }
              E.dfree(d_head); if (all != Z) E.dfree(all); E.dfree(Z); }
        }
        std::vector<u64> mu = squeeze(T, "mu_s", 2 * K - 1);
        { u64 one[TAU] = {0}; one[0] = 1; mu.insert(mu.end(), one, one + TAU); }
        HV beta = sf_to_ring(squeeze(T, "beta_s", s));
        mark("fold.challenges");
        E.eq_table(beta.data(), s, tbl(4), sc.dense.pitch, (size_t)rank() * m, m);
        mark("fold.tables");
        sc.dig = sb.dig; sc.dig_pitch = sb.dig_pitch; sc.dig_stride = sb.dig_stride;
        { HV mu_ring = sf_to_ring(mu); drv.set_mu(mu_ring.data(), 2 * K); }
        std::vector<u64> point; HV finals;
        o.msgs = run_sumcheck(drv, T, point, &finals);
        drv.free_all();
        mark("fold.sumcheck");
        HV r0 = sf_to_ring(point);
        // theta_i = f-hat_i(r_0): the sumcheck's fully folded f-hat tables (get_thetas, folding.rs:236-246)
        for (int i = 0; i < 2 * K; ++i) o.theta.emplace_back(finals.begin() + (size_t)(5 + i * TAU) * D, finals.begin() + (size_t)(5 + (i + 1) * TAU) * D);
        // eta_i = Mz_i(r_0) (get_etas, folding.rs:248-256)
        // queued without waiting: the host absorbs the thetas while the device evaluates the etas
        const u64* eta_land = eval_z_async(sb.zl, sb.zl_pitch, sb.zl_stride, 2 * K, r0);
        auto t0 = std::chrono::steady_clock::now();
        for (auto& th : o.theta) T.absorb_slice(th.data(), cnt(th));
        E.sync();
        for (int i = 0; i < 2 * K; ++i) { HV et(t * D); for (size_t j = 0; j < t; ++j) std::memcpy(&et[j * D], eta_land + (j * 2 * K + i) * D, 8 * D); o.eta.push_back(std::move(et)); }
        mark("fold.eta");
        for (auto& et : o.eta) T.absorb_slice(et.data(), cnt(et));
        // get_rhos (folding/utils.rs:116-131)
        T.absorb_tag("rho_s");
        std::vector<El> rho_coeff; HV rho((size_t)2 * K * D);
        for (int i = 0; i < 2 * K - 1; ++i) { El cfs; T.get_short_challenge(cfs.data()); rho_coeff.push_back(cfs); }
        { El one = HR::zero(); one[0] = 1; rho_coeff.push_back(one); }
        for (int i = 0; i < 2 * K; ++i) { El r = H.crt(rho_coeff[i]); std::memcpy(&rho[(size_t)i * D], r.data(), 8 * D); }
        P->timings[3] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        mark("fold.absorb_rho");
        // compute_f_0 = sum_i rho_i f_i (folding.rs:258-268)
        o.f0 = E.template dalloc<W>(pitch_of(n) * D);
        { PL pl; for (int i = 0; i < 2 * K; ++i) { pl.p[i] = sb.pieces + (size_t)i * sb.pc_stride; pl.len[i] = n; }
          if (2 * K > MAX_LIST) throw LfException(LF_ERR_UNSUPPORTED, "2K exceeds MAX_LIST");
          E.lincomb(pl, sb.pc_pitch, 2 * K, rho.data(), o.f0, pitch_of(n), n, false); }
        mark("fold.f0");
        // compute_v0_u0_x0_cm_0 (folding/utils.rs:460-521): host
        o.lc.r = r0; o.lc.v = rot_lin_combination(rho_coeff, o.theta);
        const size_t kappa = cnt(lcs[0].cm);
        o.lc.cm.assign(kappa * D, 0); o.lc.u.assign(t * D, 0); HV x0((P->l + 1) * D, 0);
        auto axpy = [&](HV& acc, size_t e, const u64* v, const El& r) { El p = HR::mul(HR::load(v), r); for (int l = 0; l < D; ++l) acc[e * D + l] = F::add(acc[e * D + l], p[l]); };
        for (int i = 0; i < 2 * K; ++i) {
            El r = HR::load(&rho[(size_t)i * D]);
            for (size_t e = 0; e < kappa && e < cnt(lcs[i].cm); ++e) axpy(o.lc.cm, e, &lcs[i].cm[e * D], r);
            for (size_t e = 0; e < t && e < cnt(o.eta[i]); ++e) axpy(o.lc.u, e, &o.eta[i][e * D], r);
            HV xh = lcs[i].x_w; xh.insert(xh.end(), lcs[i].h.begin(), lcs[i].h.end());
            for (size_t e = 0; e < P->l + 1 && e < cnt(xh); ++e) axpy(x0, e, &xh[e * D], r);
        }
        o.lc.h.assign(x0.end() - D, x0.end()); o.lc.x_w.assign(x0.begin(), x0.end() - D);
        return o;
    }

    lf_ctx* aux_ctx() {
        if (P->aux) return P->aux;
        lf_ctx* m = P->ctx; std::unique_ptr<lf_ctx> a(new lf_ctx);
        a->ring = m->ring; a->device = m->device; a->tables = m->tables; a->shared_tables = true;
        for (int i = 0; i < 2; ++i) { a->d_tab_idx[i] = m->d_tab_idx[i]; a->d_tab_val[i] = m->d_tab_val[i]; }
        LF_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
        LF_CUDA(cudaMalloc(&a->d_err, sizeof(int))); LF_CUDA(cudaMemset(a->d_err, 0, sizeof(int)));
        if (m->world > 1 && m->nccl && m->xg.on) {      // sharded: same rank / communicator, channel 1 of the peer-memory mailboxes
            a->rank = m->rank; a->world = m->world; a->nccl = m->nccl; a->xg.on = true; a->xg.cap = m->xg.cap; a->xg.parent = &m->xg;
            for (int r = 0; r < m->world; ++r) { unsigned char* b = m->xg.base[r] + Engine<Rg>::XG_CHANNEL_BYTES;
                a->xg.flags[r] = (unsigned long long*)b; a->xg.inbox[r] = (u64*)(b + Engine<Rg>::XG_FLAG_BYTES); }
        }
        P->aux = a.release(); return P->aux;
    }
    // ------------------------------------------------------------------ NIFSProver::prove (nifs.rs:48-103)
    static void put(u64*& p, const HV& v) { std::memcpy(p, v.data(), 8 * v.size()); p += v.size(); }
    static void put_lcccs(u64* p, const LCCCS& L) { put(p, L.r); put(p, L.v); put(p, L.cm); put(p, L.u); put(p, L.x_w); put(p, L.h); }
    static u64 proof_words_of(const lf_problem& P) {
        const u64 d = D, tau = TAU;
        return P.s * (P.d + 2) * d + tau * d + P.t * d + 2 * (u64)P.K * ((P.l + 1) + P.kappa + P.t + tau) * d + P.s * (2 * P.b + 1) * d + 2 * (u64)P.K * (tau + P.t) * d;
    }
    static HV load_canonical(const u64* p, size_t n, const char* what) {
        if (!p && n) throw LfException(LF_ERR_INVALID_ARG, std::string(what) + " is NULL");
        for (size_t i = 0; i < n * D; ++i) if (p[i] >= F::P) throw LfException(LF_ERR_INVALID_ARG, std::string("non-canonical field element in ") + what);
        return HV(p, p + n * D);
    }
    static LCCCS load_acc(const lf_problem& in, const lf_prover* P) {
        LCCCS a; auto ld = [&](const u64* p, size_t n) { return load_canonical(p, n, "the accumulator"); };
        a.r = ld(in.acc_r, P->s); a.v = ld(in.acc_v, TAU); a.cm = ld(in.acc_cm, P->kappa); a.u = ld(in.acc_u, P->t); a.x_w = ld(in.acc_x_w, P->l); a.h = ld(in.acc_h, 1); return a;
    }
    // derive_out = false: the caller only downloads the folded f (host-buffer entry point) -- its coefficient form and w_ccs are not derived
    lf_witness* prove(const lf_problem& in, const lf_witness* w_acc, const lf_witness* w_i, Transcript<Rg>& T, u64* out_proof, u64* out_lcccs, bool presynced = false, bool derive_out = true) {
        using clk = std::chrono::steady_clock; auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        for (double& x : P->timings) x = 0;
        if (!presynced) { E.sync(); E.arena_reset(); }      // presynced: the caller did both before queueing the witness uploads
        P->detail = std::getenv("LF_TIMING_DETAIL") != nullptr; P->marks.clear(); P->last_mark = clk::now();
        auto t_begin = clk::now();
        sanity_check();
        const int K = P->K; const size_t n = nl();
        LCCCS acc = load_acc(in, P);
        HV cm_i_cm = load_canonical(in.cm_i_cm, P->kappa, "cm_i"), x_ccs = load_canonical(in.cm_i_x_ccs, P->l, "x_ccs");
        // Schedule: the accumulator's decomposition depends on nothing the transcript produces, so its device half is
        // queued first on the auxiliary stream and runs beside the (latency-bound, host-paced) linearization sumcheck.
        StepBuffers sb;
        DevVec eq_acc; LinOut lin; DecPending pl, prr;
        // whatever a throw leaves behind (step buffers, eq tables, pending events) is released here
        struct Cleanup { Prover* self; StepBuffers* sb; DevVec* eq_acc; LinOut* lin; DecPending *a, *b; bool armed = true;
            ~Cleanup() { if (!armed) return; try { self->E.sync(); } catch (...) {}
                for (DecPending* d : {a, b}) for (cudaEvent_t& e : d->ev) if (e) { cudaEventDestroy(e); e = nullptr; }
                self->E.dfree(sb->dig); self->E.dfree(sb->pieces); self->E.dfree(sb->zl); self->E.dfree(eq_acc->p); self->E.dfree(lin->eq_r.p); } } cleanup{this, &sb, &eq_acc, &lin, &pl, &prr};
        sb.dig_pitch = (std::max(n, ml()) + 255) / 256 * 256;   // the sumcheck walks all 2^s entries
        sb.dig_stride = sb.dig_pitch * D; sb.dig = E.template dalloc<int8_t>((size_t)2 * K * sb.dig_stride);
        sb.pc_pitch = pitch_of(n); sb.pc_stride = sb.pc_pitch * D; sb.pieces = E.template dalloc<W>((size_t)2 * K * sb.pc_stride);
        sb.zl_pitch = pitch_of(zcols()); sb.zl_stride = sb.zl_pitch * D; sb.zl = E.template dalloc<W>((size_t)2 * K * sb.zl_stride);
        eq_acc.n = ((size_t)1 << cnt(acc.r)) / world(); eq_acc.pitch = pitch_of(eq_acc.n); eq_acc.p = E.template dalloc<W>(eq_acc.pitch * D);
        // sharded steps overlap too when the collectives are stream-ordered on both streams (own NCCL communicator + mailbox channels)
        const bool overlap = (world() == 1 || (E.c->nccl && E.c->xg.on)) && !P->detail && !std::getenv("LF_NO_OVERLAP");
        W* lin_tail = (overlap && world() > 1) ? gather_wccs(wp(w_i->w_ccs), w_i->w_pitch * D) : nullptr;
        const bool early_a = overlap && !std::getenv("LF_NO_EARLY_DEC");      // (A/B switch for measurements)
        {
            lf_ctx* main_ctx = E.c;
            if (overlap) {
                lf_ctx* aux = aux_ctx();
                if (P->acc_ready) { LF_CUDA(cudaStreamWaitEvent(aux->stream, P->acc_ready, 0)); }
                else { cudaEvent_t ready; LF_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming)); LF_CUDA(cudaEventRecord(ready, main_ctx->stream));
                       LF_CUDA(cudaStreamWaitEvent(aux->stream, ready, 0)); cudaEventDestroy(ready); }
                aux->arena_off = 0; aux->profiling = main_ctx->profiling;
                E.c = aux;
            }
            try {
                // on the stream that runs the decompositions: when the accumulator's decomposition starts behind `acc_ready`
                // (host-buffer entry point) it does NOT wait for what the main stream queues after that event
                LF_CUDA(cudaMemsetAsync(sb.dig, 0, (size_t)2 * K * sb.dig_stride, E.st()));   // f-hat tables are zero on [n, 2^s)
                E.eq_table(acc.r.data(), (int)cnt(acc.r), eq_acc.p, eq_acc.pitch, (size_t)rank() * eq_acc.n, eq_acc.n);
                pl = decompose_enqueue(acc, w_acc, eq_acc, sb, 0);
                // the incoming witness's digits, commitments and recomposed pieces do not wait for the linearization either
                if (early_a) {
                    // (the incoming witness may still be uploading on the main stream: everything queued there so far comes first)
                    cudaEvent_t up; LF_CUDA(cudaEventCreateWithFlags(&up, cudaEventDisableTiming)); LF_CUDA(cudaEventRecord(up, main_ctx->stream));
                    LF_CUDA(cudaStreamWaitEvent(E.st(), up, 0)); cudaEventDestroy(up);
                    HV h_one; { El one = HR::from_u64(1); h_one.assign(one.begin(), one.end()); } prr = decompose_enqueue_a(x_ccs, h_one, w_i, sb, 1);
                }
            } catch (...) { E.c = main_ctx; throw; }
            E.c = main_ctx;
        }
        mark("alloc+dec_acc_enqueue");
        // absorb_public_input (nifs.rs:175-197): after the device has been given its first work
        T.absorb_tag("acc");
        T.absorb_slice(acc.r.data(), cnt(acc.r)); T.absorb_slice(acc.v.data(), cnt(acc.v)); T.absorb_slice(acc.cm.data(), cnt(acc.cm));
        T.absorb_slice(acc.u.data(), cnt(acc.u)); T.absorb_slice(acc.x_w.data(), cnt(acc.x_w)); T.absorb(acc.h.data());
        T.absorb_tag("cm_i"); T.absorb_slice(cm_i_cm.data(), cnt(cm_i_cm)); T.absorb_slice(x_ccs.data(), cnt(x_ccs));
        auto t0 = clk::now();
        lin = linearize(cm_i_cm, x_ccs, w_i, T, lin_tail);
        mark("linearize");
        E.sync(); auto t1 = clk::now(); P->timings[0] = ms(t0, t1);
        // The second decomposition queues behind the first on the auxiliary stream (the linearization's device work is complete:
        // synchronised above), so the accumulator's results -- the ones the transcript absorbs first -- are never delayed by it.
        { lf_ctx* main_ctx = E.c;
          if (overlap) { E.c = aux_ctx(); E.c->profiling = main_ctx->profiling; }
          try { if (early_a) decompose_enqueue_b(prr, lin.lc, lin.eq_r, sb, 1); else prr = decompose_enqueue(lin.lc, w_i, lin.eq_r, sb, 1); } catch (...) { E.c = main_ctx; throw; }
          E.c = main_ctx; }
        DecOut dl = decompose_finish(pl, T);
        mark("decompose_acc");
        DecOut dr = decompose_finish(prr, T);
        mark("decompose_new");
        if (P->aux) {      // fold the auxiliary context's bookkeeping (launch count, event pairs, digit-overflow flag) into the main one
            lf_ctx* m = E.c; lf_ctx* a = P->aux;
            m->launches += a->launches; a->launches = 0; m->collectives += a->collectives; a->collectives = 0;
            for (auto& r : a->prof) m->prof.push_back(r); a->prof.clear();
            E.c = a; try { E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_witness: a coefficient does not fit K digits of base b"); } catch (...) { E.c = m; throw; } E.c = m;
        }
        E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_witness: a coefficient does not fit K digits of base b");
        E.sync(); auto t2 = clk::now(); P->timings[1] = ms(t1, t2);
        std::vector<LCCCS> lcs = dl.lc; lcs.insert(lcs.end(), dr.lc.begin(), dr.lc.end());
        FoldOut fo = fold(lcs, sb, eq_acc, lin.eq_r, T);
        mark("fold.host_tail");
        lf_witness* w_out;
        if (derive_out) w_out = witness_from_f_device(fo.f0);
        else { w_out = new lf_witness; w_out->n = nl(); w_out->pitch = pitch_of(nl()); w_out->f = ow(fo.f0); w_out->W = Wl(); w_out->w_pitch = pitch_of(w_out->W); }
        cleanup.armed = false;
        E.dfree(sb.dig); E.dfree(sb.pieces); E.dfree(sb.zl); E.dfree(eq_acc.p); E.dfree(lin.eq_r.p);
        E.sync(); auto t3 = clk::now(); P->timings[2] = ms(t2, t3);
        mark("witness_out_free");
        // serialise: lin{msgs,v,u} | dec_acc | dec_new | fold{msgs,theta,eta}
        u64* p = out_proof;
        put(p, lin.msgs); put(p, lin.lc.v); put(p, lin.lc.u);
        for (const DecOut* dd : {&dl, &dr}) for (int k = 0; k < K; ++k) { put(p, dd->x_s[k]); put(p, dd->y_s[k]); put(p, dd->u_s[k]); put(p, dd->v_s[k]); }
        put(p, fo.msgs); for (auto& v : fo.theta) put(p, v); for (auto& v : fo.eta) put(p, v);
        put_lcccs(out_lcccs, fo.lc);
        P->timings[4] = ms(t_begin, clk::now());
        return w_out;
    }
};

}  // namespace lf
