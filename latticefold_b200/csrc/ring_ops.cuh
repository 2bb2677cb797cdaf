// Ring-dependent half of the C ABI behind a small virtual interface, so that each ring's kernels are instantiated in their
// own translation unit (ring_goldilocks.cu, ring_babybear.cu, ring_frog.cu build in parallel) and lf_b200.cu only dispatches
// on the context's ring id.  Methods throw LfException; the extern "C" wrappers turn that into status codes.
#pragma once
#include "prover.cuh"
#include "verifier_host.hpp"
#include "wire_host.hpp"

struct lf_transcript { int ring; void* impl; };      // impl = lf::Transcript<Rg>*

namespace lf {

struct RingOps {
    virtual ~RingOps() {}
    virtual void describe(lf_ring_info* out) = 0;
    virtual void ctx_tables_create(lf_ctx* c) = 0;
    virtual void ctx_tables_destroy(lf_ctx* c) = 0;
    virtual void vec_free(lf_ctx* c, lf_vec* v) = 0;
    virtual void sumcheck_free(lf_sumcheck* sc) = 0;
    virtual void witness_free(lf_prover* p, lf_witness* w) = 0;
    virtual uint64_t proof_words(const lf_problem* P) = 0;
    virtual uint64_t lcccs_words(const lf_problem* P) = 0;
    virtual uint64_t wire_bytes(const lf_problem* P) = 0;
    virtual void wire_serialize(const lf_problem* P, const uint64_t* w, uint8_t* out) = 0;
    virtual void wire_deserialize(const lf_problem* P, const uint8_t* in, uint64_t n, uint64_t* w) = 0;
    virtual void* tr_new() = 0;
    virtual void* tr_clone(const void* t) = 0;
    virtual void tr_free(void* t) = 0;
    virtual void tr_absorb(void* t, const uint64_t* els, size_t count) = 0;
    virtual void tr_absorb_base(void* t, const uint64_t* limbs, size_t count) = 0;
    virtual void tr_absorb_tag(void* t, const char* tag) = 0;
    virtual void tr_get_challenge(void* t, uint64_t* out) = 0;
    virtual void tr_get_short_challenge(void* t, uint64_t* out) = 0;
    virtual uint64_t tr_permutations(const void* t) = 0;
    virtual void vec_upload(lf_ctx* c, const uint64_t* host, size_t n, int32_t form, lf_vec** out) = 0;
    virtual void vec_download(lf_ctx* c, const lf_vec* v, uint64_t* host) = 0;
    virtual void crt(lf_ctx* c, const lf_vec* in, lf_vec** out) = 0;
    virtual void icrt(lf_ctx* c, const lf_vec* in, lf_vec** out) = 0;
    virtual void gadget_decompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) = 0;
    virtual void gadget_recompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) = 0;
    virtual void decompose_to_vec(lf_ctx* c, const lf_vec* in, uint64_t b, int32_t K, lf_vec** out_k) = 0;
    virtual void fhat(lf_ctx* c, const lf_vec* in, lf_vec** out_tau) = 0;
    virtual void ajtai_create(lf_ctx* c, size_t kappa, size_t n, const uint64_t* host, lf_ajtai** out) = 0;
    virtual void commit_batch(lf_ctx* c, const lf_ajtai* a, const lf_vec* const* f, int32_t count, uint64_t* out_host) = 0;
    virtual void commit_coeff(lf_ctx* c, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t* out_host) = 0;
    virtual void decompose_and_commit(lf_ctx* c, const lf_ajtai* a, const lf_vec* v, bool ntt_form, uint64_t B, int32_t L, uint64_t* out_host) = 0;
    virtual void commit_pieces(lf_ctx* c, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t b, int32_t K, uint64_t* out_host) = 0;
    virtual void sparse_create(lf_ctx* c, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col, const uint64_t* val, lf_sparse** out) = 0;
    virtual void spmv(lf_ctx* c, const lf_sparse* m, const lf_vec* z, lf_vec** out) = 0;
    virtual void eq_table(lf_ctx* c, const uint64_t* r, int32_t s, lf_vec** out) = 0;
    virtual void mle_eval_batch(lf_ctx* c, const lf_vec* const* mles, int32_t count, int32_t nv, const uint64_t* point, int32_t point_len, uint64_t* out_host) = 0;
    virtual void lincomb(lf_ctx* c, const uint64_t* coeffs, const lf_vec* const* vecs, int32_t count, lf_vec** out) = 0;
    virtual void sumcheck_begin(lf_ctx* c, lf_vec** mles, int32_t M, int32_t nv, int32_t degree, const lf_comb* comb, lf_sumcheck** out) = 0;
    virtual void sumcheck_round(lf_sumcheck* sc, const uint64_t* prev, uint64_t* out_evals) = 0;
    virtual void sumcheck_finish(lf_sumcheck* sc, const uint64_t* last, uint64_t* out_final) = 0;
    virtual void rot_lin_combination(const uint64_t* rho, const uint64_t* theta, int32_t count, uint64_t* out) = 0;
    virtual void prover_create(lf_ctx* c, const lf_problem* sh, lf_prover** out) = 0;
    virtual void prover_upload_witness(lf_prover* p, const uint64_t* f_host, lf_witness** out) = 0;
    virtual void witness_download_f(lf_prover* p, const lf_witness* w, uint64_t* f_host) = 0;
    virtual void witness_f_from_w_ccs(lf_ctx* c, const uint64_t* w_ccs, size_t W, uint64_t B, int32_t L, uint64_t* f_host) = 0;
    virtual void linearize(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) = 0;
    virtual void nifs_prove_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_acc, const lf_witness* w_i, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, lf_witness** out_w) = 0;
    virtual void nifs_prove(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, uint64_t* out_f) = 0;
    virtual void nifs_verify(const lf_problem* in, lf_transcript* t, const uint64_t* proof, uint64_t* out_lcccs) = 0;
    virtual void linearization_verify(const lf_problem* in, lf_transcript* t, const uint64_t* lin_proof, uint64_t* out_lcccs) = 0;
    virtual void linearize_resident(lf_prover* p, const lf_problem* in, const lf_witness* w, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) = 0;
    virtual void witness_commit(lf_prover* p, const lf_witness* w, uint64_t* out_host) = 0;
};
RingOps* ring_ops_goldilocks();
RingOps* ring_ops_babybear();
RingOps* ring_ops_frog();

template <class Rg> struct RingOpsImpl final : RingOps {
    typedef typename Rg::W W; typedef PtrListT<W> PL;
    static W* wp(lf_words* p) { return reinterpret_cast<W*>(p); }
    static const W* wp(const lf_words* p) { return reinterpret_cast<const W*>(p); }
    static lf_words* ow(W* p) { return reinterpret_cast<lf_words*>(p); }
    static Transcript<Rg>& tr(lf_transcript* t) { if (t->ring != Rg::ID) throw LfException(LF_ERR_INVALID_ARG, "transcript ring differs from the context's ring"); return *(Transcript<Rg>*)t->impl; }
    void describe(lf_ring_info* out) override { out->p = Rg::F::P; out->d = Rg::D; out->n_slots = Rg::S; out->tau = Rg::TAU; out->nu = Rg::F::NU; }
    void ctx_tables_create(lf_ctx* c) override {
        auto* tab = new RingTables<Rg>(); c->tables = tab;
        const int (*idx[2])[Rg::S] = {tab->crt_idx, tab->icrt_idx}; const u64 (*val[2])[Rg::S] = {tab->crt_val, tab->icrt_val};
        for (int i = 0; i < 2; ++i) {
            LF_CUDA(cudaMalloc(&c->d_tab_idx[i], sizeof(int) * Rg::D * Rg::S)); LF_CUDA(cudaMalloc(&c->d_tab_val[i], 8 * Rg::D * Rg::S));
            LF_CUDA(cudaMemcpy(c->d_tab_idx[i], idx[i], sizeof(int) * Rg::D * Rg::S, cudaMemcpyHostToDevice));
            LF_CUDA(cudaMemcpy(c->d_tab_val[i], val[i], 8 * Rg::D * Rg::S, cudaMemcpyHostToDevice));
        }
    }
    void ctx_tables_destroy(lf_ctx* c) override { delete (RingTables<Rg>*)c->tables; c->tables = nullptr; }
    void vec_free(lf_ctx* c, lf_vec* v) override { Engine<Rg> E(c); E.vec_free(v); }
    void sumcheck_free(lf_sumcheck* sc) override { SumcheckDriver<Rg> drv(sc->ctx, sc); drv.free_all(); delete sc; }
    void witness_free(lf_prover* p, lf_witness* w) override { Prover<Rg> pr(p); pr.free_witness(w); }
    uint64_t proof_words(const lf_problem* P) override { return Prover<Rg>::proof_words_of(*P); }
    uint64_t wire_bytes(const lf_problem* P) override { return Wire<Rg>::bytes(*P); }
    void wire_serialize(const lf_problem* P, const uint64_t* w, uint8_t* out) override { Wire<Rg>::serialize(*P, w, out); }
    void wire_deserialize(const lf_problem* P, const uint8_t* in, uint64_t n, uint64_t* w) override { Wire<Rg>::deserialize(*P, in, n, w); }
    uint64_t lcccs_words(const lf_problem* P) override { return (P->s + Rg::TAU + P->kappa + P->t + P->l + 1) * (u64)Rg::D; }
    void* tr_new() override { return new Transcript<Rg>(); }
    void* tr_clone(const void* t) override { return new Transcript<Rg>(*(const Transcript<Rg>*)t); }
    void tr_free(void* t) override { delete (Transcript<Rg>*)t; }
    void tr_absorb(void* t, const uint64_t* els, size_t count) override { ((Transcript<Rg>*)t)->absorb_slice(els, count); }
    void tr_absorb_base(void* t, const uint64_t* limbs, size_t count) override { ((Transcript<Rg>*)t)->absorb_base(limbs, count); }
    void tr_absorb_tag(void* t, const char* tag) override { ((Transcript<Rg>*)t)->absorb_tag(tag); }
    void tr_get_challenge(void* t, uint64_t* out) override { ((Transcript<Rg>*)t)->get_challenge(out); }
    void tr_get_short_challenge(void* t, uint64_t* out) override { ((Transcript<Rg>*)t)->get_short_challenge(out); }
    uint64_t tr_permutations(const void* t) override { return ((const Transcript<Rg>*)t)->permutations(); }

    void vec_upload(lf_ctx* c, const uint64_t* host, size_t n, int32_t form, lf_vec** out) override {
        Engine<Rg> E(c); lf_vec* v = E.vec_alloc(n, form); E.upload_planes(host, n, wp(v->p), v->pitch); E.sync(); *out = v;
    }
    void vec_download(lf_ctx* c, const lf_vec* v, uint64_t* host) override {
        Engine<Rg> E(c); E.download_planes(wp(v->p), v->pitch, v->n, host);
    }
    void crt(lf_ctx* c, const lf_vec* in, lf_vec** out) override {
        Engine<Rg> E(c); lf_vec* o = E.vec_alloc(in->n, LF_FORM_NTT); E.crt(wp(in->p), in->pitch, wp(o->p), o->pitch, in->n, false); *out = o;
    }
    void icrt(lf_ctx* c, const lf_vec* in, lf_vec** out) override {
        Engine<Rg> E(c); lf_vec* o = E.vec_alloc(in->n, LF_FORM_COEFF); E.crt(wp(in->p), in->pitch, wp(o->p), o->pitch, in->n, true); *out = o;
    }
    void gadget_decompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) override {
        Engine<Rg> E(c); lf_vec* o = E.vec_alloc(in->n * (size_t)L, LF_FORM_COEFF);
                          try { E.gadget_decompose(wp(in->p), in->pitch, wp(o->p), o->pitch, in->n, B, L); E.check_err_flag(LF_ERR_DOES_NOT_FIT, "gadget_decompose: a coefficient does not fit L digits of base B"); }
                          catch (...) { E.vec_free(o); throw; }
                          *out = o;
    }
    void gadget_recompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) override {
        if (L < 1 || in->n % (size_t)L) throw LfException(LF_ERR_INCORRECT_LENGTH, "gadget_recompose: length is not a multiple of L");
                          Engine<Rg> E(c); lf_vec* o = E.vec_alloc(in->n / L, in->form); E.gadget_recompose(wp(in->p), in->pitch, wp(o->p), o->pitch, o->n, B, L); *out = o;
    }
    void decompose_to_vec(lf_ctx* c, const lf_vec* in, uint64_t b, int32_t K, lf_vec** out_k) override {
        Engine<Rg> E(c); const size_t n = in->n, dp = (n + 255) / 256 * 256;
        int8_t* dig = E.template dalloc<int8_t>((size_t)K * Rg::D * dp);
        try { E.digit_split(wp(in->p), in->pitch, dig, dp, n, b, K); E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_to_vec: a coefficient does not fit K digits of base b"); }
        catch (...) { E.dfree(dig); throw; }
        for (int k = 0; k < K; ++k) { lf_vec* o = E.vec_alloc(n, LF_FORM_COEFF);
            if (n) { E.launch("k_digits_to_field", [&] { k_digits_to_field<Rg><<<Engine<Rg>::blocks_for(n * Rg::D), 256, 0, E.st()>>>(dig + (size_t)k * Rg::D * dp, dp, wp(o->p), o->pitch, n); }); }
            out_k[k] = o; }
        E.dfree(dig);
    }
    void fhat(lf_ctx* c, const lf_vec* in, lf_vec** out_tau) override {
        Engine<Rg> E(c); const size_t n = in->n;
        for (int j = 0; j < Rg::TAU; ++j) { lf_vec* o = E.vec_alloc(n, LF_FORM_NTT);
            if (n) { E.launch("k_fhat", [&] { k_fhat<Rg><<<Engine<Rg>::blocks_for(n * Rg::S), 256, 0, E.st()>>>(wp(in->p) + (size_t)j * Rg::S * in->pitch, in->pitch, wp(o->p), o->pitch, n); }); }
            out_tau[j] = o; }
    }
    void ajtai_create(lf_ctx* c, size_t kappa, size_t n, const uint64_t* host, lf_ajtai** out) override {
        Engine<Rg> E(c); std::unique_ptr<lf_ajtai> a(new lf_ajtai); a->kappa = kappa; a->n = n; a->pitch = pitch_of(n);
        LF_CUDA(cudaMalloc(&a->p, std::max<size_t>(1, kappa * a->pitch * Rg::D) * sizeof(W)));
        // row by row so the staging buffer stays small (the matrix is 1.3 GB at kappa=26, n=2^18)
        for (size_t i = 0; i < kappa; ++i) E.upload_planes(host + i * n * Rg::D, n, wp(a->p) + i * a->pitch * Rg::D, a->pitch);
        E.sync(); E.ajtai_build_tiles(a.get()); *out = a.release();
    }
    void commit_batch(lf_ctx* c, const lf_ajtai* a, const lf_vec* const* f, int32_t count, uint64_t* out_host) override {
        Engine<Rg> E(c);
        for (int i = 0; i < count; ++i) if (f[i]->n != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(f[i]->n) + ", " + std::to_string(a->n) + ")");
        for (int done = 0; done < count; done += MAX_LIST) {
            const int chunk = std::min(MAX_LIST, count - done);
            PL Y; for (int i = 0; i < chunk; ++i) { Y.p[i] = wp(f[done + i]->p); Y.len[i] = a->n; }
            u64* d_out = E.small_dev(a->kappa * chunk * Rg::D);
            E.dot(wp(a->p), a->pitch * Rg::D, a->pitch, (int)a->kappa, nullptr, Y, pitch_of(a->n), chunk, a->n, d_out);
            HV all(a->kappa * chunk * Rg::D); E.download_words(d_out, all.size(), all.data());
            for (int i = 0; i < chunk; ++i) for (size_t r = 0; r < a->kappa; ++r) std::memcpy(out_host + ((size_t)(done + i) * a->kappa + r) * Rg::D, &all[(r * chunk + i) * Rg::D], 8 * Rg::D);
        }
    }
    void commit_coeff(lf_ctx* c, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t* out_host) override {
        if (f_coeff->n != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(f_coeff->n) + ", " + std::to_string(a->n) + ")");
        Engine<Rg> E(c); lf_vec* f = E.vec_alloc(f_coeff->n, LF_FORM_NTT); E.crt(wp(f_coeff->p), f_coeff->pitch, wp(f->p), f->pitch, f->n, false);
        try { const lf_vec* fp = f; commit_batch(c, a, &fp, 1, out_host); } catch (...) { E.vec_free(f); throw; }
        E.vec_free(f);
    }
    void decompose_and_commit(lf_ctx* c, const lf_ajtai* a, const lf_vec* v, bool ntt_form, uint64_t B, int32_t L, uint64_t* out_host) override {
        if (L < 1 || v->n * (size_t)L != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(v->n * (size_t)std::max(L, 0)) + ", " + std::to_string(a->n) + ")");
        Engine<Rg> E(c); lf_vec* co = nullptr; lf_vec* dg = E.vec_alloc(a->n, LF_FORM_COEFF);
        try {
            const lf_vec* src = v;
            if (ntt_form) { co = E.vec_alloc(v->n, LF_FORM_COEFF); E.crt(wp(v->p), v->pitch, wp(co->p), co->pitch, v->n, true); src = co; }
            E.gadget_decompose(wp(src->p), src->pitch, wp(dg->p), dg->pitch, src->n, B, L);
            E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_and_commit: a coefficient does not fit L digits of base B");
            commit_coeff(c, a, dg, out_host);
        } catch (...) { E.vec_free(co); E.vec_free(dg); throw; }
        E.vec_free(co); E.vec_free(dg);
    }
    void commit_pieces(lf_ctx* c, const lf_ajtai* a, const lf_vec* f_coeff, uint64_t b, int32_t K, uint64_t* out_host) override {
        if (f_coeff->n != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(f_coeff->n) + ", " + std::to_string(a->n) + ")");
        Engine<Rg> E(c); const size_t n = a->n, dp = (n + 255) / 256 * 256, ds = dp * Rg::D, kappa = a->kappa;
        int8_t* dig = E.template dalloc<int8_t>((size_t)K * ds); W* pieces = nullptr;
        try {
            LF_CUDA(cudaMemsetAsync(dig, 0, (size_t)K * ds, E.st()));
            E.digit_split(wp(f_coeff->p), f_coeff->pitch, dig, dp, n, b, K);
            E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_to_vec: a coefficient does not fit K digits of base b");
            const bool mma = E.can_commit_digits(a, 1, dp);
            if (!mma) { pieces = E.template dalloc<W>((size_t)K * pitch_of(n) * Rg::D); E.crt_digits(dig, dp, pieces, pitch_of(n), n, K, ds, pitch_of(n) * Rg::D); }
            for (int done = 0; done < K; done += 16) {
                const int chunk = std::min(16, K - done);
                u64* d_out = E.small_dev(kappa * chunk * Rg::D);
                if (mma) E.commit_digits(a, dig + (size_t)done * ds, dp, ds, chunk, d_out);
                else { PL Y; for (int i = 0; i < chunk; ++i) { Y.p[i] = pieces + (size_t)(done + i) * pitch_of(n) * Rg::D; Y.len[i] = n; }
                       E.dot(wp(a->p), a->pitch * Rg::D, a->pitch, (int)kappa, nullptr, Y, pitch_of(n), chunk, n, d_out); }
                HV all(kappa * chunk * Rg::D); E.download_words(d_out, all.size(), all.data());
                for (int i = 0; i < chunk; ++i) for (size_t r = 0; r < kappa; ++r) std::memcpy(out_host + ((size_t)(done + i) * kappa + r) * Rg::D, &all[(r * chunk + i) * Rg::D], 8 * Rg::D);
            }
        } catch (...) { E.dfree(dig); E.dfree(pieces); throw; }
        E.dfree(dig); E.dfree(pieces);
    }
    void sparse_create(lf_ctx* c, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col, const uint64_t* val, lf_sparse** out) override {
        Engine<Rg> E(c); std::unique_ptr<lf_sparse> m(new lf_sparse); m->nrows = nrows; m->ncols = ncols; m->nnz = row_ptr[nrows];
        if (m->nnz >= ((u64)1 << 32) || nrows >= ((u64)1 << 32)) throw LfException(LF_ERR_UNSUPPORTED, "sparse matrix too large for 32-bit indices");
        std::vector<u32> rp(nrows + 1), cl(m->nnz); m->eff_rows = 1;
        for (size_t i = 0; i <= nrows; ++i) rp[i] = (u32)row_ptr[i];
        for (size_t i = 0; i < nrows; ++i) if (row_ptr[i + 1] > row_ptr[i]) m->eff_rows = i + 1;
        for (size_t i = 0; i < m->nnz; ++i) { if (col[i] >= ncols) throw LfException(LF_ERR_INVALID_ARG, "column index out of range"); cl[i] = (u32)col[i]; }
        LF_CUDA(cudaMalloc(&m->row_ptr, (nrows + 1) * 4)); LF_CUDA(cudaMalloc(&m->col, std::max<size_t>(1, m->nnz) * 4));
        LF_CUDA(cudaMemcpy(m->row_ptr, rp.data(), (nrows + 1) * 4, cudaMemcpyHostToDevice)); if (m->nnz) LF_CUDA(cudaMemcpy(m->col, cl.data(), m->nnz * 4, cudaMemcpyHostToDevice));
        m->val_pitch = pitch_of(m->nnz); LF_CUDA(cudaMalloc(&m->val, m->val_pitch * Rg::D * sizeof(W)));
        E.upload_planes(val, m->nnz, wp(m->val), m->val_pitch); E.sync(); *out = m.release();
    }
    void spmv(lf_ctx* c, const lf_sparse* m, const lf_vec* z, lf_vec** out) override {
        if (z->n != m->ncols) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
        Engine<Rg> E(c); lf_vec* o = E.vec_alloc(m->nrows, LF_FORM_NTT); E.spmv(m, wp(z->p), z->n, z->pitch, wp(z->p), z->pitch, wp(o->p), o->pitch, m->nrows); *out = o;
    }
    void eq_table(lf_ctx* c, const uint64_t* r, int32_t s, lf_vec** out) override {
        if (s < 1 || s > 34) throw LfException(LF_ERR_INVALID_ARG, "r length is 0 (or too large)"); Engine<Rg> E(c); lf_vec* o = E.vec_alloc((size_t)1 << s, LF_FORM_NTT);
                          try { E.eq_table(r, s, wp(o->p), o->pitch); } catch (...) { E.vec_free(o); throw; } *out = o;
    }
    void mle_eval_batch(lf_ctx* c, const lf_vec* const* mles, int32_t count, int32_t nv, const uint64_t* point, int32_t point_len, uint64_t* out_host) override {
        if (point_len != nv) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: point length != num_vars");
        for (int i = 0; i < count; ++i) if (mles[i]->n > ((size_t)1 << nv)) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: MLE longer than 2^num_vars");
        Engine<Rg> E(c); const size_t n = (size_t)1 << nv, ep = pitch_of(n);
        W* eq = E.template dalloc<W>(ep * Rg::D); E.eq_table(point, nv, eq, ep);
        // the dot kernel treats the MLEs as columns against the eq table as the single row
        for (int done = 0; done < count; done += MAX_LIST) {
            const int chunk = std::min(MAX_LIST, count - done);
            PL Y; size_t pitch = 0;
            for (int i = 0; i < chunk; ++i) { Y.p[i] = wp(mles[done + i]->p); Y.len[i] = mles[done + i]->n; if (i && mles[done + i]->pitch != pitch) throw LfException(LF_ERR_INVALID_ARG, "MLEs of one batch must have equal length"); pitch = mles[done + i]->pitch; }
            u64* d_out = E.small_dev((size_t)chunk * Rg::D);
            E.dot(eq, 0, ep, 1, nullptr, Y, pitch, chunk, n, d_out);
            E.download_words(d_out, (size_t)chunk * Rg::D, out_host + (size_t)done * Rg::D);
        }
        E.dfree(eq);
    }
    void lincomb(lf_ctx* c, const uint64_t* coeffs, const lf_vec* const* vecs, int32_t count, lf_vec** out) override {
        if (count < 1) throw LfException(LF_ERR_INVALID_ARG, "lincomb of nothing");
        for (int i = 1; i < count; ++i) if (vecs[i]->n != vecs[0]->n) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual");
        Engine<Rg> E(c); lf_vec* o = E.vec_alloc(vecs[0]->n, vecs[0]->form);
        for (int done = 0; done < count; done += MAX_LIST) { const int chunk = std::min(MAX_LIST, count - done); PL pl; for (int i = 0; i < chunk; ++i) { pl.p[i] = wp(vecs[done + i]->p); pl.len[i] = o->n; }
            E.lincomb(pl, o->pitch, chunk, coeffs + (size_t)done * Rg::D, wp(o->p), o->pitch, o->n, done > 0); }
        *out = o;
    }
    void sumcheck_begin(lf_ctx* c, lf_vec** mles, int32_t M, int32_t nv, int32_t degree, const lf_comb* comb, lf_sumcheck** out) override {
        if (nv < 1) throw LfException(LF_ERR_SUMCHECK_MISUSE, "Attempt to prove a constant.");
        if (M < 1) throw LfException(LF_ERR_INVALID_ARG, "no MLEs");
        const size_t n = (size_t)1 << nv;
        for (int i = 0; i < M; ++i) if (mles[i]->n > n) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: MLE longer than 2^num_vars");
        Engine<Rg> E(c); std::unique_ptr<lf_sumcheck> sc(new lf_sumcheck); sc->ctx = c; sc->nv = nv; sc->deg = degree; sc->kind = comb->kind; sc->len = n;
        SumcheckDriver<Rg> drv(c, sc.get());
        int n_dense = M;
        if (comb->kind == LF_COMB_FOLD) {
            if (comb->b != 2) throw LfException(LF_ERR_UNSUPPORTED, "FOLD kernels are specialised for b = 2");
            if (degree != 4 || M != 5 + comb->n_mu * Rg::TAU) throw LfException(LF_ERR_INVALID_ARG, "FOLD: need degree 2b and 5 + n_mu*tau MLEs");
            n_dense = 5;
        } else {
            if (degree > SC_MAX_DEG || comb->n_terms < 1) throw LfException(LF_ERR_UNSUPPORTED, "PRODUCTS/LIN: degree at most 7, at least one term");
        }
        auto fill = [&](lf_sumcheck::Group& g, int first, int count) {
            SumcheckDriver<Rg>::alloc_group(E, g, count, n);
            constexpr size_t WB = sizeof(W);
            LF_CUDA(cudaMemsetAsync(g.cur, 0, (size_t)count * g.stride * WB, E.st()));
            for (int k = 0; k < count; ++k) { const lf_vec* v = mles[first + k]; if (v->n) LF_CUDA(cudaMemcpy2DAsync(wp(g.cur) + (size_t)k * g.stride, g.pitch * WB, v->p, v->pitch * WB, v->n * WB, Rg::D, cudaMemcpyDeviceToDevice, E.st())); }
        };
        fill(sc->dense, 0, n_dense);
        if (comb->kind == LF_COMB_FOLD) { fill(sc->fh, 5, M - 5); drv.set_mu(comb->mu_host, comb->n_mu); }
        else {
            std::vector<std::vector<int>> terms; int o = 0;
            for (int t = 0; t < comb->n_terms; ++t) { if (comb->idx_len[t] < 0) throw LfException(LF_ERR_INVALID_ARG, "negative term length"); terms.emplace_back(comb->idx + o, comb->idx + o + comb->idx_len[t]); o += comb->idx_len[t]; }
            drv.set_terms(M, degree, comb->kind == LF_COMB_LIN, terms);
            sc->d_coef = E.template dalloc<u64>((size_t)comb->n_terms * Rg::D);
            LF_CUDA(cudaMemcpyAsync(sc->d_coef, comb->coef_host, (size_t)comb->n_terms * Rg::D * 8, cudaMemcpyHostToDevice, E.st())); E.sync();
        }
        for (int i = 0; i < M; ++i) { E.vec_free(mles[i]); mles[i] = nullptr; }   // ownership taken, as by the reference's Vec<DenseMultilinearExtension>
        *out = sc.release();
    }
    void sumcheck_round(lf_sumcheck* sc, const uint64_t* prev, uint64_t* out_evals) override {
        SumcheckDriver<Rg> drv(sc->ctx, sc);
        if (prev) drv.apply_challenge(prev); else if (sc->round > 0) throw LfException(LF_ERR_SUMCHECK_MISUSE, "verifier message is empty");
        drv.evaluate(out_evals);
    }
    void sumcheck_finish(lf_sumcheck* sc, const uint64_t* last, uint64_t* out_final) override {
        SumcheckDriver<Rg> drv(sc->ctx, sc); if (sc->round != sc->nv) throw LfException(LF_ERR_SUMCHECK_MISUSE, "sumcheck not finished"); drv.apply_challenge(last); drv.final_values(out_final);
    }
    void rot_lin_combination(const uint64_t* rho, const uint64_t* theta, int32_t count, uint64_t* out) override {
        std::vector<typename HostRing<Rg>::El> r(count); std::vector<HV> th(count);
        for (int i = 0; i < count; ++i) { r[i] = HostRing<Rg>::load(rho + (size_t)i * Rg::D); th[i].assign(theta + (size_t)i * Rg::TAU * Rg::D, theta + (size_t)(i + 1) * Rg::TAU * Rg::D); }
        HV o = Prover<Rg>::rot_lin_combination(r, th); std::memcpy(out, o.data(), 8 * o.size());
    }
    void prover_create(lf_ctx* c, const lf_problem* sh, lf_prover** out) override {
        if (sh->ring != Rg::ID || c->ring != Rg::ID) throw LfException(LF_ERR_INVALID_ARG, "problem ring differs from the context's ring");
        if (sh->B_hi || sh->B_lo >= ((u64)1 << 62)) throw LfException(LF_ERR_UNSUPPORTED, "B >= 2^62");
        std::unique_ptr<lf_prover> p(new lf_prover); p->ctx = c; p->ring = sh->ring; p->L = sh->L; p->K = sh->K; p->B = sh->B_lo; p->b = sh->b;
        p->kappa = sh->kappa; p->n = sh->n; p->m = sh->m; p->n_ccs = sh->n_ccs; p->l = sh->l; p->t = sh->t; p->q = sh->q; p->d = sh->d; p->s = sh->s;
        int o = 0; for (u64 i = 0; i < sh->q; ++i) { p->S.emplace_back(sh->S_flat + o, sh->S_flat + o + sh->S_len[i]); o += sh->S_len[i]; }
        p->c.assign(sh->c, sh->c + sh->q * Rg::D);
        // limits of the step's kernels, checked here rather than deep inside a step (a throw on one rank of a sharded run would
        // desynchronise the peer-memory mailboxes)
        if (sh->b != 2) throw LfException(LF_ERR_UNSUPPORTED, "folding sumcheck kernels are specialised for b = 2 (every reference parameter set but Stark)");
        if (sh->K < 1 || 2 * sh->K > MAX_LIST) throw LfException(LF_ERR_UNSUPPORTED, "need 1 <= K <= " + std::to_string(MAX_LIST / 2));
        if (2 * sh->K * Rg::TAU > MAX_MU) throw LfException(LF_ERR_UNSUPPORTED, "2 K tau exceeds MAX_MU");
        if (sh->L < 1 || sh->L > 64) throw LfException(LF_ERR_UNSUPPORTED, "need 1 <= L <= 64");
        if (!sh->A) throw LfException(LF_ERR_INVALID_ARG, "Ajtai matrix is NULL");
        // sharded context: the caller passes this rank's column slice of A (kappa x n/world); CCS matrices are given whole
        // and cut to this rank's row slab here
        const size_t G = (size_t)c->world, n_loc = sh->n / G, m_loc = sh->m / G;
        if (sh->n % G || sh->m % G) throw LfException(LF_ERR_UNSUPPORTED, "n and m must be multiples of the rank count");
        ajtai_create(c, sh->kappa, n_loc, sh->A, &p->A);
        for (u64 j = 0; j < sh->t; ++j) {
            const lf_csr& M = sh->M[j]; if (M.nrows != sh->m) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "CCS matrix rows != m");
            const size_t r0 = (size_t)c->rank * m_loc; const u64 e0 = M.row_ptr[r0];
            std::vector<u64> rp(m_loc + 1); for (size_t i = 0; i <= m_loc; ++i) rp[i] = M.row_ptr[r0 + i] - e0;
            lf_sparse* m = nullptr; sparse_create(c, m_loc, M.ncols, rp.data(), M.col + e0, M.val + e0 * Rg::D, &m); p->M.push_back(m);
            // transposed image of the rank's columns (k_csc_eq): the l+1 head columns on every rank + the columns of its witness slice, all rows
            const size_t hc = sh->l + 1, W_all = sh->L ? sh->n / sh->L : 0, W_loc = W_all / G;
            if (M.ncols == hc + W_all && W_all % G == 0 && M.row_ptr[sh->m] < ((u64)1 << 32)) {
                const size_t c0 = hc + (size_t)c->rank * W_loc, nloc = hc + W_loc;
                auto local = [&](u64 col) -> long long { return col < hc ? (long long)col : (col >= c0 && col < c0 + W_loc ? (long long)(hc + (col - c0)) : -1); };
                std::vector<u32> cp(nloc + 1, 0);
                for (u64 r = 0; r < sh->m; ++r) for (u64 e = M.row_ptr[r]; e < M.row_ptr[r + 1]; ++e) { const long long lc = local(M.col[e]); if (lc >= 0) ++cp[lc + 1]; }
                for (size_t i = 0; i < nloc; ++i) cp[i + 1] += cp[i];
                const size_t tn = cp[nloc]; std::vector<u32> fill(cp.begin(), cp.end() - 1), tr(std::max<size_t>(tn, 1)); HV tv(std::max<size_t>(tn, 1) * Rg::D);
                for (u64 r = 0; r < sh->m; ++r) for (u64 e = M.row_ptr[r]; e < M.row_ptr[r + 1]; ++e) { const long long lc = local(M.col[e]); if (lc < 0) continue;
                    const u32 pos = fill[lc]++; tr[pos] = (u32)r; std::memcpy(&tv[(size_t)pos * Rg::D], M.val + e * Rg::D, 8 * Rg::D); }
                Engine<Rg> E(c);
                LF_CUDA(cudaMalloc(&m->t_col_ptr, (nloc + 1) * 4)); LF_CUDA(cudaMalloc(&m->t_row, std::max<size_t>(1, tn) * 4));
                LF_CUDA(cudaMemcpy(m->t_col_ptr, cp.data(), (nloc + 1) * 4, cudaMemcpyHostToDevice)); if (tn) LF_CUDA(cudaMemcpy(m->t_row, tr.data(), tn * 4, cudaMemcpyHostToDevice));
                m->t_val_pitch = pitch_of(tn); LF_CUDA(cudaMalloc(&m->t_val, m->t_val_pitch * Rg::D * sizeof(W)));
                E.upload_planes(tv.data(), tn, wp(m->t_val), m->t_val_pitch); E.sync();
                m->t_ncols = nloc; m->t_nnz = tn;
            } }
        *out = p.release();
    }
    void prover_upload_witness(lf_prover* p, const uint64_t* f_host, lf_witness** out) override {
        Prover<Rg> pr(p); *out = pr.upload_witness(f_host);
    }
    void witness_download_f(lf_prover* p, const lf_witness* w, uint64_t* f_host) override {
        Engine<Rg> E(p->ctx); E.download_planes(wp(w->f), w->pitch, w->n, f_host);
    }
    void witness_f_from_w_ccs(lf_ctx* c, const uint64_t* w_ccs, size_t W, uint64_t B, int32_t L, uint64_t* f_host) override {
        Engine<Rg> E(c); const size_t wpt = pitch_of(W), n = W * (size_t)L, np = pitch_of(n);
        typename Rg::W *w = E.template dalloc<typename Rg::W>(wpt * Rg::D), *wc = E.template dalloc<typename Rg::W>(wpt * Rg::D), *fc = E.template dalloc<typename Rg::W>(np * Rg::D), *f = E.template dalloc<typename Rg::W>(np * Rg::D);
        E.upload_planes(w_ccs, W, w, wpt); E.crt(w, wpt, wc, wpt, W, true); E.gadget_decompose(wc, wpt, fc, np, W, B, L); E.crt(fc, np, f, np, n, false);
        E.check_err_flag(LF_ERR_DOES_NOT_FIT, "from_w_ccs: a coefficient does not fit L digits of base B");
        E.download_planes(f, np, n, f_host); E.dfree(w); E.dfree(wc); E.dfree(fc); E.dfree(f);
    }
    void linearize(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) override {
        Prover<Rg> pr(p); lf_witness* w = pr.upload_witness(in->w_i_f);
        HV cm(in->cm_i_cm, in->cm_i_cm + p->kappa * Rg::D), x(in->cm_i_x_ccs, in->cm_i_x_ccs + p->l * Rg::D);
        auto lo = pr.linearize(cm, x, w, tr(t)); pr.E.dfree(lo.eq_r.p); pr.free_witness(w);
        Prover<Rg>::put_lcccs(out_lcccs, lo.lc);
        if (out_proof) { u64* q = out_proof; Prover<Rg>::put(q, lo.msgs); Prover<Rg>::put(q, lo.lc.v); Prover<Rg>::put(q, lo.lc.u); }
    }
    void linearize_resident(lf_prover* p, const lf_problem* in, const lf_witness* w, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) override {
        Prover<Rg> pr(p); pr.E.sync(); pr.E.arena_reset();
        HV cm = Prover<Rg>::load_canonical(in->cm_i_cm, p->kappa, "cm_i"), x = Prover<Rg>::load_canonical(in->cm_i_x_ccs, p->l, "x_ccs");
        auto lo = pr.linearize(cm, x, w, tr(t)); pr.E.dfree(lo.eq_r.p);
        Prover<Rg>::put_lcccs(out_lcccs, lo.lc);
        if (out_proof) { u64* q = out_proof; Prover<Rg>::put(q, lo.msgs); Prover<Rg>::put(q, lo.lc.v); Prover<Rg>::put(q, lo.lc.u); }
    }
    void witness_commit(lf_prover* p, const lf_witness* w, uint64_t* out_host) override {
        Engine<Rg> E(p->ctx); const lf_ajtai* a = p->A;
        if (w->n != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(w->n) + ", " + std::to_string(a->n) + ")");
        PL Y; Y.p[0] = wp(w->f); Y.len[0] = a->n;
        u64* d_out = E.small_dev(a->kappa * Rg::D);
        E.dot(wp(a->p), a->pitch * Rg::D, a->pitch, (int)a->kappa, nullptr, Y, w->pitch, 1, a->n, d_out);
        E.download_words(d_out, a->kappa * Rg::D, out_host);
    }
    void linearization_verify(const lf_problem* in, lf_transcript* t, const uint64_t* lin_proof, uint64_t* out_lcccs) override {
        Verifier<Rg> v(*in); LCCCS lc = v.verify_linearization_only(lin_proof, tr(t)); if (out_lcccs) Prover<Rg>::put_lcccs(out_lcccs, lc);
    }
    void nifs_prove_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_acc, const lf_witness* w_i, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, lf_witness** out_w) override {
        Prover<Rg> pr(p); lf_witness* w = pr.prove(*in, w_acc, w_i, tr(t), out_proof, out_lcccs); if (out_w) *out_w = w; else pr.free_witness(w);
    }
    void nifs_verify(const lf_problem* in, lf_transcript* t, const uint64_t* proof, uint64_t* out_lcccs) override {
        Verifier<Rg> v(*in); LCCCS lc = v.verify(proof, tr(t)); if (out_lcccs) Prover<Rg>::put_lcccs(out_lcccs, lc);
    }
    void nifs_prove(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, uint64_t* out_f) override {
        Prover<Rg> pr(p);
        pr.E.sync(); pr.E.arena_reset();
        lf_witness* wa = pr.upload_witness(in->w_acc_f);
        if (!p->acc_ready) LF_CUDA(cudaEventCreateWithFlags(&p->acc_ready, cudaEventDisableTiming));
        LF_CUDA(cudaEventRecord(p->acc_ready, pr.E.st()));
        lf_witness* wi = pr.upload_witness(in->w_i_f);
        lf_witness* w = nullptr;
        cudaEvent_t ev = p->acc_ready;
        try { w = pr.prove(*in, wa, wi, tr(t), out_proof, out_lcccs, true, false); } catch (...) { p->acc_ready = nullptr; cudaEventDestroy(ev); pr.free_witness(wa); pr.free_witness(wi); throw; }
        p->acc_ready = nullptr; cudaEventDestroy(ev);
        if (out_f) pr.E.download_planes(wp(w->f), w->pitch, w->n, out_f);
        pr.free_witness(w); pr.free_witness(wa); pr.free_witness(wi); pr.E.sync();
    }
};

}  // namespace lf
