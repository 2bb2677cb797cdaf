// extern "C" surface of liblf_b200.so (declared in include/lf_b200.h).  No torch types, no CPU fallback: every compute
// entry point launches sm_100a kernels on the context's stream or fails with LF_ERR_CUDA.
#include "prover.cuh"

using namespace lf;

struct lf_transcript { int ring; Transcript<GoldilocksRing> g; };

namespace {
thread_local std::string g_create_err;

template <class Fn> lf_status guard(lf_ctx* ctx, Fn&& fn) {
    try { fn(); return LF_OK; }
    catch (const LfException& e) { if (ctx) ctx->err = e.what(); else g_create_err = e.what(); return e.code; }
    catch (const std::exception& e) { if (ctx) ctx->err = e.what(); else g_create_err = e.what(); return LF_ERR_INVALID_ARG; }
}
void need_goldilocks(int ring) { if (ring != LF_RING_GOLDILOCKS) throw LfException(LF_ERR_UNSUPPORTED, "this build implements the Goldilocks ring; BabyBear / Frog descriptors are not compiled in yet"); }
typedef GoldilocksRing G;
typedef Engine<G> Eng;
constexpr int D = G::D, TAU = G::TAU;
}  // namespace

extern "C" {

lf_status lf_ring_describe(int32_t ring_id, lf_ring_info* out) {
    return guard(nullptr, [&] { need_goldilocks(ring_id); out->p = Goldilocks::P; out->d = G::D; out->n_slots = G::S; out->tau = G::TAU; out->nu = (u64)1 << Goldilocks::NU_SHIFT; });
}
const char* lf_last_error(const lf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

lf_status lf_ctx_create(int32_t ring_id, int32_t device, lf_ctx** out) {
    *out = nullptr;
    return guard(nullptr, [&] {
        need_goldilocks(ring_id);
        int ndev = 0; cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw LfException(LF_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
        if (device < 0 || device >= ndev) throw LfException(LF_ERR_INVALID_ARG, "device index out of range");
        LF_CUDA(cudaSetDevice(device));
        std::unique_ptr<lf_ctx> c(new lf_ctx); c->ring = ring_id; c->device = device;
        LF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        cudaMemPool_t pool; LF_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX; LF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));   // keep freed blocks: steady-state steps allocate nothing
        auto* tab = new RingTables<G>(); c->tables = tab;
        const int (*idx[2])[G::S] = {tab->crt_idx, tab->icrt_idx}; const u64 (*val[2])[G::S] = {tab->crt_val, tab->icrt_val};
        for (int i = 0; i < 2; ++i) {
            LF_CUDA(cudaMalloc(&c->d_tab_idx[i], sizeof(int) * D * G::S)); LF_CUDA(cudaMalloc(&c->d_tab_val[i], 8 * D * G::S));
            LF_CUDA(cudaMemcpy(c->d_tab_idx[i], idx[i], sizeof(int) * D * G::S, cudaMemcpyHostToDevice));
            LF_CUDA(cudaMemcpy(c->d_tab_val[i], val[i], 8 * D * G::S, cudaMemcpyHostToDevice));
        }
        LF_CUDA(cudaMalloc(&c->d_err, sizeof(int))); LF_CUDA(cudaMemset(c->d_err, 0, sizeof(int)));
        *out = c.release();
    });
}
void lf_ctx_destroy(lf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device); cudaStreamSynchronize(c->stream);
    if (c->nccl) NcclApi::get().CommDestroy(c->nccl);
    for (auto& kv : c->block_size) cudaFree(kv.first);
    if (!c->shared_tables) for (int i = 0; i < 2; ++i) { cudaFree(c->d_tab_idx[i]); cudaFree(c->d_tab_val[i]); }
    cudaFree(c->d_err); cudaFree(c->d_small); cudaFree(c->d_partial); if (c->h_pinned) cudaFreeHost(c->h_pinned); if (c->h_arena) cudaFreeHost(c->h_arena);
    if (!c->shared_tables) delete (RingTables<G>*)c->tables;
    cudaStreamDestroy(c->stream); delete c;
}
lf_status lf_ctx_sync(lf_ctx* c) { return guard(c, [&] { LF_CUDA(cudaStreamSynchronize(c->stream)); }); }
void* lf_ctx_stream(lf_ctx* c) { return (void*)c->stream; }
uint64_t lf_ctx_launches(const lf_ctx* c) { return c->launches; }
lf_status lf_ctx_set_shard(lf_ctx* c, int32_t rank, int32_t world, lf_collective_fn fn, void* user) {
    return guard(c, [&] { if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) throw LfException(LF_ERR_INVALID_ARG, "bad rank / world / collective");
                          c->rank = rank; c->world = world; c->coll = fn; c->coll_user = user; });
}
uint64_t lf_ctx_collectives(const lf_ctx* c) { return c->collectives; }
lf_status lf_nccl_unique_id(uint8_t* out128) {
    return guard(nullptr, [&] { NcclApi& n = NcclApi::get(); if (!n.ok()) throw LfException(LF_ERR_UNSUPPORTED, "libnccl.so.2 not found");
                                NcclApi::UniqueId id; if (n.GetUniqueId(&id) != 0) throw LfException(LF_ERR_CUDA, "ncclGetUniqueId failed"); std::memcpy(out128, id.internal, 128); });
}
lf_status lf_ctx_set_shard_nccl(lf_ctx* c, int32_t rank, int32_t world, const uint8_t* id128) {
    return guard(c, [&] { if (world < 1 || rank < 0 || rank >= world) throw LfException(LF_ERR_INVALID_ARG, "bad rank / world");
                          NcclApi& n = NcclApi::get(); if (!n.ok()) throw LfException(LF_ERR_UNSUPPORTED, "libnccl.so.2 not found");
                          LF_CUDA(cudaSetDevice(c->device));
                          NcclApi::UniqueId id; std::memcpy(id.internal, id128, 128);
                          int rc = n.CommInitRank(&c->nccl, world, id, rank);
                          if (rc != 0) throw LfException(LF_ERR_CUDA, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
                          c->rank = rank; c->world = world; });
}
lf_status lf_ctx_profile(lf_ctx* c, int32_t enable) {
    return guard(c, [&] { LF_CUDA(cudaStreamSynchronize(c->stream)); for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } c->prof.clear(); c->profiling = enable != 0; });
}
// aggregates the event pairs recorded since lf_ctx_profile(ctx, 1): one line "name count total_ms" per kernel
lf_status lf_ctx_profile_report(lf_ctx* c, char* buf, size_t buf_len) {
    return guard(c, [&] {
        LF_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<std::pair<std::string, std::pair<int, double>>> agg;
        for (auto& r : c->prof) { float ms = 0; LF_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
            bool found = false; for (auto& a : agg) if (a.first == r.name) { a.second.first++; a.second.second += ms; found = true; break; }
            if (!found) agg.push_back({r.name, {1, (double)ms}}); }
        std::string out; for (auto& a : agg) out += a.first + " " + std::to_string(a.second.first) + " " + std::to_string(a.second.second) + "\n";
        if (out.size() + 1 > buf_len) throw LfException(LF_ERR_INVALID_ARG, "profile buffer too small");
        std::memcpy(buf, out.c_str(), out.size() + 1);
    });
}

// ---- vectors
lf_status lf_vec_upload(lf_ctx* c, const uint64_t* host, size_t n, int32_t form, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { Eng E(c); lf_vec* v = E.vec_alloc(n, form); E.upload_planes(host, n, v->p, v->pitch); E.sync(); *out = v; });
}
lf_status lf_vec_download(lf_ctx* c, const lf_vec* v, uint64_t* host) { return guard(c, [&] { Eng E(c); E.download_planes(v->p, v->pitch, v->n, host); }); }
size_t lf_vec_len(const lf_vec* v) { return v->n; }
int32_t lf_vec_form(const lf_vec* v) { return v->form; }
void lf_vec_free(lf_ctx* c, lf_vec* v) { if (v) { Eng E(c); E.vec_free(v); } }

// ---- CRT / ICRT
lf_status lf_crt(lf_ctx* c, const lf_vec* in, lf_vec** out) {
    *out = nullptr; return guard(c, [&] { Eng E(c); lf_vec* o = E.vec_alloc(in->n, LF_FORM_NTT); E.crt(in->p, in->pitch, o->p, o->pitch, in->n, false); *out = o; });
}
lf_status lf_icrt(lf_ctx* c, const lf_vec* in, lf_vec** out) {
    *out = nullptr; return guard(c, [&] { Eng E(c); lf_vec* o = E.vec_alloc(in->n, LF_FORM_COEFF); E.crt(in->p, in->pitch, o->p, o->pitch, in->n, true); *out = o; });
}
// ---- decompositions
lf_status lf_gadget_decompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { Eng E(c); lf_vec* o = E.vec_alloc(in->n * (size_t)L, LF_FORM_COEFF);
                          try { E.gadget_decompose(in->p, in->pitch, o->p, o->pitch, in->n, B, L); E.check_err_flag(LF_ERR_DOES_NOT_FIT, "gadget_decompose: a coefficient does not fit L digits of base B"); }
                          catch (...) { E.vec_free(o); throw; }
                          *out = o; });
}
lf_status lf_gadget_recompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { if (L < 1 || in->n % (size_t)L) throw LfException(LF_ERR_INCORRECT_LENGTH, "gadget_recompose: length is not a multiple of L");
                          Eng E(c); lf_vec* o = E.vec_alloc(in->n / L, in->form); E.gadget_recompose(in->p, in->pitch, o->p, o->pitch, o->n, B, L); *out = o; });
}
lf_status lf_decompose_to_vec(lf_ctx* c, const lf_vec* in, uint64_t b, int32_t K, lf_vec** out_k) {
    return guard(c, [&] {
        Eng E(c); const size_t n = in->n, dp = (n + 255) / 256 * 256;
        int8_t* dig = E.dalloc<int8_t>((size_t)K * D * dp);
        try { E.digit_split(in->p, in->pitch, dig, dp, n, b, K); E.check_err_flag(LF_ERR_DOES_NOT_FIT, "decompose_to_vec: a coefficient does not fit K digits of base b"); }
        catch (...) { E.dfree(dig); throw; }
        for (int k = 0; k < K; ++k) { lf_vec* o = E.vec_alloc(n, LF_FORM_COEFF);
            if (n) { E.launch("k_digits_to_field", [&] { k_digits_to_field<G><<<Eng::blocks_for(n * D), 256, 0, E.st()>>>(dig + (size_t)k * D * dp, dp, o->p, o->pitch, n); }); }
            out_k[k] = o; }
        E.dfree(dig);
    });
}
lf_status lf_fhat(lf_ctx* c, const lf_vec* in, lf_vec** out_tau) {
    return guard(c, [&] { Eng E(c); const size_t n = in->n;
        for (int j = 0; j < TAU; ++j) { lf_vec* o = E.vec_alloc(n, LF_FORM_NTT);
            if (n) { E.launch("k_fhat", [&] { k_fhat<G><<<Eng::blocks_for(n * G::S), 256, 0, E.st()>>>(in->p + (size_t)j * G::S * in->pitch, in->pitch, o->p, o->pitch, n); }); }
            out_tau[j] = o; } });
}

// ---- Ajtai
lf_status lf_ajtai_create(lf_ctx* c, size_t kappa, size_t n, const uint64_t* host, lf_ajtai** out) {
    *out = nullptr;
    return guard(c, [&] { Eng E(c); std::unique_ptr<lf_ajtai> a(new lf_ajtai); a->kappa = kappa; a->n = n; a->pitch = pitch_of(n);
        LF_CUDA(cudaMalloc(&a->p, std::max<size_t>(1, kappa * a->pitch * D) * 8));
        // row by row so the staging buffer stays small (the matrix is 1.3 GB at kappa=26, n=2^18)
        for (size_t i = 0; i < kappa; ++i) E.upload_planes(host + i * n * D, n, a->p + i * a->pitch * D, a->pitch);
        E.sync(); *out = a.release(); });
}
void lf_ajtai_free(lf_ctx* c, lf_ajtai* a) { if (a) { cudaStreamSynchronize(c->stream); cudaFree(a->p); delete a; } }
size_t lf_ajtai_kappa(const lf_ajtai* a) { return a->kappa; }
size_t lf_ajtai_width(const lf_ajtai* a) { return a->n; }
lf_status lf_commit_batch(lf_ctx* c, const lf_ajtai* a, const lf_vec* const* f, int32_t count, uint64_t* out_host) {
    return guard(c, [&] {
        Eng E(c);
        for (int i = 0; i < count; ++i) if (f[i]->n != a->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "WrongWitnessLength(" + std::to_string(f[i]->n) + ", " + std::to_string(a->n) + ")");
        for (int done = 0; done < count; done += MAX_LIST) {
            const int chunk = std::min(MAX_LIST, count - done);
            PtrList Y; for (int i = 0; i < chunk; ++i) { Y.p[i] = f[done + i]->p; Y.len[i] = a->n; }
            u64* d_out = E.small_dev(a->kappa * chunk * D);
            E.dot(a->p, a->pitch * D, a->pitch, (int)a->kappa, nullptr, Y, pitch_of(a->n), chunk, a->n, d_out);
            HV all(a->kappa * chunk * D); E.download_words(d_out, all.size(), all.data());
            for (int i = 0; i < chunk; ++i) for (size_t r = 0; r < a->kappa; ++r) std::memcpy(out_host + ((size_t)(done + i) * a->kappa + r) * D, &all[(r * chunk + i) * D], 8 * D);
        }
    });
}
lf_status lf_commit(lf_ctx* c, const lf_ajtai* a, const lf_vec* f, uint64_t* out_host) { return lf_commit_batch(c, a, &f, 1, out_host); }

// ---- sparse
lf_status lf_sparse_create(lf_ctx* c, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col, const uint64_t* val, lf_sparse** out) {
    *out = nullptr;
    return guard(c, [&] { Eng E(c); std::unique_ptr<lf_sparse> m(new lf_sparse); m->nrows = nrows; m->ncols = ncols; m->nnz = row_ptr[nrows];
        if (m->nnz >= ((u64)1 << 32) || nrows >= ((u64)1 << 32)) throw LfException(LF_ERR_UNSUPPORTED, "sparse matrix too large for 32-bit indices");
        std::vector<u32> rp(nrows + 1), cl(m->nnz); m->eff_rows = 1;
        for (size_t i = 0; i <= nrows; ++i) rp[i] = (u32)row_ptr[i];
        for (size_t i = 0; i < nrows; ++i) if (row_ptr[i + 1] > row_ptr[i]) m->eff_rows = i + 1;
        for (size_t i = 0; i < m->nnz; ++i) { if (col[i] >= ncols) throw LfException(LF_ERR_INVALID_ARG, "column index out of range"); cl[i] = (u32)col[i]; }
        LF_CUDA(cudaMalloc(&m->row_ptr, (nrows + 1) * 4)); LF_CUDA(cudaMalloc(&m->col, std::max<size_t>(1, m->nnz) * 4));
        LF_CUDA(cudaMemcpy(m->row_ptr, rp.data(), (nrows + 1) * 4, cudaMemcpyHostToDevice)); if (m->nnz) LF_CUDA(cudaMemcpy(m->col, cl.data(), m->nnz * 4, cudaMemcpyHostToDevice));
        m->val_pitch = pitch_of(m->nnz); LF_CUDA(cudaMalloc(&m->val, m->val_pitch * D * 8));
        E.upload_planes(val, m->nnz, m->val, m->val_pitch); E.sync(); *out = m.release(); });
}
void lf_sparse_free(lf_ctx* c, lf_sparse* m) { if (m) { cudaStreamSynchronize(c->stream); cudaFree(m->row_ptr); cudaFree(m->col); cudaFree(m->val); delete m; } }
lf_status lf_spmv(lf_ctx* c, const lf_sparse* m, const lf_vec* z, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { if (z->n != m->ncols) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual(M, z)");
        Eng E(c); lf_vec* o = E.vec_alloc(m->nrows, LF_FORM_NTT); E.spmv(m, z->p, z->n, z->pitch, z->p, z->pitch, o->p, o->pitch, m->nrows); *out = o; });
}
// ---- eq table / MLE evaluation / lincomb
lf_status lf_eq_table(lf_ctx* c, const uint64_t* r, int32_t s, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { if (s < 1 || s > 34) throw LfException(LF_ERR_INVALID_ARG, "r length is 0 (or too large)"); Eng E(c); lf_vec* o = E.vec_alloc((size_t)1 << s, LF_FORM_NTT);
                          try { E.eq_table(r, s, o->p, o->pitch); } catch (...) { E.vec_free(o); throw; } *out = o; });
}
lf_status lf_mle_eval_batch(lf_ctx* c, const lf_vec* const* mles, int32_t count, int32_t nv, const uint64_t* point, int32_t point_len, uint64_t* out_host) {
    return guard(c, [&] {
        if (point_len != nv) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: point length != num_vars");
        for (int i = 0; i < count; ++i) if (mles[i]->n > ((size_t)1 << nv)) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: MLE longer than 2^num_vars");
        Eng E(c); const size_t n = (size_t)1 << nv, ep = pitch_of(n);
        u64* eq = E.dalloc<u64>(ep * D); E.eq_table(point, nv, eq, ep);
        // the dot kernel treats the MLEs as columns against the eq table as the single row
        for (int done = 0; done < count; done += MAX_LIST) {
            const int chunk = std::min(MAX_LIST, count - done);
            PtrList Y; size_t pitch = 0;
            for (int i = 0; i < chunk; ++i) { Y.p[i] = mles[done + i]->p; Y.len[i] = mles[done + i]->n; if (i && mles[done + i]->pitch != pitch) throw LfException(LF_ERR_INVALID_ARG, "MLEs of one batch must have equal length"); pitch = mles[done + i]->pitch; }
            u64* d_out = E.small_dev((size_t)chunk * D);
            E.dot(eq, 0, ep, 1, nullptr, Y, pitch, chunk, n, d_out);
            E.download_words(d_out, (size_t)chunk * D, out_host + (size_t)done * D);
        }
        E.dfree(eq);
    });
}
lf_status lf_lincomb(lf_ctx* c, const uint64_t* coeffs, const lf_vec* const* vecs, int32_t count, lf_vec** out) {
    *out = nullptr;
    return guard(c, [&] { if (count < 1) throw LfException(LF_ERR_INVALID_ARG, "lincomb of nothing");
        for (int i = 1; i < count; ++i) if (vecs[i]->n != vecs[0]->n) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "LengthsNotEqual");
        Eng E(c); lf_vec* o = E.vec_alloc(vecs[0]->n, vecs[0]->form);
        for (int done = 0; done < count; done += MAX_LIST) { const int chunk = std::min(MAX_LIST, count - done); PtrList pl; for (int i = 0; i < chunk; ++i) { pl.p[i] = vecs[done + i]->p; pl.len[i] = o->n; }
            E.lincomb(pl, o->pitch, chunk, coeffs + (size_t)done * D, o->p, o->pitch, o->n, done > 0); }
        *out = o; });
}

// ---- sumcheck
lf_status lf_sumcheck_begin(lf_ctx* c, lf_vec** mles, int32_t M, int32_t nv, int32_t degree, const lf_comb* comb, lf_sumcheck** out) {
    *out = nullptr;
    return guard(c, [&] {
        if (nv < 1) throw LfException(LF_ERR_SUMCHECK_MISUSE, "Attempt to prove a constant.");
        if (M < 1) throw LfException(LF_ERR_INVALID_ARG, "no MLEs");
        const size_t n = (size_t)1 << nv;
        for (int i = 0; i < M; ++i) if (mles[i]->n > n) throw LfException(LF_ERR_MLE_LEN, "IncorrectLength: MLE longer than 2^num_vars");
        Eng E(c); std::unique_ptr<lf_sumcheck> sc(new lf_sumcheck); sc->ctx = c; sc->nv = nv; sc->deg = degree; sc->kind = comb->kind; sc->len = n;
        SumcheckDriver<G> drv(c, sc.get());
        int n_dense = M;
        if (comb->kind == LF_COMB_FOLD) {
            if (comb->b != 2) throw LfException(LF_ERR_UNSUPPORTED, "FOLD kernels are specialised for b = 2");
            if (degree != 4 || M != 5 + comb->n_mu * TAU) throw LfException(LF_ERR_INVALID_ARG, "FOLD: need degree 2b and 5 + n_mu*tau MLEs");
            n_dense = 5;
        } else {
            if (M > SC_MAX_MLES || degree > SC_MAX_DEG || comb->n_terms > SC_MAX_TERMS || comb->n_terms < 1) throw LfException(LF_ERR_UNSUPPORTED, "PRODUCTS/LIN: at most 8 MLEs, degree 7, 4 terms");
        }
        auto fill = [&](lf_sumcheck::Group& g, int first, int count) {
            SumcheckDriver<G>::alloc_group(E, g, count, n);
            LF_CUDA(cudaMemsetAsync(g.cur, 0, (size_t)count * g.stride * 8, E.st()));
            for (int k = 0; k < count; ++k) { const lf_vec* v = mles[first + k]; if (v->n) LF_CUDA(cudaMemcpy2DAsync(g.cur + (size_t)k * g.stride, g.pitch * 8, v->p, v->pitch * 8, v->n * 8, D, cudaMemcpyDeviceToDevice, E.st())); }
        };
        fill(sc->dense, 0, n_dense);
        if (comb->kind == LF_COMB_FOLD) { fill(sc->fh, 5, M - 5); drv.set_mu(comb->mu_host, comb->n_mu); }
        else {
            sc->gen.n_mles = M; sc->gen.deg = degree; sc->gen.lin = comb->kind == LF_COMB_LIN; sc->gen.n_terms = comb->n_terms;
            int o = 0;
            for (int t = 0; t < comb->n_terms; ++t) { if (comb->idx_len[t] > SC_MAX_FACTORS) throw LfException(LF_ERR_UNSUPPORTED, "more than 4 factors in one term"); sc->gen.term_len[t] = comb->idx_len[t];
                for (int f = 0; f < comb->idx_len[t]; ++f) { int j = comb->idx[o++]; if (j < 0 || j >= M) throw LfException(LF_ERR_INVALID_ARG, "comb index outside MLE list"); sc->gen.term_idx[t][f] = j; } }
            sc->d_coef = E.dalloc<u64>((size_t)comb->n_terms * D);
            LF_CUDA(cudaMemcpyAsync(sc->d_coef, comb->coef_host, (size_t)comb->n_terms * D * 8, cudaMemcpyHostToDevice, E.st())); E.sync();
        }
        for (int i = 0; i < M; ++i) { E.vec_free(mles[i]); mles[i] = nullptr; }   // ownership taken, as by the reference's Vec<DenseMultilinearExtension>
        *out = sc.release();
    });
}
lf_status lf_sumcheck_round(lf_sumcheck* sc, const uint64_t* prev, uint64_t* out_evals) {
    return guard(sc->ctx, [&] { SumcheckDriver<G> drv(sc->ctx, sc);
        if (prev) drv.apply_challenge(prev); else if (sc->round > 0) throw LfException(LF_ERR_SUMCHECK_MISUSE, "verifier message is empty");
        drv.evaluate(out_evals); });
}
lf_status lf_sumcheck_finish(lf_sumcheck* sc, const uint64_t* last, uint64_t* out_final) {
    return guard(sc->ctx, [&] { SumcheckDriver<G> drv(sc->ctx, sc); if (sc->round != sc->nv) throw LfException(LF_ERR_SUMCHECK_MISUSE, "sumcheck not finished"); drv.apply_challenge(last); drv.final_values(out_final); });
}
void lf_sumcheck_free(lf_sumcheck* sc) { if (sc) { SumcheckDriver<G> drv(sc->ctx, sc); drv.free_all(); delete sc; } }

// ---- transcript
lf_status lf_transcript_create(int32_t ring, lf_transcript** out) { *out = nullptr; return guard(nullptr, [&] { need_goldilocks(ring); *out = new lf_transcript{ring, Transcript<G>()}; }); }
lf_status lf_transcript_clone(const lf_transcript* t, lf_transcript** out) { *out = new lf_transcript(*t); return LF_OK; }
void lf_transcript_free(lf_transcript* t) { delete t; }
void lf_transcript_absorb(lf_transcript* t, const uint64_t* els, size_t count) { t->g.absorb_slice(els, count); }
void lf_transcript_absorb_base(lf_transcript* t, const uint64_t* limbs, size_t count) { t->g.absorb_base(limbs, count); }
void lf_transcript_absorb_tag(lf_transcript* t, const char* tag) { t->g.absorb_tag(tag); }
void lf_transcript_get_challenge(lf_transcript* t, uint64_t* out) { t->g.get_challenge(out); }
void lf_transcript_get_short_challenge(lf_transcript* t, uint64_t* out) { t->g.get_short_challenge(out); }
uint64_t lf_transcript_permutations(const lf_transcript* t) { return t->g.permutations(); }

lf_status lf_rot_lin_combination(int32_t ring, const uint64_t* rho, const uint64_t* theta, int32_t count, uint64_t* out) {
    return guard(nullptr, [&] { need_goldilocks(ring); std::vector<HostRing<G>::El> r(count); std::vector<HV> th(count);
        for (int i = 0; i < count; ++i) { r[i] = HostRing<G>::load(rho + (size_t)i * D); th[i].assign(theta + (size_t)i * TAU * D, theta + (size_t)(i + 1) * TAU * D); }
        HV o = Prover<G>::rot_lin_combination(r, th); std::memcpy(out, o.data(), 8 * o.size()); });
}

// ---- prover
uint64_t lf_proof_words(const lf_problem* P) {
    const u64 d = D, tau = TAU;
    return P->s * (P->d + 2) * d + tau * d + P->t * d + 2 * (u64)P->K * ((P->l + 1) + P->kappa + P->t + tau) * d + P->s * (2 * P->b + 1) * d + 2 * (u64)P->K * (tau + P->t) * d;
}
uint64_t lf_lcccs_words(const lf_problem* P) { return (P->s + TAU + P->kappa + P->t + P->l + 1) * (u64)D; }

lf_status lf_prover_create(lf_ctx* c, const lf_problem* sh, lf_prover** out) {
    *out = nullptr;
    return guard(c, [&] {
        need_goldilocks(sh->ring);
        if (sh->B_hi || sh->B_lo >= ((u64)1 << 62)) throw LfException(LF_ERR_UNSUPPORTED, "B >= 2^62");
        std::unique_ptr<lf_prover> p(new lf_prover); p->ctx = c; p->ring = sh->ring; p->L = sh->L; p->K = sh->K; p->B = sh->B_lo; p->b = sh->b;
        p->kappa = sh->kappa; p->n = sh->n; p->m = sh->m; p->n_ccs = sh->n_ccs; p->l = sh->l; p->t = sh->t; p->q = sh->q; p->d = sh->d; p->s = sh->s;
        int o = 0; for (u64 i = 0; i < sh->q; ++i) { p->S.emplace_back(sh->S_flat + o, sh->S_flat + o + sh->S_len[i]); o += sh->S_len[i]; }
        p->c.assign(sh->c, sh->c + sh->q * D);
        if (!sh->A) throw LfException(LF_ERR_INVALID_ARG, "Ajtai matrix is NULL");
        // sharded context: the caller passes this rank's column slice of A (kappa x n/world); CCS matrices are given whole
        // and cut to this rank's row slab here
        const size_t G = (size_t)c->world, n_loc = sh->n / G, m_loc = sh->m / G;
        if (sh->n % G || sh->m % G) throw LfException(LF_ERR_UNSUPPORTED, "n and m must be multiples of the rank count");
        lf_status rc = lf_ajtai_create(c, sh->kappa, n_loc, sh->A, &p->A); if (rc) throw LfException(rc, c->err);
        for (u64 j = 0; j < sh->t; ++j) {
            const lf_csr& M = sh->M[j]; if (M.nrows != sh->m) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "CCS matrix rows != m");
            const size_t r0 = (size_t)c->rank * m_loc; const u64 e0 = M.row_ptr[r0];
            std::vector<u64> rp(m_loc + 1); for (size_t i = 0; i <= m_loc; ++i) rp[i] = M.row_ptr[r0 + i] - e0;
            lf_sparse* m = nullptr; rc = lf_sparse_create(c, m_loc, M.ncols, rp.data(), M.col + e0, M.val + e0 * D, &m);
            if (rc) throw LfException(rc, c->err); p->M.push_back(m); }
        *out = p.release();
    });
}
void lf_prover_free(lf_prover* p) { if (!p) return; if (p->aux) lf_ctx_destroy(p->aux); for (auto* m : p->M) lf_sparse_free(p->ctx, m); lf_ajtai_free(p->ctx, p->A); delete p; }
lf_status lf_prover_upload_witness(lf_prover* p, const uint64_t* f_host, lf_witness** out) { *out = nullptr; return guard(p->ctx, [&] { Prover<G> pr(p); *out = pr.upload_witness(f_host); }); }
void lf_witness_free(lf_prover* p, lf_witness* w) { Prover<G> pr(p); pr.free_witness(w); }
lf_status lf_witness_download_f(lf_prover* p, const lf_witness* w, uint64_t* f_host) { return guard(p->ctx, [&] { Eng E(p->ctx); E.download_planes(w->f, w->pitch, w->n, f_host); }); }
lf_status lf_witness_f_from_w_ccs(lf_ctx* c, const uint64_t* w_ccs, size_t W, uint64_t B, int32_t L, uint64_t* f_host) {
    return guard(c, [&] { Eng E(c); const size_t wp = pitch_of(W), n = W * (size_t)L, np = pitch_of(n);
        u64 *w = E.dalloc<u64>(wp * D), *wc = E.dalloc<u64>(wp * D), *fc = E.dalloc<u64>(np * D), *f = E.dalloc<u64>(np * D);
        E.upload_planes(w_ccs, W, w, wp); E.crt(w, wp, wc, wp, W, true); E.gadget_decompose(wc, wp, fc, np, W, B, L); E.crt(fc, np, f, np, n, false);
        E.check_err_flag(LF_ERR_DOES_NOT_FIT, "from_w_ccs: a coefficient does not fit L digits of base B");
        E.download_planes(f, np, n, f_host); E.dfree(w); E.dfree(wc); E.dfree(fc); E.dfree(f); });
}
lf_status lf_linearize(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) {
    return guard(p->ctx, [&] { Prover<G> pr(p); lf_witness* w = pr.upload_witness(in->w_i_f);
        HV cm(in->cm_i_cm, in->cm_i_cm + p->kappa * D), x(in->cm_i_x_ccs, in->cm_i_x_ccs + p->l * D);
        auto lo = pr.linearize(cm, x, w, t->g); pr.E.dfree(lo.eq_r.p); pr.free_witness(w);
        Prover<G>::put_lcccs(out_lcccs, lo.lc);
        if (out_proof) { u64* q = out_proof; Prover<G>::put(q, lo.msgs); Prover<G>::put(q, lo.lc.v); Prover<G>::put(q, lo.lc.u); } });
}
lf_status lf_nifs_prove_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_acc, const lf_witness* w_i, lf_transcript* t,
                                 uint64_t* out_proof, uint64_t* out_lcccs, lf_witness** out_w) {
    return guard(p->ctx, [&] { Prover<G> pr(p); lf_witness* w = pr.prove(*in, w_acc, w_i, t->g, out_proof, out_lcccs); if (out_w) *out_w = w; else pr.free_witness(w); });
}
lf_status lf_nifs_prove(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, uint64_t* out_f) {
    return guard(p->ctx, [&] { Prover<G> pr(p);
        lf_witness* wa = pr.upload_witness(in->w_acc_f); lf_witness* wi = pr.upload_witness(in->w_i_f);
        lf_witness* w = pr.prove(*in, wa, wi, t->g, out_proof, out_lcccs);
        if (out_f) pr.E.download_planes(w->f, w->pitch, w->n, out_f);
        pr.free_witness(w); pr.free_witness(wa); pr.free_witness(wi); pr.E.sync(); });
}
lf_status lf_prover_timing_detail(const lf_prover* p, char* buf, size_t buf_len) {
    std::string out; for (auto& m : p->marks) out += m.first + " " + std::to_string(m.second) + "\n";
    if (out.size() + 1 > buf_len) return LF_ERR_INVALID_ARG;
    std::memcpy(buf, out.c_str(), out.size() + 1); return LF_OK;
}
lf_status lf_prover_last_timings(const lf_prover* p, double* out5) { for (int i = 0; i < 5; ++i) out5[i] = p->timings[i]; return LF_OK; }

}  // extern "C"
