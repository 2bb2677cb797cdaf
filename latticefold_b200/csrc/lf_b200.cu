// extern "C" surface of liblf_b200.so (declared in include/lf_b200.h).  No torch types, no CPU fallback: every compute
// entry point launches sm_100a kernels on the context's stream or fails with LF_ERR_CUDA.  Ring-dependent work is
// dispatched through lf::RingOps (ring_ops.cuh); this file holds only the ring-independent plumbing.
#include "ring_ops.cuh"

using namespace lf;
namespace lf { void plus_forget_ctx(lf_ctx* c); }      // lfplus.cu: drops the pinned-matrix records of a context that is going away

namespace {
thread_local std::string g_create_err;

template <class Fn> lf_status guard(lf_ctx* ctx, Fn&& fn) {
    try { fn(); return LF_OK; }
    catch (const LfException& e) { if (ctx) ctx->err = e.what(); else g_create_err = e.what(); return e.code; }
    catch (const std::exception& e) { if (ctx) ctx->err = e.what(); else g_create_err = e.what(); return LF_ERR_INVALID_ARG; }
}
RingOps* ops(int ring) {
    switch (ring) {
        case LF_RING_GOLDILOCKS: return ring_ops_goldilocks();
        case LF_RING_BABYBEAR: return ring_ops_babybear();
        case LF_RING_FROG: return ring_ops_frog();
        default: throw LfException(LF_ERR_UNSUPPORTED, "unknown ring id (the Stark ring's 256-bit field is not implemented)");
    }
}
}  // namespace

extern "C" {

lf_status lf_ring_describe(int32_t ring_id, lf_ring_info* out) { return guard(nullptr, [&] { ops(ring_id)->describe(out); }); }
const char* lf_last_error(const lf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

lf_status lf_ctx_create(int32_t ring_id, int32_t device, lf_ctx** out) {
    *out = nullptr;
    return guard(nullptr, [&] {
        RingOps* o = ops(ring_id);
        int ndev = 0; cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw LfException(LF_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
        if (device < 0 || device >= ndev) throw LfException(LF_ERR_INVALID_ARG, "device index out of range");
        LF_CUDA(cudaSetDevice(device));
        std::unique_ptr<lf_ctx> c(new lf_ctx); c->ring = ring_id; c->device = device;
        // the context's own stream outranks the prover's auxiliary stream (created at default = lowest priority): the host-paced
        // sumcheck rounds on this stream are chains of small kernels that must not queue behind the decompositions' long ones
        { int least = 0, greatest = 0; LF_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
          LF_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest)); }
        o->ctx_tables_create(c.get());
        LF_CUDA(cudaMalloc(&c->d_err, sizeof(int))); LF_CUDA(cudaMemset(c->d_err, 0, sizeof(int)));
        *out = c.release();
    });
}
void lf_ctx_destroy(lf_ctx* c) {
    if (!c) return;
    lf::plus_forget_ctx(c);
    cudaSetDevice(c->device); cudaStreamSynchronize(c->stream);
    for (int r = 0; r < 8; ++r) if (c->xg.peer_region[r]) cudaIpcCloseMemHandle(c->xg.peer_region[r]);
    if (c->xg.region) cudaFree(c->xg.region);
    if (c->nccl && !c->shared_tables) NcclApi::get().CommDestroy(c->nccl);      // an auxiliary context borrows its parent's communicator
    c->deferred_free.clear();      // (they are entries of block_size and are released with it)
    for (auto& kv : c->block_size) cudaFree(kv.first);
    if (!c->shared_tables) { for (int i = 0; i < 2; ++i) { cudaFree(c->d_tab_idx[i]); cudaFree(c->d_tab_val[i]); } try { ops(c->ring)->ctx_tables_destroy(c); } catch (...) {} }
    cudaFree(c->d_err); cudaFree(c->d_small); cudaFree(c->d_partial); if (c->h_pinned) cudaFreeHost(c->h_pinned); if (c->h_arena) cudaFreeHost(c->h_arena);
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

// ---- peer-memory mailboxes for the fused reduce + all-reduce kernel (k_reduce_allreduce_p2p)
lf_status lf_ctx_p2p_export(lf_ctx* c, uint8_t* out_handle64) {
    return guard(c, [&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        LF_CUDA(cudaSetDevice(c->device));
        if (!c->xg.region) {
            const size_t cap = Engine<GoldilocksRing>::XG_CAP, bytes = Engine<GoldilocksRing>::XG_CHANNELS * Engine<GoldilocksRing>::XG_CHANNEL_BYTES;
            LF_CUDA(cudaMalloc(&c->xg.region, bytes)); LF_CUDA(cudaMemset(c->xg.region, 0, bytes)); LF_CUDA(cudaDeviceSynchronize());
            c->xg.cap = cap;
        }
        cudaIpcMemHandle_t h; LF_CUDA(cudaIpcGetMemHandle(&h, c->xg.region)); std::memcpy(out_handle64, &h, 64);
    });
}
lf_status lf_ctx_p2p_import(lf_ctx* c, int32_t rank, int32_t world, const uint8_t* handles) {
    return guard(c, [&] {
        if (!handles) { c->xg.on = false; return; }      // some rank could not map its peers: everyone stays on NCCL
        if (world < 2 || world > 8 || rank != c->rank || world != c->world) throw LfException(LF_ERR_INVALID_ARG, "p2p import: rank / world must match the shard setup (2..8 ranks)");
        if (!c->xg.region) throw LfException(LF_ERR_INVALID_ARG, "p2p import before export");
        LF_CUDA(cudaSetDevice(c->device));
        for (int r = 0; r < world; ++r) {
            void* base = c->xg.region;
            if (r != rank) {
                cudaIpcMemHandle_t h; std::memcpy(&h, handles + 64 * r, 64);
                LF_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess)); c->xg.peer_region[r] = base;
            }
            c->xg.base[r] = (unsigned char*)base;
            c->xg.flags[r] = (unsigned long long*)base;
            c->xg.inbox[r] = (u64*)((unsigned char*)base + Engine<GoldilocksRing>::XG_FLAG_BYTES);
        }
        c->xg.on = true; c->xg.calls = 0; c->xg.blocks = 0;
    });
}

lf_status lf_ctx_set_bulk_repr(lf_ctx* c, int32_t repr) {
    return guard(c, [&] { if (repr != LF_REPR_CANONICAL && repr != LF_REPR_MONTGOMERY) throw LfException(LF_ERR_INVALID_ARG, "unknown representation"); c->bulk_repr = repr; });
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
lf_status lf_vec_upload(lf_ctx* c, const uint64_t* host, size_t n, int32_t form, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->vec_upload(c, host, n, form, out); }); }
lf_status lf_vec_download(lf_ctx* c, const lf_vec* v, uint64_t* host) { return guard(c, [&] { ops(c->ring)->vec_download(c, v, host); }); }
size_t lf_vec_len(const lf_vec* v) { return v->n; }
int32_t lf_vec_form(const lf_vec* v) { return v->form; }
void lf_vec_free(lf_ctx* c, lf_vec* v) { if (v) guard(c, [&] { ops(c->ring)->vec_free(c, v); }); }
lf_status lf_crt(lf_ctx* c, const lf_vec* in, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->crt(c, in, out); }); }
lf_status lf_icrt(lf_ctx* c, const lf_vec* in, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->icrt(c, in, out); }); }
lf_status lf_gadget_decompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->gadget_decompose(c, in, B, L, out); }); }
lf_status lf_gadget_recompose(lf_ctx* c, const lf_vec* in, uint64_t B, int32_t L, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->gadget_recompose(c, in, B, L, out); }); }
lf_status lf_decompose_to_vec(lf_ctx* c, const lf_vec* in, uint64_t b, int32_t K, lf_vec** out_k) { return guard(c, [&] { ops(c->ring)->decompose_to_vec(c, in, b, K, out_k); }); }
lf_status lf_fhat(lf_ctx* c, const lf_vec* in, lf_vec** out_tau) { return guard(c, [&] { ops(c->ring)->fhat(c, in, out_tau); }); }
// ---- Ajtai
lf_status lf_ajtai_create(lf_ctx* c, size_t kappa, size_t n, const uint64_t* host, lf_ajtai** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->ajtai_create(c, kappa, n, host, out); }); }
void lf_ajtai_free(lf_ctx* c, lf_ajtai* a) { if (a) { cudaStreamSynchronize(c->stream); cudaFree(a->p); cudaFree(a->a8); cudaFree(a->epi); delete a; } }
size_t lf_ajtai_kappa(const lf_ajtai* a) { return a->kappa; }
size_t lf_ajtai_width(const lf_ajtai* a) { return a->n; }
lf_status lf_commit_batch(lf_ctx* c, const lf_ajtai* a, const lf_vec* const* f, int32_t count, uint64_t* out_host) { return guard(c, [&] { ops(c->ring)->commit_batch(c, a, f, count, out_host); }); }
lf_status lf_commit(lf_ctx* c, const lf_ajtai* a, const lf_vec* f, uint64_t* out_host) { return lf_commit_batch(c, a, &f, 1, out_host); }
lf_status lf_commit_coeff(lf_ctx* c, const lf_ajtai* a, const lf_vec* f, uint64_t* out_host) { return guard(c, [&] { ops(c->ring)->commit_coeff(c, a, f, out_host); }); }
lf_status lf_decompose_and_commit_coeff(lf_ctx* c, const lf_ajtai* a, const lf_vec* f, uint64_t B, int32_t L, uint64_t* out_host) { return guard(c, [&] { ops(c->ring)->decompose_and_commit(c, a, f, false, B, L, out_host); }); }
lf_status lf_decompose_and_commit_ntt(lf_ctx* c, const lf_ajtai* a, const lf_vec* w, uint64_t B, int32_t L, uint64_t* out_host) { return guard(c, [&] { ops(c->ring)->decompose_and_commit(c, a, w, true, B, L, out_host); }); }
lf_status lf_commit_pieces(lf_ctx* c, const lf_ajtai* a, const lf_vec* f, uint64_t b, int32_t K, uint64_t* out_host) { return guard(c, [&] { ops(c->ring)->commit_pieces(c, a, f, b, K, out_host); }); }
// ---- sparse
lf_status lf_sparse_create(lf_ctx* c, size_t nrows, size_t ncols, const uint64_t* row_ptr, const uint64_t* col, const uint64_t* val, lf_sparse** out) {
    *out = nullptr; return guard(c, [&] { ops(c->ring)->sparse_create(c, nrows, ncols, row_ptr, col, val, out); });
}
void lf_sparse_free(lf_ctx* c, lf_sparse* m) { if (m) { cudaStreamSynchronize(c->stream); cudaFree(m->row_ptr); cudaFree(m->col); cudaFree(m->val); cudaFree(m->t_col_ptr); cudaFree(m->t_row); cudaFree(m->t_val); delete m; } }
lf_status lf_spmv(lf_ctx* c, const lf_sparse* m, const lf_vec* z, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->spmv(c, m, z, out); }); }
lf_status lf_eq_table(lf_ctx* c, const uint64_t* r, int32_t s, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->eq_table(c, r, s, out); }); }
lf_status lf_mle_eval_batch(lf_ctx* c, const lf_vec* const* mles, int32_t count, int32_t nv, const uint64_t* point, int32_t point_len, uint64_t* out_host) {
    return guard(c, [&] { ops(c->ring)->mle_eval_batch(c, mles, count, nv, point, point_len, out_host); });
}
lf_status lf_lincomb(lf_ctx* c, const uint64_t* coeffs, const lf_vec* const* vecs, int32_t count, lf_vec** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->lincomb(c, coeffs, vecs, count, out); }); }
// ---- sumcheck
lf_status lf_sumcheck_begin(lf_ctx* c, lf_vec** mles, int32_t M, int32_t nv, int32_t degree, const lf_comb* comb, lf_sumcheck** out) {
    *out = nullptr; return guard(c, [&] { ops(c->ring)->sumcheck_begin(c, mles, M, nv, degree, comb, out); });
}
lf_status lf_sumcheck_round(lf_sumcheck* sc, const uint64_t* prev, uint64_t* out_evals) { return guard(sc->ctx, [&] { ops(sc->ctx->ring)->sumcheck_round(sc, prev, out_evals); }); }
lf_status lf_sumcheck_finish(lf_sumcheck* sc, const uint64_t* last, uint64_t* out_final) { return guard(sc->ctx, [&] { ops(sc->ctx->ring)->sumcheck_finish(sc, last, out_final); }); }
void lf_sumcheck_free(lf_sumcheck* sc) { if (sc) guard(sc->ctx, [&] { ops(sc->ctx->ring)->sumcheck_free(sc); }); }
// ---- transcript (host)
lf_status lf_transcript_create(int32_t ring, lf_transcript** out) { *out = nullptr; return guard(nullptr, [&] { *out = new lf_transcript{ring, ops(ring)->tr_new()}; }); }
lf_status lf_transcript_clone(const lf_transcript* t, lf_transcript** out) { *out = nullptr; return guard(nullptr, [&] { *out = new lf_transcript{t->ring, ops(t->ring)->tr_clone(t->impl)}; }); }
void lf_transcript_free(lf_transcript* t) { if (t) { guard(nullptr, [&] { ops(t->ring)->tr_free(t->impl); }); delete t; } }
void lf_transcript_absorb(lf_transcript* t, const uint64_t* els, size_t count) { guard(nullptr, [&] { ops(t->ring)->tr_absorb(t->impl, els, count); }); }
void lf_transcript_absorb_base(lf_transcript* t, const uint64_t* limbs, size_t count) { guard(nullptr, [&] { ops(t->ring)->tr_absorb_base(t->impl, limbs, count); }); }
void lf_transcript_absorb_tag(lf_transcript* t, const char* tag) { guard(nullptr, [&] { ops(t->ring)->tr_absorb_tag(t->impl, tag); }); }
void lf_transcript_get_challenge(lf_transcript* t, uint64_t* out) { guard(nullptr, [&] { ops(t->ring)->tr_get_challenge(t->impl, out); }); }
void lf_transcript_get_short_challenge(lf_transcript* t, uint64_t* out) { guard(nullptr, [&] { ops(t->ring)->tr_get_short_challenge(t->impl, out); }); }
uint64_t lf_transcript_permutations(const lf_transcript* t) { uint64_t r = 0; guard(nullptr, [&] { r = ops(t->ring)->tr_permutations(t->impl); }); return r; }
const char* lf_host_poseidon_backend(void) { return lf::poseidon_use_ifma() ? "avx512-ifma" : "scalar"; }
lf_status lf_rot_lin_combination(int32_t ring, const uint64_t* rho, const uint64_t* theta, int32_t count, uint64_t* out) { return guard(nullptr, [&] { ops(ring)->rot_lin_combination(rho, theta, count, out); }); }
// ---- prover
uint64_t lf_proof_words(const lf_problem* P) { uint64_t r = 0; guard(nullptr, [&] { r = ops(P->ring)->proof_words(P); }); return r; }
uint64_t lf_lcccs_words(const lf_problem* P) { uint64_t r = 0; guard(nullptr, [&] { r = ops(P->ring)->lcccs_words(P); }); return r; }
uint64_t lf_proof_wire_bytes(const lf_problem* P) { uint64_t r = 0; guard(nullptr, [&] { r = ops(P->ring)->wire_bytes(P); }); return r; }
lf_status lf_proof_serialize(const lf_problem* P, const uint64_t* w, uint8_t* out) { return guard(nullptr, [&] { if (!P || !w || !out) throw LfException(LF_ERR_INVALID_ARG, "null argument"); ops(P->ring)->wire_serialize(P, w, out); }); }
lf_status lf_proof_deserialize(const lf_problem* P, const uint8_t* in, uint64_t n, uint64_t* w) { return guard(nullptr, [&] { if (!P || !in || !w) throw LfException(LF_ERR_INVALID_ARG, "null argument"); ops(P->ring)->wire_deserialize(P, in, n, w); }); }
lf_status lf_prover_create(lf_ctx* c, const lf_problem* sh, lf_prover** out) { *out = nullptr; return guard(c, [&] { ops(c->ring)->prover_create(c, sh, out); }); }
void lf_prover_free(lf_prover* p) { if (!p) return; if (p->aux) lf_ctx_destroy(p->aux); for (auto* m : p->M) lf_sparse_free(p->ctx, m); lf_ajtai_free(p->ctx, p->A); delete p; }
lf_status lf_prover_upload_witness(lf_prover* p, const uint64_t* f_host, lf_witness** out) { *out = nullptr; return guard(p->ctx, [&] { ops(p->ring)->prover_upload_witness(p, f_host, out); }); }
void lf_witness_free(lf_prover* p, lf_witness* w) { guard(p->ctx, [&] { ops(p->ring)->witness_free(p, w); }); }
lf_status lf_witness_download_f(lf_prover* p, const lf_witness* w, uint64_t* f_host) { return guard(p->ctx, [&] { ops(p->ring)->witness_download_f(p, w, f_host); }); }
lf_status lf_witness_f_from_w_ccs(lf_ctx* c, const uint64_t* w_ccs, size_t W, uint64_t B, int32_t L, uint64_t* f_host) { return guard(c, [&] { ops(c->ring)->witness_f_from_w_ccs(c, w_ccs, W, B, L, f_host); }); }
lf_status lf_linearize(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) { return guard(p->ctx, [&] { ops(p->ring)->linearize(p, in, t, out_lcccs, out_proof); }); }
lf_status lf_nifs_prove_resident(lf_prover* p, const lf_problem* in, const lf_witness* w_acc, const lf_witness* w_i, lf_transcript* t,
                                 uint64_t* out_proof, uint64_t* out_lcccs, lf_witness** out_w) {
    return guard(p->ctx, [&] { ops(p->ring)->nifs_prove_resident(p, in, w_acc, w_i, t, out_proof, out_lcccs, out_w); });
}
lf_status lf_nifs_verify(const lf_problem* in, lf_transcript* t, const uint64_t* proof, uint64_t* out_lcccs) {
    return guard(nullptr, [&] { if (!in || !t || !proof) throw LfException(LF_ERR_INVALID_ARG, "null argument"); ops(in->ring)->nifs_verify(in, t, proof, out_lcccs); });
}
lf_status lf_linearization_verify(const lf_problem* in, lf_transcript* t, const uint64_t* lin_proof, uint64_t* out_lcccs) {
    return guard(nullptr, [&] { if (!in || !t || !lin_proof) throw LfException(LF_ERR_INVALID_ARG, "null argument"); ops(in->ring)->linearization_verify(in, t, lin_proof, out_lcccs); });
}
lf_status lf_linearize_resident(lf_prover* p, const lf_problem* in, const lf_witness* w, lf_transcript* t, uint64_t* out_lcccs, uint64_t* out_proof) {
    return guard(p->ctx, [&] { ops(p->ring)->linearize_resident(p, in, w, t, out_lcccs, out_proof); });
}
lf_status lf_witness_commit(lf_prover* p, const lf_witness* w, uint64_t* out_host) { return guard(p->ctx, [&] { ops(p->ring)->witness_commit(p, w, out_host); }); }
lf_status lf_nifs_prove(lf_prover* p, const lf_problem* in, lf_transcript* t, uint64_t* out_proof, uint64_t* out_lcccs, uint64_t* out_f) {
    return guard(p->ctx, [&] { ops(p->ring)->nifs_prove(p, in, t, out_proof, out_lcccs, out_f); });
}
lf_status lf_prover_timing_detail(const lf_prover* p, char* buf, size_t buf_len) {
    std::string out; for (auto& m : p->marks) out += m.first + " " + std::to_string(m.second) + "\n";
    if (out.size() + 1 > buf_len) return LF_ERR_INVALID_ARG;
    std::memcpy(buf, out.c_str(), out.size() + 1); return LF_OK;
}
lf_status lf_prover_last_timings(const lf_prover* p, double* out5) { for (int i = 0; i < 5; ++i) out5[i] = p->timings[i]; return LF_OK; }


}  // extern "C"
