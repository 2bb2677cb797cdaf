// Host drivers, host verifiers and extern "C" surface of the LatticeFold+ consumers (kernels: lfplus.cuh; declarations and image
// layouts: include/lf_b200.h).  The Fiat-Shamir transcript stays on the host as in the reference; every witness-sized loop is a kernel.
#include "ring_ops.cuh"
#include "lfplus.cuh"
#include <algorithm>
#include <numeric>
#include <chrono>
#include <mutex>
#include <cstdio>

using namespace lf;
using namespace lf::plus;

struct lf_plus_mat { u64* d = nullptr; size_t kappa = 0, n = 0; };
struct lf_plus_vec { u64* d = nullptr; size_t n = 0; };      // Vec<R> on the device: n x 16 canonical words (a block of the context's cache)
struct lf_plus_rg {
    size_t n = 0, kappa = 0, code_pitch = 0; int k = 0, l = 0; u64 b = 0;
    unsigned char* codes = nullptr;        // [k * 16 columns][code_pitch]: exponents of M_f[kk][.][c]
    unsigned char* mtau_codes = nullptr;   // [code_pitch]: exponents of m_tau
    signed char* tau = nullptr;            // [n]: the digits of split (small signed integers)
    u64* f = nullptr;                      // [n][16] canonical
    std::vector<u64> tau_host, fcoms, comM;
};

namespace {
typedef Engine<FrogRing> Eng;
typedef Transcript<FrogRing> Tr;
thread_local std::string g_plus_err;
template <class Fn> lf_status pguard(lf_ctx* ctx, Fn&& fn) {
    try { fn(); return LF_OK; }
    catch (const LfException& e) { if (ctx) ctx->err = e.what(); g_plus_err = e.what(); return e.code; }
    catch (const std::exception& e) { if (ctx) ctx->err = e.what(); g_plus_err = e.what(); return LF_ERR_INVALID_ARG; }
}
Tr& tr_of(lf_transcript* t) { if (!t || t->ring != LF_RING_FROG) throw LfException(LF_ERR_INVALID_ARG, "LatticeFold+ entry points need a transcript of LF_RING_FROG"); return *(Tr*)t->impl; }
void need_frog(lf_ctx* c) { if (!c || c->ring != LF_RING_FROG) throw LfException(LF_ERR_UNSUPPORTED, "LatticeFold+ entry points run on the Frog ring (X^16 + 1) only");
                           if (c->world > 1) throw LfException(LF_ERR_UNSUPPORTED, "LatticeFold+ entry points do not shard: run replicas (one context per GPU, each on its own instances)"); }

// transcript surface of latticefold-plus/src/transcript.rs
u64 challenge(Tr& T) { u64 c; T.squeeze_base(&c, 1); T.absorb_base(&c, 1); return c; }
void absorb_field(Tr& T, u64 c) { u64 el[PD] = {0}; el[0] = c; T.absorb_base(el, PD); }

// a sparse matrix / dense vector of ring elements grouped by columns on the device
struct DevSparse {
    u64* col_ptr = nullptr; u32 *erow = nullptr, *ecol = nullptr; u64* val = nullptr; size_t nrows = 0, ncols = 0, nnz = 0; bool borrowed = false;      // borrowed: a pinned matrix's resident copy
    void free(Eng& E) { if (borrowed) return; E.dfree(col_ptr); E.dfree(erow); E.dfree(ecol); E.dfree(val); col_ptr = nullptr; erow = ecol = nullptr; val = nullptr; }
};
void check_canonical(const u64* v, size_t n, const char* what) { for (size_t i = 0; i < n; ++i) if (v[i] >= Fm::P) throw LfException(LF_ERR_INVALID_ARG, std::string(what) + ": non-canonical field element"); }
void validate_csr(const lf_csr& m) {      // the checks of upload_by_columns without the upload
    if (!m.row_ptr || (m.row_ptr[m.nrows] && (!m.col || !m.val))) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: null arrays");
    const size_t nnz = m.row_ptr[m.nrows];
    if (m.nrows >> 32 || m.ncols >> 32) throw LfException(LF_ERR_UNSUPPORTED, "sparse matrix: more than 2^32 rows / columns");
    for (size_t r = 0; r < m.nrows; ++r) if (m.row_ptr[r + 1] < m.row_ptr[r] || m.row_ptr[r + 1] > nnz) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: row_ptr not monotone");
    for (size_t e = 0; e < nnz; ++e) if (m.col[e] >= m.ncols) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: column index out of range");
    check_canonical(m.val, nnz * PD, "sparse matrix");
}
DevSparse upload_by_columns_raw(Eng& E, const lf_csr& m) {
    if (!m.row_ptr || (m.row_ptr[m.nrows] && (!m.col || !m.val))) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: null arrays");
    const size_t nnz = m.row_ptr[m.nrows];
    if (m.nrows >> 32 || m.ncols >> 32) throw LfException(LF_ERR_UNSUPPORTED, "sparse matrix: more than 2^32 rows / columns");
    check_canonical(m.val, nnz * PD, "sparse matrix");
    std::vector<u64> cp(m.ncols + 1, 0); std::vector<u32> rows(nnz);
    for (size_t r = 0; r < m.nrows; ++r) { if (m.row_ptr[r + 1] < m.row_ptr[r] || m.row_ptr[r + 1] > nnz) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: row_ptr not monotone");
        for (u64 e = m.row_ptr[r]; e < m.row_ptr[r + 1]; ++e) { if (m.col[e] >= m.ncols) throw LfException(LF_ERR_INVALID_ARG, "sparse matrix: column index out of range"); cp[m.col[e] + 1]++; rows[e] = (u32)r; } }
    for (size_t c = 0; c < m.ncols; ++c) cp[c + 1] += cp[c];
    std::vector<u64> pos(cp.begin(), cp.end() - 1), val(std::max<size_t>(nnz, 1) * PD); std::vector<u32> er(std::max<size_t>(nnz, 1)), ec(std::max<size_t>(nnz, 1));
    for (size_t e = 0; e < nnz; ++e) { const u64 p = pos[m.col[e]]++; er[p] = rows[e]; ec[p] = (u32)m.col[e]; std::memcpy(&val[p * PD], m.val + e * PD, 8 * PD); }
    DevSparse S; S.nrows = m.nrows; S.ncols = m.ncols; S.nnz = nnz;
    S.col_ptr = E.dalloc<u64>(cp.size()); S.erow = E.dalloc<u32>(er.size()); S.ecol = E.dalloc<u32>(ec.size()); S.val = E.dalloc<u64>(val.size());
    LF_CUDA(cudaMemcpyAsync(S.col_ptr, cp.data(), cp.size() * 8, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(S.erow, er.data(), er.size() * 4, cudaMemcpyHostToDevice, E.st()));
    LF_CUDA(cudaMemcpyAsync(S.ecol, ec.data(), ec.size() * 4, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(S.val, val.data(), val.size() * 8, cudaMemcpyHostToDevice, E.st()));
    E.sync();      // the staging vectors die here
    return S;
}
DevSparse upload_dense_vector(Eng& E, const u64* v, size_t n) {
    if (!v || n >> 32) throw LfException(LF_ERR_INVALID_ARG, "vector set: null / too long");
    check_canonical(v, n * PD, "vector set");
    std::vector<u32> er(std::max<size_t>(n, 1)), ec(std::max<size_t>(n, 1), 0); std::iota(er.begin(), er.end(), 0u); const u64 cp[2] = {0, n};
    DevSparse S; S.nrows = n; S.ncols = 1; S.nnz = n;
    S.col_ptr = E.dalloc<u64>(2); S.erow = E.dalloc<u32>(er.size()); S.ecol = E.dalloc<u32>(ec.size()); S.val = E.dalloc<u64>(std::max<size_t>(n, 1) * PD);
    LF_CUDA(cudaMemcpyAsync(S.col_ptr, cp, 16, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(S.erow, er.data(), er.size() * 4, cudaMemcpyHostToDevice, E.st()));
    LF_CUDA(cudaMemcpyAsync(S.ecol, ec.data(), ec.size() * 4, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(S.val, v, n * PD * 8, cudaMemcpyHostToDevice, E.st()));
    E.sync();
    return S;
}

// one monomial set as the device sees it: exponent codes (the range check's own M_f / m_tau) or general entries (anything the caller passes)
struct DevSet { const unsigned char* codes = nullptr; size_t code_pitch = 0; const DevSparse* gen = nullptr; size_t nrows = 0, ncols = 0; };

// blocks along the rows for a weighted-sum launch of `ncols` columns: about four waves of blocks in all, so that a thread of a many-column launch
// (the double commitment: k * 16 columns) sums several rows before the block reduction instead of one
unsigned chunks_for(size_t n, int tpb, size_t ncols = 1) { return (unsigned)std::min<size_t>(std::max<size_t>((n + tpb - 1) / tpb, 1), std::max<size_t>(148 * 4 / std::max<size_t>(ncols, 1), 1)); }

// Weighted sums  out[col] = sum_x W[x] (*) entry(x, col)  are queued and flushed together: every launch writes its block partials into one
// shared [chunk][columns of all jobs][16] buffer, then ONE reduction and ONE read-back serve them all (a range check has ~30 of them, a
// commitment transformation ~100; issued one by one each costs a reduction launch and a stream synchronisation).
enum WKind { W_SCALAR_MONO, W_SCALAR_GENERAL, W_SCALAR_SMALL, W_RING_MONO, W_RING_GENERAL, W_RING_SMALL };
struct WJob { WKind kind; const u64* W; size_t nrows, cols; u64* out;
              const unsigned char* codes = nullptr; size_t code_pitch = 0;                       // *_MONO
              const u64* col_ptr = nullptr; const u32* erow = nullptr; const u64* val = nullptr;  // *_GENERAL (erow == nullptr: dense, entry e in row e)
              const signed char* tau = nullptr; };                                               // *_SMALL
struct WsumBatch {
    Eng& E; std::vector<WJob> jobs; size_t total_cols = 0;
    explicit WsumBatch(Eng& e) : E(e) {}
    void add(const WJob& j) { jobs.push_back(j); total_cols += j.cols; }
    void mono(const u64* W, bool ring_w, const unsigned char* codes, size_t pitch, size_t nrows, size_t cols, u64* out) { WJob j{ring_w ? W_RING_MONO : W_SCALAR_MONO, W, nrows, cols, out}; j.codes = codes; j.code_pitch = pitch; add(j); }
    void general(const u64* W, bool ring_w, const u64* col_ptr, const u32* erow, const u64* val, size_t rows_per_col, size_t cols, u64* out) { WJob j{ring_w ? W_RING_GENERAL : W_SCALAR_GENERAL, W, rows_per_col, cols, out}; j.col_ptr = col_ptr; j.erow = erow; j.val = val; add(j); }
    void small(const u64* W, bool ring_w, const signed char* tau, size_t nrows, u64* out) { WJob j{ring_w ? W_RING_SMALL : W_SCALAR_SMALL, W, nrows, 1, out}; j.tau = tau; add(j); }
    // scale of a result: scalar weights are Montgomery (x mono / small entries: result Montgomery; x canonical entries: canonical); ring weights are
    // canonical (x canonical general entries: the products come out times 2^-64)
    static int scale_of(WKind k) { return k == W_SCALAR_MONO || k == W_SCALAR_SMALL ? 1 : k == W_RING_GENERAL ? 2 : 0; }
    void flush() {
        if (jobs.empty()) return;
        size_t max_rows = 1; for (auto& j : jobs) max_rows = std::max(max_rows, j.nrows);
        const unsigned ch = (unsigned)std::min<size_t>(std::max<size_t>((max_rows + 1023) / 1024, 8), std::max<size_t>(148 * 4 / std::max<size_t>(total_cols / jobs.size(), 1), 8));
        u64* partial = E.partial_dev((size_t)ch * total_cols * PD); const SmallArgs sm = small_consts_cached();
        size_t off = 0;
        for (auto& j : jobs) { const POut po{partial, (unsigned)total_cols, (unsigned)off}; const dim3 g(ch, (unsigned)j.cols);
            switch (j.kind) {
                case W_SCALAR_MONO: E.launch("k_plus_wsum_scalar_mono", [&] { k_plus_wsum_scalar_mono<<<g, 256, 0, E.st()>>>(j.W, j.codes, j.code_pitch, j.nrows, po); }); break;
                case W_RING_MONO: E.launch("k_plus_wsum_ring_mono", [&] { k_plus_wsum_ring_mono<<<g, 256, 0, E.st()>>>(j.W, j.codes, j.code_pitch, j.nrows, po); }); break;
                case W_SCALAR_GENERAL: E.launch("k_plus_wsum_scalar_general", [&] { k_plus_wsum_scalar_general<<<g, 256, 0, E.st()>>>(j.W, j.col_ptr, j.erow, j.val, po); }); break;
                case W_RING_GENERAL: E.launch("k_plus_wsum_ring_general", [&] { k_plus_wsum_ring_general<<<g, 128, 0, E.st()>>>(j.W, j.col_ptr, j.erow, j.val, po); }); break;
                case W_SCALAR_SMALL: E.launch("k_plus_wsum_scalar_small", [&] { k_plus_wsum_scalar_small<<<g, 256, 0, E.st()>>>(j.W, j.tau, j.nrows, sm, po); }); break;
                case W_RING_SMALL: E.launch("k_plus_wsum_ring_small", [&] { k_plus_wsum_ring_small<<<g, 256, 0, E.st()>>>(j.W, j.tau, j.nrows, sm, po); }); break;
            }
            off += j.cols; }
        u64* d_out = E.small_dev(total_cols * PD);
        E.reduce_partials(partial, (int)ch, total_cols * PD, d_out);
        std::vector<u64> host(total_cols * PD); E.download_words(d_out, total_cols * PD, host.data());
        off = 0;
        for (auto& j : jobs) { const int sc = scale_of(j.kind); const u64* src = host.data() + off * PD;
            for (size_t i = 0; i < j.cols * PD; ++i) j.out[i] = sc == 1 ? Fm::from_mont(src[i]) : sc == 2 ? Fm::to_mont(src[i]) : src[i];
            off += j.cols; }
        jobs.clear(); total_cols = 0;
    }
    static const SmallArgs& small_consts_cached() { static const SmallArgs s = [] { SmallArgs x; for (int t = -8; t < 8; ++t) x.v[t & 15] = Fm::to_mont(t < 0 ? Fm::P - (u64)(-t) : (u64)t); return x; }(); return s; }
    // every column of a set: monomial codes or general entries
    void set(const DevSet& S, const u64* W, bool ring_w, u64* out) {
        if (S.codes) mono(W, ring_w, S.codes, S.code_pitch, S.nrows, S.ncols, out);
        else general(W, ring_w, S.gen->col_ptr, S.gen->erow, S.gen->val, std::max<size_t>(S.gen->nnz / std::max<size_t>(S.gen->ncols, 1), 1), S.gen->ncols, out);
    }
};
SmallArgs small_consts() { SmallArgs s; for (int t = -8; t < 8; ++t) s.v[t & 15] = Fm::to_mont(t < 0 ? Fm::P - (u64)(-t) : (u64)t); return s; }
void eq_table(Eng& E, const std::vector<u64>& c, u64* d_out, size_t N) {
    if (c.size() > 40) throw LfException(LF_ERR_UNSUPPORTED, "more than 40 variables");
    EqArgs a; a.nv = (int)c.size(); for (size_t i = 0; i < c.size(); ++i) { a.c[i] = Fm::to_mont(c[i]); a.omc[i] = Fm::to_mont(Fm::sub(1, c[i])); }
    E.launch("k_plus_eq", [&] { k_plus_eq<<<Eng::blocks_for(N, 256), 256, 0, E.st()>>>(d_out, N, a, Fm::r1()); });
}

struct SetCheckResult {
    int nvars = 0, n_mat = 0, ncols = 0, n_vec = 0, n_M = 0;
    std::vector<u64> r, msgs, e, b;
    u64* d_eq_r = nullptr; std::vector<u64*> d_w;      // eq(r, .) (Montgomery) and w_i = M_i^T eq(r, .) stay on the device for the range check
    void free(Eng& E) { E.dfree(d_eq_r); for (u64* p : d_w) E.dfree(p); d_eq_r = nullptr; d_w.clear(); }
    std::vector<u64> words() const {
        std::vector<u64> w = {(u64)nvars, (u64)n_mat, (u64)ncols, (u64)n_vec, (u64)n_M};
        for (auto* v : {&r, &msgs, &e, &b}) w.insert(w.end(), v->begin(), v->end());
        return w;
    }
};

// In::set_check (setchk.rs:59-262) with the tables of the sumcheck held as base-field words
SetCheckResult set_check_core(Eng& E, Tr& T, int nvars, const std::vector<DevSet>& mats, const std::vector<DevSet>& vecs, const std::vector<DevSparse>& M) {
    if (mats.empty()) throw LfException(LF_ERR_UNSUPPORTED, "set check needs at least one matrix set (setchk.rs:85)");
    if (nvars < 1 || nvars > 30) throw LfException(LF_ERR_INVALID_ARG, "set check: nvars out of range");
    const size_t N = (size_t)1 << nvars, ncols = mats[0].ncols, nrows = mats[0].nrows, nM = mats.size(), nV = vecs.size();
    if (nrows > N || !ncols) throw LfException(LF_ERR_INVALID_ARG, "set check: more rows than 2^nvars / no columns");
    for (auto& s : mats) if (s.ncols != ncols || s.nrows != nrows) throw LfException(LF_ERR_INVALID_ARG, "set check: matrix sets of different shapes");
    for (auto& s : vecs) if (s.ncols != 1 || s.nrows != nrows) throw LfException(LF_ERR_INVALID_ARG, "set check: vector set of a different length");
    for (auto& m : M) if (m.ncols != nrows || m.nrows > N) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "set check: M_i does not match the sets");
    const size_t n_tables = nM * (2 * ncols + 1) + nV * 3;
    u64* Tb = E.dalloc<u64>(n_tables * N); LF_CUDA(cudaMemsetAsync(Tb, 0, n_tables * N * 8, E.st()));
    std::vector<Group> groups; std::vector<u64> alphas; size_t base = 0, wofs = 0;
    auto one_set = [&](const DevSet& S) {      // Steps 1-2 for one set: c, beta, the tables, alpha
        std::vector<u64> c(nvars); for (auto& x : c) x = challenge(T);
        const u64 beta = challenge(T);
        PowArgs pw; u64 bp = 1;
        for (int e = 0; e < PD; ++e) { pw.v[e] = S.codes ? Fm::to_mont(bp) : Fm::to_mont(Fm::to_mont(bp)); bp = Fm::hmul(bp, beta); }
        u64* Ts = Tb + base * N;
        if (S.codes) E.launch("k_plus_tables", [&] { k_plus_tables_mono<<<dim3(Eng::blocks_for(S.nrows, 256), (unsigned)S.ncols), 256, 0, E.st()>>>(S.codes, S.code_pitch, S.nrows, Ts, N, pw); });
        else if (S.gen->nnz) E.launch("k_plus_tables", [&] { k_plus_tables_general<<<Eng::blocks_for(S.gen->nnz, 256), 256, 0, E.st()>>>(S.gen->ecol, S.gen->erow, S.gen->val, S.gen->nnz, Ts, N, pw); });
        eq_table(E, c, Ts + 2 * S.ncols * N, N);
        groups.push_back(Group{(int)base, (int)S.ncols, (int)wofs, 0}); base += 2 * S.ncols + 1; wofs += S.ncols;
        alphas.push_back(challenge(T));
    };
    for (auto& s : mats) one_set(s);
    for (auto& s : vecs) one_set(s);
    const bool have_rc = nM > 1; const u64 rc = have_rc ? challenge(T) : 1;
    // weights alpha_i^j rc^i (matrix sets), alpha rc^(n_mat + i) (vector sets).  Without rc the reference's closure returns after the
    // first matrix set (setchk.rs:169-173): only that group takes part in the sumcheck.
    std::vector<u64> w(wofs);
    for (size_t g = 0; g < groups.size(); ++g) { const u64 rcp = Fm::hpow(rc, g);
        for (int j = 0; j < groups[g].ncols; ++j) w[groups[g].wofs + j] = Fm::to_mont(Fm::hmul(g < nM ? Fm::hpow(alphas[g], j) : alphas[g], rcp)); }
    const int n_active = have_rc ? (int)groups.size() : 1;
    Group* d_groups = E.dalloc<Group>(groups.size()); u64* d_w = E.dalloc<u64>(w.size());
    E.h2d(d_groups, groups.data(), groups.size() * sizeof(Group)); E.h2d(d_w, w.data(), w.size() * 8);
    // flattened (group, column) list of the active groups for the short-table form of the round kernel
    std::vector<ColDesc> cols; for (int g = 0; g < n_active; ++g) for (int j = 0; j < groups[g].ncols; ++j) cols.push_back(ColDesc{groups[g].base + 2 * j, groups[g].base + 2 * groups[g].ncols});
    ColDesc* d_cols = E.dalloc<ColDesc>(cols.size()); E.h2d(d_cols, cols.data(), cols.size() * sizeof(ColDesc));

    SetCheckResult R; R.nvars = nvars; R.n_mat = (int)nM; R.ncols = (int)ncols; R.n_vec = (int)nV; R.n_M = (int)M.size();
    // MLSumcheck::prove_as_subprotocol (sumcheck.rs:53-80), degree 3
    absorb_field(T, (u64)nvars); absorb_field(T, 3);
    R.msgs.assign((size_t)nvars * 4 * PD, 0);
    u64* Tn = E.dalloc<u64>(n_tables * (N / 2)); u64 *cur = Tb, *nxt = Tn; size_t len = N; u64 r_prev = 0;
    for (int i = 0; i < nvars; ++i) {
        if (i > 0) {      // fix_variables with the previous challenge (prover.rs:61-72)
            const size_t n_out = len / 2; const u64 rm = Fm::to_mont(r_prev);
            E.launch("k_plus_fold", [&] { k_plus_fold<<<dim3(Eng::blocks_for(n_out, 256), (unsigned)n_tables), 256, 0, E.st()>>>(cur, len, nxt, n_out, rm); });
            std::swap(cur, nxt); len = n_out;
        }
        const size_t n_pairs = len / 2; const unsigned nblk = (unsigned)std::min<size_t>(std::max<size_t>((n_pairs + 255) / 256, 1), 148 * 16);      // one pair per thread up to 2^19 pairs: the per-pair chain is long, occupancy hides it
        u64* partial = E.partial_dev((size_t)nblk * 4); u64* d_out = E.small_dev(4);
        const bool short_tables = n_pairs < 8192;      // a warp per pair: 8 pairs per block
        const unsigned nblk_c = (unsigned)((n_pairs + 7) / 8);
        if (short_tables) { partial = E.partial_dev((size_t)nblk_c * 4); E.launch("k_plus_round", [&] { k_plus_round_cols<<<nblk_c, 256, 0, E.st()>>>(cur, len, n_pairs, d_cols, (int)cols.size(), d_w, partial); }); }
        else E.launch("k_plus_round", [&] { k_plus_round<<<nblk, 256, 0, E.st()>>>(cur, len, n_pairs, d_groups, n_active, d_w, partial); });
        E.reduce_partials(partial, (int)(short_tables ? nblk_c : nblk), 4, d_out);
        u64 h[4]; E.download_words(d_out, 4, h);
        u64* msg = R.msgs.data() + (size_t)i * 4 * PD;
        for (int X = 0; X < 4; ++X) msg[X * PD] = Fm::from_mont(h[X]);      // constants of R
        T.absorb_slice(msg, 4);
        r_prev = challenge(T); absorb_field(T, r_prev); R.r.push_back(r_prev);
    }
    E.dfree(Tb); E.dfree(Tn); E.dfree(d_groups); E.dfree(d_w); E.dfree(d_cols);
    // Step 3 (setchk.rs:199-257): e[0] = MLE(column)(r), e[1 + i] = MLE(M_i column)(r) = sum_x (M_i^T eq(r, .))[x] column[x], b = MLE(vector)(r)
    R.d_eq_r = E.dalloc<u64>(N); eq_table(E, R.r, R.d_eq_r, N);
    for (auto& m : M) { u64* wv = E.dalloc<u64>(std::max<size_t>(m.ncols, 1) * PD);
        E.launch("k_plus_mt_eq", [&] { k_plus_mt_eq<<<Eng::blocks_for(m.ncols, 128), 128, 0, E.st()>>>(R.d_eq_r, m.col_ptr, m.erow, m.val, m.ncols, wv); });
        R.d_w.push_back(wv); }
    R.e.assign((1 + M.size()) * nM * ncols * PD, 0); R.b.assign(nV * PD, 0);
    WsumBatch wb(E);
    for (size_t mi = 0; mi <= M.size(); ++mi) for (size_t i = 0; i < nM; ++i)
        wb.set(mats[i], mi == 0 ? R.d_eq_r : R.d_w[mi - 1], mi != 0, R.e.data() + (mi * nM + i) * ncols * PD);
    for (size_t i = 0; i < nV; ++i) wb.set(vecs[i], R.d_eq_r, false, R.b.data() + i * PD);
    wb.flush();
    T.absorb_slice(R.e.data(), R.e.size() / PD); T.absorb_slice(R.b.data(), R.b.size() / PD);      // absorb_evaluations, setchk.rs:346-356
    return R;
}

// ---------------------------------------------------------------- host verifiers (the reference's verifiers are host code as well)
u64 ev_host(const u64* r, u64 x) { u64 acc = 0, e = 1; for (int i = 0; i < PD; ++i) { acc = Fm::add(acc, Fm::hmul(r[i], e)); e = Fm::hmul(e, x); } return acc; }      // setchk.rs:46-57
struct SetImage { int nvars, n_mat, ncols, n_vec, n_M; const u64 *r, *msgs, *e, *b; size_t words; };
SetImage parse_set_image(const u64* w, size_t len) {
    if (!w || len < 5) throw LfException(LF_ERR_INCORRECT_LENGTH, "set-check image too short");
    if (w[0] < 1 || w[0] > 40 || w[1] < 1 || w[1] > 4096 || w[2] < 1 || w[2] > (1u << 20) || w[3] > 4096 || w[4] > 64) throw LfException(LF_ERR_INVALID_ARG, "set-check image: implausible header");
    SetImage s; s.nvars = (int)w[0]; s.n_mat = (int)w[1]; s.ncols = (int)w[2]; s.n_vec = (int)w[3]; s.n_M = (int)w[4];
    const size_t nr = s.nvars, nm = (size_t)s.nvars * 4 * PD, ne = (size_t)(1 + s.n_M) * s.n_mat * s.ncols * PD, nb = (size_t)s.n_vec * PD;
    s.words = 5 + nr + nm + ne + nb; if (len < s.words) throw LfException(LF_ERR_INCORRECT_LENGTH, "set-check image truncated");
    s.r = w + 5; s.msgs = s.r + nr; s.e = s.msgs + nm; s.b = s.e + ne;
    check_canonical(w + 5, s.words - 5, "set-check image");
    return s;
}
// Out::verify (setchk.rs:264-344); `point` receives the sumcheck's challenges
void verify_set_image(Tr& T, const SetImage& s, std::vector<u64>* point_out = nullptr) {
    const int nv = s.nvars, nclaims = s.n_mat + s.n_vec;
    struct Cba { std::vector<u64> c; u64 beta, alpha; }; std::vector<Cba> cba(nclaims);
    for (auto& x : cba) { x.c.resize(nv); for (auto& v : x.c) v = challenge(T); x.beta = challenge(T); x.alpha = challenge(T); }
    const u64 rc = s.n_mat > 1 ? challenge(T) : 1;
    // MLSumcheck::verify_as_subprotocol (sumcheck.rs:84-104) with claimed sum 0, degree 3
    absorb_field(T, (u64)nv); absorb_field(T, 3);
    std::vector<u64> point(nv);
    for (int i = 0; i < nv; ++i) { T.absorb_slice(s.msgs + (size_t)i * 4 * PD, 4); point[i] = challenge(T); absorb_field(T, point[i]); }
    u64 expected[PD] = {0};
    for (int i = 0; i < nv; ++i) { const u64* msg = s.msgs + (size_t)i * 4 * PD;
        for (int c = 0; c < PD; ++c) if (Fm::add(msg[c], msg[PD + c]) != expected[c]) throw LfException(LF_ERR_SUMCHECK_FAILED, "set check: sumcheck round sum mismatch");
        // interpolate through X = 0..3 at the challenge (verifier.rs:139-254)
        u64 lag[4];
        for (int a = 0; a < 4; ++a) { u64 num = 1, den = 1;
            for (int b = 0; b < 4; ++b) if (b != a) { num = Fm::hmul(num, Fm::sub(point[i], (u64)b)); den = Fm::hmul(den, a > b ? (u64)(a - b) : Fm::P - (u64)(b - a)); }
            lag[a] = Fm::hmul(num, Fm::hpow(den, Fm::P - 2)); }
        for (int c = 0; c < PD; ++c) { u64 v = 0; for (int a = 0; a < 4; ++a) v = Fm::add(v, Fm::hmul(msg[a * PD + c], lag[a])); expected[c] = v; }
    }
    T.absorb_slice(s.e, (size_t)(1 + s.n_M) * s.n_mat * s.ncols); T.absorb_slice(s.b, s.n_vec);
    auto eq_eval = [&](const std::vector<u64>& c) { u64 res = 1; for (int i = 0; i < nv; ++i) { const u64 xy = Fm::hmul(c[i], point[i]); res = Fm::hmul(res, Fm::add(Fm::sub(Fm::sub(Fm::add(xy, xy), c[i]), point[i]), 1)); } return res; };
    u64 ver = 0;
    for (int i = 0; i < s.n_mat; ++i) { u64 esum = 0, ap = 1; const u64 b2 = Fm::hmul(cba[i].beta, cba[i].beta);
        for (int j = 0; j < s.ncols; ++j) { const u64* ej = s.e + ((size_t)i * s.ncols + j) * PD; const u64 e1 = ev_host(ej, cba[i].beta), e2 = ev_host(ej, b2);
            esum = Fm::add(esum, Fm::hmul(Fm::sub(Fm::hmul(e1, e1), e2), ap)); ap = Fm::hmul(ap, cba[i].alpha); }
        ver = Fm::add(ver, Fm::hmul(Fm::hmul(eq_eval(cba[i].c), esum), Fm::hpow(rc, i))); }
    for (int i = 0; i < s.n_vec; ++i) { const Cba& x = cba[s.n_mat + i]; const u64* bi = s.b + (size_t)i * PD;
        const u64 e1 = ev_host(bi, x.beta), e2 = ev_host(bi, Fm::hmul(x.beta, x.beta));
        ver = Fm::add(ver, Fm::hmul(Fm::hmul(Fm::hmul(eq_eval(x.c), x.alpha), Fm::sub(Fm::hmul(e1, e1), e2)), Fm::hpow(rc, s.n_mat + i))); }
    if (expected[0] != ver) throw LfException(LF_ERR_SUMCHECK_FAILED, "set check: recomputed claim mismatch (SetCheckError::ExpectedEvaluation)");
    for (int c = 1; c < PD; ++c) if (expected[c]) throw LfException(LF_ERR_SUMCHECK_FAILED, "set check: recomputed claim mismatch (SetCheckError::ExpectedEvaluation)");
    if (point_out) *point_out = point;
}
// ct(psi * a) with psi = sum_{0<i<d/2} i (X^-i + X^i): the constant coefficient of the negacyclic product
// (psi_i = i, psi_{16-i} = -i, X^16 = -1  =>  ct = sum_i i (a_i - a_{16-i}))
u64 ct_psi(const u64* a) { u64 acc = 0; for (int i = 1; i < PD / 2; ++i) acc = Fm::add(acc, Fm::hmul((u64)i, Fm::sub(a[i], a[PD - i]))); return acc; }
// ---------------------------------------------------------------- commitment transformation (cm.rs): host pieces
void hring_mul(u64* out, const u64* a, const u64* b) {      // negacyclic product of two canonical elements
    u64 w[PD] = {0};
    for (int i = 0; i < PD; ++i) { if (!a[i]) continue;
        for (int j = 0; j < PD; ++j) { if (!b[j]) continue; const u64 pr = Fm::hmul(a[i], b[j]); const int k = i + j; if (k >= PD) w[k - PD] = Fm::sub(w[k - PD], pr); else w[k] = Fm::add(w[k], pr); } }
    std::memcpy(out, w, sizeof w);
}
void hring_add(u64* out, const u64* a, const u64* b) { for (int i = 0; i < PD; ++i) out[i] = Fm::add(a[i], b[i]); }
void hring_axpy(u64* acc, const u64* a, u64 c) { for (int i = 0; i < PD; ++i) if (a[i]) acc[i] = Fm::add(acc[i], Fm::hmul(a[i], c)); }      // acc += c a
int plus_ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) ++l; return l; }
std::vector<u64> tensor_host(const std::vector<u64>& r) {      // utils.rs:74-86
    std::vector<u64> res(1, 1);
    for (u64 ri : r) { std::vector<u64> nx; nx.reserve(res.size() * 2); for (u64 a : res) { nx.push_back(Fm::hmul(a, Fm::sub(1, ri))); nx.push_back(Fm::hmul(a, ri)); } res.swap(nx); }
    return res;
}
// short_challenge(128, transcript) (utils.rs:88-103): 16 bytes -> (byte mod 256) - 128, the decode of the Frog challenge set
void short_challenge(Tr& T, u64* out) { T.get_short_challenge(out); }
int to_small(u64 c) { return c > Fm::P / 2 ? -(int)(Fm::P - c) : (int)c; }
struct CmChallenges { std::vector<u64> s /* 3 x 16 */, sp /* k*16 x 16 */; };
CmChallenges draw_cm_challenges(Tr& T, int k) { CmChallenges c; c.s.resize(3 * PD); c.sp.resize((size_t)k * PD * PD); for (int i = 0; i < 3; ++i) short_challenge(T, &c.s[i * PD]); for (int i = 0; i < k * PD; ++i) short_challenge(T, &c.sp[(size_t)i * PD]); return c; }
// ComX (cm.rs:537-575): cm_g[L][kappa], ro[nvars][2], vo[L][1 + n_M][2]
std::vector<u64> comx_words(const u64* s, size_t L, size_t kappa, size_t nE, const u64* const* fcoms /* L pointers: cm_f | C_Mf | cm_mtau */, const u64* comh, const std::vector<u64>* evals /* [2] */, const std::vector<u64>* ro /* [2] */) {
    std::vector<u64> w; u64 acc[PD], t[PD];
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < kappa; ++i) { const u64* cm_f = fcoms[l] + i * PD; const u64* C_Mf = fcoms[l] + (kappa + i) * PD; const u64* cm_mtau = fcoms[l] + (2 * kappa + i) * PD;
        hring_mul(acc, s, C_Mf); hring_mul(t, s + PD, cm_mtau); hring_add(acc, acc, t); hring_mul(t, s + 2 * PD, cm_f); hring_add(acc, acc, t); hring_add(acc, acc, comh + (l * kappa + i) * PD);
        w.insert(w.end(), acc, acc + PD); }
    for (size_t i = 0; i < ro[0].size(); ++i) { w.push_back(ro[0][i]); w.push_back(ro[1][i]); }
    for (size_t l = 0; l < L; ++l) for (size_t i = 0; i < nE; ++i) for (int z = 0; z < 2; ++z) { const u64* e = evals[z].data() + ((l * nE + i) * 4) * PD;
        hring_mul(acc, s, e); hring_mul(t, s + PD, e + PD); hring_add(acc, acc, t); hring_mul(t, s + 2 * PD, e + 2 * PD); hring_add(acc, acc, t); hring_add(acc, acc, e + 3 * PD);
        w.insert(w.end(), acc, acc + PD); }
    return w;
}
struct DevCsr { u64 *row_ptr = nullptr, *col = nullptr, *val = nullptr; size_t nrows = 0, ncols = 0, nnz = 0; bool borrowed = false; void free(Eng& E) { if (borrowed) return; E.dfree(row_ptr); E.dfree(col); E.dfree(val); row_ptr = col = val = nullptr; } };
DevCsr upload_csr_raw(Eng& E, const lf_csr& m) {      // validated by upload_by_columns before
    DevCsr S; S.nrows = m.nrows; S.ncols = m.ncols; S.nnz = m.row_ptr[m.nrows];
    S.row_ptr = E.dalloc<u64>(m.nrows + 1); S.col = E.dalloc<u64>(std::max<size_t>(S.nnz, 1)); S.val = E.dalloc<u64>(std::max<size_t>(S.nnz, 1) * PD);
    LF_CUDA(cudaMemcpyAsync(S.row_ptr, m.row_ptr, (m.nrows + 1) * 8, cudaMemcpyHostToDevice, E.st()));
    if (S.nnz) { LF_CUDA(cudaMemcpyAsync(S.col, m.col, S.nnz * 8, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(S.val, m.val, S.nnz * PD * 8, cudaMemcpyHostToDevice, E.st())); }
    E.sync(); return S;
}

// Pinned matrices (lf_plus_csr_pin): the static matrices of a protocol run (PlusProver::init holds A and M in the reference as well) keep a resident
// copy in both orders; the entry points recognise them by their host arrays and skip validation, sorting and the copies.
struct Pinned { lf_ctx* ctx; const u64 *row_ptr, *col, *val; u64 nrows, ncols; DevSparse by_col; DevCsr by_row; };
std::vector<Pinned>& pinned_list() { static std::vector<Pinned> v; return v; }
std::mutex& pinned_mutex() { static std::mutex m; return m; }
bool find_pinned(lf_ctx* c, const lf_csr& m, Pinned* out = nullptr) {      // (a copy under the lock: another context's pin may grow the list)
    std::lock_guard<std::mutex> g(pinned_mutex());
    for (auto& p : pinned_list()) if (p.ctx == c && p.row_ptr == m.row_ptr && p.col == m.col && p.val == m.val && p.nrows == m.nrows && p.ncols == m.ncols) { if (out) *out = p; return true; }
    return false;
}
DevSparse upload_by_columns(Eng& E, const lf_csr& m) { Pinned p; if (find_pinned(E.c, m, &p)) { DevSparse s = p.by_col; s.borrowed = true; return s; } return upload_by_columns_raw(E, m); }
DevCsr upload_csr(Eng& E, const lf_csr& m) { Pinned p; if (find_pinned(E.c, m, &p)) { DevCsr s = p.by_row; s.borrowed = true; return s; } return upload_csr_raw(E, m); }

// dense ring-valued MLE (len elements, zero tail) at a base-field point, by successive halving
void mle_eval_dense(std::vector<u64> ev, int nv, const std::vector<u64>& point, u64* out) {
    size_t len = ev.size() / PD;
    for (int i = 0; i < nv; ++i) { const size_t nl = (len + 1) / 2; std::vector<u64> nx(std::max<size_t>(nl, 1) * PD, 0);
        for (size_t b = 0; b < nl; ++b) for (int c = 0; c < PD; ++c) { const u64 a = ev[(2 * b) * PD + c], hi = 2 * b + 1 < len ? ev[(2 * b + 1) * PD + c] : 0; nx[b * PD + c] = Fm::add(a, Fm::hmul(Fm::sub(hi, a), point[i])); }
        ev.swap(nx); len = std::max<size_t>(nl, 1); }
    std::memcpy(out, ev.data(), 8 * PD);
}
// Rg::range_check (rgchk.rs:75-187).  Leaves eq(r, .) and the w_i on the device (R) for the commitment transformation that follows it in Cm::prove.
std::vector<u64> range_check_core(Eng& E, Tr& T, int nvars, lf_plus_rg* const* inst, int L, const std::vector<DevSparse>& Ms, SetCheckResult& R) {
        const lf_plus_rg& I0 = *inst[0];
        for (int i = 0; i < L; ++i) if (!inst[i] || inst[i]->n != I0.n || inst[i]->k != I0.k || inst[i]->kappa != I0.kappa || inst[i]->l != I0.l || inst[i]->b != I0.b) throw LfException(LF_ERR_INVALID_ARG, "range check: instances of different shapes");
        // sets: the k matrices M_f of every instance, then every instance's m_tau (rgchk.rs:80-89)
        std::vector<DevSet> mats, vecs;
        for (int i = 0; i < L; ++i) for (int kk = 0; kk < I0.k; ++kk) { DevSet s; s.codes = inst[i]->codes + (size_t)kk * PD * inst[i]->code_pitch; s.code_pitch = inst[i]->code_pitch; s.nrows = I0.n; s.ncols = PD; mats.push_back(s); }
        for (int i = 0; i < L; ++i) { DevSet s; s.codes = inst[i]->mtau_codes; s.code_pitch = inst[i]->code_pitch; s.nrows = I0.n; s.ncols = 1; vecs.push_back(s); }
        R = set_check_core(E, T, nvars, mats, vecs, Ms);
        // evaluations at r (rgchk.rs:101-171): v = c[0] = MLE(f)(r) coefficient-wise, a[0] = MLE(tau)(r), and through w_i = M_i^T eq(r, .) the M_i-images
        const size_t nE = 1 + Ms.size(), n = I0.n;
        std::vector<u64> img = {(u64)L, (u64)I0.k, (u64)I0.l, (u64)I0.kappa, I0.b}; { auto w = R.words(); img.insert(img.end(), w.begin(), w.end()); }
        u64* cp = E.dalloc<u64>(2); const u64 cph[2] = {0, n}; E.h2d(cp, cph, 16);
        std::vector<std::vector<u64>> vv(L, std::vector<u64>(PD)), ar(L, std::vector<u64>(nE * PD)), bv(L, std::vector<u64>(nE * PD)), cv(L, std::vector<u64>(nE * PD)), av(L, std::vector<u64>(nE));
        WsumBatch wb(E);
        for (int li = 0; li < L; ++li) { const lf_plus_rg& I = *inst[li];
            wb.general(R.d_eq_r, false, cp, nullptr, I.f, n, 1, vv[li].data());                       // v = c[0] = MLE(f)(r), coefficient-wise
            wb.small(R.d_eq_r, false, I.tau, n, ar[li].data());                                       // a[0] = MLE(tau)(r)
            for (size_t mi = 0; mi < Ms.size(); ++mi) { const u64* W = R.d_w[mi];
                wb.small(W, true, I.tau, n, ar[li].data() + (1 + mi) * PD);                           // a[1 + i] = ct(MLE(M_i tau)(r))
                wb.mono(W, true, I.mtau_codes, I.code_pitch, n, 1, bv[li].data() + (1 + mi) * PD);     // b[1 + i] = MLE(M_i m_tau)(r)
                wb.general(W, true, cp, nullptr, I.f, n, 1, cv[li].data() + (1 + mi) * PD); }          // c[1 + i] = MLE(M_i f)(r)
        }
        wb.flush(); E.dfree(cp);
        for (int li = 0; li < L; ++li) { const lf_plus_rg& I = *inst[li];
            std::memcpy(bv[li].data(), R.b.data() + (size_t)li * PD, 8 * PD); std::memcpy(cv[li].data(), vv[li].data(), 8 * PD);
            for (size_t e = 0; e < nE; ++e) av[li][e] = ar[li][e * PD];
            for (auto* x : {&vv[li], &av[li], &bv[li], &cv[li]}) img.insert(img.end(), x->begin(), x->end());
            img.insert(img.end(), I.fcoms.begin(), I.fcoms.end()); }
        for (int li = 0; li < L; ++li) { for (u64 a : av[li]) absorb_field(T, a); T.absorb_slice(cv[li].data(), nE); }      // rgchk.rs:338-343
        return img;
}
// Cm::prove (cm.rs:57-203).  proof image = Dcom image | comh | two sumcheck proofs | two evaluation blocks
void cm_prove_core(Eng& E, Tr& T, int nvars, lf_plus_rg* const* inst, int L, const lf_csr* Mh, int n_M, std::vector<u64>& proof, std::vector<u64>& comx, u64* g_host, bool sum_g = false, u64* g_dev = nullptr) {
    const lf_plus_rg& I0 = *inst[0]; const size_t n = I0.n, kappa = I0.kappa, N = (size_t)1 << nvars, nE = 1 + (size_t)n_M; const int k = I0.k, l = I0.l, kd = k * PD;
    std::vector<DevSparse> Ms; std::vector<DevCsr> Mr; SetCheckResult R; std::vector<void*> blocks;
    struct Cleanup { Eng& E; std::vector<DevSparse>& a; std::vector<DevCsr>& b; SetCheckResult& r; std::vector<void*>& blk; ~Cleanup() { for (auto& s : a) s.free(E); for (auto& s : b) s.free(E); r.free(E); for (void* p : blk) E.dfree(p); } } cl{E, Ms, Mr, R, blocks};
    auto alloc = [&](size_t words) { u64* p = E.dalloc<u64>(words); blocks.push_back(p); return p; };
    for (int i = 0; i < n_M; ++i) { Ms.push_back(upload_by_columns(E, Mh[i])); Mr.push_back(upload_csr(E, Mh[i])); }
    proof = range_check_core(E, T, nvars, inst, L, Ms, R);
    const CmChallenges ch = draw_cm_challenges(T, k);
    std::vector<short> sp16((size_t)kd * PD); for (size_t i = 0; i < sp16.size(); ++i) sp16[i] = (short)to_small(ch.sp[i]);
    short* d_sp = (short*)alloc((sp16.size() * 2 + 7) / 8); E.h2d(d_sp, sp16.data(), sp16.size() * 2);
    // h_l = sum_kk M_f[kk] s'_kk (device), comh_l = sum_kk comM_f[kk] s'_kk (kappa elements, host)
    std::vector<u64*> d_h(L);
    for (int li = 0; li < L; ++li) { d_h[li] = alloc(n * PD);
        E.launch("k_plus_h", [&] { k_plus_h<<<Eng::blocks_for(n, 128), 128, (size_t)kd * PD * 2, E.st()>>>(inst[li]->codes, inst[li]->code_pitch, n, kd, d_sp, d_h[li]); }); }
    std::vector<u64> comh((size_t)L * kappa * PD, 0); u64 t[PD];
    for (int li = 0; li < L; ++li) for (int kk = 0; kk < k; ++kk) for (size_t r = 0; r < kappa; ++r) for (int c = 0; c < PD; ++c) {
        hring_mul(t, &inst[li]->comM[(((size_t)kk * kappa + r) * PD + c) * PD], &ch.sp[((size_t)kk * PD + c) * PD]); hring_add(&comh[(li * kappa + r) * PD], &comh[(li * kappa + r) * PD], t); }
    T.absorb_slice(comh.data(), (size_t)L * kappa);
    const int log_kappa = plus_ceil_log2(kappa);
    std::vector<u64> c0(log_kappa), c1(log_kappa); for (auto& x : c0) x = challenge(T); for (auto& x : c1) x = challenge(T);
    const std::vector<u64> tc0 = tensor_host(c0), tc1 = tensor_host(c1);
    const size_t nt = tc0.size() * kd * l * PD;
    if (nt > n) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "cm: t(z) longer than the witness (cm.rs:150-161)");
    // scalar tables eq(r, .) | S = sum_l tau_l (kept; every sumcheck folds copies), ring tables G | U (rebuilt per sumcheck)
    const SmallArgs sm = WsumBatch::small_consts_cached();
    u64* sc0 = alloc(2 * N); LF_CUDA(cudaMemcpyAsync(sc0, R.d_eq_r, N * 8, cudaMemcpyDeviceToDevice, E.st())); LF_CUDA(cudaMemsetAsync(sc0 + N, 0, N * 8, E.st()));
    { TauList tl; if (L > 64) throw LfException(LF_ERR_UNSUPPORTED, "cm: more than 64 instances"); tl.n = L; for (int li = 0; li < L; ++li) tl.p[li] = inst[li]->tau;
      E.launch("k_plus_tau_sum", [&] { k_plus_tau_sum<<<Eng::blocks_for(n, 256), 256, 0, E.st()>>>(tl, n, sm, sc0 + N); }); }
    u64 *GU = alloc(2 * N * PD), *GUn = alloc(N * PD), *scn = alloc(N), *scm = alloc(N / 2 + 1), *GUm = alloc(N * PD / 2 + PD), *Z = n_M ? alloc(n * PD) : nullptr, *d_scal = alloc(tc0.size() * l);
    u64* eq_ro = alloc(N); u64* cp = alloc(2); const u64 cph[2] = {0, n}; E.h2d(cp, cph, 16);
    std::vector<u64> msgs[2], evals[2], ro[2];
    for (int z = 0; z < 2; ++z) {      // Cm::sumchecker (cm.rs:205-342), twice
        const u64 rc = challenge(T);
        std::vector<u64> rcps; { u64 p = 1; for (size_t i = 0; i < (size_t)L * (4 + 4 * n_M) + 2; ++i) { rcps.push_back(p); p = Fm::hmul(p, rc); } }
        const size_t zi = (size_t)L * (4 + 4 * n_M);
        LF_CUDA(cudaMemsetAsync(GU, 0, 2 * N * PD * 8, E.st()));
        for (int li = 0; li < L; ++li) { const lf_plus_rg& I = *inst[li]; const size_t base = (size_t)li * (4 + 4 * n_M);
            Lin4 w{Fm::to_mont(rcps[base]), rcps[base + 1], Fm::to_mont(rcps[base + 2]), Fm::to_mont(rcps[base + 3])};
            E.launch("k_plus_lin4", [&] { k_plus_lin4<<<Eng::blocks_for(n * PD, 256), 256, 0, E.st()>>>(I.tau, I.mtau_codes, I.f, d_h[li], n, w, li > 0 ? 1 : 0, GU); });
            for (int mi = 0; mi < n_M; ++mi) { const size_t idx = base + 4 + 4 * mi;      // M_i (rc^a tau + rc^b m_tau + rc^c f + rc^d h): one sparse product per (instance, matrix)
                Lin4 wz{Fm::to_mont(rcps[idx]), rcps[idx + 1], Fm::to_mont(rcps[idx + 2]), Fm::to_mont(rcps[idx + 3])};
                E.launch("k_plus_lin4", [&] { k_plus_lin4<<<Eng::blocks_for(n * PD, 256), 256, 0, E.st()>>>(I.tau, I.mtau_codes, I.f, d_h[li], n, wz, 0, Z); });
                E.launch("k_plus_spmv_acc", [&] { k_plus_spmv_acc<<<Eng::blocks_for(Mr[mi].nrows, 128), 128, 0, E.st()>>>(Mr[mi].row_ptr, Mr[mi].col, Mr[mi].val, Mr[mi].nrows, Z, Fm::r2(), GU); }); }
        }
        std::vector<u64> scal(tc0.size() * l);
        for (size_t i = 0; i < tc0.size(); ++i) { u64 dpw = 1; const u64 base = Fm::add(Fm::hmul(rcps[zi], tc0[i]), Fm::hmul(rcps[zi + 1], tc1[i]));
            for (int a = 0; a < l; ++a) { scal[i * l + a] = Fm::to_mont(Fm::hmul(base, dpw)); dpw = Fm::hmul(dpw, PD / 2); } }
        E.h2d(d_scal, scal.data(), scal.size() * 8);
        E.launch("k_plus_tz", [&] { k_plus_tz<<<Eng::blocks_for(nt, 128), 128, 0, E.st()>>>(d_scal, d_sp, kd, l, nt, GU + N * PD); });
        // MLSumcheck::prove_as_subprotocol, degree 2
        absorb_field(T, (u64)nvars); absorb_field(T, 2);
        msgs[z].assign((size_t)nvars * 3 * PD, 0);
        const u64 *sc_cur = sc0, *gu_cur = GU; u64 *sc_a = scn, *sc_b = scm, *gu_a = GUn, *gu_b = GUm; size_t len = N; u64 r_prev = 0;
        for (int i = 0; i < nvars; ++i) {
            if (i > 0) { const size_t n_out = len / 2; const u64 rm = Fm::to_mont(r_prev);
                E.launch("k_plus_fold", [&] { k_plus_fold<<<dim3(Eng::blocks_for(n_out, 256), 2), 256, 0, E.st()>>>(sc_cur, len, sc_a, n_out, rm); });
                E.launch("k_plus_fold", [&] { k_plus_fold_ring<<<dim3(Eng::blocks_for(n_out * PD, 256), 2), 256, 0, E.st()>>>(gu_cur, len, gu_a, n_out, rm); });
                sc_cur = sc_a; gu_cur = gu_a; std::swap(sc_a, sc_b); std::swap(gu_a, gu_b); len = n_out; }
            const size_t n_pairs = len / 2; const unsigned nblk = (unsigned)std::min<size_t>(std::max<size_t>((n_pairs * PD + 255) / 256, 1), 148 * 16);
            u64* partial = E.partial_dev((size_t)nblk * 3 * PD); u64* d_out = E.small_dev(3 * PD);
            E.launch("k_plus_cm_round", [&] { k_plus_cm_round<<<nblk, 256, 0, E.st()>>>(sc_cur, len, gu_cur, gu_cur + len * PD, n_pairs, partial); });
            E.reduce_partials(partial, (int)nblk, 3 * PD, d_out);
            u64* msg = msgs[z].data() + (size_t)i * 3 * PD; E.download_words(d_out, 3 * PD, msg);
            T.absorb_slice(msg, 3);
            r_prev = challenge(T); absorb_field(T, r_prev); ro[z].push_back(r_prev);
        }
        // evaluations of the individual tables at ro (cm.rs:315-337), through eq(ro, .) and w_i = M_i^T eq(ro, .)
        eq_table(E, ro[z], eq_ro, N);
        evals[z].assign((size_t)L * nE * 4 * PD, 0);
        std::vector<u64*> d_w; for (int mi = 0; mi < n_M; ++mi) { u64* wv = alloc(std::max<size_t>(Ms[mi].ncols, 1) * PD);
            E.launch("k_plus_mt_eq", [&] { k_plus_mt_eq<<<Eng::blocks_for(Ms[mi].ncols, 128), 128, 0, E.st()>>>(eq_ro, Ms[mi].col_ptr, Ms[mi].erow, Ms[mi].val, Ms[mi].ncols, wv); }); d_w.push_back(wv); }
        { WsumBatch wb(E);
          for (int li = 0; li < L; ++li) { const lf_plus_rg& I = *inst[li];
            for (size_t e = 0; e < nE; ++e) { const bool ring = e > 0; const u64* W = ring ? d_w[e - 1] : eq_ro; u64* out = evals[z].data() + ((li * nE + e) * 4) * PD;
                wb.small(W, ring, I.tau, n, out); wb.mono(W, ring, I.mtau_codes, I.code_pitch, n, 1, out + PD);
                wb.general(W, ring, cp, nullptr, I.f, n, 1, out + 2 * PD); wb.general(W, ring, cp, nullptr, d_h[li], n, 1, out + 3 * PD); } }
          wb.flush(); }
        T.absorb_slice(evals[z].data(), evals[z].size() / PD);      // absorb_evaluations (cm.rs:581-588)
    }
    // g_l = s0 tau + s1 m_tau + s2 f + h (cm.rs:165-182)
    if (g_host || g_dev) { GArgs ga; for (int o = 0; o < PD; ++o) { ga.s0[o] = (short)to_small(ch.s[o]); ga.s1[o] = (short)to_small(ch.s[PD + o]); ga.s2m[o] = Fm::to_mont(ch.s[2 * PD + o]); }
        u64* d_g = g_dev ? g_dev : alloc(n * PD);      // (g_dev: the summed g stays on the device, Mlin::mlin)
        for (int li = 0; li < L; ++li) { const lf_plus_rg& I = *inst[li];
            E.launch("k_plus_g", [&] { k_plus_g<<<Eng::blocks_for(n, 128), 128, 0, E.st()>>>(I.tau, I.mtau_codes, I.f, d_h[li], n, ga, sum_g && li > 0 ? 1 : 0, d_g); });
            if (g_host && (!sum_g || li == L - 1)) { LF_CUDA(cudaMemcpyAsync(g_host + (sum_g ? 0 : (size_t)li * n * PD), d_g, n * PD * 8, cudaMemcpyDeviceToHost, E.st())); E.sync(); } } }
    for (auto* v : {&comh, &msgs[0], &msgs[1], &evals[0], &evals[1]}) proof.insert(proof.end(), v->begin(), v->end());
    std::vector<const u64*> fc(L); for (int li = 0; li < L; ++li) fc[li] = inst[li]->fcoms.data();
    comx = comx_words(ch.s.data(), L, kappa, nE, fc.data(), comh.data(), evals, ro);
}
// RgInstance::from_f (rgchk.rs:259-336)
lf_plus_rg* rg_from_f_core(Eng& E, lf_ctx* c, const lf_plus_mat* A, const uint64_t* f, uint64_t n, uint64_t b, int32_t k, int32_t l, bool f_on_device = false) {
    {
        const bool tm = std::getenv("LF_PLUS_TIMING") != nullptr; auto t_last = std::chrono::steady_clock::now();      // diagnostic: phase times on stderr
        auto mark = [&](const char* what) { if (!tm) return; E.sync(); auto now = std::chrono::steady_clock::now(); std::fprintf(stderr, "from_f %-16s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count()); t_last = now; };
        if (!A || !f || n != A->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "from_f: witness length differs from the matrix width");
        if (k < 1 || k > 16 || l < 1 || l > 64 || b < 2 || b > (1u << 20)) throw LfException(LF_ERR_INVALID_ARG, "from_f: decomposition parameters out of range");
        const size_t kappa = A->kappa, n_tau = kappa * (size_t)k * PD * l * PD;
        if (n_tau >= n) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "from_f: n must exceed kappa * k * d * l * d (utils.rs:34-40)");
        std::unique_ptr<lf_plus_rg> I(new lf_plus_rg); I->n = n; I->kappa = kappa; I->k = k; I->l = l; I->b = b; I->code_pitch = (n + 15) / 16 * 16;
        // instance buffers come from the context's block cache: a steady stream of from_f / free pairs performs no cudaMalloc
        struct Guard { Eng& E; lf_plus_rg* p; ~Guard() { if (p) { E.dfree(p->codes); E.dfree(p->mtau_codes); E.dfree(p->tau); E.dfree(p->f); } } } g{E, I.get()};
        I->f = E.dalloc<u64>(n * PD); I->codes = E.dalloc<unsigned char>((size_t)k * PD * I->code_pitch); I->mtau_codes = E.dalloc<unsigned char>(I->code_pitch); I->tau = E.dalloc<signed char>(n);
        mark("validate+alloc");
        LF_CUDA(cudaMemcpyAsync(I->f, f, n * PD * 8, f_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, E.st()));
        mark("h2d");
        // D_f = decompose_to_vec(cf(f), b, k), M_f = exp(D_f) as exponent codes
        E.launch("k_plus_digit_codes", [&] { k_plus_digit_codes<<<Eng::blocks_for(n * PD, 256), 256, 0, E.st()>>>(I->f, n, (long long)b, k, I->codes, I->code_pitch, c->d_err); });
        { int h = 0; LF_CUDA(cudaMemcpyAsync(&h, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, E.st())); E.sync();      // (the kernel also validates f: every word is read there anyway)
          if (h) { LF_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), E.st()));
                   if (h == 3) throw LfException(LF_ERR_INVALID_ARG, "from_f witness: non-canonical field element");
                   throw LfException(LF_ERR_DOES_NOT_FIT, "from_f: a coefficient of f does not fit k digits in base b inside (-d/2, d/2)"); } }
        mark("digits");
        // comM_f[kk] = A * M_f[kk]: rotations only.  com = hconcat: row r, column kk * d + c
        I->comM.assign((size_t)k * kappa * PD * PD, 0); std::vector<u64> com(kappa * (size_t)k * PD * PD);
        { WsumBatch wb(E);
          for (size_t r = 0; r < kappa; ++r) wb.mono(A->d + r * n * PD, true, I->codes, I->code_pitch, n, (size_t)k * PD, com.data() + r * k * PD * PD);
          wb.flush(); }
        for (size_t r = 0; r < kappa; ++r) for (int kk = 0; kk < k; ++kk) std::memcpy(&I->comM[(((size_t)kk * kappa + r) * PD) * PD], &com[(r * k + kk) * PD * PD], 8 * PD * PD);
        mark("A*M_f");
        // tau = split(com, n, d/2, l) (utils.rs:12-43): l balanced base-(d/2) digits of every coefficient, digit-major per entry; host work on kappa * k * d elements
        I->tau_host.assign(n, 0); std::vector<signed char> tau8(n, 0); std::vector<unsigned char> mt(I->code_pitch, 0);      // exp(0) = X^0
        int64_t dg[64];
        for (size_t e = 0; e < kappa * (size_t)k * PD; ++e) for (int cf = 0; cf < PD; ++cf) {
            const u64 v = com[e * PD + cf]; const bool neg = v > (Fm::P - 1) / 2; const u64 mag = neg ? Fm::P - v : v;
            if (!balanced_digits(neg ? -(int64_t)mag : (int64_t)mag, PD / 2, l, dg)) throw LfException(LF_ERR_DOES_NOT_FIT, "split: l digits in base d/2 do not hold a commitment coefficient");
            for (int i = 0; i < l; ++i) { const size_t pos = (e * l + i) * PD + cf; tau8[pos] = (signed char)dg[i]; I->tau_host[pos] = dg[i] < 0 ? Fm::P - (u64)(-dg[i]) : (u64)dg[i]; mt[pos] = (unsigned char)((dg[i] + PD) & (PD - 1)); }
        }
        LF_CUDA(cudaMemcpyAsync(I->tau, tau8.data(), n, cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemcpyAsync(I->mtau_codes, mt.data(), I->code_pitch, cudaMemcpyHostToDevice, E.st()));
        mark("split");
        // cm_f = A f, C_Mf = A tau, cm_mtau = A m_tau (rgchk.rs:322-327)
        I->fcoms.assign(3 * kappa * PD, 0);
        u64* cp = E.dalloc<u64>(2); const u64 cph[2] = {0, n}; E.h2d(cp, cph, 16);
        { WsumBatch wb(E);
          for (size_t r = 0; r < kappa; ++r) { const u64* Ar = A->d + r * n * PD;
            wb.general(Ar, true, cp, nullptr, I->f, n, 1, &I->fcoms[r * PD]); wb.small(Ar, true, I->tau, n, &I->fcoms[(kappa + r) * PD]); wb.mono(Ar, true, I->mtau_codes, I->code_pitch, n, 1, &I->fcoms[(2 * kappa + r) * PD]); }
          wb.flush(); }
        E.sync(); E.dfree(cp);
        mark("commitments");
        g.p = nullptr; return I.release();
    }
}

// MLSumcheck::verify_as_subprotocol (sumcheck.rs:84-104, verifier.rs:92-123) on ring-valued messages with base-field challenges
void sumcheck_verify_host(Tr& T, size_t nv, int deg, const u64* claimed, const u64* msgs, std::vector<u64>& point, u64* expected) {
    const int ne = deg + 1;
    absorb_field(T, (u64)nv); absorb_field(T, (u64)deg); point.clear();
    for (size_t i = 0; i < nv; ++i) { T.absorb_slice(msgs + i * ne * PD, ne); const u64 r = challenge(T); point.push_back(r); absorb_field(T, r); }
    std::memcpy(expected, claimed, 8 * PD);
    for (size_t i = 0; i < nv; ++i) { const u64* msg = msgs + i * ne * PD;
        for (int c = 0; c < PD; ++c) if (Fm::add(msg[c], msg[PD + c]) != expected[c]) throw LfException(LF_ERR_SUMCHECK_FAILED, "sumcheck round sum mismatch (SumCheckError::SumCheckFailed)");
        u64 lag[8];
        for (int a = 0; a < ne; ++a) { u64 num = 1, den = 1; for (int b = 0; b < ne; ++b) if (b != a) { num = Fm::hmul(num, Fm::sub(point[i], (u64)b)); den = Fm::hmul(den, a > b ? (u64)(a - b) : Fm::P - (u64)(b - a)); } lag[a] = Fm::hmul(num, Fm::hpow(den, Fm::P - 2)); }
        for (int c = 0; c < PD; ++c) { u64 v = 0; for (int a = 0; a < ne; ++a) v = Fm::add(v, Fm::hmul(msg[a * PD + c], lag[a])); expected[c] = v; } }
}
// ComR1CS::linearize (r1cs.rs:72-134).  image: [nvars] ro[nvars] messages[nvars][4][16] v | va | vb | vc
std::vector<u64> r1cs_linearize_core(Eng& E, Tr& T, const lf_csr* abc, const u64* f, size_t n, bool f_on_device = false) {
    const int nvars = plus_ceil_log2(n); const size_t N = (size_t)1 << nvars;
    std::vector<void*> blocks; std::vector<DevCsr> Mr;
    struct Cleanup { Eng& E; std::vector<DevCsr>& b; std::vector<void*>& blk; ~Cleanup() { for (auto& s : b) s.free(E); for (void* p : blk) E.dfree(p); } } cl{E, Mr, blocks};
    auto alloc = [&](size_t words) { u64* p = E.dalloc<u64>(words); blocks.push_back(p); return p; };
    if (!f_on_device) check_canonical(f, n * PD, "linearize witness");
    for (int i = 0; i < 3; ++i) { if (!find_pinned(E.c, abc[i])) validate_csr(abc[i]);      // (a pinned matrix was validated when it was pinned)
        if (abc[i].ncols != n || abc[i].nrows > N) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "linearize: R1CS matrix does not match the witness"); Mr.push_back(upload_csr(E, abc[i])); }
    u64 *d_f = alloc(n * PD), *G = alloc(3 * N * PD), *Gn = alloc(3 * N * PD / 2), *Gm = alloc(3 * N * PD / 4 + PD), *eq = alloc(N), *eqn = alloc(N / 2 + 1), *eqm = alloc(N / 4 + 1), *cp = alloc(2);
    LF_CUDA(cudaMemcpyAsync(d_f, f, n * PD * 8, f_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, E.st())); LF_CUDA(cudaMemsetAsync(G, 0, 3 * N * PD * 8, E.st()));
    for (int i = 0; i < 3; ++i) E.launch("k_plus_spmv_acc", [&] { k_plus_spmv_acc<<<Eng::blocks_for(Mr[i].nrows, 128), 128, 0, E.st()>>>(Mr[i].row_ptr, Mr[i].col, Mr[i].val, Mr[i].nrows, d_f, Fm::r2(), G + (size_t)i * N * PD); });
    std::vector<u64> r(nvars); for (auto& x : r) x = challenge(T);
    eq_table(E, r, eq, N);
    std::vector<u64> img = {(u64)nvars}, ro, msgs((size_t)nvars * 4 * PD);
    absorb_field(T, (u64)nvars); absorb_field(T, 3);
    const u64 *g_cur = G, *e_cur = eq; u64 *g_a = Gn, *g_b = Gm, *e_a = eqn, *e_b = eqm; size_t len = N; u64 r_prev = 0;
    for (int i = 0; i < nvars; ++i) {
        if (i > 0) { const size_t n_out = len / 2; const u64 rm = Fm::to_mont(r_prev);
            E.launch("k_plus_fold", [&] { k_plus_fold<<<dim3(Eng::blocks_for(n_out, 256), 1), 256, 0, E.st()>>>(e_cur, len, e_a, n_out, rm); });
            E.launch("k_plus_fold", [&] { k_plus_fold_ring<<<dim3(Eng::blocks_for(n_out * PD, 256), 3), 256, 0, E.st()>>>(g_cur, len, g_a, n_out, rm); });
            e_cur = e_a; g_cur = g_a; std::swap(e_a, e_b); std::swap(g_a, g_b); len = n_out; }
        const size_t n_pairs = len / 2; const unsigned nblk = (unsigned)std::min<size_t>(std::max<size_t>((n_pairs * PD + 255) / 256, 1), 148 * 8);
        u64* partial = E.partial_dev((size_t)nblk * 4 * PD); u64* d_out = E.small_dev(4 * PD);
        E.launch("k_plus_r1cs_round", [&] { k_plus_r1cs_round<<<nblk, 256, 0, E.st()>>>(e_cur, g_cur, len, n_pairs, Fm::r2(), partial); });
        E.reduce_partials(partial, (int)nblk, 4 * PD, d_out);
        u64* msg = msgs.data() + (size_t)i * 4 * PD; E.download_words(d_out, 4 * PD, msg);
        T.absorb_slice(msg, 4);
        r_prev = challenge(T); absorb_field(T, r_prev); ro.push_back(r_prev);
    }
    // v = MLE(f)(ro), va, vb, vc = MLE(A|B|C f)(ro)
    eq_table(E, ro, eq, N); const u64 cph[2] = {0, n}; E.h2d(cp, cph, 16);
    u64* cpN = alloc(2); const u64 cphN[2] = {0, N}; E.h2d(cpN, cphN, 16);
    std::vector<u64> v4(4 * PD);
    { WsumBatch wb(E); for (int q = 0; q < 4; ++q) wb.general(eq, false, q == 0 ? cp : cpN, nullptr, q == 0 ? d_f : G + (size_t)(q - 1) * N * PD, q == 0 ? n : N, 1, &v4[q * PD]); wb.flush(); }
    T.absorb_slice(v4.data(), 4);
    for (auto* x : {&ro, &msgs, &v4}) img.insert(img.end(), x->begin(), x->end());
    return img;
}
}  // namespace

namespace lf { void plus_forget_ctx(lf_ctx* c) {      // the resident copies are blocks of the context and are released with it
    std::lock_guard<std::mutex> g(pinned_mutex()); auto& v = pinned_list();
    for (size_t i = v.size(); i-- > 0;) if (v[i].ctx == c) v.erase(v.begin() + i);
} }

extern "C" {

void lf_transcript_get_challenge_base(lf_transcript* t, uint64_t* out1) { pguard(nullptr, [&] { *out1 = challenge(tr_of(t)); }); }

lf_status lf_plus_set_check(lf_ctx* c, lf_transcript* t, int32_t nvars, const lf_plus_set* sets, int32_t n_sets, const lf_csr* M, int32_t n_M, uint64_t* out, uint64_t out_cap, uint64_t* out_len) {
    return pguard(c, [&] {
        need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (n_sets < 1 || !sets || n_M < 0 || (n_M && !M) || !out_len) throw LfException(LF_ERR_INVALID_ARG, "set check: null / empty arguments");
        std::vector<DevSparse> store; store.reserve((size_t)n_sets + n_M); std::vector<DevSet> mats, vecs; std::vector<DevSparse> Ms; SetCheckResult R;
        struct Cleanup { Eng& E; std::vector<DevSparse>& a; std::vector<DevSparse>& b; SetCheckResult& r; ~Cleanup() { for (auto& s : a) s.free(E); for (auto& s : b) s.free(E); r.free(E); } } cl{E, store, Ms, R};
        for (int i = 0; i < n_sets; ++i) {
            if (sets[i].kind != 0 && sets[i].kind != 1) throw LfException(LF_ERR_INVALID_ARG, "set check: unknown set kind");
            store.push_back(sets[i].kind == 0 ? upload_by_columns(E, sets[i].m) : upload_dense_vector(E, sets[i].v, sets[i].n));
            DevSet s; s.gen = &store.back(); s.nrows = store.back().nrows; s.ncols = store.back().ncols; (sets[i].kind == 0 ? mats : vecs).push_back(s);
        }
        for (int i = 0; i < n_M; ++i) Ms.push_back(upload_by_columns(E, M[i]));
        R = set_check_core(E, T, nvars, mats, vecs, Ms);
        const std::vector<u64> w = R.words(); *out_len = w.size();
        if (!out || out_cap < w.size()) throw LfException(LF_ERR_INVALID_ARG, "set check: output buffer too small");
        std::memcpy(out, w.data(), w.size() * 8);
    });
}
lf_status lf_plus_set_check_verify(lf_transcript* t, const uint64_t* words, uint64_t len) {
    return pguard(nullptr, [&] { Tr& T = tr_of(t); verify_set_image(T, parse_set_image(words, len)); });
}

lf_status lf_plus_mat_create(lf_ctx* c, uint64_t kappa, uint64_t n, const uint64_t* host, lf_plus_mat** out) {
    *out = nullptr;
    return pguard(c, [&] { need_frog(c); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (!host || !kappa || !n) throw LfException(LF_ERR_INVALID_ARG, "matrix: null / empty"); check_canonical(host, kappa * n * PD, "matrix");
        std::unique_ptr<lf_plus_mat> A(new lf_plus_mat); A->kappa = kappa; A->n = n; LF_CUDA(cudaMalloc(&A->d, kappa * n * PD * 8));
        LF_CUDA(cudaMemcpyAsync(A->d, host, kappa * n * PD * 8, cudaMemcpyHostToDevice, E.st())); E.sync(); *out = A.release(); });
}
void lf_plus_mat_free(lf_ctx* c, lf_plus_mat* a) { if (a) { if (c) cudaStreamSynchronize(c->stream); cudaFree(a->d); delete a; } }

lf_status lf_plus_rg_from_f(lf_ctx* c, const lf_plus_mat* A, const uint64_t* f, uint64_t n, uint64_t b, int32_t k, int32_t l, lf_plus_rg** out) {
    *out = nullptr;
    return pguard(c, [&] { need_frog(c); LF_CUDA(cudaSetDevice(c->device)); Eng E(c); *out = rg_from_f_core(E, c, A, f, n, b, k, l); });
}
lf_status lf_plus_rg_read(const lf_plus_rg* I, uint64_t* tau, uint64_t* fcoms, uint64_t* comM) {
    return pguard(nullptr, [&] { if (!I) throw LfException(LF_ERR_INVALID_ARG, "null instance");
        if (tau) std::memcpy(tau, I->tau_host.data(), I->n * 8); if (fcoms) std::memcpy(fcoms, I->fcoms.data(), I->fcoms.size() * 8); if (comM) std::memcpy(comM, I->comM.data(), I->comM.size() * 8); });
}
void lf_plus_rg_free(lf_ctx* c, lf_plus_rg* I) { if (I && c) { Eng E(c); E.dfree(I->codes); E.dfree(I->mtau_codes); E.dfree(I->tau); E.dfree(I->f); delete I; } }      // stream order protects pending readers of the re-issued blocks

lf_status lf_plus_range_check(lf_ctx* c, lf_transcript* t, int32_t nvars, lf_plus_rg* const* inst, int32_t L, const lf_csr* M, int32_t n_M, uint64_t* out, uint64_t out_cap, uint64_t* out_len) {
    return pguard(c, [&] {
        need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (L < 1 || !inst || !inst[0] || n_M < 0 || (n_M && !M) || !out_len) throw LfException(LF_ERR_INVALID_ARG, "range check: null / empty arguments");
        std::vector<DevSparse> Ms; SetCheckResult R;
        struct Cleanup { Eng& E; std::vector<DevSparse>& a; SetCheckResult& r; ~Cleanup() { for (auto& s : a) s.free(E); r.free(E); } } cl{E, Ms, R};
        for (int i = 0; i < n_M; ++i) Ms.push_back(upload_by_columns(E, M[i]));
        const std::vector<u64> img = range_check_core(E, T, nvars, inst, L, Ms, R);
        *out_len = img.size();
        if (!out || out_cap < img.size()) throw LfException(LF_ERR_INVALID_ARG, "range check: output buffer too small");
        std::memcpy(out, img.data(), img.size() * 8);
    });
}
lf_status lf_plus_range_check_verify(lf_transcript* t, const uint64_t* w, uint64_t len) {
    return pguard(nullptr, [&] {
        Tr& T = tr_of(t);
        if (!w || len < 10) throw LfException(LF_ERR_INCORRECT_LENGTH, "range-check image too short");
        const size_t L = w[0], k = w[1], kappa = w[3];
        if (L < 1 || L > 64 || k < 1 || k > 16 || kappa < 1 || kappa > 4096) throw LfException(LF_ERR_INVALID_ARG, "range-check image: implausible header");
        const SetImage s = parse_set_image(w + 5, len - 5);
        const size_t nE = 1 + s.n_M, per = PD + nE + 2 * nE * PD + 3 * kappa * PD, off = 5 + s.words;
        if (len < off + L * per) throw LfException(LF_ERR_INCORRECT_LENGTH, "range-check image truncated");
        if ((size_t)s.n_mat != L * k || (size_t)s.n_vec != L || s.ncols != PD) throw LfException(LF_ERR_INVALID_ARG, "range-check image: set-check shape does not match L, k");
        check_canonical(w + off, L * per, "range-check image");
        verify_set_image(T, s);
        for (size_t l = 0; l < L; ++l) { const u64* p = w + off + l * per; const u64 *a = p + PD, *c = a + nE + nE * PD; for (size_t i = 0; i < nE; ++i) absorb_field(T, a[i]); T.absorb_slice(c, nE); }
        for (size_t l = 0; l < L; ++l) { const u64* p = w + off + l * per; const u64 *v = p, *a = p + PD, *b = a + nE, *c = b + nE * PD;
            for (size_t i = 0; i < nE; ++i) if (ct_psi(b + i * PD) != a[i]) throw LfException(LF_ERR_RECOMPOSED, "range check: ct(psi b) != a (RangeCheckError::PsiCheckAB)");
            for (size_t ni = 0; ni < nE; ++ni) for (int j = 0; j < PD; ++j) {
                u64 uc[PD] = {0}; u64 dpw = 1;
                for (size_t i = 0; i < k; ++i) { const u64* u = s.e + ((ni * s.n_mat + k * l + i) * s.ncols + j) * PD; for (int cf = 0; cf < PD; ++cf) uc[cf] = Fm::add(uc[cf], Fm::hmul(u[cf], dpw)); dpw = Fm::hmul(dpw, PD / 2); }
                if (ct_psi(uc) != (ni == 0 ? v[j] : c[ni * PD + j])) throw LfException(LF_ERR_RECOMPOSED, "range check: ct(psi sum d'^i u_i) != v (RangeCheckError::PsiCheckVU)");
            }
        }
    });
}
uint64_t lf_plus_comx_words(int32_t nvars, int32_t L, uint64_t kappa, int32_t n_M) { return (uint64_t)L * kappa * PD + 2 * (uint64_t)nvars + (uint64_t)L * (1 + n_M) * 2 * PD; }
lf_status lf_plus_cm_prove(lf_ctx* c, lf_transcript* t, int32_t nvars, lf_plus_rg* const* inst, int32_t L, const lf_csr* M, int32_t n_M,
                           uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* comx, uint64_t* g_host) {
    return pguard(c, [&] {
        need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (L < 1 || !inst || !inst[0] || n_M < 0 || (n_M && !M) || !proof_len) throw LfException(LF_ERR_INVALID_ARG, "cm: null / empty arguments");
        std::vector<u64> pw, xw; cm_prove_core(E, T, nvars, inst, L, M, n_M, pw, xw, g_host);
        *proof_len = pw.size();
        if (!proof || proof_cap < pw.size()) throw LfException(LF_ERR_INVALID_ARG, "cm: proof buffer too small");
        std::memcpy(proof, pw.data(), pw.size() * 8); if (comx) std::memcpy(comx, xw.data(), xw.size() * 8);
    });
}
// CmProof::verify (cm.rs:349-535), host.  comx_out (lf_plus_comx_words) receives the ComX the verifier derives
lf_status lf_plus_cm_verify(lf_transcript* t, const uint64_t* w, uint64_t len, int32_t n_M, uint64_t* comx_out) {
    return pguard(nullptr, [&] {
        Tr& T = tr_of(t);
        if (!w || len < 10) throw LfException(LF_ERR_INCORRECT_LENGTH, "cm proof image too short");
        const size_t L = w[0], k = w[1], l = w[2], kappa = w[3];
        if (L < 1 || L > 64 || k < 1 || k > 16 || l < 1 || l > 64 || kappa < 1 || kappa > 4096) throw LfException(LF_ERR_INVALID_ARG, "cm proof image: implausible header");
        const SetImage s = parse_set_image(w + 5, len - 5);
        if (s.n_M != n_M) throw LfException(LF_ERR_INVALID_ARG, "cm proof image: number of matrices differs");
        const size_t nE = 1 + s.n_M, per = PD + nE + 2 * nE * PD + 3 * kappa * PD, dlen = 5 + s.words + L * per, nv = s.nvars;
        const size_t n1 = L * kappa * PD, n2 = nv * 3 * PD, n3 = L * nE * 4 * PD;
        if (len < dlen + n1 + 2 * n2 + 2 * n3) throw LfException(LF_ERR_INCORRECT_LENGTH, "cm proof image truncated");
        check_canonical(w + dlen, n1 + 2 * n2 + 2 * n3, "cm proof image");
        lf_transcript tt{LF_RING_FROG, &T};
        const lf_status rs = lf_plus_range_check_verify(&tt, w, dlen); if (rs != LF_OK) throw LfException(rs, "cm: range check rejected");
        const u64* comh = w + dlen; const u64* msgs[2] = {comh + n1, comh + n1 + n2}; const u64* ev[2] = {comh + n1 + 2 * n2, comh + n1 + 2 * n2 + n3};
        const CmChallenges ch = draw_cm_challenges(T, (int)k);
        T.absorb_slice(comh, L * kappa);
        const int log_kappa = plus_ceil_log2(kappa);
        std::vector<u64> c0(log_kappa), c1(log_kappa); for (auto& x : c0) x = challenge(T); for (auto& x : c1) x = challenge(T);
        const std::vector<u64> tc0 = tensor_host(c0), tc1 = tensor_host(c1);
        // u[l][ni] = sum over instance l's k sets and their d columns of e * s' (cm.rs:386-404); tensor(c) . comh (cm.rs:406-430)
        std::vector<u64> u(L * nE * PD, 0), tcch0(L * PD, 0), tcch1(L * PD, 0); u64 t[PD];
        for (size_t li = 0; li < L; ++li) for (size_t ni = 0; ni < nE; ++ni) for (size_t q = 0; q < k * PD; ++q) {
            hring_mul(t, s.e + ((ni * s.n_mat + li * k) * (size_t)s.ncols + q) * PD, &ch.sp[q * PD]); hring_add(&u[(li * nE + ni) * PD], &u[(li * nE + ni) * PD], t); }
        for (size_t li = 0; li < L; ++li) for (size_t i = 0; i < std::min(tc0.size(), kappa); ++i) { hring_axpy(&tcch0[li * PD], comh + (li * kappa + i) * PD, tc0[i]); hring_axpy(&tcch1[li * PD], comh + (li * kappa + i) * PD, tc1[i]); }
        // t(z) as dense vectors (cm.rs:590-601), evaluated at each sumcheck's point below
        auto t_z = [&](const std::vector<u64>& tc) { std::vector<u64> out(tc.size() * k * PD * l * PD * PD, 0); size_t idx = 0;
            for (u64 tci : tc) for (size_t j = 0; j < k * PD; ++j) { u64 dpw = 1; for (size_t a = 0; a < l; ++a) { const u64 sc = Fm::hmul(tci, dpw); dpw = Fm::hmul(dpw, PD / 2);
                for (int b = 0; b < PD; ++b, ++idx) for (int o = 0; o < PD; ++o) { const u64 v = ch.sp[j * PD + ((o - b) & (PD - 1))]; out[idx * PD + o] = Fm::hmul(sc, o >= b ? v : Fm::neg(v)); } } }
            return out; };
        const std::vector<u64> t0 = t_z(tc0), t1 = t_z(tc1);
        if (t0.size() / PD > ((size_t)1 << nv)) throw LfException(LF_ERR_INVALID_SIZE_BOUNDS, "cm: t(z) longer than 2^nvars");
        const u64* dcom_evals = w + 5 + s.words; std::vector<u64> ro[2]; std::vector<u64> evv[2];
        for (int z = 0; z < 2; ++z) {
            const u64 rc = challenge(T); const size_t zi = L * (4 + 4 * (size_t)s.n_M);
            std::vector<u64> rcps; { u64 p = 1; for (size_t i = 0; i < zi + 2; ++i) { rcps.push_back(p); p = Fm::hmul(p, rc); } }
            u64 claimed[PD] = {0};
            for (size_t li = 0; li < L; ++li) { const u64* p = dcom_evals + li * per; const u64 *a = p + PD, *b = a + nE, *cc = b + nE * PD; const size_t base = li * (4 + 4 * (size_t)s.n_M);
                for (size_t i = 0; i < nE; ++i) { const size_t idx = base + 4 * i;
                    claimed[0] = Fm::add(claimed[0], Fm::hmul(a[i], rcps[idx])); hring_axpy(claimed, b + i * PD, rcps[idx + 1]); hring_axpy(claimed, cc + i * PD, rcps[idx + 2]); hring_axpy(claimed, &u[(li * nE + i) * PD], rcps[idx + 3]); }
                hring_axpy(claimed, &tcch0[li * PD], rcps[zi]); hring_axpy(claimed, &tcch1[li * PD], rcps[zi + 1]); }
            // MLSumcheck::verify_as_subprotocol, degree 2
            absorb_field(T, (u64)nv); absorb_field(T, 2);
            for (size_t i = 0; i < nv; ++i) { T.absorb_slice(msgs[z] + i * 3 * PD, 3); const u64 r = challenge(T); ro[z].push_back(r); absorb_field(T, r); }
            u64 expected[PD]; std::memcpy(expected, claimed, sizeof expected);
            for (size_t i = 0; i < nv; ++i) { const u64* msg = msgs[z] + i * 3 * PD;
                for (int cf = 0; cf < PD; ++cf) if (Fm::add(msg[cf], msg[PD + cf]) != expected[cf]) throw LfException(LF_ERR_SUMCHECK_FAILED, "cm: sumcheck round sum mismatch");
                u64 lag[3]; for (int a = 0; a < 3; ++a) { u64 num = 1, den = 1; for (int b = 0; b < 3; ++b) if (b != a) { num = Fm::hmul(num, Fm::sub(ro[z][i], (u64)b)); den = Fm::hmul(den, a > b ? (u64)(a - b) : Fm::P - (u64)(b - a)); } lag[a] = Fm::hmul(num, Fm::hpow(den, Fm::P - 2)); }
                for (int cf = 0; cf < PD; ++cf) { u64 v = 0; for (int a = 0; a < 3; ++a) v = Fm::add(v, Fm::hmul(msg[a * PD + cf], lag[a])); expected[cf] = v; } }
            u64 t0_ro[PD], t1_ro[PD]; mle_eval_dense(t0, (int)nv, ro[z], t0_ro); mle_eval_dense(t1, (int)nv, ro[z], t1_ro);
            T.absorb_slice(ev[z], L * nE * 4);
            u64 eq = 1; for (size_t i = 0; i < nv; ++i) { const u64 xy = Fm::hmul(s.r[i], ro[z][i]); eq = Fm::hmul(eq, Fm::add(Fm::sub(Fm::sub(Fm::add(xy, xy), s.r[i]), ro[z][i]), 1)); }
            u64 evl[PD] = {0};
            for (size_t li = 0; li < L; ++li) { const u64* el = ev[z] + li * nE * 4 * PD; const size_t base = li * (4 + 4 * (size_t)s.n_M); u64 in[PD] = {0}, pr[PD];
                for (size_t q = 0; q < 4 * nE; ++q) hring_axpy(in, el + q * PD, rcps[base + q]);
                hring_axpy(evl, in, eq);
                hring_mul(pr, t0_ro, el); hring_axpy(evl, pr, rcps[zi]); hring_mul(pr, t1_ro, el); hring_axpy(evl, pr, rcps[zi + 1]); }
            if (std::memcmp(evl, expected, sizeof evl) != 0) throw LfException(LF_ERR_SUMCHECK_FAILED, "cm: evaluation claim mismatch (cm.rs:521)");
            evv[z].assign(ev[z], ev[z] + n3);
        }
        if (comx_out) { std::vector<const u64*> fc(L); for (size_t li = 0; li < L; ++li) fc[li] = dcom_evals + li * per + PD + nE + 2 * nE * PD;
            const std::vector<u64> xw = comx_words(ch.s.data(), L, kappa, nE, fc.data(), comh, evv, ro); std::memcpy(comx_out, xw.data(), xw.size() * 8); }
    });
}
// Mlin::mlin (mlin.rs:41-106): from_f on every witness, Cm::prove, the sums over the instances.  linb2x = cm_g[kappa][16] | ro[nvars][2] | vo[1 + n_M][2][16]
static void mlin_core(lf_ctx* c, lf_transcript* t, const lf_plus_mat* A, const uint64_t* const* srcs, bool on_device, int32_t L, uint64_t n, uint64_t b, int32_t k, int32_t l, const lf_csr* M, int32_t n_M,
                      uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* linb2x, uint64_t* g_host, uint64_t* g_dev) {
        need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (L < 1 || L > 64 || !srcs || !A || n_M < 0 || (n_M && !M) || !proof_len || !linb2x) throw LfException(LF_ERR_INVALID_ARG, "mlin: null / empty arguments");
        std::vector<lf_plus_rg*> inst; struct Cleanup { lf_ctx* c; std::vector<lf_plus_rg*>& v; ~Cleanup() { for (auto* p : v) lf_plus_rg_free(c, p); } } cl{c, inst};
        for (int i = 0; i < L; ++i) inst.push_back(rg_from_f_core(E, c, A, srcs[i], n, b, k, l, on_device));
        const int nvars = plus_ceil_log2(n); const size_t kappa = A->kappa, nE = 1 + (size_t)n_M;
        std::vector<u64> pw, xw; cm_prove_core(E, T, nvars, inst.data(), L, M, n_M, pw, xw, g_host, true, g_dev);
        // LinB2X: sums of the per-instance cm_g and vo, ro as it is
        const u64 *cmg = xw.data(), *ro = cmg + (size_t)L * kappa * PD, *vo = ro + 2 * (size_t)nvars;
        std::vector<u64> x(kappa * PD + 2 * (size_t)nvars + nE * 2 * PD, 0);
        for (int li = 0; li < L; ++li) { for (size_t i = 0; i < kappa * PD; ++i) x[i] = Fm::add(x[i], cmg[(size_t)li * kappa * PD + i]);
            for (size_t i = 0; i < nE * 2 * PD; ++i) x[kappa * PD + 2 * (size_t)nvars + i] = Fm::add(x[kappa * PD + 2 * (size_t)nvars + i], vo[(size_t)li * nE * 2 * PD + i]); }
        std::memcpy(&x[kappa * PD], ro, 2 * (size_t)nvars * 8);
        *proof_len = pw.size();
        if (!proof || proof_cap < pw.size()) throw LfException(LF_ERR_INVALID_ARG, "mlin: proof buffer too small");
        std::memcpy(proof, pw.data(), pw.size() * 8); std::memcpy(linb2x, x.data(), x.size() * 8);
}
lf_status lf_plus_mlin(lf_ctx* c, lf_transcript* t, const lf_plus_mat* A, const uint64_t* fs, int32_t L, uint64_t n, uint64_t b, int32_t k, int32_t l, const lf_csr* M, int32_t n_M,
                       uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* linb2x, uint64_t* g_host) {
    return pguard(c, [&] { if (!fs || L < 1 || L > 64) throw LfException(LF_ERR_INVALID_ARG, "mlin: null / empty arguments");
        std::vector<const uint64_t*> srcs(L); for (int i = 0; i < L; ++i) srcs[i] = fs + (size_t)i * n * PD;
        mlin_core(c, t, A, srcs.data(), false, L, n, b, k, l, M, n_M, proof, proof_cap, proof_len, linb2x, g_host, nullptr); });
}
// the same on device-resident witnesses (lf_plus_vec): g stays on the device as a new vector
lf_status lf_plus_mlin_v(lf_ctx* c, lf_transcript* t, const lf_plus_mat* A, const lf_plus_vec* const* fs, int32_t L, uint64_t b, int32_t k, int32_t l, const lf_csr* M, int32_t n_M,
                         uint64_t* proof, uint64_t proof_cap, uint64_t* proof_len, uint64_t* linb2x, lf_plus_vec** g_out) {
    if (g_out) *g_out = nullptr;
    return pguard(c, [&] { need_frog(c); if (!fs || L < 1 || L > 64 || !fs[0] || !g_out) throw LfException(LF_ERR_INVALID_ARG, "mlin: null / empty arguments");
        const size_t n = fs[0]->n; std::vector<const uint64_t*> srcs(L); for (int i = 0; i < L; ++i) { if (!fs[i] || fs[i]->n != n) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "mlin: witnesses of different lengths"); srcs[i] = fs[i]->d; }
        Eng E(c); std::unique_ptr<lf_plus_vec> g(new lf_plus_vec); g->n = n; g->d = E.dalloc<u64>(n * PD);
        try { mlin_core(c, t, A, srcs.data(), true, L, n, b, k, l, M, n_M, proof, proof_cap, proof_len, linb2x, nullptr, g->d); } catch (...) { E.dfree(g->d); throw; }
        *g_out = g.release(); });
}
// Decomp::decompose (decomp.rs:32-99).  r_pairs: nvars x 2 field elements (the points are constants of R on every path of the reference).
// proof = C0[kappa][16] | C1 | v0[1 + n_M][2][16] | v1; F_host (2 x n x 16: the two LinB witnesses) may be NULL
static void decompose_core(lf_ctx* c, const lf_plus_mat* A, const uint64_t* f, bool on_device, uint64_t n, const uint64_t* r_pairs, const lf_csr* M, int32_t n_M, uint64_t B, uint64_t* proof, uint64_t* F_host, u64* F_dev0, u64* F_dev1) {
    {
        need_frog(c); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (!A || !f || !r_pairs || !proof || n_M < 0 || (n_M && !M)) throw LfException(LF_ERR_INVALID_ARG, "decompose: null arguments");
        if (n != A->n) throw LfException(LF_ERR_WRONG_WITNESS_LEN, "decompose: witness length differs from the matrix width");
        if (B < 2 || B >> 40) throw LfException(LF_ERR_INVALID_ARG, "decompose: base out of range");
        const int nvars = plus_ceil_log2(n); const size_t N = (size_t)1 << nvars, kappa = A->kappa, nE = 1 + (size_t)n_M; check_canonical(r_pairs, 2 * (size_t)nvars, "decompose point");
        std::vector<DevSparse> Ms; std::vector<void*> blocks;
        struct Cleanup { Eng& E; std::vector<DevSparse>& a; std::vector<void*>& blk; ~Cleanup() { for (auto& s : a) s.free(E); for (void* p : blk) E.dfree(p); } } cl{E, Ms, blocks};
        auto alloc = [&](size_t words) { u64* p = E.dalloc<u64>(words); blocks.push_back(p); return p; };
        for (int i = 0; i < n_M; ++i) { Ms.push_back(upload_by_columns(E, M[i])); if (Ms.back().ncols != n || Ms.back().nrows > N) throw LfException(LF_ERR_LENGTHS_NOT_EQUAL, "decompose: M_i does not match the witness"); }
        u64 *d_f = alloc(n * PD), *F = alloc(2 * n * PD), *eq = alloc(2 * N), *cp = alloc(2); const u64 cph[2] = {0, n}; E.h2d(cp, cph, 16);
        LF_CUDA(cudaMemcpyAsync(d_f, f, n * PD * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, E.st()));
        E.launch("k_plus_split2", [&] { k_plus_split2<<<Eng::blocks_for(n * PD, 256), 256, 0, E.st()>>>(d_f, n * PD, (long long)B, F, F + n * PD, c->d_err); });
        { int h = 0; LF_CUDA(cudaMemcpyAsync(&h, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, E.st())); E.sync();
          if (h) { LF_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), E.st())); if (h == 3) throw LfException(LF_ERR_INVALID_ARG, "decompose: non-canonical field element"); throw LfException(LF_ERR_DOES_NOT_FIT, "decompose: a coefficient needs more than two digits in base B"); } }
        std::vector<u64> pt[2]; for (int i = 0; i < nvars; ++i) { pt[0].push_back(r_pairs[2 * i]); pt[1].push_back(r_pairs[2 * i + 1]); }
        eq_table(E, pt[0], eq, N); eq_table(E, pt[1], eq + N, N);
        std::vector<u64*> w[2]; for (int q = 0; q < 2; ++q) for (auto& m : Ms) { u64* wv = alloc(n * PD);
            E.launch("k_plus_mt_eq", [&] { k_plus_mt_eq<<<Eng::blocks_for(m.ncols, 128), 128, 0, E.st()>>>(eq + q * N, m.col_ptr, m.erow, m.val, m.ncols, wv); }); w[q].push_back(wv); }
        u64 *C = proof, *v = proof + 2 * kappa * PD;
        { WsumBatch wb(E);
          for (int z = 0; z < 2; ++z) { const u64* Fz = F + (size_t)z * n * PD;
            for (size_t r = 0; r < kappa; ++r) wb.general(A->d + r * n * PD, true, cp, nullptr, Fz, n, 1, C + ((size_t)z * kappa + r) * PD);      // C_z = A F_z
            for (size_t e = 0; e < nE; ++e) for (int q = 0; q < 2; ++q)      // (MLE(.)(r_a), MLE(.)(r_b)) of F_z and of every M_j F_z
                wb.general(e == 0 ? eq + q * N : w[q][e - 1], e != 0, cp, nullptr, Fz, n, 1, v + (((size_t)z * nE + e) * 2 + q) * PD); }
          wb.flush(); }
        if (F_host) { LF_CUDA(cudaMemcpyAsync(F_host, F, 2 * n * PD * 8, cudaMemcpyDeviceToHost, E.st())); E.sync(); }
        if (F_dev0) LF_CUDA(cudaMemcpyAsync(F_dev0, F, n * PD * 8, cudaMemcpyDeviceToDevice, E.st()));
        if (F_dev1) LF_CUDA(cudaMemcpyAsync(F_dev1, F + n * PD, n * PD * 8, cudaMemcpyDeviceToDevice, E.st()));
    }
}
lf_status lf_plus_decompose(lf_ctx* c, const lf_plus_mat* A, const uint64_t* f, uint64_t n, const uint64_t* r_pairs, const lf_csr* M, int32_t n_M, uint64_t B, uint64_t* proof, uint64_t* F_host) {
    return pguard(c, [&] { decompose_core(c, A, f, false, n, r_pairs, M, n_M, B, proof, F_host, nullptr, nullptr); });
}
// the same on a device-resident witness: the two digit vectors (the next accumulator, plus.rs:111-114) stay on the device
lf_status lf_plus_decompose_v(lf_ctx* c, const lf_plus_mat* A, const lf_plus_vec* f, const uint64_t* r_pairs, const lf_csr* M, int32_t n_M, uint64_t B, uint64_t* proof, lf_plus_vec** F0, lf_plus_vec** F1) {
    if (F0) *F0 = nullptr; if (F1) *F1 = nullptr;
    return pguard(c, [&] { need_frog(c); if (!f || !F0 || !F1) throw LfException(LF_ERR_INVALID_ARG, "decompose: null arguments");
        Eng E(c); std::unique_ptr<lf_plus_vec> a(new lf_plus_vec), b(new lf_plus_vec); a->n = b->n = f->n; a->d = E.dalloc<u64>(f->n * PD); b->d = E.dalloc<u64>(f->n * PD);
        try { decompose_core(c, A, f->d, true, f->n, r_pairs, M, n_M, B, proof, nullptr, a->d, b->d); } catch (...) { E.dfree(a->d); E.dfree(b->d); throw; }
        *F0 = a.release(); *F1 = b.release(); });
}
// DecompProof::verify (decomp.rs:102-126), host: recompose([C0, C1], B) = cm_f and the same for every evaluation pair.  LF_ERR_RECOMPOSED on mismatch
lf_status lf_plus_decompose_verify(const uint64_t* proof, uint64_t kappa, int32_t n_M, const uint64_t* cm_f, const uint64_t* v, uint64_t B) {
    return pguard(nullptr, [&] {
        if (!proof || !cm_f || !v || n_M < 0 || !kappa) throw LfException(LF_ERR_INVALID_ARG, "decompose verify: null arguments");
        const size_t nc = kappa * PD, nv = (size_t)(1 + n_M) * 2 * PD; const u64 Bm = B % Fm::P;
        check_canonical(proof, 2 * nc + 2 * nv, "decomposition proof"); check_canonical(cm_f, nc, "cm_f"); check_canonical(v, nv, "v");
        for (size_t i = 0; i < nc; ++i) if (Fm::add(proof[i], Fm::hmul(proof[nc + i], Bm)) != cm_f[i]) throw LfException(LF_ERR_RECOMPOSED, "decompose verify: commitments do not recompose");
        for (size_t i = 0; i < nv; ++i) if (Fm::add(proof[2 * nc + i], Fm::hmul(proof[2 * nc + nv + i], Bm)) != v[i]) throw LfException(LF_ERR_RECOMPOSED, "decompose verify: evaluations do not recompose");
    });
}
// ComR1CS::linearize (r1cs.rs:72-134) on the R1CS matrices abc[3] = A, B, C (n columns) and the committed witness f (n x 16)
lf_status lf_plus_r1cs_linearize(lf_ctx* c, lf_transcript* t, const lf_csr* abc, const uint64_t* f, uint64_t n, uint64_t* out, uint64_t out_cap, uint64_t* out_len) {
    return pguard(c, [&] { need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (!abc || !f || !out_len || n < 2) throw LfException(LF_ERR_INVALID_ARG, "linearize: null / empty arguments");
        const std::vector<u64> img = r1cs_linearize_core(E, T, abc, f, n); *out_len = img.size();
        if (!out || out_cap < img.size()) throw LfException(LF_ERR_INVALID_ARG, "linearize: output buffer too small");
        std::memcpy(out, img.data(), img.size() * 8); });
}
// Vec<R> resident on the device (the LinB witnesses that travel between the sub-protocols of PlusProver::prove)
lf_status lf_plus_vec_upload(lf_ctx* c, const uint64_t* host, uint64_t n, lf_plus_vec** out) {
    *out = nullptr;
    return pguard(c, [&] { need_frog(c); if (!host || !n) throw LfException(LF_ERR_INVALID_ARG, "vector: null / empty"); check_canonical(host, n * PD, "vector"); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        std::unique_ptr<lf_plus_vec> v(new lf_plus_vec); v->n = n; v->d = E.dalloc<u64>(n * PD);
        LF_CUDA(cudaMemcpyAsync(v->d, host, n * PD * 8, cudaMemcpyHostToDevice, E.st())); E.sync(); *out = v.release(); });
}
lf_status lf_plus_vec_download(lf_ctx* c, const lf_plus_vec* v, uint64_t* host) {
    return pguard(c, [&] { if (!v || !host) throw LfException(LF_ERR_INVALID_ARG, "vector: null"); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        LF_CUDA(cudaMemcpyAsync(host, v->d, v->n * PD * 8, cudaMemcpyDeviceToHost, E.st())); E.sync(); });
}
uint64_t lf_plus_vec_len(const lf_plus_vec* v) { return v ? v->n : 0; }
void lf_plus_vec_free(lf_ctx* c, lf_plus_vec* v) { if (v && c) { Eng E(c); E.dfree(v->d); delete v; } }
lf_status lf_plus_r1cs_linearize_v(lf_ctx* c, lf_transcript* t, const lf_csr* abc, const lf_plus_vec* f, uint64_t* out, uint64_t out_cap, uint64_t* out_len) {
    return pguard(c, [&] { need_frog(c); Tr& T = tr_of(t); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (!abc || !f || !out_len || f->n < 2) throw LfException(LF_ERR_INVALID_ARG, "linearize: null / empty arguments");
        const std::vector<u64> img = r1cs_linearize_core(E, T, abc, f->d, f->n, true); *out_len = img.size();
        if (!out || out_cap < img.size()) throw LfException(LF_ERR_INVALID_ARG, "linearize: output buffer too small");
        std::memcpy(out, img.data(), img.size() * 8); });
}
// ComR1CSProof::verify (r1cs.rs:136-162), host
lf_status lf_plus_r1cs_linearize_verify(lf_transcript* t, const uint64_t* w, uint64_t len) {
    return pguard(nullptr, [&] { Tr& T = tr_of(t);
        if (!w || len < 1 || w[0] < 1 || w[0] > 40) throw LfException(LF_ERR_INCORRECT_LENGTH, "linearization image: bad header");
        const size_t nv = w[0]; if (len < 1 + nv + nv * 4 * PD + 4 * PD) throw LfException(LF_ERR_INCORRECT_LENGTH, "linearization image truncated");
        check_canonical(w + 1, nv + nv * 4 * PD + 4 * PD, "linearization image");
        const u64 *msgs = w + 1 + nv, *v4 = msgs + nv * 4 * PD;
        std::vector<u64> r(nv); for (auto& x : r) x = challenge(T);
        std::vector<u64> point; u64 zero[PD] = {0}, expected[PD];
        sumcheck_verify_host(T, nv, 3, zero, msgs, point, expected);
        T.absorb_slice(v4, 4);
        u64 e = 1; for (size_t i = 0; i < nv; ++i) { const u64 xy = Fm::hmul(r[i], point[i]); e = Fm::hmul(e, Fm::add(Fm::sub(Fm::sub(Fm::add(xy, xy), r[i]), point[i]), 1)); }
        u64 pr[PD]; hring_mul(pr, v4 + PD, v4 + 2 * PD);
        for (int c = 0; c < PD; ++c) if (Fm::hmul(Fm::sub(pr[c], v4[3 * PD + c]), e) != expected[c]) throw LfException(LF_ERR_SUMCHECK_FAILED, "linearization: evaluation claim mismatch (r1cs.rs:159)"); });
}
// keep a resident copy of a static matrix; the host arrays must stay alive and unchanged until lf_plus_csr_unpin
lf_status lf_plus_csr_pin(lf_ctx* c, const lf_csr* m) {
    return pguard(c, [&] { need_frog(c); if (!m) throw LfException(LF_ERR_INVALID_ARG, "pin: null matrix"); LF_CUDA(cudaSetDevice(c->device)); Eng E(c);
        if (find_pinned(c, *m)) return;
        Pinned p{c, m->row_ptr, m->col, m->val, m->nrows, m->ncols, upload_by_columns_raw(E, *m), upload_csr_raw(E, *m)};
        std::lock_guard<std::mutex> g(pinned_mutex()); pinned_list().push_back(p); });
}
lf_status lf_plus_csr_unpin(lf_ctx* c, const lf_csr* m) {
    return pguard(c, [&] { if (!c || !m) throw LfException(LF_ERR_INVALID_ARG, "unpin: null argument"); Eng E(c); E.sync();
        std::lock_guard<std::mutex> g(pinned_mutex()); auto& v = pinned_list();
        for (size_t i = 0; i < v.size(); ++i) if (v[i].ctx == c && v[i].row_ptr == m->row_ptr && v[i].col == m->col && v[i].val == m->val) { v[i].by_col.free(E); v[i].by_row.free(E); v.erase(v.begin() + i); return; } });
}
lf_status lf_plus_tensor(const uint64_t* r, int32_t n, uint64_t* out) {
    return pguard(nullptr, [&] { if (!r || !out || n < 0 || n > 30) throw LfException(LF_ERR_INVALID_ARG, "tensor: bad arguments"); check_canonical(r, n, "tensor");
        std::vector<u64> res(1, 1);
        for (int i = 0; i < n; ++i) { std::vector<u64> nx; nx.reserve(res.size() * 2); for (u64 a : res) { nx.push_back(Fm::hmul(a, Fm::sub(1, r[i]))); nx.push_back(Fm::hmul(a, r[i])); } res.swap(nx); }
        std::memcpy(out, res.data(), res.size() * 8); });
}

}  // extern "C"
