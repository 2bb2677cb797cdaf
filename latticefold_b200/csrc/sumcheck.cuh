// Device-resident ring sumcheck prover state: MLSumcheck / IPForMLSumcheck of the reference
// (crates/latticefold/src/utils/sumcheck.rs:53-80, utils/sumcheck/prover.rs:19-162).  The Fiat-Shamir transcript stays on
// the host: one round = evaluate on the device -> (deg+1) ring elements to the host -> caller hashes -> challenge comes
// back and is applied (fix_variables) before the next evaluation.
#pragma once
#include "engine.cuh"
#include "sumcheck_wide.cuh"

struct lf_sumcheck {
    lf_ctx* ctx = nullptr;
    int nv = 0, deg = 0, round = 0, kind = 0;
    // table groups, each [count][D planes][pitch]; ping-pong buffers (cur -> nxt on every applied challenge)
    struct Group { lf_words *cur = nullptr, *nxt = nullptr, *alt = nullptr; size_t pitch = 0, stride = 0, nxt_cap = 0, alt_cap = 0; int count = 0; bool cur_owned = true; };
    Group dense;          // PRODUCTS/LIN: all MLEs.  FOLD: the first five
    Group fh;             // FOLD: the 2K*tau f-hat tables (slot-field valued); empty while still in digit form
    const int8_t* dig = nullptr; size_t dig_pitch = 0, dig_stride = 0;   // FOLD round 1 in the prover: borrowed int8 digits
    int n_f = 0;
    lf::u64* d_mu_pow = nullptr;      // n_f x TAU
    lf::u64* d_coef = nullptr;        // PRODUCTS/LIN term coefficients, n_terms x D
    // general term list (any number of tables / terms / factors: k_sc_terms); empty when the shape fits the register-resident kernel
    int* d_term_off = nullptr; int* d_term_idx = nullptr; int n_terms_general = 0;
    lf::ScGenericArgsT<lf::u64> gen;  // term structure (the table pointers are filled per launch, in the ring's word type)
    size_t len = 0;                   // current LOCAL table length (2^(nv - applied challenges) / ranks while sharded)
    int applied = 0;
    bool sharded = false;             // tables hold this rank's slab of the hypercube (high bits = rank)
    // FOLD from digits: after the first challenge the f-hat tables stay in digit form (round 2 is evaluated straight from the digits,
    // k_fold_sc_round2) and are materialised only after the second challenge (k_fold_digits2)
    bool fh_deferred = false; lf::u64 r1[16] = {0};
};

namespace lf {

template <class Rg> struct SumcheckDriver {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef HostRing<Rg> HR; typedef typename Rg::W W;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU;
    static W* wp(lf_words* p) { return reinterpret_cast<W*>(p); }
    static lf_words* ow(W* p) { return reinterpret_cast<lf_words*>(p); }
    Engine<Rg> E; lf_sumcheck* sc;
    SumcheckDriver(lf_ctx* c, lf_sumcheck* s) : E(c), sc(s) {}

    static void alloc_group(Engine<Rg>& E, lf_sumcheck::Group& g, int count, size_t len) {
        g.count = count; g.pitch = pitch_of(len); g.stride = g.pitch * D; g.cur = ow(E.template dalloc<W>((size_t)count * g.stride)); g.cur_owned = true;
    }
    void free_all() {
        for (lf_sumcheck::Group* g : {&sc->dense, &sc->fh}) { if (g->cur_owned && g->cur != g->nxt && g->cur != g->alt) E.dfree(g->cur); E.dfree(g->nxt); E.dfree(g->alt); g->cur = g->nxt = g->alt = nullptr; g->nxt_cap = g->alt_cap = 0; }
        E.dfree(sc->d_mu_pow); E.dfree(sc->d_coef); sc->d_mu_pow = sc->d_coef = nullptr;
        E.dfree(sc->d_term_off); E.dfree(sc->d_term_idx); sc->d_term_off = sc->d_term_idx = nullptr;
    }
    // mu (n_mu ring elements, slot-constant) -> mu_k^{d+1} for d < tau as slot-field elements (folding/utils.rs:293-322)
    void set_mu(const u64* mu_host, int n_mu) {
        sc->n_f = n_mu * TAU; if (sc->n_f > MAX_MU) throw LfException(LF_ERR_UNSUPPORTED, "FOLD: 2K*tau exceeds MAX_MU");
        std::vector<u64> pw((size_t)sc->n_f * TAU);
        for (int k = 0; k < n_mu; ++k) {
            const u64* m = mu_host + (size_t)k * D;     // slot 0 of a slot-constant element
            for (int s = 1; s < S; ++s) if (std::memcmp(m, m + s * TAU, 8 * TAU) != 0) throw LfException(LF_ERR_UNSUPPORTED, "FOLD: mu must be slot-constant (a sumcheck challenge)");
            u64 acc[TAU]; std::memcpy(acc, m, 8 * TAU);
            for (int d = 0; d < TAU; ++d) { std::memcpy(&pw[((size_t)k * TAU + d) * TAU], acc, 8 * TAU); SF::mul(acc, acc, m); }
        }
        sc->d_mu_pow = E.template dalloc<u64>(pw.size());
        E.h2d(sc->d_mu_pow, pw.data(), pw.size() * 8);
    }

    // PRODUCTS / LIN term structure: term t multiplies the tables idx[off[t] .. off[t+1]) (and its coefficient).  Shapes within
    // SC_MAX_MLES / SC_MAX_TERMS / SC_MAX_FACTORS run on the register-resident kernel, anything else on the general one.
    void set_terms(int n_mles, int deg, bool lin, const std::vector<std::vector<int>>& terms) {
        if (deg < 1 || deg > SC_MAX_DEG) throw LfException(LF_ERR_UNSUPPORTED, "sumcheck degree above SC_MAX_DEG (7)");
        sc->gen.n_mles = n_mles; sc->gen.deg = deg; sc->gen.lin = lin ? 1 : 0; sc->gen.n_terms = (int)terms.size();
        bool small = n_mles <= SC_MAX_MLES && (int)terms.size() <= SC_MAX_TERMS;
        for (auto& t : terms) { small = small && (int)t.size() <= SC_MAX_FACTORS; for (int j : t) if (j < 0 || j >= n_mles) throw LfException(LF_ERR_INVALID_ARG, "comb index outside MLE list"); }
        if (small) { for (size_t t = 0; t < terms.size(); ++t) { sc->gen.term_len[t] = (int)terms[t].size(); for (size_t f = 0; f < terms[t].size(); ++f) sc->gen.term_idx[t][f] = terms[t][f]; } return; }
        std::vector<int> off(1, 0), idx; for (auto& t : terms) { idx.insert(idx.end(), t.begin(), t.end()); off.push_back((int)idx.size()); }
        sc->n_terms_general = (int)terms.size();
        sc->d_term_off = E.template dalloc<int>(off.size()); sc->d_term_idx = E.template dalloc<int>(std::max<size_t>(idx.size(), 1));
        E.h2d(sc->d_term_off, off.data(), off.size() * sizeof(int)); if (!idx.empty()) E.h2d(sc->d_term_idx, idx.data(), idx.size() * sizeof(int));
    }

    // prove_round's evaluation half: out_host = (deg+1) x D limbs
    void evaluate(u64* out_host) {
        if (sc->applied != sc->round) throw LfException(LF_ERR_SUMCHECK_MISUSE, "verifier message is empty");
        if (sc->round >= sc->nv) throw LfException(LF_ERR_SUMCHECK_MISUSE, "Prover is not active");
        if (sc->sharded && sc->len == 1) gather_tables();      // last log2(world) rounds run replicated on every rank
        const size_t n_pairs = sc->len / 2; const int ne = sc->deg + 1;
        unsigned nblk; u64* partial;
        if (sc->kind == LF_COMB_FOLD) {
            const bool round1 = sc->dig && sc->applied == 0, round2d = sc->dig && sc->applied == 1 && sc->fh_deferred;
            const unsigned gx1 = (unsigned)((n_pairs + 127) / 128), gx2 = (unsigned)((n_pairs + 63) / 64);      // rounds >= 2: two lanes per pair
            // rounds >= 3 on short tables: the 2K*tau tables are cut into table slices (blockIdx.z) so that about four waves of blocks exist whatever the length
            unsigned tslices = 1;
            if (!round1 && !round2d && !std::getenv("LF_FOLD_NO_TSLICE")) { const int want = (int)((148 * 4 + gx2 * S - 1) / (gx2 * S)); int ts = want < 12 ? want : 12; if (sc->n_f / 4 < ts) ts = sc->n_f / 4; tslices = (unsigned)(ts < 1 ? 1 : ts); }
            nblk = round1 ? gx1 : round2d ? (unsigned)((n_pairs + 32 * R2_GROUPS - 1) / (32 * R2_GROUPS)) : gx2 * tslices;      // round 2 from digits: four lanes per pair, R2_GROUPS groups of 32 pairs per block
            partial = E.partial_dev((size_t)nblk * 5 * D);
            FoldScArgsT<W> a; a.dense = wp(sc->dense.cur); a.dense_pitch = sc->dense.pitch; a.dense_stride = sc->dense.stride; a.mu_pow = sc->d_mu_pow; a.n_f = sc->n_f;
            a.n_pairs = n_pairs; a.partial = partial; a.dig = sc->dig; a.dig_pitch = sc->dig_pitch; a.dig_stride = sc->dig_stride;
            a.fh = wp(sc->fh.cur); a.fh_pitch = sc->fh.pitch; a.fh_stride = sc->fh.stride;
            for (int l = 0; l < TAU; ++l) a.r1[l] = sc->r1[l];
            if (round1) E.launch("k_fold_sc_round1", [&] { k_fold_sc_round1<Rg><<<dim3(gx1, S), 128, 0, E.st()>>>(a); });
            else if (round2d) {
                // up to MAX_MU tables: 64 KB of dynamic shared memory (above the 48 KB default; per device, so set on every use)
                LF_CUDA(cudaFuncSetAttribute(k_fold_sc_round2<Rg>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_MU * (TAU * 8 + 128)));
                E.launch("k_fold_sc_round2", [&] { k_fold_sc_round2<Rg><<<dim3(nblk, S), 128, (size_t)sc->n_f * (TAU * 8 + 128), E.st()>>>(a); }); }
            else E.launch("k_fold_sc_round", [&] { k_fold_sc_round<Rg><<<dim3(gx2, S, tslices), 128, 0, E.st()>>>(a); });
        } else {
            // one thread per (pair, evaluation point): see k_sc_points
            const int ppb = 128 / ne;
            nblk = (unsigned)std::min<size_t>((n_pairs + ppb - 1) / ppb, 148 * 16);
            bool wide = false;
            if constexpr (std::is_same<Rg, BabyBearRing>::value)
                wide = !sc->n_terms_general && n_pairs >= (size_t)SCW_TILE && n_pairs % SCW_TILE == 0 && ne <= 8 && !std::getenv("LF_SC_NARROW");      // (the narrow kernels are kept for A/B measurements and short tables)
            size_t wide_smem = 0;
            if constexpr (std::is_same<Rg, BabyBearRing>::value) if (wide) {
                // persistent CTAs: as many as are resident on the chip at once, spread over the S slots
                wide_smem = (size_t)sc->gen.n_mles * TAU * (SCW_STAGES * 2 * SCW_TILE + std::max(ne - 2, 0) * SCW_TILE) * sizeof(W);      // tile stages + values at the points >= 2
                auto kern = ne <= 5 ? k_sc_wide_bb<160, 5> : k_sc_wide_bb<256, 3>;      // up to degree 4 (the degree-three CCS): five CTAs of five warps per SM
                if (wide_smem > 48 * 1024) LF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
                if (!E.c->sm_count) LF_CUDA(cudaDeviceGetAttribute(&E.c->sm_count, cudaDevAttrMultiProcessorCount, E.c->device));
                int per_sm = 1; LF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * ne, wide_smem));
                nblk = (unsigned)std::min<size_t>(n_pairs / SCW_TILE, (size_t)std::max(1, per_sm * E.c->sm_count / S));
            }
            partial = E.partial_dev((size_t)nblk * ne * D);
            if (wide) {
                if constexpr (std::is_same<Rg, BabyBearRing>::value) {
                    ScGenericArgsT<W> a; const auto& gen = sc->gen;
                    a.n_mles = gen.n_mles; a.deg = gen.deg; a.n_terms = gen.n_terms; a.lin = gen.lin;
                    for (int t = 0; t < SC_MAX_TERMS; ++t) { a.term_len[t] = gen.term_len[t]; for (int f = 0; f < SC_MAX_FACTORS; ++f) a.term_idx[t][f] = gen.term_idx[t][f]; }
                    a.pitch = sc->dense.pitch; a.n_pairs = n_pairs; a.partial = partial; a.coef = sc->d_coef;
                    for (int k = 0; k < SC_MAX_MLES; ++k) a.mle[k] = wp(sc->dense.cur) + (size_t)std::min(k, a.n_mles - 1) * sc->dense.stride;
                    E.launch("k_sc_wide", [&] { if (ne <= 5) k_sc_wide_bb<160, 5><<<dim3(nblk, S), 32 * ne, wide_smem, E.st()>>>(a); else k_sc_wide_bb<256, 3><<<dim3(nblk, S), 32 * ne, wide_smem, E.st()>>>(a); });
                }
            } else if (sc->n_terms_general) {
                ScTermsArgsT<W> g; g.base = wp(sc->dense.cur); g.stride = sc->dense.stride; g.pitch = sc->dense.pitch; g.n_mles = sc->gen.n_mles; g.deg = sc->gen.deg;
                g.n_terms = sc->n_terms_general; g.lin = sc->gen.lin; g.term_off = sc->d_term_off; g.idx = sc->d_term_idx; g.coef = sc->d_coef; g.n_pairs = n_pairs; g.partial = partial;
                E.launch("k_sc_generic", [&] { k_sc_terms<Rg><<<dim3(nblk, S), 128, 0, E.st()>>>(g); });
            } else {
            ScGenericArgsT<W> a; const auto& gen = sc->gen;
            a.n_mles = gen.n_mles; a.deg = gen.deg; a.n_terms = gen.n_terms; a.lin = gen.lin;
            for (int t = 0; t < SC_MAX_TERMS; ++t) { a.term_len[t] = gen.term_len[t]; for (int f = 0; f < SC_MAX_FACTORS; ++f) a.term_idx[t][f] = gen.term_idx[t][f]; }
            a.pitch = sc->dense.pitch; a.n_pairs = n_pairs; a.partial = partial; a.coef = sc->d_coef;
            for (int k = 0; k < a.n_mles; ++k) a.mle[k] = wp(sc->dense.cur) + (size_t)k * sc->dense.stride;
            dim3 g(nblk, S);
            for (int k = a.n_mles; k < SC_MAX_MLES; ++k) a.mle[k] = a.mle[0];
            const bool legacy = std::getenv("LF_SC_LEGACY") != nullptr;       // the one-thread-per-pair kernel, kept for A/B measurements
            E.launch("k_sc_generic", [&] {
                if (legacy) { if (a.n_mles <= 2) k_sc_generic<Rg, 2><<<g, 128, 0, E.st()>>>(a); else if (a.n_mles <= 4) k_sc_generic<Rg, 4><<<g, 128, 0, E.st()>>>(a); else k_sc_generic<Rg, 8><<<g, 128, 0, E.st()>>>(a); }
                else if (a.n_mles <= 2) k_sc_points<Rg, 2><<<g, 128, 0, E.st()>>>(a);
                else if (a.n_mles <= 4) k_sc_points<Rg, 4><<<g, 128, 0, E.st()>>>(a);
                else if (a.n_mles <= 5) k_sc_points<Rg, 5><<<g, 128, 0, E.st()>>>(a);      // the degree-three CCS: four matrices + eq
                else k_sc_points<Rg, 8><<<g, 128, 0, E.st()>>>(a);
            });
            }
        }
        u64* d_out = E.small_dev((size_t)ne * D);
        if (sc->sharded) E.reduce_partials_allreduce(partial, (int)nblk, (size_t)ne * D, d_out);      // one all-reduce of (deg+1) ring elements per round, fused with the reduction
        else E.reduce_partials(partial, (int)nblk, (size_t)ne * D, d_out);
        E.download_words(d_out, (size_t)ne * D, out_host);
        sc->round += 1;
    }
    // ping-pong targets are allocated once (sizes len/2 and len/4 of the first fold) and reused by all later rounds
    W* pingpong_target(lf_sumcheck::Group& g, size_t n_out) {
        const size_t need = (size_t)g.count * pitch_of(n_out) * D;
        lf_words*& slot = (g.cur == g.nxt) ? g.alt : g.nxt;
        size_t& cap = (g.cur == g.nxt) ? g.alt_cap : g.nxt_cap;
        if (cap < need) { E.dfree(slot); slot = ow(E.template dalloc<W>(need)); cap = need; }
        return wp(slot);
    }
    void fold_group(lf_sumcheck::Group& g, const u64* r_sf, size_t n_out) {
        if (!g.count || !g.cur) return;
        const size_t np = pitch_of(n_out);
        W* out = pingpong_target(g, n_out);
        bool wide = false;
        if constexpr (std::is_same<Rg, BabyBearRing>::value) {
            if (!std::getenv("LF_SC_NARROW")) {
                wide = true;
                FoldWideArgs a; a.in = wp(g.cur); a.out = out; a.in_pitch = g.pitch; a.out_pitch = np; a.in_stride = g.stride; a.out_stride = np * D; a.n_out = n_out; a.r = BbBal::fixed(r_sf);
                E.launch("k_fold", [&] { k_fold_wide_bb<><<<dim3(Engine<Rg>::blocks_for(n_out, 128), S, g.count), 128, 0, E.st()>>>(a); });
            }
        }
        if (!wide) {
        FoldArgsT<W> a; a.in = wp(g.cur); a.out = out; a.in_pitch = g.pitch; a.out_pitch = np; a.in_stride = g.stride; a.out_stride = np * D; a.n_out = n_out;
        for (int l = 0; l < TAU; ++l) a.r[l] = r_sf[l];
        E.launch("k_fold", [&] { k_fold<Rg><<<dim3(Engine<Rg>::blocks_for(n_out, 128), S, g.count), 128, 0, E.st()>>>(a); });
        }
        if (g.cur_owned && g.cur != g.nxt && g.cur != g.alt) E.dfree(g.cur);     // the caller-provided / initial table set
        g.cur = ow(out); g.cur_owned = false; g.pitch = np; g.stride = np * D;
    }
    // every rank holds one entry per table: all-gather them (summing into a zeroed buffer) so the remaining variables,
    // which index the ranks, can be bound on every rank redundantly
    void gather_group(lf_sumcheck::Group& g) {
        if (!g.count || !g.cur) return;
        const int G = E.c->world; const size_t np = pitch_of(G), rows = (size_t)g.count * D;
        // (summed as u64 lanes: every word is non-zero on exactly one rank, so packed 4-byte words cannot carry into each other)
        W* out = E.template dalloc<W>(rows * np);
        LF_CUDA(cudaMemsetAsync(out, 0, rows * np * sizeof(W), E.st()));
        E.launch("k_scatter_entry", [&] { k_scatter_entry<W><<<Engine<Rg>::blocks_for(rows), 256, 0, E.st()>>>(wp(g.cur), g.pitch, out, np, rows, E.c->rank); });
        E.collective(0, reinterpret_cast<u64*>(out), rows * np * sizeof(W) / 8);
        if (g.cur_owned && g.cur != g.nxt && g.cur != g.alt) E.dfree(g.cur);
        g.cur = ow(out); g.cur_owned = true; g.pitch = np; g.stride = np * D;
    }
    void gather_tables() {
        if (sc->kind == LF_COMB_FOLD && sc->dig && sc->applied == 0) throw LfException(LF_ERR_UNSUPPORTED, "sharded FOLD sumcheck needs at least 2 local entries");
        gather_group(sc->dense); if (sc->kind == LF_COMB_FOLD) gather_group(sc->fh);
        sc->len = (size_t)E.c->world; sc->sharded = false;
    }
    // fix_variables with the verifier's challenge (prover.rs:61-72)
    void apply_challenge(const u64* r_sf) {
        if (sc->applied >= sc->round) throw LfException(LF_ERR_SUMCHECK_MISUSE, "first round should be prover first.");
        const size_t n_out = sc->len / 2;
        fold_group(sc->dense, r_sf, n_out);
        if (sc->kind == LF_COMB_FOLD) {
            if (sc->dig && sc->applied == 0 && sc->nv >= 2 && n_out >= 2 && !std::getenv("LF_FOLD_R2_LEGACY")) {
                sc->fh_deferred = true; sc->fh.count = sc->n_f; sc->fh.cur = nullptr;
                for (int l = 0; l < TAU; ++l) sc->r1[l] = r_sf[l];
            } else if (sc->dig && sc->applied == 1 && sc->fh_deferred) {
                // T2 = a0 + r1 e0 + r2 (a1 - a0) + r1 r2 (e1 - e0) from four digits per entry
                sc->fh.pitch = pitch_of(n_out); sc->fh.stride = sc->fh.pitch * D;
                sc->fh.cur = ow(pingpong_target(sc->fh, n_out)); sc->fh.cur_owned = false; sc->fh_deferred = false;
                u64 cs[4 * TAU], r12[TAU]; SF::mul(r12, sc->r1, r_sf);
                for (int l = 0; l < TAU; ++l) { cs[l] = sc->r1[l]; cs[TAU + l] = r_sf[l]; cs[2 * TAU + l] = r12[l];
                    const u64 two = F::add(F::add(sc->r1[l], sc->r1[l]), F::add(r_sf[l], r_sf[l])), r12x2 = F::add(r12[l], r12[l]); cs[3 * TAU + l] = F::add(two, F::add(r12x2, r12x2)); }
                u64* d_cs = E.template dalloc<u64>(4 * TAU); E.h2d(d_cs, cs, sizeof cs);
                E.launch("k_fold_digits", [&] { k_fold_digits2<Rg><<<dim3(Engine<Rg>::blocks_for(n_out, 128), S, sc->n_f), 128, 0, E.st()>>>(sc->dig, sc->dig_pitch, sc->dig_stride, sc->n_f, wp(sc->fh.cur), sc->fh.pitch, sc->fh.stride, n_out, d_cs); });
                E.dfree(d_cs);
            } else if (sc->dig && sc->applied == 0) {
                sc->fh.count = sc->n_f; sc->fh.pitch = pitch_of(n_out); sc->fh.stride = sc->fh.pitch * D; sc->fh.cur = nullptr;
                sc->fh.cur = ow(pingpong_target(sc->fh, n_out)); sc->fh.cur_owned = false;
                FoldArgsT<W> a; for (int l = 0; l < TAU; ++l) a.r[l] = r_sf[l];
                E.launch("k_fold_digits", [&] { k_fold_digits<Rg><<<dim3(Engine<Rg>::blocks_for(n_out, 128), S, sc->n_f), 128, 0, E.st()>>>(sc->dig, sc->dig_pitch, sc->dig_stride, sc->n_f, wp(sc->fh.cur), sc->fh.pitch, sc->fh.stride, n_out, a); });
            } else fold_group(sc->fh, r_sf, n_out);
        }
        sc->len = n_out; sc->applied += 1;
    }
    // after the last challenge every table has one entry: mle_k(r).  out: (dense.count + n_f) x D limbs on the host
    void final_values(u64* out_host) {
        if (sc->len != 1 || sc->sharded) throw LfException(LF_ERR_SUMCHECK_MISUSE, "sumcheck not finished");
        const int total = sc->dense.count + (sc->kind == LF_COMB_FOLD ? sc->n_f : 0);
        std::vector<W> tmp;
        auto grab = [&](const lf_sumcheck::Group& g, u64* dst) {
            if (!g.count) return;
            tmp.resize((size_t)g.count * g.stride + 8 / sizeof(W));
            E.download_words(reinterpret_cast<const u64*>(g.cur), ((size_t)g.count * g.stride * sizeof(W) + 7) / 8, reinterpret_cast<u64*>(tmp.data()));
            for (int k = 0; k < g.count; ++k) for (int l = 0; l < D; ++l) dst[(size_t)k * D + l] = (u64)tmp[(size_t)k * g.stride + (size_t)l * g.pitch];
        };
        grab(sc->dense, out_host);
        if (sc->kind == LF_COMB_FOLD) grab(sc->fh, out_host + (size_t)sc->dense.count * D);
        (void)total;
    }
};

}  // namespace lf
