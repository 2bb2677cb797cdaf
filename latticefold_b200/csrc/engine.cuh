// Host-side engine: device memory, launches and the per-op drivers behind the C ABI (include/lf_b200.h).
#pragma once
#include "kernels.cuh"
#include "commit_mma.cuh"
#include "transcript_host.hpp"
#include "../../include/lf_b200.h"
#include <vector>
#include <string>
#include <memory>
#include <chrono>
#include <cstdlib>
#include <dlfcn.h>
#include <unordered_map>

namespace lf {

#define LF_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) throw LfException(LF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)

typedef std::vector<u64> HV;   // host vector of ring elements, D limbs each

inline size_t pitch_of(size_t n) { return ((n ? n : 1) + 31) / 32 * 32; }   // planes start 256-byte aligned
inline int ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) ++l; return l; }

}  // namespace lf

// NCCL is reached through dlopen (the process already holds torch's bundled libnccl.so.2; nothing links against it at build
// time), with the handful of prototypes the sharded path needs.  ncclSum = 0, ncclUint64 = 5 (nccl.h).
namespace lf {
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    static NcclApi& get() {
        static NcclApi api; static bool tried = false;
        if (!tried) { tried = true;
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (h) { api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId"); api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
                     api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy"); api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
                     api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather"); api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString"); } }
        return api;
    }
    bool ok() const { return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather; }
};
}  // namespace lf

struct lf_ctx {
    int ring = 0, device = 0, sm_count = 0;      // sm_count: filled on first use (persistent-grid kernels)
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    // ring tables on the device: [0] CRT, [1] ICRT
    int* d_tab_idx[2] = {nullptr, nullptr}; lf::u64* d_tab_val[2] = {nullptr, nullptr};
    void* tables = nullptr;        // RingTables<Rg>*
    bool shared_tables = false;    // auxiliary context: ring tables belong to the parent
    lf::u64* h_pinned = nullptr; size_t h_pinned_words = 0;       // D2H landing zone / H2D staging
    lf::u64* d_small = nullptr; size_t d_small_words = 0;          // small device results / parameters
    lf::u64* d_partial = nullptr; size_t d_partial_words = 0;      // block partial sums
    int* d_err = nullptr;
    // size-keyed cache of device blocks.  Every use is on this context's single stream, so a block handed back by dfree can be
    // re-issued at once (stream order serialises the old and the new user); steady-state prover steps allocate nothing.
    std::unordered_multimap<size_t, void*> block_cache; std::unordered_map<void*, size_t> block_size; size_t cached_bytes = 0;
    // blocks whose last use may still be pending on this stream when ANOTHER stream (the prover's auxiliary one) could be handed them:
    // they return to the cache only at the next synchronisation of this stream (upload staging buffers, see Engine::dfree_later)
    std::vector<void*> deferred_free;
    // pinned bump arena: staging for small async H2D copies and landing zone for async D2H results; reset per prover step
    unsigned char* h_arena = nullptr; size_t arena_size = 0, arena_off = 0;
    // column / hypercube sharding across the GPUs of one box: rank, world and the collective the host side
    // provides (torch.distributed over NCCL in bench.py).  op 0: in-place sum of u64 lanes; op 1: in-place all-gather
    // (buffer = world x words, this rank's part at rank * words).  The pointer is device memory on this context's device.
    int rank = 0, world = 1; lf_collective_fn coll = nullptr; void* coll_user = nullptr; uint64_t collectives = 0;
    void* nccl = nullptr;          // ncclComm_t when the collectives run on this context's stream (no host round trip)
    // peer-memory mailboxes (k_reduce_allreduce_p2p): this rank's region and the IPC mappings of every peer's
    // Two channels per region: channel 0 serves this context's stream, channel 1 the prover's auxiliary stream (the accumulator's
    // decomposition runs beside the linearization), each with its own call sequence.
    struct XGpu { bool on = false; void* region = nullptr; void* peer_region[8] = {nullptr}; unsigned char* base[8] = {nullptr};
                  lf::u64* inbox[8] = {nullptr}; unsigned long long* flags[8] = {nullptr};
                  size_t cap = 0; unsigned long long calls = 0, blocks = 0;
                  // channel 1 outlives any one prover's auxiliary context (the flags are monotone counters that are never reset), so
                  // its call sequence is kept here, in the owning context; an auxiliary context points at its parent
                  unsigned long long aux_calls = 0, aux_blocks = 0; XGpu* parent = nullptr; } xg;
    int bulk_repr = 0;             // LF_REPR_CANONICAL / LF_REPR_MONTGOMERY for witness-sized host vectors (lf_ctx_set_bulk_repr)
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
};
// Limb planes are arrays of the ring's word type (Rg::W: u64, or packed u32 for the 31-bit BabyBear prime).  The ring-independent
// handle structs hold them behind an incomplete type, so that any pointer arithmetic outside the typed accessors (Engine::wp) is a
// compile error rather than a wrong stride.
struct lf_words;
struct lf_vec { lf_words* p = nullptr; size_t n = 0, pitch = 0; int form = 0; };
struct lf_ajtai { lf_words* p = nullptr; size_t kappa = 0, n = 0, pitch = 0;
                  // byte-limb tiles of the matrix for the tensor-core digit commit (commit_mma.cuh); absent on rings that do not use it
                  uint8_t* a8 = nullptr; int a8_tiles = 0, a8_chunks = 0; void* epi = nullptr; };
struct lf_sparse { lf::u32 *row_ptr = nullptr, *col = nullptr; lf_words* val = nullptr; size_t nrows = 0, ncols = 0, nnz = 0, val_pitch = 0, eff_rows = 0;
                   // transposed image of the rank's columns (k_csc_eq): all rows of the whole matrix, local column indices
                   lf::u32 *t_col_ptr = nullptr, *t_row = nullptr; lf_words* t_val = nullptr; size_t t_ncols = 0, t_nnz = 0, t_val_pitch = 0; };

namespace lf {

template <class Rg> struct Engine {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef HostRing<Rg> HR; typedef typename HR::El El; typedef typename Rg::W W;
    typedef PtrListT<W> PL;
    static constexpr int D = Rg::D, S = Rg::S, TAU = Rg::TAU, TPB_MA = matrix_apply_tpb<Rg>();
    static W* wp(lf_words* p) { return reinterpret_cast<W*>(p); }
    static const W* wp(const lf_words* p) { return reinterpret_cast<const W*>(p); }
    static lf_words* ow(W* p) { return reinterpret_cast<lf_words*>(p); }
    lf_ctx* c;
    explicit Engine(lf_ctx* ctx) : c(ctx) {}
    const RingTables<Rg>& tab() const { return *(const RingTables<Rg>*)c->tables; }
    cudaStream_t st() const { return c->stream; }

    // ---------------------------------------------------------------- memory
    template <class T> T* dalloc(size_t count) {
        size_t bytes = ((count ? count : 1) * sizeof(T) + 511) / 512 * 512;
        auto it = c->block_cache.find(bytes);
        if (it != c->block_cache.end()) { void* p = it->second; c->block_cache.erase(it); c->cached_bytes -= bytes; return (T*)p; }
        void* p = nullptr; cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {      // out of memory: drop the cache and retry once
            cudaGetLastError(); sync(); for (auto& kv : c->block_cache) { cudaFree(kv.second); c->block_size.erase(kv.second); } c->block_cache.clear(); c->cached_bytes = 0;
            LF_CUDA(cudaMalloc(&p, bytes));
        }
        c->block_size[p] = bytes; return (T*)p;
    }
    void dfree(void* p) { if (!p) return; auto it = c->block_size.find(p); if (it == c->block_size.end()) return; c->block_cache.emplace(it->second, p); c->cached_bytes += it->second; }
    lf_vec* vec_alloc(size_t n, int form) { lf_vec* v = new lf_vec; v->n = n; v->pitch = pitch_of(n); v->form = form; v->p = ow(dalloc<W>(v->pitch * D)); return v; }
    void vec_free(lf_vec* v) { if (v) { dfree(v->p); delete v; } }
    u64* pinned(size_t words) {
        if (c->h_pinned_words < words) { if (c->h_pinned) { LF_CUDA(cudaStreamSynchronize(st())); cudaFreeHost(c->h_pinned); } size_t w = std::max(words, (size_t)1 << 16); LF_CUDA(cudaMallocHost(&c->h_pinned, w * 8)); c->h_pinned_words = w; }
        return c->h_pinned;
    }
    u64* small_dev(size_t words) {
        if (c->d_small_words < words) { if (c->d_small) { LF_CUDA(cudaStreamSynchronize(st())); cudaFree(c->d_small); } size_t w = std::max(words, (size_t)1 << 16); LF_CUDA(cudaMalloc(&c->d_small, w * 8)); c->d_small_words = w; }
        return c->d_small;
    }
    u64* partial_dev(size_t words) {
        if (c->d_partial_words < words) { if (c->d_partial) { LF_CUDA(cudaStreamSynchronize(st())); cudaFree(c->d_partial); } size_t w = std::max(words, (size_t)1 << 20); LF_CUDA(cudaMalloc(&c->d_partial, w * 8)); c->d_partial_words = w; }
        return c->d_partial;
    }
    void sync() { LF_CUDA(cudaStreamSynchronize(st())); for (void* p : c->deferred_free) dfree(p); c->deferred_free.clear(); }
    void dfree_later(void* p) { if (p) c->deferred_free.push_back(p); }
    // ---- pinned arena (no synchronisation: the staged bytes stay untouched until arena_reset at the next step)
    void* arena_alloc(size_t bytes, bool may_wrap = true) {
        if (!c->h_arena) { c->arena_size = (size_t)64 << 20; LF_CUDA(cudaMallocHost(&c->h_arena, c->arena_size)); c->arena_off = 0; }
        bytes = (bytes + 63) / 64 * 64;
        if (bytes > c->arena_size) throw LfException(LF_ERR_INVALID_ARG, "pinned arena too small for this transfer");
        if (c->arena_off + bytes > c->arena_size) {
            if (!may_wrap) throw LfException(LF_ERR_INVALID_ARG, "pinned arena exhausted by pending downloads");
            sync(); c->arena_off = 0;      // wrap: every staged upload has been consumed by now
        }
        void* p = c->h_arena + c->arena_off; c->arena_off += bytes; return p;
    }
    void arena_reset() { c->arena_off = 0; }
    void h2d(void* dev, const void* host, size_t bytes) {          // small async upload from pageable memory
        if (!bytes) return;
        void* stage = arena_alloc(bytes); std::memcpy(stage, host, bytes);
        LF_CUDA(cudaMemcpyAsync(dev, stage, bytes, cudaMemcpyHostToDevice, st()));
    }
    const u64* d2h_async(const u64* dev, size_t words) {           // valid after the next sync / event on this stream
        u64* land = (u64*)arena_alloc(std::max<size_t>(words, 1) * 8, false);
        if (words) LF_CUDA(cudaMemcpyAsync(land, dev, words * 8, cudaMemcpyDeviceToHost, st()));
        return land;
    }
    // every kernel launch goes through here: counts launches (bench.py's gpu_launches) and, in profiling mode, brackets
    // the launch with CUDA events on the context's stream so bench.py can attribute device time per kernel.
    template <class Fn> void launch(const char* name, Fn&& fn) {
        if (c->profiling) {
            cudaEvent_t a, b; LF_CUDA(cudaEventCreate(&a)); LF_CUDA(cudaEventCreate(&b));
            LF_CUDA(cudaEventRecord(a, st())); fn(); LF_CUDA(cudaEventRecord(b, st()));
            c->prof.push_back({name, a, b});
        } else fn();
        ++c->launches; LF_CUDA(cudaGetLastError());
    }
    void check_err_flag(int code, const char* msg) {
        int h = 0; LF_CUDA(cudaMemcpyAsync(&h, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, st())); sync();
        if (h) { LF_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), st()));
                 if (h == 2) throw LfException(LF_ERR_CUDA, "peer-memory all-reduce timed out waiting for another rank");
                 throw LfException(code, msg); }
    }
    // host elements -> small SoA device vector, staged through the pinned arena (asynchronous)
    void upload_small(const u64* host, size_t n, W* dev, size_t pitch) {
        W* soa = (W*)arena_alloc(pitch * D * sizeof(W)); std::memset(soa, 0, pitch * D * sizeof(W));
        for (size_t i = 0; i < n; ++i) for (int l = 0; l < D; ++l) soa[(size_t)l * pitch + i] = (W)host[i * D + l];
        LF_CUDA(cudaMemcpyAsync(dev, soa, pitch * D * sizeof(W), cudaMemcpyHostToDevice, st()));
    }
    // download `words` u64 from the device into a host vector (through the pinned landing zone)
    void download_words(const u64* dev, size_t words, u64* host) {
        u64* pz = pinned(words);
        LF_CUDA(cudaMemcpyAsync(pz, dev, words * 8, cudaMemcpyDeviceToHost, st())); sync();
        std::memcpy(host, pz, words * 8);
    }

    // ---------------------------------------------------------------- collectives
    bool sharded() const { return c->world > 1; }
    void collective(int op, u64* dev, size_t words) {
        if (c->nccl) {      // enqueued on the context's stream: ordered with the producing / consuming kernels, no synchronisation
            NcclApi& n = NcclApi::get(); ++c->collectives;
            int rc = op == 0 ? n.AllReduce(dev, dev, words, 5, 0, c->nccl, st()) : n.AllGather(dev + (size_t)c->rank * words, dev, words, 5, c->nccl, st());
            if (rc != 0) throw LfException(LF_ERR_CUDA, std::string("NCCL: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
            return;
        }
        if (!c->coll) throw LfException(LF_ERR_INVALID_ARG, "sharded context without a collective");
        sync(); ++c->collectives;
        if (c->coll(c->coll_user, op, dev, words) != 0) throw LfException(LF_ERR_CUDA, "collective callback failed");
    }
    static constexpr size_t XG_CAP = 16384, XG_FLAG_BYTES = 256;      // mailbox words per (parity, source); flag area in front
    static constexpr size_t XG_CHANNEL_BYTES = XG_FLAG_BYTES + 2 * 8 * XG_CAP * sizeof(u64), XG_CHANNELS = 2;
    // out[j] = sum over ranks of (sum_b partial[b * nout + j]): one kernel with peer stores when the mailboxes are mapped,
    // otherwise the local reduction followed by the NCCL all-reduce
    void reduce_partials_allreduce(const u64* partial, int nblk, size_t nout, u64* d_out) {
        if (sharded() && c->xg.on && nout <= c->xg.cap) {
            // many block partials (large shards): sum them with the wide reduction first and let the exchange kernel work in place --
            // its own column sum walks the blocks serially, one thread per output
            if (nblk >= 64) { reduce_partials(partial, nblk, nout, d_out); partial = nullptr; }
            const unsigned grid = blocks_for(nout, 128);
            XgArgs x; for (int r = 0; r < 8; ++r) { x.inbox[r] = c->xg.inbox[r]; x.flags[r] = c->xg.flags[r]; }
            unsigned long long& calls = c->xg.parent ? c->xg.parent->aux_calls : c->xg.calls;
            unsigned long long& blocks = c->xg.parent ? c->xg.parent->aux_blocks : c->xg.blocks;
            x.rank = c->rank; x.world = c->world; x.parity = (unsigned)(calls & 1); x.cap = c->xg.cap; x.err = c->d_err;
            blocks += grid; calls += 1; x.expected = blocks; ++c->collectives;
            launch("k_reduce_allreduce_p2p", [&] { k_reduce_allreduce_p2p<F><<<grid, 128, 0, st()>>>(partial, nblk, (int)nout, d_out, x); });
            return;
        }
        reduce_partials(partial, nblk, nout, d_out);
        allreduce_field(d_out, nout);
    }
    void reduce_partials(const u64* partial, int nblk, size_t nout, u64* d_out) {
        if (nblk > 512) {      // two levels: chunks of 32 blocks first
            const int chunk = 32, nch = (nblk + chunk - 1) / chunk;
            u64* tmp = dalloc<u64>((size_t)nch * nout);
            launch("k_reduce_partials", [&] { k_reduce_chunks<F><<<dim3(blocks_for(nout, 128), nch), 128, 0, st()>>>(partial, nblk, (int)nout, chunk, tmp); });
            reduce_partials(tmp, nch, nout, d_out); dfree(tmp); return;
        }
        if (nblk >= 64 && nout <= 16384) launch("k_reduce_partials", [&] { k_reduce_partials_wide<F><<<blocks_for(nout, 32), 256, 0, st()>>>(partial, nblk, (int)nout, d_out); });
        else launch("k_reduce_partials", [&] { k_reduce_partials<F><<<blocks_for(nout, 128), 128, 0, st()>>>(partial, nblk, (int)nout, d_out); });
    }
    // sum mod p across ranks of `words` field elements at dev (in place)
    void allreduce_field(u64* dev, size_t words) {
        if (!sharded() || !words) return;
        u64* tmp = dalloc<u64>(2 * words);
        launch("k_split_limbs", [&] { k_split_limbs<0><<<blocks_for(words), 256, 0, st()>>>(dev, tmp, words); });
        collective(0, tmp, 2 * words);
        launch("k_combine_limbs", [&] { k_combine_limbs<F><<<blocks_for(words), 256, 0, st()>>>(tmp, dev, words); });
        dfree(tmp);
    }

    // ---------------------------------------------------------------- layout
    // 2^64 mod p (to Montgomery form) or its inverse (from Montgomery form): ark-ff's R for one-limb fields
    static u64 mont_factor(bool to_mont) { const u64 r = (u64)((((u128)1) << 64) % F::P); return to_mont ? r : F::inv(r); }
    void upload_planes(const u64* host, size_t n, W* dev, size_t pitch) {   // host AoS (u64 limbs) -> device planes
        if (!n) return;
        u64* stage = dalloc<u64>(n * D);
        LF_CUDA(cudaMemcpyAsync(stage, host, n * D * 8, cudaMemcpyHostToDevice, st()));
        launch("k_aos_to_soa", [&] { k_aos_to_soa<F, D, W><<<(unsigned)((n + 63) / 64), 256, 0, st()>>>(stage, dev, n, pitch, c->bulk_repr ? mont_factor(false) : 0); });
        // not back into the cache yet: the prover allocates buffers that its auxiliary stream writes first, and that stream does not
        // wait for this copy (the accumulator's decomposition starts while the incoming witness is still uploading)
        dfree_later(stage);
    }
    void download_planes(const W* dev, size_t pitch, size_t n, u64* host) {
        if (!n) return;
        u64* stage = dalloc<u64>(n * D);
        launch("k_soa_to_aos", [&] { k_soa_to_aos<F, D, W><<<(unsigned)((n + 63) / 64), 256, 0, st()>>>(dev, stage, n, pitch, c->bulk_repr ? mont_factor(true) : 0); });
        LF_CUDA(cudaMemcpyAsync(host, stage, n * D * 8, cudaMemcpyDeviceToHost, st())); sync();
        dfree(stage);
    }

    // ---------------------------------------------------------------- elementwise ops
    void crt(const W* in, size_t in_pitch, W* out, size_t out_pitch, size_t n, bool inverse) {
        if (!n) return;
        launch("k_matrix_apply", [&] { k_matrix_apply<Rg, W><<<(unsigned)((n + TPB_MA - 1) / TPB_MA), TPB_MA, 0, st()>>>(in, in_pitch, out, out_pitch, n, c->d_tab_idx[inverse], c->d_tab_val[inverse], 0, 0); });
    }
    // CRT of `batch` digit vectors (in/out strides between consecutive vectors) in one launch
    void crt_digits(const int8_t* in, size_t in_pitch, W* out, size_t out_pitch, size_t n, int batch = 1, size_t in_stride = 0, size_t out_stride = 0) {
        if (!n || !batch) return;
        launch("k_matrix_apply", [&] { k_matrix_apply<Rg, int8_t><<<dim3((unsigned)((n + TPB_MA - 1) / TPB_MA), batch), TPB_MA, 0, st()>>>(in, in_pitch, out, out_pitch, n, c->d_tab_idx[0], c->d_tab_val[0], in_stride, out_stride); });
    }
    static unsigned blocks_for(size_t work, int bs = 256) { return (unsigned)((work + bs - 1) / bs); }
    void gadget_decompose(const W* in, size_t in_pitch, W* out, size_t out_pitch, size_t n, u64 B, int L) {
        if (B < 2 || B >= ((u64)1 << 62) || L < 1 || L > 64) throw LfException(LF_ERR_UNSUPPORTED, "gadget_decompose: need 2 <= B < 2^62, 1 <= L <= 64");
        if (!n) return;
        launch("k_gadget_decompose", [&] { k_gadget_decompose<Rg><<<blocks_for(n * D), 256, 0, st()>>>(in, in_pitch, out, out_pitch, n, (int64_t)B, L, c->d_err); });
    }
    void gadget_recompose(const W* in, size_t in_pitch, W* out, size_t out_pitch, size_t n_out, u64 B, int L, int batch = 1, size_t in_stride = 0, size_t out_stride = 0) {
        if (!n_out || !batch) return;
        launch("k_gadget_recompose", [&] { k_gadget_recompose<Rg><<<dim3(blocks_for(n_out * D), batch), 256, 0, st()>>>(in, in_pitch, out, out_pitch, n_out, B % F::P, L, in_stride, out_stride); });
    }
    void digit_split(const W* in, size_t in_pitch, int8_t* out, size_t out_pitch, size_t n, u64 b, int K) {
        if (b < 2 || b > 254 || K < 1 || K > 64) throw LfException(LF_ERR_UNSUPPORTED, "decompose_to_vec: need 2 <= b <= 254 (int8 digits), 1 <= K <= 64");
        if (!n) return;
        if (b == 2 && out_pitch % 4 == 0 && out_pitch >= (n + 3) / 4 * 4)      // the 4-byte stores of the last group stay inside the plane's padding
            launch("k_digit_split", [&] { k_digit_split_b2<Rg><<<dim3(blocks_for((n + 3) / 4), D), 256, 0, st()>>>(in, in_pitch, out, out_pitch, n, K, c->d_err); });
        else launch("k_digit_split", [&] { k_digit_split<Rg><<<blocks_for(n * D), 256, 0, st()>>>(in, in_pitch, out, out_pitch, n, (int64_t)b, K, c->d_err); });
    }

    // ---------------------------------------------------------------- batched dot products (commit, MLE evaluation)
    // result: nrows x ncols x D limbs on the device (d_out)
    void dot(const W* X, size_t x_row_stride, size_t x_pitch, int nrows, const size_t* x_len_dev,
             const PL& Y, size_t y_pitch, int ncols, size_t n, u64* d_out, const char* name = "k_dot") {
        if (nrows == 0 || ncols == 0) return;
        if (ncols > MAX_LIST) throw LfException(LF_ERR_INVALID_ARG, "dot: too many columns in one launch");
        DotArgsT<W> a; a.X = X; a.x_row_stride = x_row_stride; a.x_pitch = x_pitch; a.nrows = nrows; a.Y = Y; a.y_pitch = y_pitch; a.ncols = ncols;
        a.x_len = x_len_dev; a.n = n;
        // x tile: long enough to amortise the per-warp reduction, short enough to fill 148 SMs
        // measured on B200 (tools/dot_ab.py, kappa=26, n=2^18, 15 pieces): CT=2 with 4-8 warps per block runs at 87% of the
        // IMAD.WIDE-bound multiply-accumulate peak (4.7 ms vs 4.1 ms); CT=4 is register-starved (7.9 ms), CT=1 reloads too much (6.6 ms)
        int ct = ncols >= 2 ? 2 : 1;
        if (const char* e = std::getenv("LF_DOT_CT")) { int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && v <= ncols) ct = v; }
        const int units = nrows * ((ncols + ct - 1) / ct);
        int wmax = 4; if (const char* e = std::getenv("LF_DOT_WPB")) { int v = atoi(e); if (v >= 1 && v <= 16) wmax = v; }
        int wpb = 1; { int best = 1 << 30; for (int w = wmax; w >= std::min(3, wmax); --w) { int waste = (units + w - 1) / w * w - units; if (waste < best) { best = waste; wpb = w; } } if (units < 4) wpb = std::min(units, wmax); }
        const int groups = (units + wpb - 1) / wpb;
        size_t xpb = 128 * 32;
        while (xpb > 256 && (size_t)groups * S * ((n + xpb - 1) / xpb) < 148 * 4) xpb /= 2;
        a.x_per_block = (int)xpb;
        const unsigned xt = (unsigned)std::max<size_t>(1, (n + xpb - 1) / xpb);
        const size_t nout = (size_t)nrows * ncols * D;
        a.partial = partial_dev((size_t)xt * nout);
        launch(name, [&] {
            dim3 g((unsigned)groups, xt, S);
            if (wpb <= 8) {
                if (ct == 4) k_dot<Rg, 4, 256><<<g, wpb * 32, 0, st()>>>(a); else if (ct == 2) k_dot<Rg, 2, 256><<<g, wpb * 32, 0, st()>>>(a); else k_dot<Rg, 1, 256><<<g, wpb * 32, 0, st()>>>(a);
            } else {
                if (ct == 4) k_dot<Rg, 4, 512><<<g, wpb * 32, 0, st()>>>(a); else if (ct == 2) k_dot<Rg, 2, 512><<<g, wpb * 32, 0, st()>>>(a); else k_dot<Rg, 1, 512><<<g, wpb * 32, 0, st()>>>(a);
            }
        });
        reduce_partials_allreduce(a.partial, (int)xt, nout, d_out);      // x axis sharded across ranks: one small all-reduce per batched dot
    }
    // ---------------------------------------------------------------- tensor-core commit of digit pieces (commit_mma.cuh)
    static bool commit_mma_supported() { return Rg::ID == 0 && !(std::getenv("LF_COMMIT_MMA") && std::getenv("LF_COMMIT_MMA")[0] == '0'); }
    // byte-limb tiles + epilogue tables of an uploaded matrix (once per matrix)
    void ajtai_build_tiles(lf_ajtai* A) {
        if constexpr (Rg::ID == 0) {
            if (!commit_mma_supported() || !A->kappa || !A->n || A->n >= ((size_t)1 << 23)) return;
            const int g_total = (int)(S * TAU * A->kappa), ntiles = (g_total + cmma::GROUPS - 1) / cmma::GROUPS, nchunks = (int)((A->n + cmma::J - 1) / cmma::J);
            LF_CUDA(cudaMalloc(&A->a8, (size_t)ntiles * nchunks * cmma::A_STAGE_BYTES));
            const size_t work = (size_t)ntiles * cmma::GROUPS * nchunks * (cmma::J / 16);
            launch("k_a8_tile", [&] { cmma::k_a8_tile<Rg><<<blocks_for(work), 256, 0, st()>>>(wp(A->p), A->pitch * D, A->pitch, A->n, (int)A->kappa, g_total, nchunks, (size_t)ntiles * cmma::GROUPS, A->a8); });
            cmma::EpiTables<Rg> t; const RingTables<Rg>& rt = tab();
            for (int s = 0; s < S; ++s) {
                for (int r = 0; r < TAU; ++r) { t.perm[s][r] = (rt.k[s] * r) % TAU; t.corr[s][r] = 0; }
                for (int c = 0; c < D; ++c) { t.val[s][c] = rt.crt[s * TAU + t.perm[s][c % TAU]][c]; t.corr[s][c % TAU] = F::add(t.corr[s][c % TAU], t.val[s][c]); }
                for (int r = 0; r < TAU; ++r) t.corr[s][r] = F::mul(t.corr[s][r], (u64)1 << 31);
            }
            LF_CUDA(cudaMalloc(&A->epi, sizeof t)); LF_CUDA(cudaMemcpyAsync(A->epi, &t, sizeof t, cudaMemcpyHostToDevice, st())); sync();
            A->a8_tiles = ntiles; A->a8_chunks = nchunks;
            LF_CUDA(cudaFuncSetAttribute(cmma::k_commit_mma<Rg>, cudaFuncAttributeMaxDynamicSharedMemorySize, commit_mma_smem()));      // per device
        }
    }
    static constexpr int commit_mma_smem() { return cmma::STAGES * (cmma::A_STAGE_BYTES + D * cmma::MAX_PIECES * cmma::J); }
    bool can_commit_digits(const lf_ajtai* A, int ncols, size_t dig_pitch) const { return A->a8 && ncols >= 1 && ncols <= cmma::MAX_PIECES && dig_pitch >= (size_t)A->a8_chunks * cmma::J && dig_pitch % 16 == 0; }
    // y[i][p] = sum_j A[i][j] * CRT(digit piece p)[j] for `ncols` consecutive digit pieces (planes zero beyond n); result kappa x ncols x D limbs
    void commit_digits(const lf_ajtai* A, const int8_t* dig, size_t dig_pitch, size_t dig_stride, int ncols, u64* d_out) {
        if constexpr (Rg::ID == 0) {
            const int nchunks = A->a8_chunks, ntiles = A->a8_tiles;
            const size_t d_stage = (size_t)D * ncols * cmma::J;
            int8_t* d8 = dalloc<int8_t>(d_stage * nchunks);
            launch("k_d8_tile", [&] { cmma::k_d8_tile<Rg><<<blocks_for((size_t)ncols * D * (cmma::J / 16) * nchunks), 256, 0, st()>>>(dig, dig_pitch, dig_stride, ncols, nchunks, d8); });
            // split the witness axis so that the grid fills whole waves of 148 SMs (one CTA per SM: 512 TMEM columns each)
            // (every CTA pays a fixed prologue + epilogue, so among the well-filled grids the one with the fewest splits wins: score = wave
            // efficiency minus 1 % per split; measured at C2: 11 splits 0.50 ms, 15 0.54 ms, 27 0.60 ms, 53 0.71 ms per batch)
            int best = 1; double best_score = -1;
            for (int sp = 1; sp <= 64 && sp <= nchunks; ++sp) { const int ctas = ntiles * sp, waves = (ctas + 147) / 148; const double score = (double)ctas / (148.0 * waves) - 0.01 * sp;
                if (score > best_score) { best_score = score; best = sp; } }
            if (const char* e = std::getenv("LF_COMMIT_SPLITS")) { int v = atoi(e); if (v >= 1 && v <= nchunks) best = v; }
            const int nsplits = best, cps = (nchunks + nsplits - 1) / nsplits, nsp = (nchunks + cps - 1) / cps;
            const int npad = (ncols + 1) / 2 * 2, ntot = D * npad;
            cmma::Args a; a.A8 = A->a8; a.D8 = d8; a.nchunks = nchunks; a.chunks_per_split = cps; a.ncols = ncols;
            a.n_mma1 = std::min(ntot, 192); a.n_mma2 = ntot - a.n_mma1; a.kappa = (int)A->kappa; a.g_total = (int)(S * TAU * A->kappa);
            a.d_stage_bytes = (u32)d_stage; a.tables = A->epi;
            const size_t nout = A->kappa * (size_t)ncols * D;
            a.partial = partial_dev((size_t)nsp * TAU * nout);
            launch("k_commit_mma", [&] { cmma::k_commit_mma<Rg><<<dim3(ntiles, nsp), cmma::THREADS, commit_mma_smem(), st()>>>(a); });
            dfree(d8);
            reduce_partials_allreduce(a.partial, nsp * TAU, nout, d_out);
        } else throw LfException(LF_ERR_UNSUPPORTED, "tensor-core commit is built for the Goldilocks ring");
    }
    // f-hat evaluation from coefficient planes; result nvec x TAU x D limbs on the device
    template <class TIn> void coeff_eval(const TIn* coeff, size_t c_pitch, size_t c_vec_stride, int nvec, const W* eq, size_t eq_pitch, size_t n, u64* d_out) {
        if (!nvec) return;
        const int xpb = 128 * 8; const unsigned xt = (unsigned)std::max<size_t>(1, (n + xpb - 1) / xpb);
        const size_t nout = (size_t)nvec * TAU * D;
        u64* partial = partial_dev((size_t)xt * nout);
        launch("k_coeff_eval", [&] { k_coeff_eval<Rg, TIn><<<dim3(xt, S, nvec * (TAU / coeff_eval_jb<Rg>())), 128, 0, st()>>>(coeff, c_pitch, c_vec_stride, eq, eq_pitch, n, xpb, nvec, partial); });
        reduce_partials_allreduce(partial, (int)xt, nout, d_out);
    }
    void spmv(const lf_sparse* M, const W* head, size_t head_len, size_t head_pitch, const W* tail, size_t tail_pitch, W* out, size_t out_pitch, size_t nrows,
              size_t tail_chunk = ~(size_t)0, size_t tail_chunk_stride = 0, int batch = 1, size_t head_batch_stride = 0, size_t tail_batch_stride = 0, size_t out_batch_stride = 0, bool accumulate = false) {
        if (!nrows || !batch) return;
        launch("k_spmv", [&] { k_spmv<Rg><<<dim3(blocks_for(nrows, 128), S, batch), 128, 0, st()>>>(M->row_ptr, M->col, wp(M->val), M->val_pitch, head, head_len, head_pitch, tail, tail_pitch, tail_chunk, tail_chunk_stride, out, out_pitch, nrows, head_batch_stride, tail_batch_stride, out_batch_stride, accumulate ? 1 : 0); });
    }
    // the two half tables of eq(., r): lo over variables [0, h), hi over [h, s); every rank builds them whole (2^(s/2) entries each)
    struct EqHalves { W *lo = nullptr, *hi = nullptr; size_t plo = 0, phi = 0; int h = 0; };
    EqHalves eq_halves(const u64* r_host, int s) {
        if (s < 1 || s > 40) throw LfException(LF_ERR_INVALID_ARG, "eq_table: r length is 0 or too large");
        std::vector<u64> pair((size_t)s * 2 * D);
        El one = HR::from_u64(1);
        for (int i = 0; i < s; ++i) { El r = HR::load(r_host + (size_t)i * D), m = HR::sub(one, r); std::memcpy(&pair[((size_t)i * 2) * D], m.data(), 8 * D); std::memcpy(&pair[((size_t)i * 2 + 1) * D], r.data(), 8 * D); }
        u64* d_pair = dalloc<u64>(pair.size()); h2d(d_pair, pair.data(), pair.size() * 8);
        EqHalves q; q.h = std::max(1, s / 2); const int sh = s - q.h;      // s = 1: lo over the single variable, hi = the empty product
        const size_t nlo = (size_t)1 << q.h, nhi = (size_t)1 << sh; q.plo = pitch_of(nlo); q.phi = pitch_of(nhi);
        q.lo = dalloc<W>(q.plo * D); q.hi = dalloc<W>(q.phi * D);
        launch("k_eq_table", [&] { k_eq_table<Rg><<<dim3(blocks_for(nlo, 128), S), 128, (size_t)q.h * 2 * TAU * 8, st()>>>(d_pair, q.h, q.lo, q.plo, nlo, 0); });
        if (sh > 0) launch("k_eq_table", [&] { k_eq_table<Rg><<<dim3(blocks_for(nhi, 128), S), 128, (size_t)sh * 2 * TAU * 8, st()>>>(d_pair + (size_t)q.h * 2 * D, sh, q.hi, q.phi, nhi, 0); });
        else upload_small(one.data(), 1, q.hi, q.phi);
        dfree(d_pair); return q;
    }
    void free_halves(EqHalves& q) { dfree(q.lo); dfree(q.hi); q.lo = q.hi = nullptr; }
    // v = M^T eq(., r) on the rank's columns
    void csc_eq(const lf_sparse* M, const EqHalves& q, W* out, size_t out_pitch) {
        if (!M->t_ncols) return;
        launch("k_csc_eq", [&] { k_csc_eq<Rg><<<dim3(blocks_for(M->t_ncols, 128), S), 128, 0, st()>>>(M->t_col_ptr, M->t_row, wp(M->t_val), M->t_val_pitch, q.lo, q.plo, q.hi, q.phi, q.h, out, out_pitch, M->t_ncols); });
    }
    // eq(., r) for r given as s ring elements on the host
    // x_offset / n_local: the slab [x_offset, x_offset + n_local) of the table (hypercube sharding); default = whole table
    void eq_table(const u64* r_host, int s, W* out, size_t out_pitch, size_t x_offset = 0, size_t n_local = 0) {
        if (s < 1 || s > 40) throw LfException(LF_ERR_INVALID_ARG, "eq_table: r length is 0 or too large");
        std::vector<u64> pair((size_t)s * 2 * D);
        El one = HR::from_u64(1);
        for (int i = 0; i < s; ++i) { El r = HR::load(r_host + (size_t)i * D), m = HR::sub(one, r); std::memcpy(&pair[((size_t)i * 2) * D], m.data(), 8 * D); std::memcpy(&pair[((size_t)i * 2 + 1) * D], r.data(), 8 * D); }
        u64* d_pair = dalloc<u64>(pair.size());
        h2d(d_pair, pair.data(), pair.size() * 8);
        const size_t n = n_local ? n_local : (size_t)1 << s;
        if (s < 8) {
            launch("k_eq_table", [&] { k_eq_table<Rg><<<dim3(blocks_for(n, 128), S), 128, (size_t)s * 2 * TAU * 8, st()>>>(d_pair, s, out, out_pitch, n, x_offset); });
        } else {      // two half tables, then one multiply per entry
            const int h = s / 2; const size_t nlo = (size_t)1 << h, nhi = (size_t)1 << (s - h), plo = pitch_of(nlo), phi = pitch_of(nhi);
            W *lo = dalloc<W>(plo * D), *hi = dalloc<W>(phi * D);
            launch("k_eq_table", [&] { k_eq_table<Rg><<<dim3(blocks_for(nlo, 128), S), 128, (size_t)h * 2 * TAU * 8, st()>>>(d_pair, h, lo, plo, nlo, 0); });
            launch("k_eq_table", [&] { k_eq_table<Rg><<<dim3(blocks_for(nhi, 128), S), 128, (size_t)(s - h) * 2 * TAU * 8, st()>>>(d_pair + (size_t)h * 2 * D, s - h, hi, phi, nhi, 0); });
            launch("k_eq_combine", [&] { k_eq_combine<Rg><<<dim3(blocks_for(n, 128), S), 128, 0, st()>>>(lo, plo, hi, phi, h, out, out_pitch, n, x_offset); });
            dfree(lo); dfree(hi);
        }
        dfree(d_pair);
    }
    // out (+)= sum_i coef_i (.) vecs_i ; coef on the host (count x D)
    void lincomb(const PL& vecs, size_t v_pitch, int count, const u64* coef_host, W* out, size_t out_pitch, size_t n, bool accumulate) {
        int done = 0;
        while (done < count || (count == 0 && !accumulate && done == 0)) {
            const int chunk = std::min(MAX_LIST, count - done);
            PL pl; for (int i = 0; i < chunk; ++i) { pl.p[i] = vecs.p[done + i]; pl.len[i] = vecs.len[done + i]; }
            u64* d_coef = dalloc<u64>((size_t)std::max(chunk, 1) * D);
            if (chunk) h2d(d_coef, coef_host + (size_t)done * D, (size_t)chunk * D * 8);
            launch("k_lincomb", [&] { k_lincomb<Rg><<<dim3(blocks_for(n, 128), S), 128, 0, st()>>>(pl, v_pitch, chunk, d_coef, out, out_pitch, n, (accumulate || done > 0) ? 1 : 0); });
            dfree(d_coef);
            done += chunk; if (count == 0) break;
        }
    }
};

}  // namespace lf
