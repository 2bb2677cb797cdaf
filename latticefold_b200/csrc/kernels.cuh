// sm_100a kernels of the LatticeFold prover hot path.  All of them are integer modular arithmetic over limb planes:
// a device vector of n ring elements is D planes of n u64 (plane = limb index, pitch-aligned), so that a warp reading
// one limb of 32 consecutive elements issues one fully coalesced request and a thread that owns (element, slot) has a
// whole slot-field element in registers.  Tensor cores are not used (nothing here is a dense FP contraction).
// Each kernel cites the reference loop it replaces; byte counts per unit are in DESIGN.md.
#pragma once
#include <type_traits>
#include "field.cuh"
#include "ring_host.hpp"
#include <cuda_runtime.h>

namespace lf {

constexpr int MAX_LIST = 32;          // pointer-list capacity of one launch (pieces, MLEs, ...)
constexpr int MAX_MU = 320;           // 2K * tau (K = 16 on the BabyBear ring: 288)
constexpr int SC_MAX_MLES = 8, SC_MAX_TERMS = 4, SC_MAX_FACTORS = 4, SC_MAX_DEG = 7;

template <class W> struct PtrListT { const W* p[MAX_LIST]; size_t len[MAX_LIST]; };
typedef PtrListT<u64> PtrList;
struct PtrList8 { const int8_t* p[MAX_LIST]; };

// Limb planes hold words of type Rg::W (u64, or packed u32 for the 31-bit ring); arithmetic is on u64 values.
// Two / four consecutive words with one vector load (index even / multiple of four: planes are 128-byte aligned).
__device__ __forceinline__ void ld_pair(const u64* p, u64& a, u64& b) { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p); a = v.x; b = v.y; }
__device__ __forceinline__ void ld_pair(const u32* p, u64& a, u64& b) { const uint2 v = *reinterpret_cast<const uint2*>(p); a = v.x; b = v.y; }
__device__ __forceinline__ void ldg_pair(const u64* p, u64& a, u64& b) { const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p)); a = v.x; b = v.y; }
__device__ __forceinline__ void ldg_pair(const u32* p, u64& a, u64& b) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); a = v.x; b = v.y; }
__device__ __forceinline__ void ldg_quad(const u64* p, u64* o) { const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(p)), b = __ldg(reinterpret_cast<const ulonglong2*>(p) + 1); o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; }
__device__ __forceinline__ void ldg_quad(const u32* p, u64* o) { const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)); o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; }

// ------------------------------------------------------------------------------------------------ reductions
// sum over the block of NV field elements per thread; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
template <class F, int NV> __device__ __forceinline__ void block_reduce_add(u64* v, u64* smem /* NV * 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        u64 x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = F::add(x, __shfl_down_sync(0xffffffffu, x, o));
        v[i] = x;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            u64 x = lane < nw ? smem[i * 32 + lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = F::add(x, __shfl_down_sync(0xffffffffu, x, o));
            v[i] = x;
        }
    }
    __syncthreads();
}

// out[j] = sum_b partial[b * nout + j]
template <class F> __global__ void k_reduce_partials(const u64* __restrict__ partial, int nblk, int nout, u64* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nout) return;
    u64 acc = 0;
    for (int b = 0; b < nblk; ++b) acc = F::add(acc, partial[(size_t)b * nout + j]);
    out[j] = acc;
}
// first level for very many block partials (large shards: a FOLD round at 2^20 pairs leaves 16384 of them): the blocks are cut
// into chunks of `chunk`, thread (output j, chunk c) sums its chunk with loads that are coalesced across j
template <class F> __global__ void k_reduce_chunks(const u64* __restrict__ partial, int nblk, int nout, int chunk, u64* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (j >= nout) return;
    const int b0 = c * chunk, b1 = min(nblk, b0 + chunk);
    u64 acc = 0;
    for (int b = b0; b < b1; ++b) acc = F::add(acc, partial[(size_t)b * nout + j]);
    out[(size_t)c * nout + j] = acc;
}
// the same sum for few outputs and many blocks (a sumcheck round: 120 outputs, up to ~1200 block partials): 8 lanes per output
// walk the blocks and combine by shuffles, so a warp still reads 4 x 8 consecutive words per request
template <class F> __global__ void __launch_bounds__(256) k_reduce_partials_wide(const u64* __restrict__ partial, int nblk, int nout, u64* __restrict__ out) {
    const int g = blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8, sub = threadIdx.x & 7;      // output g, lane group member sub
    u64 acc = 0;
    if (g < nout) for (int b = sub; b < nblk; b += 8) acc = F::add(acc, partial[(size_t)b * nout + g]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc = F::add(acc, __shfl_down_sync(0xffffffffu, acc, o, 8));
    if (g < nout && sub == 0) out[g] = acc;
}

// ------------------------------------------------------------------------------------------------ multi-GPU helpers
// NCCL has no "sum mod p": a field element is sent as two 32-bit halves in u64 lanes, summed with ncclSum (world * 2^32
// cannot wrap) and folded back mod p.  Bit-exact because integer addition is associative.
template <int = 0> __global__ void k_split_limbs(const u64* __restrict__ in, u64* __restrict__ out, size_t words) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) { const u64 v = in[i]; out[2 * i] = v & 0xFFFFFFFFULL; out[2 * i + 1] = v >> 32; }
}
template <class F> __global__ void k_combine_limbs(const u64* __restrict__ in, u64* __restrict__ out, size_t words) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) { const u64 lo = in[2 * i], hi = in[2 * i + 1];        // each < world * 2^32
        out[i] = F::from_split(lo, hi); }
}
// Fused block-partial reduction + all-reduce over NVLink peer memory (every collective of the sharded step is a few
// KB, so latency is everything).  Every rank owns a "mailbox" region that all peers map through CUDA IPC:
//   flags[src]                        monotonically increasing count of blocks that have delivered, per source rank
//   inbox[parity][src][cap]           the source's reduced values for the call of that parity
// One launch per call on every rank, in the same order (the prover's call sequence is replicated):  thread j sums its column
// of block partials, STORES the sum straight into every rank's inbox over NVLink, the block publishes with a system-scope
// atomic on every rank's flag, waits until all blocks of all sources have delivered to it, and adds the world values mod p in
// rank order (exact arithmetic: any order gives the same canonical element).  Two parities: a rank can only be one call ahead of
// a peer (call k+1 cannot complete before every peer has finished call k), so slot k+2 never overwrites data still being read.
struct XgArgs { u64* inbox[8]; unsigned long long* flags[8]; int rank, world; unsigned parity; size_t cap; unsigned long long expected; int* err; };
template <class F> __global__ void __launch_bounds__(128)
k_reduce_allreduce_p2p(const u64* __restrict__ partial, int nblk, int nout, u64* __restrict__ out, const XgArgs x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nout) {
        u64 acc = 0;
        if (partial) { for (int b = 0; b < nblk; ++b) acc = F::add(acc, partial[(size_t)b * nout + j]); } else acc = out[j];
        const size_t slot = ((size_t)x.parity * x.world + x.rank) * x.cap + j;
        for (int r = 0; r < x.world; ++r) x.inbox[r][slot] = acc;              // peer stores (NVLink) for r != rank
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < x.world) {
        atomicAdd_system(x.flags[threadIdx.x] + x.rank, 1ULL);                  // publish to rank threadIdx.x
        const volatile unsigned long long* f = x.flags[x.rank] + threadIdx.x;   // and wait for source threadIdx.x
        // bounded spin (a rank that died must not hang the others' GPUs): 20 s on the global timer, then the error flag (2) is
        // raised and the host reports LF_ERR_CUDA at its next flag check
        unsigned long long t0 = 0, now = 0; unsigned spins = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f < x.expected) {
            if ((++spins & 1023u) == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); if (now - t0 > 20000000000ULL) { atomicExch(x.err, 2); break; } }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (j < nout) {
        u64 acc = 0;
        for (int src = 0; src < x.world; ++src) acc = F::add(acc, __ldcv(x.inbox[x.rank] + ((size_t)x.parity * x.world + src) * x.cap + j));
        out[j] = acc;
    }
}

// entry 0 of every (table, plane) -> column `rank` of a zeroed [rows][pitch_out] buffer (all-gather by summation)
template <class W> __global__ void k_scatter_entry(const W* __restrict__ in, size_t in_pitch, W* __restrict__ out, size_t out_pitch, size_t rows, int rank) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) out[i * out_pitch + rank] = in[i * in_pitch];
}

// ------------------------------------------------------------------------------------------------ layout changes
// host "Vec<R>" image (element-major) <-> limb planes.  64 elements per block through shared memory so both sides coalesce.
// mont != 0: the host limbs are arkworks Montgomery representatives a * 2^64 mod p (ark-ff 0.4 Fp64<MontBackend>, one u64 limb): the
// factor is taken out on the way in (mont = 2^-64 mod p) and put back on the way out (mont = 2^64 mod p), so a Rust caller can pass
// the memory of a Vec<R> as it is (LF_REPR_MONTGOMERY) instead of converting 50 MB per witness with into_bigint().
template <class F, int D, class W> __global__ void k_aos_to_soa(const u64* __restrict__ aos, W* __restrict__ soa, size_t n, size_t pitch, u64 mont) {
    __shared__ u64 tile[64][D + 1];
    size_t base = (size_t)blockIdx.x * 64; int cnt = (int)min((size_t)64, n - base);
    for (int i = threadIdx.x; i < cnt * D; i += blockDim.x) { u64 v = aos[base * D + i]; if (mont) v = F::mul(v % F::P, mont); tile[i / D][i % D] = v; }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * D; i += blockDim.x) { int l = i / 64, e = i % 64; if (e < cnt) soa[(size_t)l * pitch + base + e] = (W)tile[e][l]; }
}
template <class F, int D, class W> __global__ void k_soa_to_aos(const W* __restrict__ soa, u64* __restrict__ aos, size_t n, size_t pitch, u64 mont) {
    __shared__ u64 tile[64][D + 1];
    size_t base = (size_t)blockIdx.x * 64; int cnt = (int)min((size_t)64, n - base);
    for (int i = threadIdx.x; i < 64 * D; i += blockDim.x) { int l = i / 64, e = i % 64; if (e < cnt) tile[e][l] = soa[(size_t)l * pitch + base + e]; }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * D; i += blockDim.x) { u64 v = tile[i / D][i % D]; if (mont) v = F::mul(v, mont); aos[base * D + i] = v; }
}

// ------------------------------------------------------------------------------------------------ K2/K3 CRT / ICRT
// CRT::elementwise_crt / ICRT::elementwise_icrt (reference call sites arith.rs:232,238,300,327).  One thread per element;
// the D input limbs are staged in shared memory ([limb][thread], conflict free) because the sparse table indexes them
// dynamically.  tab_idx/tab_val: D rows x NNZ entries.  TIn = u64 (field elements) or int8_t (balanced digits).
template <class Rg> constexpr int matrix_apply_tpb() { return Rg::D > 32 ? 64 : 128; }      // [D][tpb] u64 staging must fit 48 KB
template <class Rg, class TIn> __global__ void __launch_bounds__(128)
k_matrix_apply(const TIn* __restrict__ in, size_t in_pitch, typename Rg::W* __restrict__ out, size_t out_pitch, size_t n,
               const int* __restrict__ tab_idx, const u64* __restrict__ tab_val, size_t in_batch_stride, size_t out_batch_stride) {
    typedef typename Rg::F F; constexpr int D = Rg::D, NNZ = Rg::S;
    in += (size_t)blockIdx.y * in_batch_stride; out += (size_t)blockIdx.y * out_batch_stride;      // blockIdx.y = vector of a batch
    __shared__ u64 s_in[D][matrix_apply_tpb<Rg>()];
    __shared__ int s_idx[D * NNZ]; __shared__ u64 s_val[D * NNZ];
    __shared__ u64 s_corr[D];      // digits only: 128 * (sum of the row's table values), see below
    for (int i = threadIdx.x; i < D * NNZ; i += blockDim.x) { s_idx[i] = tab_idx[i]; s_val[i] = tab_val[i]; }
    constexpr bool DIGITS = sizeof(TIn) == 1;
    if (DIGITS) {
        // balanced digits are small signed integers: shifted to d + 128 >= 0 they feed the two-multiply small MAC, and the
        // shift leaves as 128 * rowsum (a full 64 x 64 MAC with the field image of -1 = p - 1 would cost twice as much)
        for (int r = threadIdx.x; r < D; r += blockDim.x) { u64 sum = 0; for (int c = 0; c < NNZ; ++c) sum = F::add(sum, tab_val[r * NNZ + c]); s_corr[r] = F::mul(sum, 128); }
    }
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
#pragma unroll
        for (int l = 0; l < D; ++l) {
            if (DIGITS) s_in[l][threadIdx.x] = (u64)((int)(int8_t)in[(size_t)l * in_pitch + e] + 128);
            else s_in[l][threadIdx.x] = (u64)in[(size_t)l * in_pitch + e];
        }
    }
    __syncthreads();
    if (e >= n) return;
#pragma unroll 4
    for (int r = 0; r < D; ++r) {
        typename F::Acc a; a.clear();
#pragma unroll
        for (int c = 0; c < NNZ; ++c) {
            if (DIGITS) a.mac_small((u32)s_in[s_idx[r * NNZ + c]][threadIdx.x], s_val[r * NNZ + c]);
            else a.mac(s_val[r * NNZ + c], s_in[s_idx[r * NNZ + c]][threadIdx.x]);
        }
        out[(size_t)r * out_pitch + e] = (typename Rg::W)(DIGITS ? F::sub(F::reduce(a), s_corr[r]) : F::reduce(a));
    }
}

// ------------------------------------------------------------------------------------------------ K4/K5 digits
// gadget_decompose(B, L) (arith.rs:235): coefficient c of element i -> digits l = 0..L-1 at element i*L + l.
template <class Rg> __global__ void k_gadget_decompose(const typename Rg::W* __restrict__ in, size_t in_pitch, typename Rg::W* __restrict__ out, size_t out_pitch,
                                                       size_t n, int64_t B, int L, int* __restrict__ err) {
    typedef typename Rg::F F;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * Rg::D) return;
    size_t i = t % n; int c = (int)(t / n);
    int64_t dg[64];
    if (!balanced_digits(F::to_signed(in[(size_t)c * in_pitch + i]), B, L, dg)) atomicExch(err, 1);
    for (int l = 0; l < L; ++l) out[(size_t)c * out_pitch + i * L + l] = (typename Rg::W)F::from_i64(dg[l]);
}
// decompose_to_vec(b, K).transpose() (decomposition/utils.rs:45-49) into K int8 digit planes sets: out[k][c][i]
template <class Rg> __global__ void k_digit_split(const typename Rg::W* __restrict__ in, size_t in_pitch, int8_t* __restrict__ out, size_t out_pitch,
                                                  size_t n, int64_t b, int K, int* __restrict__ err) {
    typedef typename Rg::F F;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * Rg::D) return;
    size_t i = t % n; int c = (int)(t / n);
    int64_t dg[64];
    if (!balanced_digits(F::to_signed(in[(size_t)c * in_pitch + i]), b, K, dg)) atomicExch(err, 1);
    for (int k = 0; k < K; ++k) out[((size_t)k * Rg::D + c) * out_pitch + i] = (int8_t)dg[k];
}
// b = 2 (every reference parameter set of the 64/31-bit rings): the balanced digits of v are sign(v) times the bits of |v| -- no
// divisions, no digit array.  Four consecutive elements per thread, one 4-byte store per digit plane; grid.y = coefficient plane.
template <class Rg> __global__ void __launch_bounds__(256) k_digit_split_b2(const typename Rg::W* __restrict__ in, size_t in_pitch, int8_t* __restrict__ out, size_t out_pitch,
                                                                           size_t n, int K, int* __restrict__ err) {
    typedef typename Rg::F F;
    const size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); const int c = blockIdx.y;
    if (i >= n) return;
    u64 mag[4]; int sgn[4]; bool bad = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t v = i + q < n ? F::to_signed(in[(size_t)c * in_pitch + i + q]) : 0;
        sgn[q] = v < 0 ? -1 : 1; mag[q] = (u64)(v < 0 ? -v : v);
        bad = bad || (K < 64 && (mag[q] >> K) != 0);
    }
    if (bad) atomicExch(err, 1);
    for (int k = 0; k < K; ++k) {
        char4 d;
        d.x = (signed char)(sgn[0] * (int)((mag[0] >> k) & 1)); d.y = (signed char)(sgn[1] * (int)((mag[1] >> k) & 1));
        d.z = (signed char)(sgn[2] * (int)((mag[2] >> k) & 1)); d.w = (signed char)(sgn[3] * (int)((mag[3] >> k) & 1));
        *reinterpret_cast<char4*>(out + ((size_t)k * Rg::D + c) * out_pitch + i) = d;
    }
}
template <class Rg> __global__ void k_digits_to_field(const int8_t* __restrict__ in, size_t in_pitch, typename Rg::W* __restrict__ out, size_t out_pitch, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * Rg::D) return;
    size_t i = t % n; int c = (int)(t / n);
    out[(size_t)c * out_pitch + i] = (typename Rg::W)Rg::F::from_i64((int64_t)in[(size_t)c * in_pitch + i]);
}
// gadget_recompose(B, L): out[i] = sum_l in[i*L + l] * B^l, limb-wise (B is an integer scalar; arith.rs:305,330)
template <class Rg> __global__ void k_gadget_recompose(const typename Rg::W* __restrict__ in, size_t in_pitch, typename Rg::W* __restrict__ out, size_t out_pitch,
                                                       size_t n_out, u64 Bmod, int L, size_t in_batch_stride, size_t out_batch_stride) {
    typedef typename Rg::F F;
    in += (size_t)blockIdx.y * in_batch_stride; out += (size_t)blockIdx.y * out_batch_stride;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out * Rg::D) return;
    size_t i = t % n_out; int c = (int)(t / n_out);
    u64 acc = 0, pw = 1;
    for (int l = 0; l < L; ++l) { acc = F::add(acc, F::mul(in[(size_t)c * in_pitch + i * L + l], pw)); pw = F::mul(pw, Bmod); }
    out[(size_t)c * out_pitch + i] = (typename Rg::W)acc;
}
// get_fhat (arith.rs:273-297): MLE j, slot k, limb 0 = coefficient j*S + k; other limbs 0.  `in` points at plane j*S.
template <class Rg> __global__ void k_fhat(const typename Rg::W* __restrict__ in, size_t in_pitch, typename Rg::W* __restrict__ out, size_t out_pitch, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * Rg::S) return;
    size_t i = t % n; int k = (int)(t / n);
    out[(size_t)(k * Rg::TAU) * out_pitch + i] = in[(size_t)k * in_pitch + i];
    for (int l = 1; l < Rg::TAU; ++l) out[(size_t)(k * Rg::TAU + l) * out_pitch + i] = 0;
}

// ------------------------------------------------------------------------------------------------ K1 / K10 batched dot products
// out[r][c] = sum_x X_r[x] (.) Y_c[x]   (slot-wise product), the shape of both
//   AjtaiCommitmentScheme::commit  (rows X_r = matrix rows, Y_c = the K-1 witness pieces; commitment_scheme.rs:45-51), and
//   evaluate_mles                  (rows X_r = MLE tables, Y_0 = eq(., r) table; mle_helpers.rs:65-88).
// Work unit = (row r, tile of CT columns); one WARP per unit, lanes over x.  The warps of a block take consecutive units
// (column tile fastest) over the SAME x range, so a row is fetched from L2 once per block and re-served from L1 to the
// warps that share it, and likewise for the column vectors.  grid = (unit groups, x tiles, slots); blocks that share an
// x tile are adjacent in launch order, which keeps HBM traffic at one pass over X and Y (ncu: 2.1 GB for the 2.06 GB
// algorithmic at kappa=26, n=2^18, 15 pieces).  Accumulators are lazily reduced (F::Acc): the inner loop is
// 9 * CT 64-bit multiply-accumulates per x with no modular reduction and no branch.
template <class W> struct DotArgsT {
    const W* X; size_t x_row_stride, x_pitch; int nrows;      // rows: X + r * x_row_stride
    PtrListT<W> Y; size_t y_pitch; int ncols;                 // columns: separate vectors, len[] = effective length
    const size_t* x_len;                                      // optional per-row effective length (device), else n
    size_t n; int x_per_block;
    u64* partial;                                             // [x tile][row][col][D]
};
template <class Rg, int CT, int MAXT> __global__ void __launch_bounds__(MAXT)
k_dot(const DotArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef typename Rg::W W; constexpr int TAU = Rg::TAU;
    const int col_tiles = (a.ncols + CT - 1) / CT, n_units = a.nrows * col_tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int unit = blockIdx.x * wpb + warp, slot = blockIdx.z;
    if (unit >= n_units) return;
    const int row = unit / col_tiles, c0 = (unit % col_tiles) * CT;
    const size_t x_begin = (size_t)blockIdx.y * a.x_per_block, x_stop = min(a.n, x_begin + a.x_per_block);
    const size_t rlen = a.x_len ? a.x_len[row] : a.n;
    const W* xp = a.X + (size_t)row * a.x_row_stride + (size_t)(slot * TAU) * a.x_pitch;
    const W* yp[CT]; size_t ylen[CT];
#pragma unroll
    for (int j = 0; j < CT; ++j) { const int c = min(c0 + j, a.ncols - 1); yp[j] = a.Y.p[c] + (size_t)(slot * TAU) * a.y_pitch; ylen[j] = (c0 + j < a.ncols) ? a.Y.len[c] : 0; }
    constexpr int NA = SF::NDOT;
    typename F::Acc acc[CT][NA];
#pragma unroll
    for (int j = 0; j < CT; ++j)
#pragma unroll
        for (int l = 0; l < NA; ++l) acc[j][l].clear();
    // everything below min(lengths) needs no per-element bounds test
    size_t safe = min(x_stop, rlen);
#pragma unroll
    for (int j = 0; j < CT; ++j) if (c0 + j < a.ncols) safe = min(safe, ylen[j]);
    size_t x = x_begin + lane;
#pragma unroll 2
    for (; x < safe; x += 32) {
        u64 xv[TAU], yv[CT][TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) xv[l] = __ldg(xp + (size_t)l * a.x_pitch + x);
#pragma unroll
        for (int j = 0; j < CT; ++j)
#pragma unroll
            for (int l = 0; l < TAU; ++l) yv[j][l] = __ldg(yp[j] + (size_t)l * a.y_pitch + x);
        const typename SF::DotPrepped pr = SF::dot_prep(xv);
#pragma unroll
        for (int j = 0; j < CT; ++j) SF::dot_mac(acc[j], yv[j], pr);
    }
    for (; x < x_stop; x += 32) {            // ragged tail: operands past their effective length are zero
        if (x >= rlen) break;
        u64 xv[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) xv[l] = xp[(size_t)l * a.x_pitch + x];
        const typename SF::DotPrepped pr = SF::dot_prep(xv);
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            if (x >= ylen[j]) continue;
            u64 yv[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) yv[l] = yp[j][(size_t)l * a.y_pitch + x];
            SF::dot_mac(acc[j], yv, pr);
        }
    }
#pragma unroll
    for (int j = 0; j < CT; ++j) {
        u64 c[TAU]; SF::dot_finish(c, acc[j]);
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            u64 v = c[l];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = F::add(v, __shfl_down_sync(0xffffffffu, v, o));
            if (lane == 0 && c0 + j < a.ncols) a.partial[(((size_t)blockIdx.y * a.nrows + row) * a.ncols + (c0 + j)) * Rg::D + slot * TAU + l] = v;
        }
    }
}

// evaluate f-hat MLEs straight from coefficient planes: out[v][j][slot][l] = sum_x eq[x][slot][l] * coeff_v[x][j*S + slot]
// (compute_v_s, decomposition.rs:204-211; linearization.rs:126-131).  TIn = int8_t (digit pieces) or u64 (field coefficients).
// grid = (x tiles, slots, vectors * (TAU / JB)); partial: [x tile][vector][tau_j][D].  A block covers JB of the TAU coefficient planes of
// its slot: all of them on the narrow slot fields; THREE on the wide BabyBear field, whose 9 x 9 lazily reduced accumulators (243
// registers) spilled -- 27 per thread stay in registers and the eq limbs are re-read by the three blocks of a tile through L2
// (one plane per block, nine re-reads, measured L2-bound: 4.2 ms at configs[2] against 5.4 ms spilling).
template <class Rg> constexpr int coeff_eval_jb() { return Rg::TAU > 4 ? 3 : Rg::TAU; }
template <class Rg, class TIn> __global__ void __launch_bounds__(128)
k_coeff_eval(const TIn* __restrict__ coeff, size_t c_pitch, size_t c_vec_stride, const typename Rg::W* __restrict__ eq, size_t eq_pitch,
             size_t n, int x_per_block, int nvec, u64* __restrict__ partial) {
    typedef typename Rg::F F; constexpr int TAU = Rg::TAU, S = Rg::S, JB = coeff_eval_jb<Rg>(), JG = TAU / JB;
    __shared__ u64 red[JB * TAU * 32];
    const int slot = blockIdx.y, vec = blockIdx.z / JG, j0 = (blockIdx.z % JG) * JB;
    const TIn* cv = coeff + (size_t)vec * c_vec_stride + (size_t)j0 * S * c_pitch;
    typename F::Acc acc[JB][TAU];
#pragma unroll
    for (int j = 0; j < JB; ++j)
#pragma unroll
        for (int l = 0; l < TAU; ++l) acc[j][l].clear();
    const size_t x_begin = (size_t)blockIdx.x * x_per_block, x_end = min(n, x_begin + x_per_block);
    u64 v[JB * TAU];
    if constexpr (sizeof(TIn) == 1) {
        // digit planes: four consecutive x per thread (one 4-byte digit load per plane, two 16-byte eq loads per limb); the signed
        // digits are shifted to d + 128 >= 0 for the two-multiply small MAC and the shift leaves as 128 * sum_x eq[x]
        typename F::Sum se[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) se[l].clear();
        const size_t x_vec_end = x_begin + ((x_end - x_begin) & ~(size_t)3);       // x_begin is a multiple of x_per_block (a multiple of 4)
        for (size_t x = x_begin + 4 * (size_t)threadIdx.x; x < x_vec_end; x += 4 * (size_t)blockDim.x) {
            u64 e[TAU][4];
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                ldg_quad(eq + (size_t)(slot * TAU + l) * eq_pitch + x, e[l]);
#pragma unroll
                for (int q = 0; q < 4; ++q) se[l].add(e[l][q]);
            }
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const char4 d = *reinterpret_cast<const char4*>(cv + (size_t)(j * S + slot) * c_pitch + x);
                const u32 g[4] = {(u32)((int)d.x + 128), (u32)((int)d.y + 128), (u32)((int)d.z + 128), (u32)((int)d.w + 128)};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int l = 0; l < TAU; ++l) acc[j][l].mac_small(g[q], e[l][q]);
            }
        }
        for (size_t x = x_vec_end + threadIdx.x; x < x_end; x += blockDim.x) {      // ragged tail
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                const u64 el = eq[(size_t)(slot * TAU + l) * eq_pitch + x]; se[l].add(el);
#pragma unroll
                for (int j = 0; j < JB; ++j) acc[j][l].mac_small((u32)((int)(int8_t)cv[(size_t)(j * S + slot) * c_pitch + x] + 128), el);
            }
        }
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            const u64 corr = F::mul(F::reduce(se[l]), 128);
#pragma unroll
            for (int j = 0; j < JB; ++j) v[j * TAU + l] = F::sub(F::reduce(acc[j][l]), corr);
        }
    } else {
        for (size_t x = x_begin + threadIdx.x; x < x_end; x += blockDim.x) {
            u64 e[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) e[l] = eq[(size_t)(slot * TAU + l) * eq_pitch + x];
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const u64 c = (u64)cv[(size_t)(j * S + slot) * c_pitch + x];
#pragma unroll
                for (int l = 0; l < TAU; ++l) acc[j][l].mac(c, e[l]);
            }
        }
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int l = 0; l < TAU; ++l) v[j * TAU + l] = F::reduce(acc[j][l]);
    }
    block_reduce_add<F, JB * TAU>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int l = 0; l < TAU; ++l) partial[(((size_t)blockIdx.x * nvec + vec) * TAU + j0 + j) * Rg::D + slot * TAU + l] = v[j * TAU + l];
    }
}

// ------------------------------------------------------------------------------------------------ K7 sparse mat-vec
// mat_vec_mul (arith/utils.rs:52-65): out[row] = sum (val, col) val (.) z[col].  thread = (row, slot).  z may be the
// concatenation z = head || tail (x_s[k] || w_ccs_k, decomposition.rs:238-246) without materialising it.
template <class Rg> __global__ void k_spmv(const u32* __restrict__ row_ptr, const u32* __restrict__ col, const typename Rg::W* __restrict__ val, size_t val_pitch,
                                           const typename Rg::W* __restrict__ z_head, size_t head_len, size_t head_pitch,
                                           const typename Rg::W* __restrict__ z_tail, size_t tail_pitch, size_t tail_chunk, size_t tail_chunk_stride,
                                           typename Rg::W* __restrict__ out, size_t out_pitch, size_t nrows,
                                           size_t head_batch_stride, size_t tail_batch_stride, size_t out_batch_stride, int accumulate) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (row >= nrows) return;
    // blockIdx.z = which z vector of a batch (the K pieces of one decomposition share the matrix)
    z_head += (size_t)blockIdx.z * head_batch_stride; z_tail += (size_t)blockIdx.z * tail_batch_stride; out += (size_t)blockIdx.z * out_batch_stride;
    typename F::Acc acc[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) { acc[l].clear(); if (accumulate) acc[l].add((u64)out[(size_t)(slot * TAU + l) * out_pitch + row]); }
    for (u32 e = row_ptr[row]; e < row_ptr[row + 1]; ++e) {
        u64 v[TAU], z[TAU]; const size_t c = col[e];
        // the tail may be the all-gathered concatenation of per-rank slabs: chunk r lives at z_tail + r * tail_chunk_stride
        // (one 64-bit division per non-zero, skipped on the unsharded path where the tail is one chunk)
        const size_t tc = c - head_len;
        const size_t toff = tail_chunk == ~(size_t)0 ? tc : (tc / tail_chunk) * tail_chunk_stride + (tc % tail_chunk);
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            v[l] = val[(size_t)(slot * TAU + l) * val_pitch + e];
            z[l] = c < head_len ? z_head[(size_t)(slot * TAU + l) * head_pitch + c] : z_tail[toff + (size_t)(slot * TAU + l) * tail_pitch];
        }
        SF::mac(acc, v, SF::prep(z));
    }
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + row] = (typename Rg::W)F::reduce(acc[l]);
}

// Transposed evaluation.  An Mz MLE evaluated at a point r is  (M z)(r) = sum_row eq(row, r) (M z)[row] = sum_col (M^T eq(., r))[col] (.) z[col]
// (arith/utils.rs:52-65 followed by evaluate_mles, mle_helpers.rs:65-88, regrouped: exact arithmetic).  v = M^T eq depends only on the
// matrix and the point, so the 2K t evaluations u_s / eta of a step become t sparse products with the transposed matrix plus dot
// products over the COLUMN axis -- the axis the witness is sharded along: every rank works on its own columns, nothing is gathered
// and the per-piece Mz tables are never materialised.  eq at an arbitrary row comes from two half tables (lo over the low h
// variables, hi over the rest), which every rank holds whole.  CSC of the rank's columns: col_ptr / row (global row index) / val planes.
// thread = (local column, slot)
template <class Rg> __global__ void k_csc_eq(const u32* __restrict__ col_ptr, const u32* __restrict__ row, const typename Rg::W* __restrict__ val, size_t val_pitch,
                                             const typename Rg::W* __restrict__ lo, size_t lo_pitch, const typename Rg::W* __restrict__ hi, size_t hi_pitch, int h,
                                             typename Rg::W* __restrict__ out, size_t out_pitch, size_t ncols) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (c >= ncols) return;
    u64 acc[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) acc[l] = 0;
    for (u32 e = col_ptr[c]; e < col_ptr[c + 1]; ++e) {
        const size_t r = row[e], rlo = r & (((size_t)1 << h) - 1), rhi = r >> h;
        u64 v[TAU], a[TAU], b[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) { v[l] = val[(size_t)(slot * TAU + l) * val_pitch + e]; a[l] = lo[(size_t)(slot * TAU + l) * lo_pitch + rlo]; b[l] = hi[(size_t)(slot * TAU + l) * hi_pitch + rhi]; }
        SF::mul(a, a, b); SF::mul(a, a, v); SF::add(acc, acc, a);
    }
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + c] = (typename Rg::W)acc[l];
}

// ------------------------------------------------------------------------------------------------ K8 eq table
// build_eq_x_r (sumcheck/utils.rs:100-170): eq[x] = prod_i (x_i r_i + (1 - x_i)(1 - r_i)), r[0] on bit 0.
// r_pair: s x 2 x D limbs on the device = (1 - r_i, r_i) per variable.  thread = (x, slot).
// The table is built as lo(x mod 2^h) * hi(x div 2^h) from two half tables held in shared memory would save multiplies;
// at s <= 24 the direct product is s-1 slot-field multiplies per entry and is not on the critical path.
template <class Rg> __global__ void k_eq_table(const u64* __restrict__ r_pair, int s, typename Rg::W* __restrict__ out, size_t out_pitch, size_t n, size_t x_offset) {
    typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU, D = Rg::D;
    extern __shared__ u64 s_r[];   // s * 2 * TAU for this slot
    const int slot = blockIdx.y;
    for (int i = threadIdx.x; i < s * 2 * TAU; i += blockDim.x) { int v = i / (2 * TAU), w = (i / TAU) & 1, l = i % TAU; s_r[i] = r_pair[((size_t)v * 2 + w) * D + slot * TAU + l]; }
    __syncthreads();
    const size_t xl = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // local index; x = global hypercube index
    if (xl >= n) return;
    const size_t x = xl + x_offset;
    u64 acc[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) acc[l] = s_r[((x & 1) ? TAU : 0) + l];
    for (int v = 1; v < s; ++v) SF::mul(acc, acc, &s_r[(v * 2 + ((x >> v) & 1)) * TAU]);
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + xl] = (typename Rg::W)acc[l];
}

// eq[x] = lo[x mod 2^h] * hi[x div 2^h]: the table over s variables from two half tables (one multiply per entry instead of s-1)
template <class Rg> __global__ void k_eq_combine(const typename Rg::W* __restrict__ lo, size_t lo_pitch, const typename Rg::W* __restrict__ hi, size_t hi_pitch, int h,
                                                 typename Rg::W* __restrict__ out, size_t out_pitch, size_t n, size_t x_offset) {
    typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    const size_t xl = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (xl >= n) return;
    const size_t x = xl + x_offset, xlo = x & (((size_t)1 << h) - 1), xhi = x >> h;
    u64 a[TAU], b[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) { a[l] = __ldg(lo + (size_t)(slot * TAU + l) * lo_pitch + xlo); b[l] = __ldg(hi + (size_t)(slot * TAU + l) * hi_pitch + xhi); }
    SF::mul(a, a, b);
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + xl] = (typename Rg::W)a[l];
}

// ------------------------------------------------------------------------------------------------ K11 linear combinations
// out[x] (+)= sum_i c_i (.) v_i[x]   (compute_f_0 folding.rs:258-268; zeta-Horner combination of Mz MLEs folding.rs:208-226)
// coef: count x D limbs on the device.  thread = (x, slot).
template <class Rg> __global__ void k_lincomb(const PtrListT<typename Rg::W> vecs, size_t v_pitch, int count, const u64* __restrict__ coef,
                                              typename Rg::W* __restrict__ out, size_t out_pitch, size_t n, int accumulate) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU, D = Rg::D;
    __shared__ typename SF::Prepped s_c[MAX_LIST];
    const int slot = blockIdx.y;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_c[i] = SF::prep(coef + (size_t)i * D + slot * TAU);
    __syncthreads();
    size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    typename F::Acc acc[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) { acc[l].clear(); if (accumulate) acc[l].add(out[(size_t)(slot * TAU + l) * out_pitch + x]); }
    for (int i = 0; i < count; ++i) {
        if (x >= vecs.len[i]) continue;
        u64 v[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) v[l] = vecs.p[i][(size_t)(slot * TAU + l) * v_pitch + x];
        SF::mac(acc, v, s_c[i]);
    }
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + x] = (typename Rg::W)F::reduce(acc[l]);
}
// out[x][slot] (+)= sum_k sum_j w[k][j] * digit_k[x][j*S + slot]     (prepare_g1_and_3_k_mles_list, folding/utils.rs:524-546:
// the alpha-Horner combination of the f-hat MLEs, computed from the int8 digits).  w: K x TAU slot-field elements.
template <class Rg> __global__ void k_digit_lincomb(const int8_t* __restrict__ dig, size_t d_pitch, size_t d_vec_stride, int K,
                                                    const u64* __restrict__ w, typename Rg::W* __restrict__ out, size_t out_pitch, size_t n, int accumulate) {
    typedef typename Rg::F F; constexpr int TAU = Rg::TAU, S = Rg::S;
    __shared__ u64 s_w[MAX_MU * TAU];
    for (int i = threadIdx.x; i < K * TAU * TAU; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (x >= n) return;
    typename F::Acc acc[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) { acc[l].clear(); if (accumulate) acc[l].add(out[(size_t)(slot * TAU + l) * out_pitch + x]); }
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < TAU; ++j) {
            u64 c = F::from_i64((int64_t)dig[(size_t)k * d_vec_stride + (size_t)(j * S + slot) * d_pitch + x]);
#pragma unroll
            for (int l = 0; l < TAU; ++l) acc[l].mac(c, s_w[(k * TAU + j) * TAU + l]);
        }
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * out_pitch + x] = (typename Rg::W)F::reduce(acc[l]);
}

// ------------------------------------------------------------------------------------------------ K9 sumcheck
// fix_variables on a list of tables (sumcheck/prover.rs:61-72): new[b] = old[2b] + r (old[2b+1] - old[2b]).
// in/out may alias only through separate buffers (ping-pong).  r_sf: TAU limbs (slot-constant challenge).
// grid = (b tiles, slots, tables)
template <class W> struct FoldArgsT { const W* in; W* out; size_t in_pitch, out_pitch, in_stride, out_stride; size_t n_out; u64 r[16]; };      // r: TAU limbs of the challenge (TAU <= 9)
template <class Rg> __global__ void k_fold(const FoldArgsT<typename Rg::W> a) {
    typedef SlotField<Rg> SF; typedef typename Rg::W W; constexpr int TAU = Rg::TAU;
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (b >= a.n_out) return;
    const W* in = a.in + (size_t)blockIdx.z * a.in_stride; W* out = a.out + (size_t)blockIdx.z * a.out_stride;
    u64 f0[TAU], f1[TAU], t[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) ld_pair(in + (size_t)(slot * TAU + l) * a.in_pitch + 2 * b, f0[l], f1[l]);
    SF::sub(t, f1, f0); SF::mul(t, t, a.r); SF::add(t, t, f0);
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)(slot * TAU + l) * a.out_pitch + b] = (W)t[l];
}

// generic round evaluation for PRODUCTS / LIN combination functions (prove_round, sumcheck/prover.rs:111-143):
// evals[e] = sum_b comb(v_k(2b) + e (v_k(2b+1) - v_k(2b))), e = 0..deg.   thread = (b, slot); partial: [b tile][deg+1][D]
template <class W> struct ScGenericArgsT {
    const W* mle[SC_MAX_MLES]; size_t pitch; int n_mles, deg, n_terms, lin;
    int term_len[SC_MAX_TERMS]; int term_idx[SC_MAX_TERMS][SC_MAX_FACTORS];
    const u64* coef;           // n_terms x D on the device
    size_t n_pairs; u64* partial;
};
template <class Rg, int NM> __global__ void __launch_bounds__(128)
k_sc_generic(const ScGenericArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    __shared__ u64 red[(SC_MAX_DEG + 1) * TAU * 32];
    const int slot = blockIdx.y;
    u64 ev[SC_MAX_DEG + 1][TAU];
#pragma unroll
    for (int e = 0; e <= SC_MAX_DEG; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) ev[e][l] = 0;
    u64 cf[SC_MAX_TERMS][TAU];
#pragma unroll
    for (int t = 0; t < SC_MAX_TERMS; ++t)
#pragma unroll
        for (int l = 0; l < TAU; ++l) cf[t][l] = t < a.n_terms ? a.coef[(size_t)t * Rg::D + slot * TAU + l] : 0;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.n_pairs; b += (size_t)gridDim.x * blockDim.x) {
        u64 val[NM][TAU], step[NM][TAU];
#pragma unroll
        for (int k = 0; k < NM; ++k)
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                u64 p1; ld_pair(a.mle[k] + (size_t)(slot * TAU + l) * a.pitch + 2 * b, val[k][l], p1);
                step[k][l] = F::sub(p1, val[k][l]);
            }
#pragma unroll
        for (int e = 0; e <= SC_MAX_DEG; ++e) {
            if (e > a.deg) break;
            u64 res[TAU] = {0};
#pragma unroll
            for (int t = 0; t < SC_MAX_TERMS; ++t) {
                if (t >= a.n_terms) break;
                u64 term[TAU];
#pragma unroll
                for (int l = 0; l < TAU; ++l) term[l] = cf[t][l];
#pragma unroll
                for (int f = 0; f < SC_MAX_FACTORS; ++f) {
                    if (f >= a.term_len[t]) break;
                    const int idx = a.term_idx[t][f];
                    u64 fac[TAU];
#pragma unroll
                    for (int k = 0; k < NM; ++k) if (k == idx) {
#pragma unroll
                        for (int l = 0; l < TAU; ++l) fac[l] = val[k][l];
                    }
                    SF::mul(term, term, fac);
                }
                SF::add(res, res, term);
            }
            if (a.lin) {
                u64 last[TAU];
#pragma unroll
                for (int k = 0; k < NM; ++k) if (k == a.n_mles - 1) {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) last[l] = val[k][l];
                }
                SF::mul(res, res, last);
            }
            SF::add(ev[e], ev[e], res);
#pragma unroll
            for (int k = 0; k < NM; ++k) SF::add(val[k], val[k], step[k]);
        }
    }
    u64 v[(SC_MAX_DEG + 1) * TAU];
#pragma unroll
    for (int e = 0; e <= SC_MAX_DEG; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) v[e * TAU + l] = ev[e][l];
    block_reduce_add<F, (SC_MAX_DEG + 1) * TAU>(v, red);
    if (threadIdx.x == 0)
        for (int e = 0; e <= a.deg; ++e)
#pragma unroll
            for (int l = 0; l < TAU; ++l) a.partial[((size_t)blockIdx.x * (a.deg + 1) + e) * Rg::D + slot * TAU + l] = v[e * TAU + l];
}

// The same round evaluation with one thread per (pair, evaluation point): (deg+1) times the parallelism and a (deg+1)-th of the
// live state per thread -- the wide slot field of the BabyBear ring (9 limbs: 45 table words per point) spilled and ran at two warps
// per scheduler in the one-thread-per-pair form above, and the late, small rounds of every ring are latency bound.  Thread
// (pair b, point e) forms v_k(e) = v_k(2b) + e (v_k(2b+1) - v_k(2b)) for every table and evaluates the combination once.
template <class Rg, int NM> __global__ void __launch_bounds__(128)
k_sc_points(const ScGenericArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef typename Rg::W E; constexpr int TAU = Rg::TAU;
    __shared__ u64 red[TAU * 128];
    const int slot = blockIdx.y, npts = a.deg + 1, ppb = blockDim.x / npts;
    const int e = threadIdx.x % npts, pl = threadIdx.x / npts;
    E cf[SC_MAX_TERMS][TAU];
#pragma unroll
    for (int t = 0; t < SC_MAX_TERMS; ++t)
#pragma unroll
        for (int l = 0; l < TAU; ++l) cf[t][l] = t < a.n_terms ? (E)a.coef[(size_t)t * Rg::D + slot * TAU + l] : (E)0;
    E ev[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) ev[l] = 0;
    if (pl < ppb)
    for (size_t b = (size_t)blockIdx.x * ppb + pl; b < a.n_pairs; b += (size_t)gridDim.x * ppb) {
        E val[NM][TAU];
#pragma unroll
        for (int k = 0; k < NM; ++k) {
            if (k >= a.n_mles) break;
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                u64 v0, v1; ld_pair(a.mle[k] + (size_t)(slot * TAU + l) * a.pitch + 2 * b, v0, v1);
                const u64 st = F::sub(v1, v0); u64 x = v0;
#pragma unroll
                for (int i = 0; i < SC_MAX_DEG; ++i) if (i < e) x = F::add(x, st);
                val[k][l] = (E)x;
            }
        }
        E res[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) res[l] = 0;
#pragma unroll 1
        for (int t = 0; t < a.n_terms; ++t) {
            E term[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) term[l] = cf[t][l];
#pragma unroll 1
            for (int f = 0; f < a.term_len[t]; ++f) {
                const int idx = a.term_idx[t][f];
                E fac[TAU];
#pragma unroll
                for (int k = 0; k < NM; ++k) if (k == idx) {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) fac[l] = val[k][l];
                }
                SF::mul_inl(term, term, fac);
            }
            SF::add(res, res, term);
        }
        if (a.lin) {
            E last[TAU];
#pragma unroll
            for (int k = 0; k < NM; ++k) if (k == a.n_mles - 1) {
#pragma unroll
                for (int l = 0; l < TAU; ++l) last[l] = val[k][l];
            }
            SF::mul_inl(res, res, last);
        }
        SF::add(ev, ev, res);
    }
#pragma unroll
    for (int l = 0; l < TAU; ++l) red[l * 128 + threadIdx.x] = (u64)ev[l];
    __syncthreads();
    if (threadIdx.x < npts * TAU) {
        const int pe = threadIdx.x / TAU, l = threadIdx.x % TAU;
        u64 acc = 0;
        for (int q = 0; q < ppb; ++q) acc = F::add(acc, red[l * 128 + q * npts + pe]);
        a.partial[((size_t)blockIdx.x * npts + pe) * Rg::D + slot * TAU + l] = acc;
    }
}

// The same per-point evaluation for ARBITRARY sums of products: any number of tables (one contiguous group: table k at base + k *
// stride), any number of terms and factors -- the reference's `comb_fn` is an arbitrary closure; every one it ships is a sum of products
// with ring coefficients (CCS of any shape, linearization/utils.rs:90-107; LatticeFold+ v0 (v1 v2 - v3), latticefold-plus/src/r1cs.rs:92;
// the set-check and commitment-transformation batches, setchk.rs:155-186, cm.rs:285-307).  The term list lives in device memory
// (term_off[n_terms + 1], idx[]); factors are re-read per term (L1 resident) instead of being held in registers.
template <class W> struct ScTermsArgsT {
    const W* base; size_t stride, pitch; int n_mles, deg, n_terms, lin;
    const int* term_off; const int* idx; const u64* coef;      // coef: n_terms x D
    size_t n_pairs; u64* partial;
};
template <class Rg> __global__ void __launch_bounds__(128)
k_sc_terms(const ScTermsArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; typedef typename Rg::W E; constexpr int TAU = Rg::TAU;
    __shared__ u64 red[TAU * 128];
    const int slot = blockIdx.y, npts = a.deg + 1, ppb = blockDim.x / npts;
    const int e = threadIdx.x % npts, pl = threadIdx.x / npts;
    E ev[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) ev[l] = 0;
    auto at_point = [&](int k, size_t b, E* out) {
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            u64 v0, v1; ldg_pair(a.base + (size_t)k * a.stride + (size_t)(slot * TAU + l) * a.pitch + 2 * b, v0, v1);
            const u64 st = F::sub(v1, v0); u64 x = v0;
#pragma unroll
            for (int i = 0; i < SC_MAX_DEG; ++i) if (i < e) x = F::add(x, st);
            out[l] = (E)x;
        }
    };
    if (pl < ppb)
    for (size_t b = (size_t)blockIdx.x * ppb + pl; b < a.n_pairs; b += (size_t)gridDim.x * ppb) {
        E res[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) res[l] = 0;
#pragma unroll 1
        for (int t = 0; t < a.n_terms; ++t) {
            E term[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) term[l] = (E)a.coef[(size_t)t * Rg::D + slot * TAU + l];
#pragma unroll 1
            for (int f = a.term_off[t]; f < a.term_off[t + 1]; ++f) { E fac[TAU]; at_point(a.idx[f], b, fac); SF::mul_inl(term, term, fac); }
            SF::add(res, res, term);
        }
        if (a.lin) { E last[TAU]; at_point(a.n_mles - 1, b, last); SF::mul_inl(res, res, last); }
        SF::add(ev, ev, res);
    }
#pragma unroll
    for (int l = 0; l < TAU; ++l) red[l * 128 + threadIdx.x] = (u64)ev[l];
    __syncthreads();
    if (threadIdx.x < npts * TAU) {
        const int pe = threadIdx.x / TAU, l = threadIdx.x % TAU;
        u64 acc = 0;
        for (int q = 0; q < ppb; ++q) acc = F::add(acc, red[l * 128 + q * npts + pe]);
        a.partial[((size_t)blockIdx.x * npts + pe) * Rg::D + slot * TAU + l] = acc;
    }
}

// ---- FOLD combination function, b = 2 (folding/utils.rs:273-325):
//   g(x) = v0 v1 + v2 v3 + v4 * h(x),   h = sum_{k<2K} sum_{d<tau} mu_k^{d+1} (f_{k,d}^3 - f_{k,d})
// h is a cubic along the line through a pair, so 4 points determine it; the degree-4 message needs 5 points of g.
// "dense" = the five tables [eq(r_acc), G_acc, eq(r_new), G_new, eq(beta)].
template <class W> struct FoldScArgsT {
    const W* dense; size_t dense_pitch, dense_stride;       // 5 tables
    const u64* mu_pow;                                      // n_f x TAU limbs: mu_k^{d+1} (slot-constant)
    int n_f;                                                // 2K * tau
    size_t n_pairs; u64* partial;                           // [b tile][5][D]
    // round 1: int8 digits, piece k at dig + k * dig_stride, coefficient plane c at c * dig_pitch
    const int8_t* dig; size_t dig_pitch, dig_stride;
    // rounds >= 2: slot-field tables, f-hat (k,d) at fh + (k*tau+d) * fh_stride
    const W* fh; size_t fh_pitch, fh_stride;
    u64 r1[16];                                             // round 2 from digits: the first challenge (TAU limbs)
};
template <class Rg, bool WITH_PRODUCTS = true> __device__ __forceinline__ void fold_sc_tail(const FoldScArgsT<typename Rg::W>& a, size_t b, bool active, int slot, const u64 (*h)[Rg::TAU] /* h(0..4) */, u64* red, size_t partial_block = ~(size_t)0, bool rt_products = true) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    u64 ev[5][TAU];
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) ev[e][l] = 0;
    if (active) {
        u64 val[5][TAU], step[5][TAU];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (!WITH_PRODUCTS && k < 4) continue;
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                u64 p1; ld_pair(a.dense + (size_t)k * a.dense_stride + (size_t)(slot * TAU + l) * a.dense_pitch + 2 * b, val[k][l], p1);
                step[k][l] = F::sub(p1, val[k][l]);
            }
        }
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            // eq(beta) h + eq(r_acc) G_acc + eq(r_new) G_new as one lazily reduced sum of products: three reductions instead of nine
            typename F::Acc acc3[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) acc3[l].clear();
            SF::mac(acc3, val[4], SF::prep(h[e]));
            if (WITH_PRODUCTS && rt_products) { SF::mac(acc3, val[0], SF::prep(val[1])); SF::mac(acc3, val[2], SF::prep(val[3])); }
#pragma unroll
            for (int l = 0; l < TAU; ++l) ev[e][l] = F::reduce(acc3[l]);
#pragma unroll
            for (int k = 0; k < 5; ++k) { if (!WITH_PRODUCTS && k < 4) continue; SF::add(val[k], val[k], step[k]); }
        }
    }
    u64 v[5 * TAU];
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) v[e * TAU + l] = ev[e][l];
    block_reduce_add<F, 5 * TAU>(v, red);
    const size_t pb = partial_block == ~(size_t)0 ? (size_t)blockIdx.x : partial_block;
    if (threadIdx.x == 0)
#pragma unroll
        for (int e = 0; e < 5; ++e)
#pragma unroll
            for (int l = 0; l < TAU; ++l) a.partial[(pb * 5 + e) * Rg::D + slot * TAU + l] = v[e * TAU + l];
}
// round 1: every f-hat entry is a balanced digit in {-1,0,1} embedded in the base field (arith.rs:283-289), so
// f^3 - f vanishes at X = 0, 1 and is a small integer at X = 2, 3; h(4) follows from the cubic's finite differences.
template <class Rg> __global__ void __launch_bounds__(128, Rg::TAU <= 3 ? 3 : 1)
k_fold_sc_round1(const FoldScArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; constexpr int TAU = Rg::TAU, S = Rg::S;
    __shared__ u64 red[5 * TAU * 32];
    __shared__ u64 s_mu[MAX_MU * TAU];
    __shared__ u64 s_msum[TAU];                              // sum of all mu (the same for every pair: formed once per block)
    __shared__ u32 s_g[9];                                   // (g2 + 24) | (g3 + 120) << 16 by digit pair
    for (int i = threadIdx.x; i < a.n_f * TAU; i += blockDim.x) s_mu[i] = a.mu_pow[i];
    if (threadIdx.x < 9) { const int x = (int)threadIdx.x % 3 - 1, y = (int)threadIdx.x / 3 - 1, f2 = 2 * y - x, f3 = 3 * y - 2 * x;
        s_g[threadIdx.x] = (u32)(f2 * f2 * f2 - f2 + 24) | ((u32)(f3 * f3 * f3 - f3 + 120) << 16); }
    __syncthreads();
    if (threadIdx.x < TAU) { typename F::Sum sm; sm.clear(); for (int kd = 0; kd < a.n_f; ++kd) sm.add(s_mu[kd * TAU + threadIdx.x]); s_msum[threadIdx.x] = F::reduce(sm); }
    __syncthreads();
    const int slot = blockIdx.y;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const bool active = b < a.n_pairs;
    u64 h[5][TAU];
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) h[e][l] = 0;
    if (active) {
        // g(X) = f(X)^3 - f(X) is one of 0, +-6, +-24 at X = 2 and 0, +-6, ..., +-120 at X = 3.  The signed weights are shifted to
        // non-negative ones (g + 24, g + 120) so that every table feeds the same branch-free small-multiplier MACs -- no lane
        // divergence on the digits, and the loads of several tables can be in flight together -- and the shift is taken out
        // again as 24 / 120 times the sum of all mu.  The weights come from a nine-entry table indexed by the digit pair (the
        // integer multiplies of the closed form sit on the same FMA-heavy pipe as the wide multiplies that bound the kernel).
        typename F::AccS a2[TAU], a3[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) { a2[l].clear(); a3[l].clear(); }
#pragma unroll 4
        for (int kd = 0; kd < a.n_f; ++kd) {
            const int k = kd / TAU, d = kd - k * TAU;
            const char2 dd = *reinterpret_cast<const char2*>(a.dig + (size_t)k * a.dig_stride + (size_t)(d * S + slot) * a.dig_pitch + 2 * b);
            const u32 g = s_g[(dd.x + 1) + 3 * (dd.y + 1)], g2 = g & 0xffffu, g3 = g >> 16;
            const u64* mu = &s_mu[kd * TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) { a2[l].mac_small(g2, mu[l]); a3[l].mac_small(g3, mu[l]); }
        }
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            const u64 ms = s_msum[l];
            const u64 h2 = F::sub(F::reduce(a2[l]), F::mul(ms, 24)), h3 = F::sub(F::reduce(a3[l]), F::mul(ms, 120));
            h[2][l] = h2; h[3][l] = h3;
            // h(0) = h(1) = 0 and third differences constant: h(4) = 4 h(3) - 6 h(2)
            const u64 h3x2 = F::add(h3, h3), h3x4 = F::add(h3x2, h3x2), h2x2 = F::add(h2, h2), h2x6 = F::add(F::add(h2x2, h2x2), h2x2);
            h[4][l] = F::sub(h3x4, h2x6);
        }
    }
    fold_sc_tail<Rg>(a, b, active, slot, h, red);
}
// after the first challenge r: f-hat tables become slot-field valued: new[b] = d0 + r (d1 - d0)
template <class Rg> __global__ void k_fold_digits(const int8_t* __restrict__ dig, size_t dig_pitch, size_t dig_stride, int n_f,
                                                  typename Rg::W* __restrict__ out, size_t out_pitch, size_t out_stride, size_t n_out, const FoldArgsT<typename Rg::W> r) {
    typedef typename Rg::F F; constexpr int TAU = Rg::TAU, S = Rg::S;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y, kd = blockIdx.z;
    if (b >= n_out) return;
    const int k = kd / TAU, d = kd % TAU;
    const char2 dd = *reinterpret_cast<const char2*>(dig + (size_t)k * dig_stride + (size_t)(d * S + slot) * dig_pitch + 2 * b);
    const u64 st = F::from_i64((int64_t)dd.y - dd.x), d0 = F::from_i64((int64_t)dd.x);
    typename Rg::W* o = out + (size_t)kd * out_stride;
#pragma unroll
    for (int l = 0; l < TAU; ++l) { u64 v = F::mul(st, r.r[l]); if (l == 0) v = F::add(v, d0); o[(size_t)(slot * TAU + l) * out_pitch + b] = (typename Rg::W)v; }
}
// Round 2 straight from the digits.  After the first challenge r an f-hat entry is T1[i] = d[2i] + r (d[2i+1] - d[2i]), so along the
// line through the pair (T1[2b], T1[2b+1]) the value at an integer point X is f = P(X) + r Q(X) with small INTEGERS
//   P = a0 + X (a1 - a0),  Q = e0 + X (e1 - e0),    a0 = d[4b], e0 = d[4b+1] - d[4b], a1 = d[4b+2], e1 = d[4b+3] - d[4b+2],
// and  f^3 - f = (P^3 - P) + r Q (3 P^2 - 1) + r^2 3 P Q^2 + r^3 Q^3.   Summed over the tables with the weights mu this is
//   h(X) = S0 + r S1 + r^2 S2 + r^3 S3,   S_j = sum_kd mu_kd c_j(kd, X)   with |c_j| < 2048,
// i.e. four sums of mu with small integer weights per point -- two-multiply MACs on 96-bit accumulators, no slot-field product per
// table -- instead of the 66 full multiply-accumulates per table of the general round kernel, and the 2.4 GB of T1 tables are never
// written or read.  Four lanes per pair take the points X = 0..3 (h is a cubic: h(4) follows from its finite differences).
// A block walks R2_GROUPS groups of 32 pairs; the round message's five points are spread over the four lanes of a pair for the dense
// part too -- lane X evaluates g at point X (its own h(X) times eq(beta) plus the two eq * G products, one lazily reduced sum of three
// slot-field products), lane 0 also point 4 -- and the block's partial is reduced once.  (ncu r02y of the one-group form with the
// whole dense part on lane 0: 6426 of a warp's 13 200 instructions were that tail and the 15-value block reduction.)
constexpr int R2_GROUPS = 4;
template <class Rg> __global__ void __launch_bounds__(128, Rg::TAU <= 3 ? 4 : 1)      // four blocks per SM on the narrow slot field (132 B of spills in the dense part)
k_fold_sc_round2(const FoldScArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU, S = Rg::S;
    __shared__ u64 s_corr[TAU];                              // 2048 * sum of all mu
    __shared__ ushort4 s_lut[81 * 4];
    extern __shared__ __align__(16) unsigned char dyn_smem[];   // n_f * TAU mu limbs, then the group's digits [table][32 pairs]
    u64* s_mu = reinterpret_cast<u64*>(dyn_smem);
    char4 (*s_dig)[32] = reinterpret_cast<char4 (*)[32]>(dyn_smem + (size_t)a.n_f * TAU * 8);
    const int slot = blockIdx.y, X = threadIdx.x & 3, pl = threadIdx.x >> 2;
    for (int i = threadIdx.x; i < a.n_f * TAU; i += blockDim.x) s_mu[i] = a.mu_pow[i];
    // the four weights (+ 2048) by digit quadruple and point: 81 x 4 entries (one dp4a and one load per table instead of the closed forms)
    for (int i = threadIdx.x; i < 81 * 4; i += blockDim.x) {
        const int q = i >> 2, Xi = i & 3, d0 = q % 3 - 1, d1 = (q / 3) % 3 - 1, d2 = (q / 9) % 3 - 1, d3 = q / 27 - 1;
        const int a0 = d0, e0 = d1 - d0, a1 = d2, e1 = d3 - d2, P = a0 + Xi * (a1 - a0), Q = e0 + Xi * (e1 - e0);
        s_lut[i] = make_ushort4((unsigned short)(P * P * P - P + 2048), (unsigned short)(Q * (3 * P * P - 1) + 2048), (unsigned short)(3 * P * Q * Q + 2048), (unsigned short)(Q * Q * Q + 2048));
    }
    __syncthreads();
    if (threadIdx.x < TAU) { typename F::Sum sm; sm.clear(); for (int kd = 0; kd < a.n_f; ++kd) sm.add(s_mu[kd * TAU + threadIdx.x]); s_corr[threadIdx.x] = F::mul(F::reduce(sm), 2048); }
    u64 evx[TAU], ev4[TAU];                                  // running sums of g(X) (every lane) and g(4) (lane 0 of each quad)
#pragma unroll
    for (int l = 0; l < TAU; ++l) evx[l] = ev4[l] = 0;
    for (int grp = 0; grp < R2_GROUPS; ++grp) {
        const size_t b0 = ((size_t)blockIdx.x * R2_GROUPS + grp) * 32, b = b0 + pl; const bool active = b < a.n_pairs;
        if (b0 >= a.n_pairs) break;
        __syncthreads();                                     // the previous group's digits have been consumed (first pass: s_corr is complete)
        // all digit loads of the group are issued up front (n_f / 4 independent 4-byte loads per thread, one 128-byte row per warp and table)
        for (int kd = threadIdx.x >> 5; kd < a.n_f; kd += 4) {
            const int k = kd / TAU, d = kd - k * TAU; const int lane = threadIdx.x & 31;
            char4 v = make_char4(0, 0, 0, 0);
            if (b0 + lane < a.n_pairs) v = *reinterpret_cast<const char4*>(a.dig + (size_t)k * a.dig_stride + (size_t)(d * S + slot) * a.dig_pitch + 4 * (b0 + lane));
            s_dig[kd][lane] = v;
        }
        __syncthreads();
        u64 hx[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) hx[l] = 0;
        if (active) {
            typename F::AccS acc[4][TAU];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int l = 0; l < TAU; ++l) acc[j][l].clear();
#pragma unroll 2
            for (int kd = 0; kd < a.n_f; ++kd) {
                const int q = __dp4a(reinterpret_cast<const int*>(&s_dig[kd][0])[pl], 0x1B090301, 40);      // (d0 + 1) + 3 (d1 + 1) + 9 (d2 + 1) + 27 (d3 + 1)
                const ushort4 cw = s_lut[q * 4 + X];
                const u32 c0 = cw.x, c1 = cw.y, c2 = cw.z, c3 = cw.w;
                const u64* mu = &s_mu[kd * TAU];
#pragma unroll
                for (int l = 0; l < TAU; ++l) { const u64 m = mu[l]; acc[0][l].mac_small(c0, m); acc[1][l].mac_small(c1, m); acc[2][l].mac_small(c2, m); acc[3][l].mac_small(c3, m); }
            }
            u64 sj[4][TAU];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int l = 0; l < TAU; ++l) sj[j][l] = F::sub(F::reduce(acc[j][l]), s_corr[l]);
            // Horner in r
            SF::mul(hx, sj[3], a.r1); SF::add(hx, hx, sj[2]); SF::mul(hx, hx, a.r1); SF::add(hx, hx, sj[1]); SF::mul(hx, hx, a.r1); SF::add(hx, hx, sj[0]);
        }
        // h(4) = 4 h(3) - 6 h(2) + 4 h(1) - h(0) (h is a cubic), formed in every lane of the quad (lane 0 uses it)
        u64 h4[TAU];
        const int base = (threadIdx.x & 31) & ~3;
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            const u64 h0 = __shfl_sync(0xffffffffu, hx[l], base), h1 = __shfl_sync(0xffffffffu, hx[l], base + 1), h2 = __shfl_sync(0xffffffffu, hx[l], base + 2), h3 = __shfl_sync(0xffffffffu, hx[l], base + 3);
            const u64 t31 = F::add(h3, h1), t31x2 = F::add(t31, t31), t31x4 = F::add(t31x2, t31x2), h2x2 = F::add(h2, h2), h2x6 = F::add(F::add(h2x2, h2x2), h2x2);
            h4[l] = F::sub(F::sub(t31x4, h2x6), h0);
        }
        if (active) {
            // g at this lane's point: v_k(X) = v_k(0) + X step_k, then eq(beta) h + eq(r_acc) G_acc + eq(r_new) G_new as ONE lazily reduced sum
            auto g_at = [&](const u64 (*v)[TAU], const u64* hh, u64* out) {
                typename F::Acc acc3[TAU];
#pragma unroll
                for (int l = 0; l < TAU; ++l) acc3[l].clear();
                SF::mac(acc3, v[4], SF::prep(hh)); SF::mac(acc3, v[0], SF::prep(v[1])); SF::mac(acc3, v[2], SF::prep(v[3]));
#pragma unroll
                for (int l = 0; l < TAU; ++l) out[l] = F::reduce(acc3[l]);
            };
            // the pair's dense entries (the same addresses in the four lanes of a quad: one broadcast request)
            u64 val[5][TAU], stp[5][TAU];
#pragma unroll
            for (int k = 0; k < 5; ++k)
#pragma unroll
                for (int l = 0; l < TAU; ++l) { u64 p1; ld_pair(a.dense + (size_t)k * a.dense_stride + (size_t)(slot * TAU + l) * a.dense_pitch + 2 * b, val[k][l], p1); stp[k][l] = F::sub(p1, val[k][l]); }
            u64 v[5][TAU], g[TAU];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
#pragma unroll
                for (int l = 0; l < TAU; ++l) v[k][l] = val[k][l];
#pragma unroll
                for (int i = 0; i < 3; ++i) if (i < X) SF::add(v[k], v[k], stp[k]);
            }
            g_at(v, hx, g);
            SF::add(evx, evx, g);
            if (X == 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k) { u64 s2[TAU]; SF::add(s2, stp[k], stp[k]); SF::add(s2, s2, s2); SF::add(v[k], val[k], s2); }
                g_at(v, h4, g);
                SF::add(ev4, ev4, g);
            }
        }
    }
    // quads of a warp -> lanes 0..3 hold the warp's sums of point X (lane X) and point 4 (lane 0); warps through shared memory
#pragma unroll
    for (int l = 0; l < TAU; ++l) {
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) { evx[l] = F::add(evx[l], __shfl_xor_sync(0xffffffffu, evx[l], o)); ev4[l] = F::add(ev4[l], __shfl_xor_sync(0xffffffffu, ev4[l], o)); }
    }
    __shared__ u64 s_w[4][5][TAU];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 4) {
#pragma unroll
        for (int l = 0; l < TAU; ++l) { s_w[warp][lane][l] = evx[l]; if (lane == 0) s_w[warp][4][l] = ev4[l]; }
    }
    __syncthreads();
    if (threadIdx.x < 5 * TAU) {
        const int e = threadIdx.x / TAU, l = threadIdx.x % TAU;
        const u64 sum = F::add(F::add(s_w[0][e][l], s_w[1][e][l]), F::add(s_w[2][e][l], s_w[3][e][l]));
        a.partial[((size_t)blockIdx.x * 5 + e) * Rg::D + slot * TAU + l] = sum;
    }
}
// after the second challenge the f-hat tables are materialised for the first time, again from the digits:
//   T2[b] = T1[2b] + r2 (T1[2b+1] - T1[2b]) = a0 + r1 e0 + r2 (a1 - a0) + r1 r2 (e1 - e0)
// cs (device): TAU x 3 limbs (r1, r2, r1 r2) followed by TAU limbs of the correction 2 r1 + 2 r2 + 4 r1 r2 (the signed weights are shifted
// to non-negative ones for the two-multiply MACs).  grid = (b tiles, slots, tables)
template <class Rg> __global__ void k_fold_digits2(const int8_t* __restrict__ dig, size_t dig_pitch, size_t dig_stride, int n_f,
                                                   typename Rg::W* __restrict__ out, size_t out_pitch, size_t out_stride, size_t n_out, const u64* __restrict__ cs) {
    typedef typename Rg::F F; constexpr int TAU = Rg::TAU, S = Rg::S;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y, kd = blockIdx.z;
    if (b >= n_out) return;
    const int k = kd / TAU, d = kd % TAU;
    const char4 dd = *reinterpret_cast<const char4*>(dig + (size_t)k * dig_stride + (size_t)(d * S + slot) * dig_pitch + 4 * b);
    const int a0 = dd.x, e0 = dd.y - dd.x, al = dd.z - dd.x, be = (dd.w - dd.z) - e0;      // |e0| <= 2, |al| <= 2, |be| <= 4
    typename Rg::W* o = out + (size_t)kd * out_stride;
#pragma unroll
    for (int l = 0; l < TAU; ++l) {
        typename F::AccS x; x.clear();
        x.mac_small((u32)(e0 + 2), cs[l]); x.mac_small((u32)(al + 2), cs[TAU + l]); x.mac_small((u32)(be + 4), cs[2 * TAU + l]);
        u64 v = F::sub(F::reduce(x), cs[3 * TAU + l]);
        if (l == 0) v = F::add(v, F::from_i64((int64_t)a0));
        o[(size_t)(slot * TAU + l) * out_pitch + b] = (typename Rg::W)v;
    }
}
// rounds >= 2.  Along the pair's line f(X) = u + X s:
//   f^3 - f = (u^3 - u) + 3X u^2 s + 3X^2 u s^2 + X^3 (s^3 - s) + (X^3 - X) s,
// so h(X) = A + 3X B + 3X^2 C + X^3 D + (X^3 - X) Es with five mu-weighted sums over all 2K*tau tables that stay lazily reduced.
// (History: one thread per pair with all five sums needed 211 registers; two thread sets over {A,B,Es} / {C,D} -- 95 MACs per
// (pair, slot, table), 10.0 ms per step -- were the round-1 form before the two-lane kernel below: 7.8 ms.)
// Rounds >= 2, two lanes per pair.  With w = mu u and z = mu s (one slot-field product each) the five sums become
//   A = sum (w u^2 - w),  B = sum z u^2,  C = sum w s^2,  D = sum (z s^2 - z),  Es = sum z
// -- 68 instead of 95 multiply-accumulates and 12 instead of 15 reductions per (pair, slot, table).  Holding all of it in one
// thread needs 12 lazily reduced accumulators (168 registers with spills, 12 warps per SM: measured 9.5 ms, latency bound), so
// the work is split over two ADJACENT LANES that run the same code on different operands: lane 0 owns t = u, lane 1 owns t = s;
// each forms mine = mu t and q = t^2, swaps `mine` with its neighbour by one shuffle per limb, and accumulates mine*q, other*q
// and the plain sum of mine.  Lane 0 thus holds {A, B}, lane 1 {D, C, Es}; the round message is linear in them.
// 128 registers (4 blocks of 128 threads per SM) with the table loop not unrolled measured best: 152 registers / 3 blocks 8.3 ms,
// 96 registers / 5 blocks 7.9 ms, this 7.8 ms.
template <class Rg> __global__ void __launch_bounds__(128, 4)
k_fold_sc_round(const FoldScArgsT<typename Rg::W> a) {
    typedef typename Rg::F F; typedef SlotField<Rg> SF; constexpr int TAU = Rg::TAU;
    __shared__ u64 red[5 * TAU * 32];
    // mu and its nu-multiples, prepared once per block where that fits the static shared-memory budget (5 words per table on the
    // Goldilocks ring); the wide-slot rings keep mu itself and prepare it per table
    constexpr bool PREP_SMEM = sizeof(typename SF::Prepped) * MAX_MU <= 16384;
    typedef typename std::conditional<PREP_SMEM, typename SF::Prepped, u64>::type MuT;
    __shared__ MuT s_mu[PREP_SMEM ? MAX_MU : MAX_MU * TAU];
    if constexpr (PREP_SMEM) { for (int i = threadIdx.x; i < a.n_f; i += blockDim.x) s_mu[i] = SF::prep(a.mu_pow + (size_t)i * TAU); }
    else { for (int i = threadIdx.x; i < a.n_f * TAU; i += blockDim.x) s_mu[i] = a.mu_pow[i]; }
    __syncthreads();
    const int slot = blockIdx.y, role = threadIdx.x & 1;
    const size_t b = (size_t)blockIdx.x * (blockDim.x / 2) + (threadIdx.x >> 1); const bool active = b < a.n_pairs;
    const unsigned lanes = __ballot_sync(0xffffffffu, active);      // both lanes of a pair are active together
    u64 h[5][TAU];
#pragma unroll
    for (int e = 0; e < 5; ++e)
#pragma unroll
        for (int l = 0; l < TAU; ++l) h[e][l] = 0;
    if (active) {
        typename F::Acc s_mine[TAU], s_other[TAU]; typename F::Sum sum_mine[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) { s_mine[l].clear(); s_other[l].clear(); sum_mine[l].clear(); }
        // the next table's pair is requested before the current one is consumed (the loop is not unrolled: register budget)
        // short tables: the table loop is cut over blockIdx.z (the sums are linear in the tables, every slice writes its own row of block partials);
        // one thread walking all 2K*tau tables of its pair is a serial chain of dependent loads -- 100 us per late round whatever its size (ncu launch list r02z)
        const int per = (a.n_f + (int)gridDim.z - 1) / (int)gridDim.z, k0 = (int)blockIdx.z * per, k1 = min(a.n_f, k0 + per);
        u64 nx0[TAU], nx1[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) { nx0[l] = nx1[l] = 0; if (k0 < k1) ldg_pair(a.fh + (size_t)k0 * a.fh_stride + (size_t)(slot * TAU + l) * a.fh_pitch + 2 * b, nx0[l], nx1[l]); }
#pragma unroll 1
        for (int kd = k0; kd < k1; ++kd) {
            u64 t[TAU], mine[TAU], other[TAU], q[TAU];
#pragma unroll
            for (int l = 0; l < TAU; ++l) { const u64 sl = F::sub(nx1[l], nx0[l]); t[l] = role ? sl : nx0[l]; }
            if (kd + 1 < k1) {
#pragma unroll
                for (int l = 0; l < TAU; ++l) ldg_pair(a.fh + (size_t)(kd + 1) * a.fh_stride + (size_t)(slot * TAU + l) * a.fh_pitch + 2 * b, nx0[l], nx1[l]);
            }
            if constexpr (PREP_SMEM) SF::mul_prepped(mine, t, s_mu[kd]); else SF::mul_prepped(mine, t, SF::prep(&s_mu[kd * TAU]));
#pragma unroll
            for (int l = 0; l < TAU; ++l) { other[l] = __shfl_xor_sync(lanes, mine[l], 1); sum_mine[l].add(mine[l]); }
            SF::sqr(q, t); const typename SF::Prepped qp = SF::prep(q);
            SF::mac(s_mine, mine, qp); SF::mac(s_other, other, qp);
        }
        auto mulc = [](u64 v, u64 c) { return F::mul(v, c); };
#pragma unroll
        for (int l = 0; l < TAU; ++l) {
            const u64 sm = F::reduce(sum_mine[l]), x0 = F::sub(F::reduce(s_mine[l]), sm), x1 = F::reduce(s_other[l]);
            if (role == 0) {            // x0 = A, x1 = B:  A + 3X B
                h[0][l] = x0; h[1][l] = F::add(x0, mulc(x1, 3)); h[2][l] = F::add(x0, mulc(x1, 6)); h[3][l] = F::add(x0, mulc(x1, 9)); h[4][l] = F::add(x0, mulc(x1, 12));
            } else {                    // x0 = D, x1 = C, sm = Es:  3X^2 C + X^3 D + (X^3 - X) Es
                h[0][l] = 0;
                h[1][l] = F::add(mulc(x1, 3), x0);
                h[2][l] = F::add(F::add(mulc(x1, 12), mulc(x0, 8)), mulc(sm, 6));
                h[3][l] = F::add(F::add(mulc(x1, 27), mulc(x0, 27)), mulc(sm, 24));
                h[4][l] = F::add(F::add(mulc(x1, 48), mulc(x0, 64)), mulc(sm, 60));
            }
        }
    }
    fold_sc_tail<Rg, true>(a, b, active, slot, h, red, (size_t)blockIdx.z * gridDim.x + blockIdx.x, role == 0 && blockIdx.z == 0);
}

}  // namespace lf
