// K9 on the WIDE slot field of the BabyBear ring (Fq9 = Fq[Y]/(Y^9 - nu), 31-bit prime, packed 4-byte limb planes): round
// evaluation and fix_variables of the ring sumcheck (sumcheck/prover.rs:56-162) for PRODUCTS / LIN combination functions, the
// kernels BASELINE configs[2] (2^20-constraint degree-three CCS, arith/ccs.rs:14-43) spends its time in.
//
// Why a second implementation next to k_sc_points / k_fold: one Fq9 product is 81 base-field multiply-accumulates, a table entry
// is 36 B per slot, and the degree-three CCS needs five tables at every evaluation point.  Held per thread that state goes to
// local memory (ncu r02f: long-scoreboard stalls, 25 % of the warps resident, 43 % ALU / 13 % FMA pipe).  Here
//  * a CTA walks 32-pair tiles of the hypercube; the tile of every table (n_mles x 9 limb rows x 256 B) is brought into shared
//    memory by bulk asynchronous copies (cp.async.bulk -> mbarrier, UBLKCP in SASS), double buffered: the next tile streams in
//    while this one is evaluated, and no thread holds table words in registers;
//  * warp e of the CTA evaluates point e for the 32 pairs of the tile (lanes = consecutive pairs: conflict-free 8-byte shared
//    loads; the point is warp-uniform, so v(e) = v0 + e (v1 - v0) is straight-line code);
//  * arithmetic is on balanced representatives with signed 64-bit accumulators (field.cuh: BbBal) -- one IMAD.WIDE per
//    multiply-accumulate, 17 reductions per product, 9 when the right operand is the round's challenge;
//  * coefficients +-1 (every CCS the reference ships: c = [1, -1]) cost a sign, not a product.
// Outputs are the same block partials [tile block][point][D] that reduce_partials / the peer-memory all-reduce consume, and the
// same canonical limbs as the generic kernels (exact arithmetic), which tests/test_gpu_parity.py holds against the oracle.
#pragma once
#include "commit_mma.cuh"

namespace lf {

constexpr int SCW_TILE = 32;            // pairs per tile (one per lane)

template <int = 0> __global__ void __launch_bounds__(256)
k_sc_wide_bb(const ScGenericArgsT<u32> a) {
    typedef BbBal B; constexpr int TAU = 9; constexpr int ROW = 2 * SCW_TILE;      // words per (table, limb) row of a tile
    extern __shared__ __align__(128) u32 scw_smem[];                                // [2 stages][n_mles * 9][ROW]
    __shared__ __align__(8) unsigned long long bars[2];
    const int slot = blockIdx.y, e = threadIdx.x >> 5, lane = threadIdx.x & 31, rows = a.n_mles * TAU;
    const size_t n_tiles = a.n_pairs / SCW_TILE;
    const u32 bar0 = cmma::smem_addr(&bars[0]), stage_bytes = (u32)rows * ROW * 4;
    if (threadIdx.x == 0) { cmma::mbar_init(bar0, 1); cmma::mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](size_t tile, int stage) {          // one thread: n_mles * 9 row copies of 256 B
        const u32 bar = bar0 + 8 * stage, dst = cmma::smem_addr(scw_smem) + stage * stage_bytes;
        cmma::mbar_expect_tx(bar, stage_bytes);
        for (int k = 0; k < a.n_mles; ++k)
            for (int l = 0; l < TAU; ++l) cmma::bulk_g2s(dst + (u32)(k * TAU + l) * ROW * 4, a.mle[k] + (size_t)(slot * TAU + l) * a.pitch + tile * ROW, ROW * 4, bar);
    };
    if (threadIdx.x == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    // term coefficients of this slot: +-1 is a sign
    unsigned unit_pos = 0, unit_neg = 0;      // bit t
    for (int t = 0; t < a.n_terms; ++t) {
        if (a.term_len[t] <= 0) continue;
        const u64* c = a.coef + (size_t)t * BabyBearRing::D + slot * TAU; bool rest0 = true;
        for (int l = 1; l < TAU; ++l) rest0 = rest0 && c[l] == 0;
        if (rest0 && c[0] == 1) unit_pos |= 1u << t; else if (rest0 && c[0] == (u64)B::P - 1) unit_neg |= 1u << t;
    }
    long long ev[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) ev[l] = 0;
    int it = 0;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x, stage ^ 1);      // stage^1 was released by the barrier that ended the previous iteration
        cmma::mbar_wait(bar0 + 8 * stage, (it >> 1) & 1);
        const u32* sm = scw_smem + (size_t)stage * rows * ROW + 2 * lane;
        auto at_point = [&](int k, int* out) {
#pragma unroll
            for (int l = 0; l < TAU; ++l) {
                const uint2 v = *reinterpret_cast<const uint2*>(sm + (k * TAU + l) * ROW);
                if (e == 0) out[l] = B::bal(v.x);
                else if (e == 1) out[l] = B::bal(v.y);
                else out[l] = B::red_small((long long)e * ((int)v.y - (int)v.x) + (long long)v.x);
            }
        };
        int res[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) res[l] = 0;
#pragma unroll 1
        for (int t = 0; t < a.n_terms; ++t) {
            int term[TAU], fac[TAU];
            const int len = a.term_len[t]; int f = 0;
            if ((unit_pos | unit_neg) >> t & 1) {
                at_point(a.term_idx[t][0], term); f = 1;
                if (unit_neg >> t & 1) {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) term[l] = -term[l]; }
            } else {
                const u64* c = a.coef + (size_t)t * BabyBearRing::D + slot * TAU;
#pragma unroll
                for (int l = 0; l < TAU; ++l) term[l] = B::bal((u32)c[l]);
            }
#pragma unroll 1
            for (; f < len; ++f) { at_point(a.term_idx[t][f], fac); B::mul(term, term, fac); }
#pragma unroll
            for (int l = 0; l < TAU; ++l) res[l] = B::fix(res[l] + term[l]);
        }
        if (a.lin) { int last[TAU]; at_point(a.n_mles - 1, last); B::mul(res, res, last); }
#pragma unroll
        for (int l = 0; l < TAU; ++l) ev[l] += res[l];
        __syncthreads();                                  // every warp is done with this stage before it is refilled
    }
    // 32 lanes -> one canonical sum per (point, limb)
#pragma unroll
    for (int l = 0; l < TAU; ++l) {
        u32 v = B::canon(B::red(ev[l]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); if (v >= (u32)B::P) v -= (u32)B::P; }
        if (lane == 0) a.partial[((size_t)blockIdx.x * (a.deg + 1) + e) * BabyBearRing::D + slot * TAU + l] = v;
    }
}

// fix_variables (sumcheck/prover.rs:61-72) on the wide slot field: new[b] = old[2b] + r (old[2b+1] - old[2b]) with the challenge
// and its nu-multiples as kernel parameters (constant-bank operands of the 81 multiply-accumulates); 9 reductions per entry.
struct FoldWideArgs { const u32* in; u32* out; size_t in_pitch, out_pitch, in_stride, out_stride, n_out; BbBal::Fixed r; };
template <int = 0> __global__ void __launch_bounds__(128)
k_fold_wide_bb(const FoldWideArgs a) {
    typedef BbBal B; constexpr int TAU = 9;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int slot = blockIdx.y;
    if (b >= a.n_out) return;
    const u32* in = a.in + (size_t)blockIdx.z * a.in_stride + (size_t)slot * TAU * a.in_pitch + 2 * b;
    u32* out = a.out + (size_t)blockIdx.z * a.out_stride + (size_t)slot * TAU * a.out_pitch + b;
    int f0[TAU], d[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) { const uint2 v = *reinterpret_cast<const uint2*>(in + (size_t)l * a.in_pitch); f0[l] = B::bal(v.x); d[l] = B::fix(B::bal(v.y) - f0[l]); }
    B::mul_fixed_add(d, d, a.r, f0);
#pragma unroll
    for (int l = 0; l < TAU; ++l) out[(size_t)l * a.out_pitch] = B::canon(d[l]);
}

}  // namespace lf
