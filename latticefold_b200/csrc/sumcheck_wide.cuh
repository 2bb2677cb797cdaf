// K9 on the WIDE slot field of the BabyBear ring (Fq9 = Fq[Y]/(Y^9 - nu), 31-bit prime, packed 4-byte limb planes): round
// evaluation and fix_variables of the ring sumcheck (sumcheck/prover.rs:56-162) for PRODUCTS / LIN combination functions, the
// kernels BASELINE configs[2] (2^20-constraint degree-three CCS, arith/ccs.rs:14-43) spends its time in.
//
// Why a second implementation next to k_sc_points / k_fold: one Fq9 product is 81 base-field multiply-accumulates, a table entry
// is 36 B per slot, and the degree-three CCS needs five tables at every evaluation point.  Held per thread that state goes to
// local memory (ncu r02f: long-scoreboard stalls, 25 % of the warps resident, 43 % ALU / 13 % FMA pipe).  Here
//  * a persistent CTA walks 32-pair tiles of the hypercube; the tile of every table (n_mles x 9 limb rows x 256 B) is brought into
//    shared memory by bulk asynchronous copies (cp.async.bulk -> mbarrier, UBLKCP in SASS), three stages deep: tiles stream in while
//    earlier ones are evaluated and no thread holds table words in registers.  (A dedicated producer warp with full / empty
//    barriers was measured and lost 30 %: the idle warp costs a fifth of the resident compute warps and its polling steals issue slots.)
//  * warp e of the CTA evaluates point e for the 32 pairs of the tile (lanes = consecutive pairs: conflict-free shared loads).
//    The table values at the points e >= 2, v(e) = v(e-1) + (v1 - v0), are formed ONCE per tile by all warps together (row r of
//    the tile by warp r mod warps, one add and one conditional correction per point) and parked in shared memory: evaluated
//    inside the point's own warp they cost a wide multiply and a reduction each and leave the warps of points 0 and 1 idle at
//    the tile barrier (ncu r02r: 20 % barrier stalls);
//  * arithmetic is on balanced representatives with signed 64-bit accumulators (field.cuh: BbBal) -- one IMAD.WIDE per
//    multiply-accumulate, 17 reductions per product, 9 when the right operand is the round's challenge;
//  * coefficients +-1 (every CCS the reference ships: c = [1, -1]) cost a sign, not a product.
// Outputs are the same block partials [tile block][point][D] that reduce_partials / the peer-memory all-reduce consume, and the
// same canonical limbs as the generic kernels (exact arithmetic), which tests/test_gpu_parity.py holds against the oracle.
#pragma once
#include "commit_mma.cuh"

namespace lf {

constexpr int SCW_TILE = 32;            // pairs per tile (one per lane)
constexpr int SCW_STAGES = 2;           // tiles in flight per CTA

// block = (deg + 1) warps, warp e evaluates point e; grid = (CTAs resident on the chip / S, S), every CTA walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...  The lanes of warp 0 issue the row copies of the tile two iterations ahead.
template <int MAXT, int MINB> __global__ void __launch_bounds__(MAXT, MINB)
k_sc_wide_bb(const ScGenericArgsT<u32> a) {
    typedef BbBal B; constexpr int TAU = 9; constexpr int ROW = 2 * SCW_TILE;      // words per (table, limb) row of a tile
    extern __shared__ __align__(128) u32 scw_smem[];                                // [stages][n_mles * 9][ROW] tiles, then [n_mles * 9][deg - 1][32] point values
    __shared__ __align__(8) unsigned long long bars[SCW_STAGES];
    const int slot = blockIdx.y, e = threadIdx.x >> 5, lane = threadIdx.x & 31, npts = a.deg + 1, rows = a.n_mles * TAU;
    const size_t n_tiles = a.n_pairs / SCW_TILE;
    const u32 bar0 = cmma::smem_addr(&bars[0]), stage_bytes = (u32)rows * ROW * 4;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SCW_STAGES; ++s) cmma::mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t tile, int stage) {          // warp 0: lane r copies rows r, r + 32, ... (256 B each)
        const u32 bar = bar0 + 8 * stage, dst = cmma::smem_addr(scw_smem) + stage * stage_bytes;
        if (lane == 0) cmma::mbar_expect_tx(bar, stage_bytes);
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
            const int k = r / TAU, l = r - k * TAU;
            cmma::bulk_g2s(dst + (u32)r * ROW * 4, a.mle[k] + (size_t)(slot * TAU + l) * a.pitch + tile * ROW, ROW * 4, bar);
        }
    };
    if (e == 0)
        for (int s = 0; s < SCW_STAGES - 1; ++s) if (blockIdx.x + (size_t)s * gridDim.x < n_tiles) issue(blockIdx.x + (size_t)s * gridDim.x, s);
    // term coefficients of this slot: +-1 is a sign, not a product
    unsigned unit_pos = 0, unit_neg = 0;      // bit t
    for (int t = 0; t < a.n_terms; ++t) {
        if (a.term_len[t] <= 0) continue;
        const u64* c = a.coef + (size_t)t * BabyBearRing::D + slot * TAU; bool rest0 = true;
        for (int l = 1; l < TAU; ++l) rest0 = rest0 && c[l] == 0;
        if (rest0 && c[0] == 1) unit_pos |= 1u << t; else if (rest0 && c[0] == (u64)B::P - 1) unit_neg |= 1u << t;
    }
    const int n_pass = a.n_terms + (a.lin ? 1 : 0);      // the LIN factor is one more pass: (sum of the terms) * last table
    long long ev[TAU];
#pragma unroll
    for (int l = 0; l < TAU; ++l) ev[l] = 0;
    int it = 0;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it % SCW_STAGES;
        // the stage refilled here was released by the barrier that ended the previous iteration
        if (e == 0 && tile + (size_t)(SCW_STAGES - 1) * gridDim.x < n_tiles) issue(tile + (size_t)(SCW_STAGES - 1) * gridDim.x, (it + SCW_STAGES - 1) % SCW_STAGES);
        cmma::mbar_wait(bar0 + 8 * stage, (it / SCW_STAGES) & 1);
        const u32* sm = scw_smem + (size_t)stage * rows * ROW + 2 * lane;
        // table values at the points 2 .. deg (balanced), row r by warp r mod npts
        int* pts = reinterpret_cast<int*>(scw_smem + (size_t)SCW_STAGES * rows * ROW);
        for (int r = e; r < rows; r += npts) {
            const uint2 v = *reinterpret_cast<const uint2*>(sm + r * ROW);
            const int a0 = B::bal(v.x); int x = B::bal(v.y); const int d = B::fix(x - a0);
            for (int q = 0; q < npts - 2; ++q) { x = B::fix(x + d); pts[(r * (npts - 2) + q) * SCW_TILE + lane] = x; }
        }
        __syncthreads();
        int res[TAU], term[TAU];
#pragma unroll
        for (int l = 0; l < TAU; ++l) res[l] = 0;
#pragma unroll 1
        for (int t = 0; t < n_pass; ++t) {
            const bool last = t == a.n_terms;
            const int len = last ? 1 : a.term_len[t];
            const bool unit = !last && ((unit_pos | unit_neg) >> t & 1), neg = !last && (unit_neg >> t & 1);
            if (last) {
#pragma unroll
                for (int l = 0; l < TAU; ++l) term[l] = res[l];
            } else if (!unit) {
                const u64* c = a.coef + (size_t)t * BabyBearRing::D + slot * TAU;
#pragma unroll
                for (int l = 0; l < TAU; ++l) term[l] = B::bal((u32)c[l]);
            }
#pragma unroll 1
            for (int f = 0; f < len; ++f) {
                const int k = last ? a.n_mles - 1 : a.term_idx[t][f];
                int fac[TAU];
                if (e >= 2) {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) fac[l] = pts[((k * TAU + l) * (npts - 2) + (e - 2)) * SCW_TILE + lane];
                } else {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) { const uint2 v = *reinterpret_cast<const uint2*>(sm + (k * TAU + l) * ROW); fac[l] = B::bal(e == 0 ? v.x : v.y); }
                }
                if (unit && f == 0) {
#pragma unroll
                    for (int l = 0; l < TAU; ++l) term[l] = neg ? -fac[l] : fac[l];
                } else B::mul(term, term, fac);
            }
            if (!last) {
#pragma unroll
                for (int l = 0; l < TAU; ++l) res[l] = B::fix(res[l] + term[l]);
            }
        }
#pragma unroll
        for (int l = 0; l < TAU; ++l) ev[l] += a.lin ? term[l] : res[l];
        __syncthreads();                                  // every warp is done with this stage before it is refilled
    }
    // 32 lanes -> one canonical sum per (point, limb)
#pragma unroll
    for (int l = 0; l < TAU; ++l) {
        u32 v = B::canon(B::red(ev[l]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); if (v >= (u32)B::P) v -= (u32)B::P; }
        if (lane == 0) a.partial[((size_t)blockIdx.x * npts + e) * BabyBearRing::D + slot * TAU + l] = v;
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
