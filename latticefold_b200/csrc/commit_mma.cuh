// Ajtai commitments of decomposed witnesses on the 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM).
//
// commit_witnesses (crates/latticefold/src/nifs/decomposition.rs:178-201) commits the K-1 pieces of decompose_to_vec(b, K)
// (decomposition.rs:162-167): y_p = A * CRT(d_p) with d_p a vector of ring elements whose coefficients are balanced base-b
// digits.  CRT is linear and the digits are tiny, so the contraction over the witness axis is regrouped (exact arithmetic:
// any regrouping yields the identical canonical field elements, commitment_scheme.rs:45-51):
//
//   y_p[i][slot] = sum_j A_ij[slot] * CRT(d_pj)[slot]
//                = sum_m Y^m  sum_c  X^c|_slot  *  ( sum_j a_m[i][slot][j] * d_p[j][c] )          a_m = limb m of the slot-field element
//
// and the innermost sum is an integer GEMM: the 64-bit field limb a_m is cut into eight unsigned bytes a_{m,u}, the digit is one
// signed byte, and T[(slot,m,i,u)][(p,c)] = sum_j a_{m,u}[j] * d_p[j][c] accumulates exactly in s32 (|T| <= 255 * n < 2^31 for
// n < 2^23).  One MMA tile is M = 128 rows = 16 row groups (slot, m, i) x 8 byte limbs, N = D * pieces columns (360 for the
// Goldilocks ring with 15 pieces: two instructions N = 192 + 168), K = 32 witness positions per instruction.
//
// Operand images in HBM are stored pre-tiled in exactly the shared-memory layout the MMA reads (K-major, no swizzle: 8 x 16-byte
// core matrices, LBO = 128 B between K-adjacent core matrices, SBO = 512 B between 8-row groups), so one stage of the pipeline is
// two bulk asynchronous copies (cp.async.bulk, UBLKCP in SASS) completing on an mbarrier:
//   A8  [tile][chunk][16 groups][4 x (8 limbs x 16 j)]     8 KB per stage, written once when the matrix is uploaded
//   D8  [chunk][D*pieces/8 groups][4 x (8 coeffs x 16 j)]  23 KB per stage for 15 pieces, written by k_d8_tile after the digit split
// The matrix is read from HBM exactly once per batch (1.31 GB at kappa = 26, n = 2^18); the digit tiles (94 MB) are shared by all
// row tiles and stay in L2.
//
// Warp roles (192 threads, one CTA per SM, 512 TMEM columns): warps 0-3 epilogue (one TMEM lane quarter each), warp 4 copy producer,
// warp 5 MMA issuer + TMEM allocation.  Epilogue per lane (slot, m, i, u) and piece p: the D coefficients' sums are contracted with
// the CRT table column of the slot (the one non-zero per coefficient), scaled by 2^(8u), summed over the eight limb lanes by
// shuffles and written as block partials [split * TAU + m][i][p][slot * TAU + (m + l) mod TAU] (times nu on wrap-around), the layout
// the existing partial reduction / peer-memory all-reduce consumes.
#pragma once
#include "kernels.cuh"

namespace lf {
namespace cmma {

constexpr int J = 64;                  // witness positions per pipeline stage
constexpr int STAGES = 4;
constexpr int GROUPS = 16;             // 8-row groups per M tile (M = 128)
constexpr int MAX_PIECES = 16;
constexpr int A_STAGE_BYTES = GROUPS * 8 * J;                    // 8192
constexpr int TMEM_COLS = 512;
constexpr int THREADS = 192;

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    u32 done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (version 1 = Blackwell): start >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32
__device__ __forceinline__ u64 smem_desc(u32 addr, u32 lbo, u32 sbo) {
    return (u64)((addr & 0x3FFFFu) >> 4) | ((u64)(lbo >> 4) << 16) | ((u64)(sbo >> 4) << 32) | ((u64)1 << 46);
}
// instruction descriptor of kind::i8: D = s32, A = unsigned bytes, B = signed bytes, both K-major, M = 128
__host__ __device__ constexpr u32 idesc_i8(int n) { return (2u << 4) | (0u << 7) | (1u << 10) | ((u32)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_i8(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(u32 bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld8(u32 addr, u32* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}

// per-ring constants of the epilogue: for slot s and coefficient c the single non-zero CRT entry of column c among the slot's rows
// (value val[s][c], in row s*TAU + perm[s][c % TAU]); corr[s][r] = 2^31 * sum_q val[s][TAU q + r]
template <class Rg> struct EpiTables { u64 val[Rg::S][Rg::D]; u64 corr[Rg::S][Rg::TAU]; int perm[Rg::S][Rg::TAU]; };

struct Args {
    const uint8_t* A8; const int8_t* D8;
    int nchunks, chunks_per_split, ncols, n_mma1, n_mma2;      // ncols = real pieces; MMA widths (multiples of 16, n_mma2 may be 0)
    int kappa, g_total;                                        // g = (slot * TAU + m) * kappa + i
    u32 d_stage_bytes;                                         // D * ncols * J
    const void* tables;                                        // EpiTables<Rg> on the device
    u64* partial;                                              // [split * TAU + m][kappa][ncols][D]
};

template <class Rg> __global__ void __launch_bounds__(THREADS, 1) k_commit_mma(const Args a) {
    typedef typename Rg::F F; constexpr int D = Rg::D, TAU = Rg::TAU, S = Rg::S, Q = D / TAU;
    constexpr int D_STAGE_MAX = D * MAX_PIECES * J, STAGE_BYTES = A_STAGE_BYTES + D_STAGE_MAX;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_full[STAGES], s_empty[STAGES], s_done;
    __shared__ u32 s_tmem;
    __shared__ EpiTables<Rg> s_tab;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int chunk0 = split * a.chunks_per_split, nch = min(a.nchunks, chunk0 + a.chunks_per_split) - chunk0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_addr(&s_full[s]), 1); mbar_init(smem_addr(&s_empty[s]), 1); }
        mbar_init(smem_addr(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"((u32)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    { const u64* src = reinterpret_cast<const u64*>(a.tables); u64* dst = reinterpret_cast<u64*>(&s_tab);
      for (int i = threadIdx.x; i < (int)(sizeof(EpiTables<Rg>) / 8); i += blockDim.x) dst[i] = src[i]; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = s_tmem;

    if (warp == 4) {
        // ---------------------------------------------------------------- producer: two bulk copies per stage
        if (lane == 0) {
            const uint8_t* srcA = a.A8 + ((size_t)tile * a.nchunks + chunk0) * A_STAGE_BYTES;
            const int8_t* srcD = a.D8 + (size_t)chunk0 * a.d_stage_bytes;
            for (int it = 0; it < nch; ++it) {
                const int s = it % STAGES; const u32 ph = (u32)(it / STAGES) & 1u;
                mbar_wait(smem_addr(&s_empty[s]), ph ^ 1u);
                const u32 full = smem_addr(&s_full[s]), dst = smem_addr(smem + (size_t)s * STAGE_BYTES);
                mbar_expect_tx(full, A_STAGE_BYTES + a.d_stage_bytes);
                bulk_g2s(dst, srcA + (size_t)it * A_STAGE_BYTES, A_STAGE_BYTES, full);
                bulk_g2s(dst + A_STAGE_BYTES, srcD + (size_t)it * a.d_stage_bytes, a.d_stage_bytes, full);
            }
        }
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer (one thread)
        if (lane == 0) {
            const u32 id1 = idesc_i8(a.n_mma1), id2 = idesc_i8(a.n_mma2 ? a.n_mma2 : 16);
            for (int it = 0; it < nch; ++it) {
                const int s = it % STAGES; const u32 ph = (u32)(it / STAGES) & 1u;
                mbar_wait(smem_addr(&s_full[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 sa = smem_addr(smem + (size_t)s * STAGE_BYTES), sd = sa + A_STAGE_BYTES;
#pragma unroll
                for (int kb = 0; kb < J / 32; ++kb) {
                    const u32 acc = (it > 0 || kb > 0) ? 1u : 0u;
                    const u64 ad = smem_desc(sa + kb * 256, 128, (J / 16) * 128);
                    mma_i8(tmem, ad, smem_desc(sd + kb * 256, 128, (J / 16) * 128), id1, acc);
                    if (a.n_mma2) mma_i8(tmem + a.n_mma1, ad, smem_desc(sd + (a.n_mma1 / 8) * ((J / 16) * 128) + kb * 256, 128, (J / 16) * 128), id2, acc);
                }
                mma_commit(smem_addr(&s_empty[s]));          // arrives when the MMAs that read this stage have completed
            }
            mma_commit(smem_addr(&s_done));
        }
    } else {
        // ---------------------------------------------------------------- epilogue: TMEM lane = threadIdx.x = (row group, byte limb)
        const int u = threadIdx.x & 7, g = tile * GROUPS + (threadIdx.x >> 3);
        const bool valid = g < a.g_total;
        const int gg = valid ? g : 0, i = gg % a.kappa, sm = gg / a.kappa, m = sm % TAU, slot = sm / TAU;
        mbar_wait(smem_addr(&s_done), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 lane_addr = tmem + ((u32)(warp * 32) << 16);
        for (int p = 0; p < a.ncols; ++p) {
            u32 t[D];
#pragma unroll
            for (int c8 = 0; c8 < D / 8; ++c8) tmem_ld8(lane_addr + (u32)(p * D + c8 * 8), &t[c8 * 8]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            u64* out = a.partial + ((((size_t)split * TAU + m) * a.kappa + i) * a.ncols + p) * D + slot * TAU;
#pragma unroll
            for (int r = 0; r < TAU; ++r) {
                typename F::Acc acc; acc.clear();
#pragma unroll
                for (int q = 0; q < Q; ++q) acc.mac_small(t[TAU * q + r] ^ 0x80000000u, s_tab.val[slot][TAU * q + r]);      // T + 2^31 >= 0
                u64 w = F::sub(F::reduce(acc), s_tab.corr[slot][r]);
                w = F::mul(w, (u64)1 << (8 * u));
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) w = F::add(w, __shfl_xor_sync(0xffffffffu, w, o));
                if (u == 0 && valid) {
                    int k = m + s_tab.perm[slot][r];
                    if (k >= TAU) { k -= TAU; w = F::mul_nu(w); }
                    out[k] = w;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)TMEM_COLS) : "memory");
    }
}

// Ajtai matrix limb planes [row i][D planes][pitch] -> A8 (see the header).  One thread per (row group g, 16 witness positions).
template <class Rg> __global__ void __launch_bounds__(256)
k_a8_tile(const u64* __restrict__ A, size_t row_stride, size_t pitch, size_t n, int kappa, int g_total, int nchunks, size_t g_pad, uint8_t* __restrict__ out) {
    constexpr int TAU = Rg::TAU;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t jblocks = (size_t)nchunks * (J / 16);
    if (t >= g_pad * jblocks) return;
    const size_t g = t / jblocks, jbg = t % jblocks, chunk = jbg / (J / 16); const int jb = (int)(jbg % (J / 16));
    const size_t j0 = jbg * 16;
    u64 v[16];
    if (g < (size_t)g_total) {
        const int i = (int)(g % kappa), sm = (int)(g / kappa);        // plane sm = slot * TAU + m
        const u64* src = A + (size_t)i * row_stride + (size_t)sm * pitch + j0;
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = j0 + e < n ? src[e] : 0;
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0;
    }
    uint8_t* dst = out + (((g / GROUPS) * nchunks + chunk) * GROUPS + g % GROUPS) * (size_t)(8 * J) + (size_t)jb * 128;
#pragma unroll
    for (int ub = 0; ub < 8; ++ub) {
        u32 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (u32)((v[4 * q] >> (8 * ub)) & 0xFF) | ((u32)((v[4 * q + 1] >> (8 * ub)) & 0xFF) << 8) | ((u32)((v[4 * q + 2] >> (8 * ub)) & 0xFF) << 16) | ((u32)((v[4 * q + 3] >> (8 * ub)) & 0xFF) << 24);
        *reinterpret_cast<uint4*>(dst + ub * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    (void)TAU;
}

// digit planes [piece][D coefficient planes][pitch] -> D8 (see the header).  One thread per (chunk, piece, coefficient, 16 positions);
// the planes are zero beyond n (memset at the start of the step) and pitch >= nchunks * J.
template <class Rg> __global__ void __launch_bounds__(256)
k_d8_tile(const int8_t* __restrict__ dig, size_t dig_pitch, size_t dig_stride, int ncols, int nchunks, int8_t* __restrict__ out) {
    constexpr int D = Rg::D;
    // thread = one 16-byte row of a core matrix, in the order of the output image (coalesced stores)
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_chunk = (size_t)ncols * D * (J / 16);
    if (t >= per_chunk * nchunks) return;
    const size_t chunk = t / per_chunk; const size_t r = t % per_chunk;
    const int r8 = (int)(r % 8), jb = (int)((r / 8) % (J / 16)), row = (int)(r / (8 * (J / 16))) * 8 + r8;
    const int c = row % D, p = row / D;
    const uint4 v = *reinterpret_cast<const uint4*>(dig + (size_t)p * dig_stride + (size_t)c * dig_pitch + chunk * J + (size_t)jb * 16);
    *reinterpret_cast<uint4*>(out + chunk * ((size_t)ncols * D * J) + r * 16) = v;
}

}  // namespace cmma
}  // namespace lf
