// LatticeFold+ consumers of the commitment / sumcheck path on the coefficient-form ring
// R = Z_q[X]/(X^16 + 1) (`stark_rings::cyclotomic_ring::models::frog_ring::RqPoly`, BaseRing = Fq):
//   monomial set check      crates/latticefold-plus/src/setchk.rs:59-262   (In::set_check)
//   double commitment       crates/latticefold-plus/src/rgchk.rs:259-336   (RgInstance::from_f: decompose, exp, A * M_f, split, the three commitments)
//   range check             crates/latticefold-plus/src/rgchk.rs:75-187    (Rg::range_check)
//   transcript              crates/latticefold-plus/src/transcript.rs:16-56
// The reference runs all of this on dense ring-valued MLEs (16 coefficients per entry) with ring products.  What the data IS:
//   * every table of the set check's sumcheck holds constants of R (ev(entry, beta), its square, eq(c, .)), so the sumcheck runs on
//     ONE base-field word per entry -- 1/16 of the reference's bytes and one field product where the reference does a ring product;
//   * the matrices M_f and the vector m_tau hold monomials exp(a) = X^(a mod 16): ONE byte per entry (the exponent) instead of 128;
//     a ring product with such an entry is a negacyclic rotation, so the double commitment A * M_f and the evaluations
//     MLE(M * column)(r) need no multiplications at all;
//   * MLE(M_i * v)(r) = sum_x (M_i^T eq(r, .))[x] * v[x]: one pass over M_i per point instead of one sparse mat-vec per column.
// Field arithmetic is exact, so these regroupings give the reference's values bit for bit (tests/test_gpu_plus.py against oracle/lfplus.hpp).
// Representation: device-resident sumcheck tables, weights and eq tables are in Montgomery form (Fm, R = 2^64); ring data that
// crosses the boundary (A, f, M_i, results) is canonical.  mont x canonical products are canonical, mont x mont products Montgomery.
#pragma once
#include "engine.cuh"
#include "transcript_host.hpp"

namespace lf { namespace plus {

constexpr int PD = 16;                 // ring dimension
constexpr unsigned char CODE_ZERO = 0xFF;      // a zero entry of a monomial set (codes 0..15 are X^code)

// 64-bit Montgomery arithmetic for the Frog prime (p > 2^63: the REDC sum can carry out of 64 bits)
constexpr u64 inv_mod_2_64(u64 p) { u64 x = 1; for (int i = 0; i < 6; ++i) x *= 2 - p * x; return x; }      // p^-1 mod 2^64 (Newton, p odd)
struct Fm {
    static constexpr u64 P = FrogField::P;
    static constexpr u64 NINV = ~inv_mod_2_64(FrogField::P) + 1;
    static LF_HD u64 add(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= P) s -= P; return s; }
    static LF_HD u64 sub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
    static LF_HD u64 neg(u64 a) { return a ? P - a : 0; }
    static LF_HD u64 mul(u64 a, u64 b) {      // a b 2^-64 mod p
        u64 lo, hi, mlo, mhi; mul_wide(a, b, lo, hi);
        const u64 m = lo * NINV; mul_wide(m, P, mlo, mhi);      // lo + mlo = 0 mod 2^64: the carry is (lo != 0)
        u64 t = hi + mhi; const bool c1 = t < hi; const u64 t2 = t + (lo != 0 ? 1 : 0); const bool c2 = t2 < t;
        return (c1 || c2 || t2 >= P) ? t2 - P : t2;
    }
    // host-side conversions
    static u64 r1() { return (u64)((((u128)1) << 64) % P); }
    static u64 r2() { const u128 r = r1(); return (u64)(r * r % P); }
    static u64 to_mont(u64 a) { return (u64)((u128)(a % P) * r1() % P); }
    static u64 from_mont(u64 a) { return mul(a, 1); }
    static u64 hmul(u64 a, u64 b) { return (u64)((u128)a * b % P); }      // canonical product on the host
    static u64 hpow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = hmul(r, a); a = hmul(a, a); e >>= 1; } return r; }
};

// ---------------------------------------------------------------- kernels
struct PowArgs { u64 v[PD]; };
// tables of one monomial set held as exponent codes: T[2j][x] = beta^code, T[2j+1][x] = its square   (setchk.rs:100-113)
__global__ void k_plus_tables_mono(const unsigned char* __restrict__ codes, size_t code_pitch, size_t nrows, u64* __restrict__ T, size_t stride, PowArgs bpow /* beta^e R */) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const int j = blockIdx.y;
    if (x >= nrows) return;
    const unsigned char c = codes[(size_t)j * code_pitch + x];
    u64 m = 0;
#pragma unroll
    for (int e = 0; e < PD; ++e) if (c == e) m = bpow.v[e];
    T[(size_t)(2 * j) * stride + x] = m; T[(size_t)(2 * j + 1) * stride + x] = Fm::mul(m, m);
}
// the same for a set given as general sparse entries (possibly not monomials): m = sum_i coef_i beta^i
__global__ void k_plus_tables_general(const u32* __restrict__ ecol, const u32* __restrict__ erow, const u64* __restrict__ val, size_t nnz, u64* __restrict__ T, size_t stride, PowArgs bpow2 /* beta^i R^2 */) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    u64 m = 0;
#pragma unroll
    for (int i = 0; i < PD; ++i) m = Fm::add(m, Fm::mul(val[e * PD + i], bpow2.v[i]));
    const size_t j = ecol[e], x = erow[e];
    T[(2 * j) * stride + x] = m; T[(2 * j + 1) * stride + x] = Fm::mul(m, m);
}
struct EqArgs { u64 c[40], omc[40]; int nv; };      // Montgomery c_i and 1 - c_i
__global__ void k_plus_eq(u64* __restrict__ out, size_t n, EqArgs a, u64 one /* R */) {      // eq(x, c), c[0] on bit 0 (sumcheck/utils.rs:100-170)
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 v = one;
    for (int i = 0; i < a.nv; ++i) v = Fm::mul(v, ((x >> i) & 1) ? a.c[i] : a.omc[i]);
    out[x] = v;
}
struct Group { int base, ncols, wofs, pad; };      // tables base .. base + 2 ncols (m_j, m'_j pairs, then eq); weights w[wofs + j] = alpha^j rc^i
// one round of the set check's sumcheck (comb of setchk.rs:157-189, degree 3): h(X) = sum_g eq_g(X) sum_j w_gj (m_gj(X)^2 - m'_gj(X))
__global__ void __launch_bounds__(256) k_plus_round(const u64* __restrict__ T, size_t stride, size_t n_pairs, const Group* __restrict__ groups, int n_groups,
                                                     const u64* __restrict__ w, u64* __restrict__ partial) {
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_pairs; b += (size_t)gridDim.x * blockDim.x) {
        for (int g = 0; g < n_groups; ++g) {
            const Group G = groups[g];
            u64 s[4] = {0, 0, 0, 0};
            for (int j = 0; j < G.ncols; ++j) {
                const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(T + (size_t)(G.base + 2 * j) * stride + 2 * b);
                const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(T + (size_t)(G.base + 2 * j + 1) * stride + 2 * b);
                const u64 wj = w[G.wofs + j], dm = Fm::sub(m.y, m.x), dq = Fm::sub(q.y, q.x);
                u64 mx = m.x, qx = q.x;
#pragma unroll
                for (int X = 0; X < 4; ++X) {
                    s[X] = Fm::add(s[X], Fm::mul(wj, Fm::sub(Fm::mul(mx, mx), qx)));
                    mx = Fm::add(mx, dm); qx = Fm::add(qx, dq);
                }
            }
            const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(T + (size_t)(G.base + 2 * G.ncols) * stride + 2 * b);
            const u64 de = Fm::sub(e.y, e.x); u64 ex = e.x;
#pragma unroll
            for (int X = 0; X < 4; ++X) { acc[X] = Fm::add(acc[X], Fm::mul(ex, s[X])); ex = Fm::add(ex, de); }
        }
    }
    __shared__ u64 sh[8][4];
#pragma unroll
    for (int X = 0; X < 4; ++X) {
        u64 v = acc[X];
        for (int o = 16; o; o >>= 1) v = Fm::add(v, __shfl_down_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][X] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) { u64 v = 0; for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = Fm::add(v, sh[wv][threadIdx.x]); partial[(size_t)blockIdx.x * 4 + threadIdx.x] = v; }
}
// the same round for SHORT tables: one warp per pair, the lanes over the flattened (group, column) list, so that the late rounds of the sumcheck
// (a handful of pairs, ~70 columns) are a few parallel steps instead of one thread walking every column (ncu: 70 us per tiny round before).
// cols[c] = {table of m_c, table of eq of its group}; w[c] the weight.  Exact arithmetic: the regrouped sum is the same element.
struct ColDesc { int mt, et; };
__global__ void __launch_bounds__(256) k_plus_round_cols(const u64* __restrict__ T, size_t stride, size_t n_pairs, const ColDesc* __restrict__ cols, int n_cols,
                                                          const u64* __restrict__ w, u64* __restrict__ partial) {
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); const int lane = threadIdx.x & 31;
    u64 acc[4] = {0, 0, 0, 0};
    if (b < n_pairs)
        for (int c = lane; c < n_cols; c += 32) {
            const ColDesc cd = cols[c];
            const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(T + (size_t)cd.mt * stride + 2 * b);
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(T + (size_t)(cd.mt + 1) * stride + 2 * b);
            const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(T + (size_t)cd.et * stride + 2 * b);
            const u64 wj = w[c], dm = Fm::sub(m.y, m.x), dq = Fm::sub(q.y, q.x), de = Fm::sub(e.y, e.x);
            u64 mx = m.x, qx = q.x, ex = Fm::mul(e.x, wj); const u64 dex = Fm::mul(de, wj);      // (w eq)(X) is linear in X as well
#pragma unroll
            for (int X = 0; X < 4; ++X) {
                acc[X] = Fm::add(acc[X], Fm::mul(ex, Fm::sub(Fm::mul(mx, mx), qx)));
                mx = Fm::add(mx, dm); qx = Fm::add(qx, dq); ex = Fm::add(ex, dex);
            }
        }
    __shared__ u64 sh[8][4];
#pragma unroll
    for (int X = 0; X < 4; ++X) {
        u64 v = acc[X];
        for (int o = 16; o; o >>= 1) v = Fm::add(v, __shfl_down_sync(0xffffffffu, v, o));
        if (lane == 0) sh[threadIdx.x >> 5][X] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) { u64 v = 0; for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = Fm::add(v, sh[wv][threadIdx.x]); partial[(size_t)blockIdx.x * 4 + threadIdx.x] = v; }
}
// fix_variables for every table at once: out[t][b] = in[t][2b] + r (in[t][2b+1] - in[t][2b])
__global__ void k_plus_fold(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ out, size_t n_out, u64 r) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t t = blockIdx.y;
    if (b >= n_out) return;
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(in + t * in_stride + 2 * b);
    out[t * n_out + b] = Fm::add(v.x, Fm::mul(r, Fm::sub(v.y, v.x)));
}

// ---- weighted sums  out[col] = sum_x W[x] (*) entry(x, col)  with block partials  partial[chunk][col][16]
// where a weighted-sum launch puts its block partials: several launches share one [chunk][stride_cols][16] buffer (one reduction, one read-back for all)
struct POut { u64* p; unsigned stride_cols, col_off; };
template <class Body> __device__ __forceinline__ void wsum_finish(u64 (&acc)[PD], const POut po) {
    __shared__ u64 sh[8][PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) {
        u64 v = acc[c];
        for (int o = 16; o; o >>= 1) v = Fm::add(v, __shfl_down_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < PD) { u64 v = 0; for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = Fm::add(v, sh[wv][threadIdx.x]); po.p[((size_t)blockIdx.x * po.stride_cols + po.col_off + blockIdx.y) * PD + threadIdx.x] = v; }
}
// scalar weights (eq(r, .), Montgomery), monomial entries: coefficient `code` of the result collects the weights      setchk.rs:199-214, 251-257
__global__ void __launch_bounds__(256) k_plus_wsum_scalar_mono(const u64* __restrict__ W, const unsigned char* __restrict__ codes, size_t code_pitch, size_t nrows, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    const unsigned char* col = codes + (size_t)blockIdx.y * code_pitch;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nrows; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned char cd = col[x]; const u64 wv = W[x];
#pragma unroll
        for (int c = 0; c < PD; ++c) acc[c] = Fm::add(acc[c], cd == c ? wv : 0);
    }
    wsum_finish<void>(acc, po);
}
// scalar weights, general entries grouped by column (col_ptr / row / val); row == nullptr: a dense vector, entry e sits in row e
__global__ void __launch_bounds__(256) k_plus_wsum_scalar_general(const u64* __restrict__ W, const u64* __restrict__ col_ptr, const u32* __restrict__ erow, const u64* __restrict__ val, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    const size_t e0 = col_ptr[blockIdx.y], e1 = col_ptr[blockIdx.y + 1];
    for (size_t e = e0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += (size_t)gridDim.x * blockDim.x) {
        const u64 wv = W[erow ? erow[e] : e - e0];
#pragma unroll
        for (int c = 0; c < PD; ++c) acc[c] = Fm::add(acc[c], Fm::mul(wv, val[e * PD + c]));
    }
    wsum_finish<void>(acc, po);
}
// ring-valued weights W[x] (16 words each), monomial entries: W[x] X^code is a negacyclic rotation                  rgchk.rs:297-304 (A * M_f), setchk.rs:217-240
__global__ void __launch_bounds__(256) k_plus_wsum_ring_mono(const u64* __restrict__ W, const unsigned char* __restrict__ codes, size_t code_pitch, size_t nrows, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    const unsigned char* col = codes + (size_t)blockIdx.y * code_pitch;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nrows; x += (size_t)gridDim.x * blockDim.x) {
        const unsigned cd = col[x];
        if (cd >= PD) continue;
        const u64* wx = W + x * PD;
#pragma unroll
        for (int o = 0; o < PD; ++o) {      // coefficient o of W X^cd: +W[o - cd] for o >= cd, -W[o - cd + 16] below (X^16 = -1)
            const u64 v = wx[(o - cd) & (PD - 1)];
            acc[o] = (unsigned)o >= cd ? Fm::add(acc[o], v) : Fm::sub(acc[o], v);
        }
    }
    wsum_finish<void>(acc, po);
}
// ring-valued weights, general entries: negacyclic products (both operands canonical: the sum comes out times 2^-64)   rgchk.rs:322 (A f), 167-172 (M f)
__global__ void __launch_bounds__(128) k_plus_wsum_ring_general(const u64* __restrict__ W, const u64* __restrict__ col_ptr, const u32* __restrict__ erow, const u64* __restrict__ val, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    const size_t e0 = col_ptr[blockIdx.y], e1 = col_ptr[blockIdx.y + 1];
    for (size_t e = e0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += (size_t)gridDim.x * blockDim.x) {
        const u64* wx = W + (size_t)(erow ? erow[e] : e - e0) * PD;
        u64 v[PD];
#pragma unroll
        for (int j = 0; j < PD; ++j) v[j] = val[e * PD + j];
#pragma unroll 1
        for (int i = 0; i < PD; ++i) {      // acc += W_i (X^i v); X^i v is kept in registers by rotating v one step per iteration (X^16 = -1)
            const u64 wi = wx[i];
            if (wi) {
#pragma unroll
                for (int j = 0; j < PD; ++j) acc[j] = Fm::add(acc[j], Fm::mul(wi, v[j]));
            }
            const u64 top = v[PD - 1];
#pragma unroll
            for (int j = PD - 1; j > 0; --j) v[j] = v[j - 1];
            v[0] = Fm::neg(top);
        }
    }
    wsum_finish<void>(acc, po);
}
// ring-valued weights, small signed scalar entries (tau, |tau| < 8) given as Montgomery constants by value: W[x] * tau[x]   rgchk.rs:323-325 (A tau), 146-156
struct SmallArgs { u64 v[16]; };      // Montgomery form of -8 .. 7 at index (t & 15)
__global__ void __launch_bounds__(256) k_plus_wsum_ring_small(const u64* __restrict__ W, const signed char* __restrict__ tau, size_t nrows, SmallArgs sm, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nrows; x += (size_t)gridDim.x * blockDim.x) {
        const int t = tau[x];
        if (!t) continue;
        u64 s = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) if ((t & 15) == k) s = sm.v[k];
        const u64* wx = W + x * PD;
#pragma unroll
        for (int c = 0; c < PD; ++c) acc[c] = Fm::add(acc[c], Fm::mul(s, wx[c]));
    }
    wsum_finish<void>(acc, po);
}
// scalar weights, small scalar entries: sum_x W[x] tau[x] (coefficient 0 of the partial; Montgomery)                 rgchk.rs:131-135
__global__ void __launch_bounds__(256) k_plus_wsum_scalar_small(const u64* __restrict__ W, const signed char* __restrict__ tau, size_t nrows, SmallArgs sm, const POut po) {
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nrows; x += (size_t)gridDim.x * blockDim.x) {
        const int t = tau[x];
        if (!t) continue;
        u64 s = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) if ((t & 15) == k) s = sm.v[k];
        acc[0] = Fm::add(acc[0], Fm::mul(s, W[x]));
    }
    wsum_finish<void>(acc, po);
}
// w = M^T eq(r, .) for a sparse matrix of ring elements held by columns: thread x sums its column                      setchk.rs:217-240 regrouped
__global__ void k_plus_mt_eq(const u64* __restrict__ eq, const u64* __restrict__ col_ptr, const u32* __restrict__ erow, const u64* __restrict__ val, size_t ncols, u64* __restrict__ w) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ncols) return;
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    for (size_t e = col_ptr[x]; e < col_ptr[x + 1]; ++e) {
        const u64 q = eq[erow[e]];
#pragma unroll
        for (int c = 0; c < PD; ++c) acc[c] = Fm::add(acc[c], Fm::mul(q, val[e * PD + c]));
    }
#pragma unroll
    for (int c = 0; c < PD; ++c) w[x * PD + c] = acc[c];
}
// cf(f) -> k balanced base-b digits per coefficient -> exponent codes of M_f = exp(D_f): codes[kk][coefficient][x]       rgchk.rs:262-295
__global__ void k_plus_digit_codes(const u64* __restrict__ f, size_t n, long long b, int k, unsigned char* __restrict__ codes, size_t code_pitch, int* __restrict__ err) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * PD) return;
    const size_t x = i / PD; const int c = (int)(i % PD);
    const u64 v = f[i];
    if (v >= Fm::P) { atomicExch(err, 3); return; }      // non-canonical input
    const bool negv = v > (Fm::P - 1) / 2; const u64 mag = negv ? Fm::P - v : v;
    if (mag >> 62) { atomicExch(err, 1); return; }
    int64_t dg[16];
    if (!balanced_digits(negv ? -(int64_t)mag : (int64_t)mag, b, k, dg)) { atomicExch(err, 1); return; }
    for (int kk = 0; kk < k; ++kk) {
        if (dg[kk] <= -(PD / 2) || dg[kk] >= PD / 2) { atomicExch(err, 1); return; }      // exp is defined on (-d/2, d/2)
        codes[((size_t)kk * PD + c) * code_pitch + x] = (unsigned char)((dg[kk] + PD) & (PD - 1));
    }
}

// ---------------------------------------------------------------- commitment transformation (cm.rs)
LF_HD u64 small_to_field(int v) { return v >= 0 ? (u64)v : Fm::P - (u64)(-v); }
// h[x] = sum_kk M_f[kk][x] . s'_kk = sum_{kk, c} X^code s'[kk][c]  (cm.rs:83-103): rotations of the short challenges, exact in 32-bit integers
__global__ void __launch_bounds__(128) k_plus_h(const unsigned char* __restrict__ codes, size_t code_pitch, size_t nrows, int kd /* k * 16 */, const short* __restrict__ sp /* kd x 16 */, u64* __restrict__ h) {
    extern __shared__ short sps[];
    for (int i = threadIdx.x; i < kd * PD; i += blockDim.x) sps[i] = sp[i];
    __syncthreads();
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nrows) return;
    int acc[PD];
#pragma unroll
    for (int o = 0; o < PD; ++o) acc[o] = 0;
    for (int q = 0; q < kd; ++q) {
        const unsigned cd = codes[(size_t)q * code_pitch + x];
        if (cd >= PD) continue;
        const short* sq = sps + q * PD;
#pragma unroll
        for (int o = 0; o < PD; ++o) { const int v = sq[(o - cd) & (PD - 1)]; acc[o] += (unsigned)o >= cd ? v : -v; }
    }
#pragma unroll
    for (int o = 0; o < PD; ++o) h[x * PD + o] = small_to_field(acc[o]);
}
// U = rho0 t0 + rho1 t1 with t(z) = tensor(c_z) (x) s' (x) (1, d', ..) (x) (1, X, ..) (cm.rs:590-601): entry ((i kd + j) l + a) 16 + b is
// scal[i][a] * (s'_j X^b), scal[i][a] = (rho0 tensor(c0)_i + rho1 tensor(c1)_i) d'^a in Montgomery form
__global__ void k_plus_tz(const u64* __restrict__ scal, const short* __restrict__ sp, int kd, int l, size_t nt, u64* __restrict__ U) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nt) return;
    const unsigned b = idx % PD; const int a = (int)((idx / PD) % l); const int j = (int)((idx / PD / l) % kd); const size_t i = idx / PD / l / kd;
    const u64 sc = scal[i * l + a]; const short* sj = sp + j * PD;
#pragma unroll
    for (int o = 0; o < PD; ++o) { const int v = sj[(o - b) & (PD - 1)]; U[idx * PD + o] = Fm::mul(sc, small_to_field((unsigned)o >= b ? v : -v)); }
}
// out[x] (+)= ra tau[x] + rb m_tau[x] + rc f[x] + rd h[x] as ring elements (the rc-weighted sum of one instance's four tables, cm.rs:285-300)
struct Lin4 { u64 ra_m, rb, rc_m, rd_m; };      // ra, rc, rd in Montgomery form, rb canonical
__global__ void k_plus_lin4(const signed char* __restrict__ tau, const unsigned char* __restrict__ mcode, const u64* __restrict__ f, const u64* __restrict__ h, size_t n, Lin4 w, int accumulate, u64* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * PD) return;
    const size_t x = i / PD; const unsigned o = (unsigned)(i % PD);
    u64 v = Fm::add(Fm::mul(w.rc_m, f[i]), Fm::mul(w.rd_m, h[i]));
    if (o == 0) v = Fm::add(v, Fm::mul(w.ra_m, small_to_field(tau[x])));
    if (o == mcode[x]) v = Fm::add(v, w.rb);
    out[i] = accumulate ? Fm::add(out[i], v) : v;
}
// G[y] += sum_e M[y][col_e] * Z[col_e] (row-major CSR, general ring products; M's coefficients are lifted to Montgomery form so the products are canonical)
__global__ void __launch_bounds__(128) k_plus_spmv_acc(const u64* __restrict__ row_ptr, const u64* __restrict__ col, const u64* __restrict__ val, size_t nrows, const u64* __restrict__ Z, u64 r2, u64* __restrict__ G) {
    const size_t y = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nrows) return;
    u64 acc[PD];
#pragma unroll
    for (int c = 0; c < PD; ++c) acc[c] = 0;
    for (u64 e = row_ptr[y]; e < row_ptr[y + 1]; ++e) {
        u64 v[PD];
#pragma unroll
        for (int j = 0; j < PD; ++j) v[j] = Z[col[e] * PD + j];
#pragma unroll 1
        for (int i = 0; i < PD; ++i) {
            const u64 mi = val[e * PD + i];
            if (mi) { const u64 mm = Fm::mul(mi, r2);
#pragma unroll
                for (int j = 0; j < PD; ++j) acc[j] = Fm::add(acc[j], Fm::mul(mm, v[j])); }
            const u64 top = v[PD - 1];
#pragma unroll
            for (int j = PD - 1; j > 0; --j) v[j] = v[j - 1];
            v[0] = Fm::neg(top);
        }
    }
#pragma unroll
    for (int c = 0; c < PD; ++c) G[y * PD + c] = Fm::add(G[y * PD + c], acc[c]);
}
// one round of the cm sumcheck (cm.rs:285-307 collapsed): h(X)[o] = sum_b eq(X) G(X)[o] + S(X) U(X)[o], X = 0, 1, 2; thread = (pair, coefficient)
__global__ void __launch_bounds__(256) k_plus_cm_round(const u64* __restrict__ sc /* eq | S, stride */, size_t stride, const u64* __restrict__ G, const u64* __restrict__ U, size_t n_pairs, u64* __restrict__ partial) {
    u64 acc[3] = {0, 0, 0};
    const unsigned o = threadIdx.x & (PD - 1);
    for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; b < n_pairs; b += ((size_t)gridDim.x * blockDim.x) >> 4) {
        const u64 e0 = sc[2 * b], e1 = sc[2 * b + 1], s0 = sc[stride + 2 * b], s1 = sc[stride + 2 * b + 1];
        const u64 g0 = G[(2 * b) * PD + o], g1 = G[(2 * b + 1) * PD + o], u0 = U[(2 * b) * PD + o], u1 = U[(2 * b + 1) * PD + o];
        acc[0] = Fm::add(acc[0], Fm::add(Fm::mul(e0, g0), Fm::mul(s0, u0)));
        acc[1] = Fm::add(acc[1], Fm::add(Fm::mul(e1, g1), Fm::mul(s1, u1)));
        const u64 e2 = Fm::add(e1, Fm::sub(e1, e0)), s2 = Fm::add(s1, Fm::sub(s1, s0)), g2 = Fm::add(g1, Fm::sub(g1, g0)), u2 = Fm::add(u1, Fm::sub(u1, u0));
        acc[2] = Fm::add(acc[2], Fm::add(Fm::mul(e2, g2), Fm::mul(s2, u2)));
    }
    __shared__ u64 sh[8][3][PD];
#pragma unroll
    for (int X = 0; X < 3; ++X) { u64 v = acc[X]; v = Fm::add(v, __shfl_xor_sync(0xffffffffu, v, 16)); if ((threadIdx.x & 31) < PD) sh[threadIdx.x >> 5][X][o] = v; }
    __syncthreads();
    if (threadIdx.x < 3 * PD) { const int X = threadIdx.x / PD, oo = threadIdx.x % PD; u64 v = 0; for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = Fm::add(v, sh[wv][X][oo]); partial[(size_t)blockIdx.x * 3 * PD + threadIdx.x] = v; }
}
// fix_variables on ring-valued tables (coefficient-wise): out[t][b][o] = in[t][2b][o] + r (in[t][2b+1][o] - in[t][2b][o])
__global__ void k_plus_fold_ring(const u64* __restrict__ in, size_t in_len, u64* __restrict__ out, size_t n_out, u64 r) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t t = blockIdx.y;
    if (i >= n_out * PD) return;
    const size_t b = i / PD, o = i % PD; const u64* base = in + t * in_len * PD;
    const u64 a = base[(2 * b) * PD + o], c = base[(2 * b + 1) * PD + o];
    out[t * n_out * PD + i] = Fm::add(a, Fm::mul(r, Fm::sub(c, a)));
}
// g[x] = s0 tau[x] + s1 m_tau[x] + s2 f[x] + h[x] (cm.rs:165-182): s small (|.| <= 128); s2 f is the only real ring product
struct GArgs { short s0[PD], s1[PD]; u64 s2m[PD]; };      // s2 in Montgomery form
__global__ void __launch_bounds__(128) k_plus_g(const signed char* __restrict__ tau, const unsigned char* __restrict__ mcode, const u64* __restrict__ f, const u64* __restrict__ h, size_t n, GArgs a, int accumulate, u64* __restrict__ g) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 acc[PD], v[PD];
    const int t = tau[x]; const unsigned cd = mcode[x];
#pragma unroll
    for (int o = 0; o < PD; ++o) { const int s1v = a.s1[(o - cd) & (PD - 1)]; acc[o] = Fm::add(h[x * PD + o], small_to_field(a.s0[o] * t + ((unsigned)o >= cd ? s1v : -s1v))); v[o] = f[x * PD + o]; }
#pragma unroll 1
    for (int i = 0; i < PD; ++i) {      // + s2_i (X^i f)
        const u64 si = a.s2m[i];
        if (si) {
#pragma unroll
            for (int j = 0; j < PD; ++j) acc[j] = Fm::add(acc[j], Fm::mul(si, v[j])); }
        const u64 top = v[PD - 1];
#pragma unroll
        for (int j = PD - 1; j > 0; --j) v[j] = v[j - 1];
        v[0] = Fm::neg(top);
    }
#pragma unroll
    for (int o = 0; o < PD; ++o) g[x * PD + o] = accumulate ? Fm::add(g[x * PD + o], acc[o]) : acc[o];      // (Mlin::mlin sums the instances' g, mlin.rs:97-103)
}
// S[x] = sum_l tau_l[x] in Montgomery form (the scalar factor of the t(z) terms, cm.rs:303-304)
struct TauList { const signed char* p[64]; int n; };
__global__ void k_plus_tau_sum(TauList tl, size_t n, SmallArgs sm, u64* __restrict__ S) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 v = 0;
    for (int l = 0; l < tl.n; ++l) { const int t = tl.p[l][x]; u64 s = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) if ((t & 15) == k) s = sm.v[k];
        v = Fm::add(v, s); }
    S[x] = v;
}

// Decomp::decompose (decomp.rs:37-40): every coefficient of f -> two balanced base-B digits, F0 and F1 as canonical field elements
__global__ void k_plus_split2(const u64* __restrict__ f, size_t words, long long B, u64* __restrict__ F0, u64* __restrict__ F1, int* __restrict__ err) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= words) return;
    const u64 v = f[i];
    if (v >= Fm::P) { atomicExch(err, 3); return; }
    const bool negv = v > (Fm::P - 1) / 2; const u64 mag = negv ? Fm::P - v : v;
    int64_t dg[2];
    if (!balanced_digits(negv ? -(int64_t)mag : (int64_t)mag, B, 2, dg)) { atomicExch(err, 1); return; }
    F0[i] = dg[0] < 0 ? Fm::P - (u64)(-dg[0]) : (u64)dg[0]; F1[i] = dg[1] < 0 ? Fm::P - (u64)(-dg[1]) : (u64)dg[1];
}

// one round of the R1CS linearization sumcheck (r1cs.rs:92: eq (ga gb - gc), degree 3) on ring-valued tables: the only LatticeFold+ sumcheck with
// real ring products.  16 lanes per pair, lane k owns coefficient k; ga(X) (lifted to Montgomery form) and gb(X) are exchanged through shared
// memory and lane k forms coefficient k of the negacyclic product.  partial[block][4][16]
__global__ void __launch_bounds__(256) k_plus_r1cs_round(const u64* __restrict__ eq, const u64* __restrict__ G /* [3][len][16] */, size_t len, size_t n_pairs, u64 r2, u64* __restrict__ partial) {
    __shared__ u64 sa[16][PD], sb[16][PD];      // per pair slot of the block
    const unsigned k = threadIdx.x & (PD - 1), slot = threadIdx.x >> 4;
    u64 acc[4] = {0, 0, 0, 0};
    const size_t pairs_per_pass = ((size_t)gridDim.x * blockDim.x) >> 4;
    for (size_t b0 = 0; b0 < n_pairs; b0 += pairs_per_pass) {
        const size_t b = b0 + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4); const bool live = b < n_pairs;
        u64 a = 0, da = 0, bb = 0, db = 0, c = 0, dc = 0, e = 0, de = 0;
        if (live) { const u64 a1 = G[(2 * b + 1) * PD + k], b1 = G[(len + 2 * b + 1) * PD + k], c1 = G[(2 * len + 2 * b + 1) * PD + k], e1 = eq[2 * b + 1];
            a = G[(2 * b) * PD + k]; bb = G[(len + 2 * b) * PD + k]; c = G[(2 * len + 2 * b) * PD + k]; e = eq[2 * b];
            da = Fm::sub(a1, a); db = Fm::sub(b1, bb); dc = Fm::sub(c1, c); de = Fm::sub(e1, e); }
#pragma unroll 1
        for (int X = 0; X < 4; ++X) {
            __syncwarp();
            sa[slot][k] = Fm::mul(a, r2); sb[slot][k] = bb;
            __syncwarp();
            u64 p = 0;
#pragma unroll
            for (int i = 0; i < PD; ++i) { const u64 pr = Fm::mul(sa[slot][i], sb[slot][(k - i) & (PD - 1)]); p = (unsigned)i <= k ? Fm::add(p, pr) : Fm::sub(p, pr); }
            if (live) acc[X] = Fm::add(acc[X], Fm::mul(e, Fm::sub(p, c)));
            a = Fm::add(a, da); bb = Fm::add(bb, db); c = Fm::add(c, dc); e = Fm::add(e, de);
        }
    }
    __shared__ u64 sh[8][4][PD];
#pragma unroll
    for (int X = 0; X < 4; ++X) { u64 v = acc[X]; v = Fm::add(v, __shfl_xor_sync(0xffffffffu, v, 16)); if ((threadIdx.x & 31) < PD) sh[threadIdx.x >> 5][X][k] = v; }
    __syncthreads();
    if (threadIdx.x < 4 * PD) { const int X = threadIdx.x / PD, kk = threadIdx.x % PD; u64 v = 0; for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = Fm::add(v, sh[wv][X][kk]); partial[(size_t)blockIdx.x * 4 * PD + threadIdx.x] = v; }
}

} }  // namespace lf::plus
