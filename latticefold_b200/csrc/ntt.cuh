// K12: batched negacyclic number-theoretic transform over Z_p[X]/(X^N + 1), N = 2^8 .. 2^16, for the two word-sized
// base fields of the reference's rings (Goldilocks p = 2^64 - 2^32 + 1 as canonical u64, BabyBear p = 15*2^27 + 1 as
// canonical u32).  BASELINE.json configs[4] / SURVEY.md 8 row C5(b): the reference has no transform of this shape (its
// "NTT form" is the CRT of the degree-24/72 rings, k_matrix_apply), so the definition is the textbook one and is fixed here:
//
//     forward   A[k] = sum_j a[j] * psi^(j * (2k + 1))            k = 0 .. N-1, natural order in, natural order out
//     inverse   a[j] = N^-1 * sum_k A[k] * psi^(-j * (2k + 1))
//
// psi = psi_N is the primitive 2N-th root of unity rho^(2^A / 2N), where 2^A is the 2-adic order used (A = 32 for
// Goldilocks, 27 for BabyBear) and rho = r0^u, r0 = g^((p-1)/2^A) for the field's customary generator g (7 / 31), u the
// smallest odd exponent with rho^(2^A / 32) = 2^6 for Goldilocks (so that every radix-16 butterfly's internal twiddle is a
// power of two: omega_16 = 2^12, 2^96 = -1) and u = 1 for BabyBear.  lf_ntt_root() returns psi_N; the oracle restates the rule.
//
// Kernel shape (B200): one CTA transforms one polynomial (several for N < 4096) entirely on chip: the polynomial arrives in
// shared memory with ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier), N/16 threads each keep 16 coefficients in
// registers and run ceil(log2 N / 4) Stockham autosort passes (radix 16, the first pass radix 2/4/8 when log2 N is not a
// multiple of 4) -- read 16 strided words, twiddle, radix-R DFT in registers, barrier, write 16 words to the autosort position
// (shared-memory index padded by 1/16 so the radix-16 scatter is bank-conflict free) -- and the last pass stores straight to
// global memory, coalesced, in natural order.  HBM traffic = the algorithmic 2 * N * sizeof(T) per polynomial.
// N = 2^15, 2^16 run as a four-step transform (N = N1 * 4096): strided 4096-point sub-transforms into a scratch that is
// the size of the batch, then a radix-N1 cross pass (the kernels are integer-pipe bound: the second trip through HBM is cheaper
// than the launch tails of L2-sized chunks were).
#pragma once
#include "field.cuh"

namespace lf { namespace ntt {

// ------------------------------------------------------------------------------------------------ field adaptors
struct GlF {
    typedef u64 T;
    static constexpr int ID = 0, TWO_ADICITY = 32;
    static constexpr u64 P = Goldilocks::P, GEN = 7;
    static LF_HD T add(T a, T b) { return Goldilocks::add(a, b); }
    static LF_HD T sub(T a, T b) { return Goldilocks::sub(a, b); }
    static LF_HD T mul_tw(T a, T w) { return Goldilocks::mul(a, w); }      // table twiddles are canonical
    static u64 to_tw(u64 w) { return w; }
    // a * 2^s mod p for a compile-time-foldable s in [0, 192): 2^64 = 2^32 - 1, 2^96 = -1
    static LF_HD T mul_pow2(T a, int s) {
        bool neg = false;
        if (s >= 96) { s -= 96; neg = true; }
        T r;
        if (s == 0) r = a;
        else if (s < 64) r = Goldilocks::reduce128(a << s, a >> (64 - s));
        else { int t = s - 64; u64 ylo = a << t, yhi = t ? a >> (64 - t) : 0; r = Goldilocks::sub(Goldilocks::reduce128(0, ylo), yhi << 32); }
        return neg ? Goldilocks::neg(r) : r;
    }
    // a * omega_16^e, omega_16 = 2^12
    static LF_D T mul_w16(T a, int e) { return mul_pow2(a, 12 * (e & 15)); }
};

struct BbF {
    typedef u32 T;
    static constexpr int ID = 1, TWO_ADICITY = 27;
    static constexpr u64 P = 2013265921ULL, GEN = 31;
    static constexpr u32 PINV = 2281701377u;      // p^-1 mod 2^32
    static LF_HD T add(T a, T b) { u32 s = a + b, t = s - (u32)P; return t < s ? t : s; }           // min(s, s - p) as unsigned
    static LF_HD T sub(T a, T b) { u32 d = a - b, t = d + (u32)P; return t < d ? t : d; }           // min(d, d + p)
    // a canonical, w in Montgomery form (w * 2^32 mod p): returns a * w mod p, canonical
    static LF_HD T mul_tw(T a, T w) {
        u64 t = (u64)a * w; u32 m = (u32)t * PINV; u32 u = (u32)(((u64)m * (u32)P) >> 32), h = (u32)(t >> 32);
        u32 r = h - u; return h < u ? r + (u32)P : r;
    }
    static u64 to_tw(u64 w) { return (u64)(((u128)w << 32) % P); }
#if defined(__CUDACC__)
    static LF_D T mul_w16(T a, int e);            // Montgomery constants in __constant__ memory (ntt.cu)
#endif
};

// ------------------------------------------------------------------------------------------------ plan geometry
// N = 2^LOGN = R1 * 16^(P-1): P passes, the first of radix R1 in {2,4,8,16}
template <int LOGN, bool STRIDED = false> struct Geo {
    static constexpr int N = 1 << LOGN, P = (LOGN + 3) / 4, R1 = 1 << (LOGN - 4 * (P - 1));
    static constexpr int TP = N / 16;                               // threads per polynomial
    static constexpr int PP = STRIDED ? 4 : (TP >= 256 ? 1 : 256 / TP);   // polynomials per CTA
    static constexpr int THREADS = TP * PP;
    static constexpr int PADN = N + N / 16;                         // padded shared-memory words per polynomial
    static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : 4);
};

}}  // namespace lf::ntt
