// Host Poseidon, dense (MDS) layer for the Goldilocks field on AVX-512 IFMA hosts: out[i] = sum_j M[i][j] * st[j] mod p for the
// width-24 state of the Fiat-Shamir transcript (transcript_host.hpp; reference: crates/latticefold/src/transcript/poseidon.rs:29-75,
// ark-crypto-primitives 0.4.0 PoseidonSponge::permute).  Host code only, compiled by g++ (nvcc's front end does not know the
// AVX-512 builtins); selected at run time by poseidon_ifma_supported(), the scalar add/adc path of transcript_host.hpp remains for
// hosts without IFMA and is bit-identical (tests/test_cabi_cpu.py runs both).
//
// Every 64-bit operand is cut into a 52-bit and a 12-bit limb.  vpmadd52luq / vpmadd52huq add the low / high 52 bits of a
// 52 x 52 product to a 64-bit lane, so the four limb products of one 64 x 64 product take 7 instructions per 8 lanes
// (lo+hi of m0 s0, m0 s1, m1 s0; lo of m1 s1 < 2^24) instead of 5 scalar micro-ops per product, and land in three sums of
// weight 2^0, 2^52, 2^104 (each < 2^59 after 24 products).  The sums are regrouped into 32-bit digits d0..d4 and folded with
// 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 (mod p = 2^64 - 2^32 + 1).
#include <immintrin.h>
#include <cstdint>
#include <cstring>

namespace lf {

typedef uint64_t u64;
static constexpr int W = 24;
static constexpr u64 P = 0xFFFFFFFF00000001ULL, EPS = 0xFFFFFFFFULL, M52 = (1ULL << 52) - 1;

struct PoseidonIfmaMatrix { alignas(64) u64 limb[W][2][3][8]; };   // [column j][limb of M][zmm z][lane]: limb of M[8 z + lane][j]

bool poseidon_ifma_supported() {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512dq") &&
           __builtin_cpu_supports("avx512vl");
}

void poseidon_ifma_prepare(const u64* m, PoseidonIfmaMatrix* out) {   // m: W x W row-major, canonical
    for (int j = 0; j < W; ++j) for (int z = 0; z < 3; ++z) for (int l = 0; l < 8; ++l) {
        const u64 v = m[(8 * z + l) * W + j];
        out->limb[j][0][z][l] = v & M52; out->limb[j][1][z][l] = v >> 52;
    }
}

#define LF_IFMA_TARGET __attribute__((target("avx512f,avx512ifma,avx512dq,avx512vl")))

// st: W lanes, any u64 representatives on entry, canonical on return
LF_IFMA_TARGET void poseidon_ifma_dense(const PoseidonIfmaMatrix* M, u64* st) {
    alignas(64) u64 sl[2][W];
    const __m512i m52 = _mm512_set1_epi64((long long)M52);
    for (int z = 0; z < 3; ++z) {
        const __m512i v = _mm512_loadu_si512(st + 8 * z);
        _mm512_store_si512(&sl[0][8 * z], _mm512_and_si512(v, m52));
        _mm512_store_si512(&sl[1][8 * z], _mm512_srli_epi64(v, 52));
    }
    __m512i acc[3][3];
    for (int k = 0; k < 3; ++k) for (int z = 0; z < 3; ++z) acc[k][z] = _mm512_setzero_si512();
    for (int j = 0; j < W; ++j) {
        const __m512i s0 = _mm512_set1_epi64((long long)sl[0][j]), s1 = _mm512_set1_epi64((long long)sl[1][j]);
#pragma GCC unroll 3
        for (int z = 0; z < 3; ++z) {
            const __m512i m0 = _mm512_load_si512(M->limb[j][0][z]), m1 = _mm512_load_si512(M->limb[j][1][z]);
            acc[0][z] = _mm512_madd52lo_epu64(acc[0][z], m0, s0);
            acc[1][z] = _mm512_madd52hi_epu64(acc[1][z], m0, s0);
            acc[1][z] = _mm512_madd52lo_epu64(acc[1][z], m0, s1);
            acc[1][z] = _mm512_madd52lo_epu64(acc[1][z], m1, s0);
            acc[2][z] = _mm512_madd52hi_epu64(acc[2][z], m0, s1);
            acc[2][z] = _mm512_madd52hi_epu64(acc[2][z], m1, s0);
            acc[2][z] = _mm512_madd52lo_epu64(acc[2][z], m1, s1);
        }
    }
    // R = acc_0 + acc_1 2^52 + acc_2 2^104 = d0 + d1 2^32 + d2 2^64 + d3 2^96 + d4 2^128 with every d < 2^36 (unnormalised 32-bit digits);
    // R = (d0 - d2 - d3) + (d1 + d2 - d4) 2^32 (mod p).  Both brackets are made positive by the offsets 2^40 and 2^36 * 2^32,
    // whose sum mod p is subtracted at the end.
    const __m512i m32 = _mm512_set1_epi64((long long)EPS), eps = m32, pp = _mm512_set1_epi64((long long)P);
    const u64 off_a = 1ULL << 40, off_b = 1ULL << 36;
    // (2^40 + 2^68) mod p: 2^68 = 2^4 * 2^64 = 2^4 (2^32 - 1)
    const __m512i cst = _mm512_set1_epi64((long long)(off_a + (16ULL << 32) - 16ULL));
    for (int z = 0; z < 3; ++z) {
        const __m512i a0 = acc[0][z], a1 = acc[1][z], a2 = acc[2][z];     // a0 < 2^57, a1 < 2^59, a2 < 2^29
        const __m512i d0 = _mm512_and_si512(a0, m32);
        const __m512i d1 = _mm512_add_epi64(_mm512_srli_epi64(a0, 32), _mm512_slli_epi64(_mm512_and_si512(a1, _mm512_set1_epi64(0xFFF)), 20));
        const __m512i d2 = _mm512_and_si512(_mm512_srli_epi64(a1, 12), m32);
        const __m512i d3 = _mm512_add_epi64(_mm512_srli_epi64(a1, 44), _mm512_slli_epi64(_mm512_and_si512(a2, _mm512_set1_epi64(0xFFFFFF)), 8));
        const __m512i d4 = _mm512_srli_epi64(a2, 24);
        // A = d0 - d2 - d3 + 2^40 in (0, 2^41); B = d1 + d2 - d4 + 2^36 in (0, 2^38)
        const __m512i A = _mm512_sub_epi64(_mm512_add_epi64(d0, _mm512_set1_epi64((long long)off_a)), _mm512_add_epi64(d2, d3));
        const __m512i B = _mm512_sub_epi64(_mm512_add_epi64(_mm512_add_epi64(d1, d2), _mm512_set1_epi64((long long)off_b)), d4);
        // B 2^32 = bl 2^32 + bh 2^64 = bl 2^32 + bh (2^32 - 1)
        const __m512i bl = _mm512_and_si512(B, m32), bh = _mm512_srli_epi64(B, 32);
        const __m512i lowpart = _mm512_add_epi64(A, _mm512_sub_epi64(_mm512_slli_epi64(bh, 32), bh));   // < 2^42
        const __m512i hipart = _mm512_slli_epi64(bl, 32);                                                // <= 2^64 - 2^32
        __m512i t = _mm512_add_epi64(hipart, lowpart);
        const __mmask8 cy = _mm512_cmplt_epu64_mask(t, hipart);
        t = _mm512_mask_add_epi64(t, cy, t, eps);                     // wrapped past 2^64 = eps (mod p); cannot wrap twice
        // subtract the offsets, then canonicalise
        __m512i r = _mm512_sub_epi64(t, cst);
        const __mmask8 bw = _mm512_cmplt_epu64_mask(t, cst);
        r = _mm512_mask_sub_epi64(r, bw, r, eps);                     // borrowed 2^64: give eps back, i.e. r = t - cst + p
        const __mmask8 ge = _mm512_cmpge_epu64_mask(r, pp);
        r = _mm512_mask_sub_epi64(r, ge, r, pp);
        _mm512_storeu_si512(st + 8 * z, r);
    }
}

}  // namespace lf
