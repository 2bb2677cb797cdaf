// Base-field and slot-field arithmetic for the LatticeFold rings, usable from host and device code.
//
// Goldilocks  p = 2^64 - 2^32 + 1, ring Z_p[X]/(X^24 - X^12 + 1) = 8 slots of Fq3 = Fq[Y]/(Y^3 - nu), nu = 2^40
//   (reference type aliases: crates/cyclotomic-rings/src/rings/goldilocks.rs:9-20; the arithmetic itself is the
//    un-vendored stark-rings crate -- see DESIGN.md "Conventions" for what is and is not pinned).
// Elements are canonical (value in [0, p)), never Montgomery: 2^64 = 2^32 - 1 (mod p) makes the reduction of a
// 128-bit product a handful of 32/64-bit adds, cheaper on the GPU's 32-bit integer pipes than a Montgomery REDC.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LF_HD __host__ __device__ __forceinline__
#define LF_D __device__ __forceinline__
#define LF_HD_CALL __host__ __device__ __noinline__      // real calls: keeps the generic (parity-only) rings' code size and build time sane
#else
#define LF_HD inline
#define LF_D inline
#define LF_HD_CALL __attribute__((noinline))
#endif

namespace lf {

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

// 64 x 64 -> 128.  On the device the product is written on 32-bit halves with explicit carry chains: ptxas fuses each
// mad.lo.cc / madc.hi.cc pair into one IMAD.WIDE.U32 with a carry predicate (4 IMAD.WIDE + 2 IADD3.X in SASS), whereas
// `a * b` and `__umul64hi(a, b)` are lowered separately and recompute the partial products (~2x the instructions).
LF_HD void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
    asm("mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %7;\n\t"
        "mul.hi.u32 %3, %5, %7;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    lo = ((u64)r1 << 32) | r0; hi = ((u64)r3 << 32) | r2;
#else
    u128 x = (u128)a * b; lo = (u64)x; hi = (u64)(x >> 64);
#endif
}

// Accumulator for lazily reduced sums of 128-bit products.  Device layout: the even-aligned partial products
// (a0 b0 at bit 0, a1 b1 at bit 64) and the odd-aligned ones (a0 b1 + a1 b0 at bit 32) are summed in separate word
// chains, so one 64 x 64 multiply-accumulate is 4 IMAD.WIDE.U32 + 2-3 carry adds and no chain depends on another.
// Capacity: 2^31 products.  words() returns the 192-bit value (w2 < 2^32).
struct Acc192 {
#if defined(__CUDA_ARCH__)
    u32 e0, e1, e2, e3, e4, o1, o2, o3;
    LF_HD void clear() { e0 = e1 = e2 = e3 = e4 = o1 = o2 = o3 = 0; }
    LF_HD void mac(u64 a, u64 b) {
        u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
        asm("mad.lo.cc.u32 %0, %8, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %5, %8, %11, %5;\n\t"
            "madc.hi.cc.u32 %6, %8, %11, %6;\n\t"
            "addc.u32 %7, %7, 0;\n\t"
            "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
            "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
            "addc.u32 %7, %7, 0;"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o1), "+r"(o2), "+r"(o3)
            : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    }
    // a < 2^32: two wide multiplies instead of four
    LF_HD void mac_small(u32 a, u64 b) {
        u32 b0 = (u32)b, b1 = (u32)(b >> 32);
        asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %5, %8, %10, %5;\n\t"
            "madc.hi.cc.u32 %6, %8, %10, %6;\n\t"
            "addc.u32 %7, %7, 0;"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o1), "+r"(o2), "+r"(o3)
            : "r"(a), "r"(b0), "r"(b1));
    }
    LF_HD void add(u64 a) {
        u32 a0 = (u32)a, a1 = (u32)(a >> 32);
        asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4) : "r"(a0), "r"(a1));
    }
    LF_HD void words(u64& w0, u64& w1, u32& w2) const {
        u32 r0 = e0, r1, r2, r3, r4;
        asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, %9;\n\taddc.cc.u32 %2, %6, %10;\n\taddc.u32 %3, %7, 0;"
            : "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4) : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(o1), "r"(o2), "r"(o3));
        w0 = ((u64)r1 << 32) | r0; w1 = ((u64)r3 << 32) | r2; w2 = r4;
    }
#else
    u64 h0, h1, h2;
    LF_HD void clear() { h0 = h1 = h2 = 0; }
    LF_HD void mac(u64 a, u64 b) {
        u128 x = (u128)a * b; u128 s = (u128)h0 + (u64)x; h0 = (u64)s;
        s = (u128)h1 + (u64)(x >> 64) + (u64)(s >> 64); h1 = (u64)s; h2 += (u64)(s >> 64);
    }
    LF_HD void mac_small(u32 a, u64 b) { mac((u64)a, b); }
    LF_HD void add(u64 a) { u128 s = (u128)h0 + a; h0 = (u64)s; s = (u128)h1 + (u64)(s >> 64); h1 = (u64)s; h2 += (u64)(s >> 64); }
    LF_HD void words(u64& w0, u64& w1, u32& w2) const { w0 = h0; w1 = h1; w2 = (u32)h2; }
#endif
};

struct Goldilocks {
    static constexpr u64 P = 0xFFFFFFFF00000001ULL;
    static constexpr u64 EPS = 0xFFFFFFFFULL;  // 2^64 mod p
    static constexpr int NU_SHIFT = 40;        // nu = 2^40 (a primitive 24th root of unity)
    static constexpr u64 NU = (u64)1 << NU_SHIFT;
    typedef Acc192 Acc;

    // Device: straight-line carry-flag code, no compares or selects (ncu showed the 64-bit compare/select form the compiler
    // emits for `if (s < a || s >= P) s -= P` saturating the ALU pipe): a borrow becomes the mask 0xFFFFFFFF = 2^64 - p,
    // so "+= p on borrow" is one more masked subtract.  sub = 5 instructions, add(a, b) = a - (p - b) = 7.
    // Host: mask arithmetic -- the comparisons are data dependent coin flips and a mispredicted branch costs more than the
    // whole reduction (the Poseidon transcript runs ~10^7 of these per prover step).
    static LF_HD u64 sub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
        u32 d0, d1, m;
        asm("sub.cc.u32 %0, %3, %5;\n\t"
            "subc.cc.u32 %1, %4, %6;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %2;\n\t"
            "subc.u32 %1, %1, 0;"
            : "=&r"(d0), "=&r"(d1), "=&r"(m) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
        return ((u64)d1 << 32) | d0;
#else
        u64 d = a - b; d += (0 - (u64)(a < b)) & P; return d;
#endif
    }
    static LF_HD u64 add(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
        u32 n0, n1, d0, d1, m;
        asm("sub.cc.u32 %0, 1, %7;\n\t"             // n = p - b  (b < p: no borrow out)
            "subc.u32 %1, 0xffffffff, %8;\n\t"
            "sub.cc.u32 %2, %5, %0;\n\t"            // d = a - n
            "subc.cc.u32 %3, %6, %1;\n\t"
            "subc.u32 %4, 0, 0;\n\t"
            "sub.cc.u32 %2, %2, %4;\n\t"
            "subc.u32 %3, %3, 0;"
            : "=&r"(n0), "=&r"(n1), "=&r"(d0), "=&r"(d1), "=&r"(m) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
        return ((u64)d1 << 32) | d0;
#else
        u64 s = a + b; s -= (0 - (u64)((s < a) | (s >= P))) & P; return s;
#endif
    }
    static LF_HD u64 neg(u64 a) {
#if defined(__CUDA_ARCH__)
        return sub(0, a);
#else
        return a ? P - a : 0;
#endif
    }
    // (hi:lo) mod p, canonical.  2^64 = EPS, 2^96 = -1.
    static LF_HD u64 reduce128(u64 lo, u64 hi) {
#if defined(__CUDA_ARCH__)
        // hi = hh*2^32 + hl:  r = (lo - hh) - (p - hl*(2^32 - 1)), every borrow repaired by the mask trick, then one
        // canonicalising subtract of p.  Subtract chains only: ptxas keeps the hardware carry sense when a subc follows an
        // add.cc, so mixed chains do not compute what the PTX text says.  18 instructions, no compares.
        u32 t0, t1, n0, n1, m, v0, v1;
        asm("sub.cc.u32 %0, %7, %10;\n\t"           // t = lo - hh
            "subc.cc.u32 %1, %8, 0;\n\t"
            "subc.u32 %4, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %4;\n\t"            // wrapped by 2^64 = eps too much
            "subc.u32 %1, %1, 0;\n\t"
            "not.b32 %3, %9;\n\t"                   // n = p - hl * eps = (~hl) * 2^32 + hl + 1
            "add.cc.u32 %2, %9, 1;\n\t"
            "addc.u32 %3, %3, 0;\n\t"
            "sub.cc.u32 %0, %0, %2;\n\t"            // t -= n
            "subc.cc.u32 %1, %1, %3;\n\t"
            "subc.u32 %4, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %4;\n\t"            // borrowed: += p
            "subc.u32 %1, %1, 0;\n\t"
            "sub.cc.u32 %5, %0, 1;\n\t"             // canonicalise: v = t - p, keep t if that borrows
            "subc.cc.u32 %6, %1, 0xffffffff;\n\t"
            "subc.u32 %4, 0, 0;\n\t"
            "sub.cc.u32 %0, %5, %4;\n\t"
            "subc.u32 %1, %6, 0;"
            : "=&r"(t0), "=&r"(t1), "=&r"(n0), "=&r"(n1), "=&r"(m), "=&r"(v0), "=&r"(v1)
            : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
        return ((u64)t1 << 32) | t0;
#else
        u64 hh = hi >> 32, hl = hi & EPS;
        u64 t1 = (hl << 32) - hl;                      // hl * EPS < 2^64
        u64 t0 = lo - hh; t0 -= (0 - (u64)(lo < hh)) & EPS;
        u64 r = t0 + t1; r += (0 - (u64)(r < t1)) & EPS;
        r -= (0 - (u64)(r >= P)) & P;
        return r;
#endif
    }
    static LF_HD u64 mul(u64 a, u64 b) { u64 lo, hi; mul_wide(a, b, lo, hi); return reduce128(lo, hi); }
    static LF_HD u64 sqr(u64 a) { return mul(a, a); }
    // (w2:w1:w0) mod p with w2 < 2^32; 2^128 = -2^32 and w2 * 2^32 < p is already canonical
    static LF_HD u64 reduce192(const Acc192& a) {
        u64 w0, w1; u32 w2; a.words(w0, w1, w2);
        return sub(reduce128(w0, w1), (u64)w2 << 32);
    }
    static LF_HD u64 reduce(const Acc192& a) { return reduce192(a); }
    // slim accumulator for plain sums of field elements (no products): 96 bits in three registers, 2^32 terms
    struct Sum {
#if defined(__CUDA_ARCH__)
        u32 a0, a1, a2;
        LF_HD void clear() { a0 = a1 = a2 = 0; }
        LF_HD void add(u64 v) { asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+r"(a0), "+r"(a1), "+r"(a2) : "r"((u32)v), "r"((u32)(v >> 32))); }
        LF_HD u64 lo() const { return ((u64)a1 << 32) | a0; }
        LF_HD u64 hi() const { return a2; }
#else
        u128 v;
        LF_HD void clear() { v = 0; }
        LF_HD void add(u64 x) { v += x; }
        LF_HD u64 lo() const { return (u64)v; }
        LF_HD u64 hi() const { return (u64)(v >> 64); }
#endif
    };
    static LF_HD u64 reduce(const Sum& a) { return reduce128(a.lo(), a.hi()); }
    // slim accumulator for sums of (small multiplier) x (field element); holds 2^20 terms with multipliers below 2^12 (the
    // digit-weighted sums of the FOLD sumcheck's first two rounds).  Two independent 64-bit sums, one per 32-bit half of the field
    // element: each multiply-accumulate is two chained IMAD.WIDE on their own even-aligned register pairs and nothing else.  (The
    // 96-bit three-register form this replaces made the second wide multiply accumulate into the pair (e1, e2), which cannot be
    // even-aligned together with (e0, e1): ptxas inserted ~3.6 register moves per multiply-accumulate, all on the FMA-heavy pipe
    // the wide multiplies already saturate -- 87 of 190 instructions in the round-2 table loop, ncu r02w.)
    struct AccS {
#if defined(__CUDA_ARCH__)
        u64 L, H;      // (64-bit operands: the register allocator keeps each sum in one even-aligned pair)
        LF_HD void clear() { L = H = 0; }
        LF_HD void mac_small(u32 a, u64 b) {
            asm("mad.wide.u32 %0, %2, %3, %0;\n\tmad.wide.u32 %1, %2, %4, %1;" : "+l"(L), "+l"(H) : "r"(a), "r"((u32)b), "r"((u32)(b >> 32)));
        }
        // value = L + H 2^32
        LF_HD u64 lo() const { return L + (H << 32); }
        LF_HD u64 hi() const { const u64 s = L + (H << 32); return (H >> 32) + (s < L ? 1 : 0); }
#else
        u128 v;
        LF_HD void clear() { v = 0; }
        LF_HD void mac_small(u32 a, u64 b) { v += (u128)a * b; }
        LF_HD u64 lo() const { return (u64)v; }
        LF_HD u64 hi() const { return (u64)(v >> 64); }
#endif
    };
    static LF_HD u64 reduce(const AccS& a) { return reduce128(a.lo(), a.hi()); }
    static LF_HD u64 mul_nu(u64 a) { return reduce128(a << NU_SHIFT, a >> (64 - NU_SHIFT)); }
    // (lo + hi * 2^32) mod p for lo, hi < 2^40: recombination of the split-limb all-reduce
    static LF_HD u64 from_split(u64 lo, u64 hi) { return add(reduce128(lo, 0), reduce128(hi << 32, hi >> 32)); }
    // (lo + hi * 2^64) mod p for 128-bit lo, hi < 2^72 (host: sums of products in the Poseidon layers)
    static inline u64 reduce_wide(u128 lo, u128 hi) { u128 t = hi + (u64)(lo >> 64); return sub(reduce128((u64)lo, (u64)t), (u64)(t >> 64) << 32); }
    static LF_HD u64 from_i64(int64_t v) { return v >= 0 ? (u64)v : P - (u64)(-v); }   // |v| < p
    static LF_HD int64_t to_signed(u64 a) { return a <= (P - 1) / 2 ? (int64_t)a : -(int64_t)(P - a); }
    static inline u64 pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; } return r; }
    static inline u64 inv(u64 a) { return pow(a, P - 2); }
};

// Generic prime field for the reference's other rings (BabyBear p = 15*2^27+1, Frog p = 15912092521325583641): plain
// `%`-based arithmetic with an eagerly reduced accumulator.  These rings are carried for parity, not tuned: the headline
// workload is Goldilocks.  SMALL: p < 2^32, products fit in 64 bits; otherwise 128-bit intermediates.
template <u64 P_, u64 NU_, bool SMALL> struct ModField {
    static constexpr u64 P = P_, NU = NU_;
    static LF_HD u64 add(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= P) s -= P; return s; }
    static LF_HD u64 sub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
    static LF_HD u64 neg(u64 a) { return a ? P - a : 0; }
    static LF_HD_CALL u64 mul(u64 a, u64 b) { if (SMALL) return (a * b) % P; return (u64)(((u128)a * b) % P); }
    static LF_HD u64 sqr(u64 a) { return mul(a, a); }
    static LF_HD u64 mul_nu(u64 a) { return mul(a, NU); }
    struct Acc { u64 v; LF_HD void clear() { v = 0; } LF_HD void mac(u64 a, u64 b) { v = ModField::add(v, ModField::mul(a, b)); } LF_HD void mac_small(u32 a, u64 b) { mac((u64)a, b); } LF_HD void add(u64 a) { v = ModField::add(v, a % P); } };
    typedef Acc Sum; typedef Acc AccS;
    static LF_HD u64 reduce(const Acc& a) { return a.v; }
    static LF_HD_CALL u64 reduce128(u64 lo, u64 hi) { return (u64)((((u128)hi << 64) | lo) % P); }
    static LF_HD_CALL u64 from_split(u64 lo, u64 hi) { return (u64)((((u128)hi << 32) + lo) % P); }
    static inline u64 reduce_wide(u128 lo, u128 hi) { u128 t = ((hi % P) * (((u128)1 << 64) % P)) % P; return (u64)((t + lo % P) % P); }
    static LF_HD u64 from_i64(int64_t v) { return v >= 0 ? (u64)v % P : (P - ((u64)(-v) % P)) % P; }
    static LF_HD int64_t to_signed(u64 a) { return a <= (P - 1) / 2 ? (int64_t)a : -(int64_t)(P - a); }
    static inline u64 pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; } return r; }
    static inline u64 inv(u64 a) { return pow(a, P - 2); }
};
// nu = the primitive g-th root of unity h^((p-1)/g) for the smallest h of exact order g (same rule as the oracle's
// find_nu; this convention is "unpinned"): 1398021245 for BabyBear (g = 24, struct below), 2755067726615789629 for Frog (g = 8)
// BabyBear p = 15 * 2^27 + 1 (31 bits).  Same canonical-u64 interface as ModField, but sums of products stay lazily reduced:
// a product of two canonical values is < 2^62, one multiply-accumulate is one IMAD.WIDE.U32 with carry-out plus one carry
// add into a 96-bit accumulator (capacity 2^34 products), and the accumulator is folded mod p once
// (2^32 = 2^28 - 2 and 2^64 = (2^28 - 2)^2 mod p, then a single 64-bit remainder).  The eager ModField accumulator spent a
// full `%` per product: the commit's dot products and the FOLD sumcheck are ~81 MACs per slot-field product on this ring.
struct BabyBear {
    static constexpr u64 P = 2013265921ULL, NU = 1398021245ULL;
    static constexpr u64 C32 = ((u64)1 << 32) % P, C64 = (C32 * C32) % P;
    static LF_HD u64 add(u64 a, u64 b) { u64 s = a + b; if (s >= P) s -= P; return s; }
    static LF_HD u64 sub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
    static LF_HD u64 neg(u64 a) { return a ? P - a : 0; }
    // x mod p for any 64-bit x without a 64-bit division or multiply-high (the compiler's `x % P` is a 64 x 64 multiply-high on
    // 32-bit pipes): the upper word is brought below p, folded with 2^32 = 2^28 - 2 (mod p) into y < 2^60, and a Barrett quotient
    // estimate floor((y >> 28) * floor(2^62 / p) / 2^34) -- one 32-bit multiply-high, off by at most one -- leaves a 32-bit remainder.
    static constexpr u32 M62 = (u32)(((u128)1 << 62) / P);
    static LF_HD u64 red60(u64 y) {
        const u32 t = (u32)(y >> 28);
#if defined(__CUDA_ARCH__)
        const u32 q = __umulhi(t, M62) >> 2;
#else
        const u32 q = (u32)(((u64)t * M62) >> 34);
#endif
        u32 r = (u32)y - q * (u32)P;
        if (r >= (u32)P) r -= (u32)P;
        return r;
    }
    static LF_HD u64 red64(u64 x) {
        u32 xh = (u32)(x >> 32);
        if (xh >= 2 * (u32)P) xh -= 2 * (u32)P;
        if (xh >= (u32)P) xh -= (u32)P;
        return red60((u64)xh * (u32)C32 + (u32)x);
    }
    static LF_HD u64 mul(u64 a, u64 b) { const u64 x = a * b; return red60((x >> 32) * (u32)C32 + (u32)x); }      // a, b < p: upper word < 2^30
    static LF_HD u64 sqr(u64 a) { return mul(a, a); }
    static LF_HD u64 mul_nu(u64 a) { return mul(a, NU); }
    struct Acc {
#if defined(__CUDA_ARCH__)
        u32 e0, e1, e2;
        LF_HD void clear() { e0 = e1 = e2 = 0; }
        LF_HD void mac(u64 a, u64 b) {
            asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(e0), "+r"(e1), "+r"(e2) : "r"((u32)a), "r"((u32)b));
        }
        LF_HD void mac_small(u32 a, u64 b) { mac((u64)a, b); }
        LF_HD void add(u64 a) { asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+r"(e0), "+r"(e1), "+r"(e2) : "r"((u32)a), "r"((u32)(a >> 32))); }
        LF_HD u64 fold() const { return BabyBear::red64((u64)e2 * C64 + (u64)e1 * C32 + e0); }      // < 2^63 + 2^60 + 2^32
#else
        u128 v;
        LF_HD void clear() { v = 0; }
        LF_HD void mac(u64 a, u64 b) { v += (u128)a * b; }
        LF_HD void mac_small(u32 a, u64 b) { mac((u64)a, b); }
        LF_HD void add(u64 a) { v += a; }
        LF_HD u64 fold() const { return (u64)(v % P); }
#endif
    };
    typedef Acc Sum; typedef Acc AccS;
    static LF_HD u64 reduce(const Acc& a) { return a.fold(); }
    static LF_HD u64 reduce128(u64 lo, u64 hi) { return (u64)((((u128)hi << 64) | lo) % P); }
    static LF_HD u64 from_split(u64 lo, u64 hi) { return (u64)((((u128)hi << 32) + lo) % P); }
    // (lo + hi * 2^64) mod p word by word: 64-bit remainders by the constant p only (the host Poseidon calls this once per state lane
    // and layer; the 128-bit `%` it replaces is a library call)
    static inline u64 reduce_wide(u128 lo, u128 hi) {
        constexpr u64 C128 = (C64 * C64) % P;
        const u64 w0 = (u64)lo % P, w1 = ((u64)(lo >> 64) % P + (u64)hi % P) % P, w2 = (u64)(hi >> 64) % P;
        return (w0 + (w1 * C64) % P + (w2 * C128) % P) % P;
    }
    static LF_HD u64 from_i64(int64_t v) { return v >= 0 ? (u64)v % P : (P - ((u64)(-v) % P)) % P; }
    static LF_HD int64_t to_signed(u64 a) { return a <= (P - 1) / 2 ? (int64_t)a : -(int64_t)(P - a); }
    static inline u64 pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; } return r; }
    static inline u64 inv(u64 a) { return pow(a, P - 2); }
};
typedef ModField<15912092521325583641ULL, 2755067726615789629ULL, false> FrogField;

// ---------------------------------------------------------------------------------------------------------------
// Ring descriptor: slot field Fq[Y]/(Y^TAU - nu).  Rg::F is the base field.
struct GoldilocksRing {
    typedef Goldilocks F; typedef u64 W;      // W: word type of one limb in device planes
    static constexpr int ID = 0, D = 24, S = 8, TAU = 3, G = 24;
    static constexpr bool TRINOMIAL = true;   // X^24 = X^12 - 1
    static constexpr int CS_BYTES = 18;
};

struct BabyBearRing {      // Z_p[X]/(X^72 - X^36 + 1), 8 slots of Fq9   (crates/cyclotomic-rings/src/rings/babybear.rs:9-20)
    typedef BabyBear F; typedef u32 W;        // 31-bit prime: limb planes are packed 4-byte words (288 B per ring element)
    static constexpr int ID = 1, D = 72, S = 8, TAU = 9, G = 24;
    static constexpr bool TRINOMIAL = true;
    static constexpr int CS_BYTES = 18;
};
struct FrogRing {          // Z_p[X]/(X^16 + 1), 4 slots of Fq4           (crates/cyclotomic-rings/src/rings/frog.rs:9-20)
    typedef FrogField F; typedef u64 W;
    static constexpr int ID = 2, D = 16, S = 4, TAU = 4, G = 8;
    static constexpr bool TRINOMIAL = false;
    static constexpr int CS_BYTES = 16;
};

// c = a * b in the slot field Fq[Y]/(Y^TAU - nu) (fully reduced operands and result).  Generic form for any TAU;
// the Goldilocks TAU = 3 specialisation below keeps everything in registers with lazily reduced 192-bit sums.
template <class Rg> struct SlotField {
    typedef typename Rg::F F;
    static constexpr int TAU = Rg::TAU;
    struct Prepped { u64 b[Rg::TAU], bn[Rg::TAU]; };      // b and nu * b
    static LF_HD_CALL Prepped prep(const u64* b) { Prepped p;
#pragma unroll
        for (int i = 0; i < TAU; ++i) { p.b[i] = b[i]; p.bn[i] = F::mul_nu(b[i]); } return p; }
    // acc[k] += sum_{i+j=k} a_i b_j + nu sum_{i+j=k+TAU} a_i b_j
    static LF_HD_CALL void mac(typename F::Acc* acc, const u64* a, const Prepped& p) {
        for (int k = 0; k < TAU; ++k)
            for (int i = 0; i < TAU; ++i) { if (i <= k) acc[k].mac(a[i], p.b[k - i]); else acc[k].mac(a[i], p.bn[k + TAU - i]); }
    }
    static LF_HD_CALL void mul(u64* c, const u64* a, const u64* b) {
        const Prepped p = prep(b); typename F::Acc acc[Rg::TAU];
#pragma unroll
        for (int k = 0; k < TAU; ++k) acc[k].clear();
        mac(acc, a, p);
#pragma unroll
        for (int k = 0; k < TAU; ++k) c[k] = F::reduce(acc[k]);
    }
    static LF_HD_CALL void mul_prepped(u64* c, const u64* a, const Prepped& p) {
        typename F::Acc acc[Rg::TAU];
#pragma unroll
        for (int k = 0; k < TAU; ++k) acc[k].clear();
        mac(acc, a, p);
#pragma unroll
        for (int k = 0; k < TAU; ++k) c[k] = F::reduce(acc[k]);
    }
    static LF_HD_CALL void sqr(u64* c, const u64* a) { u64 t[Rg::TAU];
#pragma unroll
        for (int k = 0; k < TAU; ++k) t[k] = a[k]; mul(c, a, t); }
    static LF_HD void mul_inl(u64* c, const u64* a, const u64* b) { mul(c, a, b); }
    static LF_HD void add(u64* c, const u64* a, const u64* b) {
#pragma unroll
        for (int i = 0; i < TAU; ++i) c[i] = F::add(a[i], b[i]); }
    static LF_HD void sub(u64* c, const u64* a, const u64* b) {
#pragma unroll
        for (int i = 0; i < TAU; ++i) c[i] = F::sub(a[i], b[i]); }
    // accumulating dot products (k_dot): NDOT accumulators per output, one fixed operand prepared once per x
    static constexpr int NDOT = Rg::TAU;
    typedef Prepped DotPrepped;
    static LF_HD DotPrepped dot_prep(const u64* x) { return prep(x); }
    static LF_HD void dot_mac(typename F::Acc* acc, const u64* y, const DotPrepped& x) { mac(acc, y, x); }
    static LF_HD void dot_finish(u64* c, const typename F::Acc* acc) {
#pragma unroll
        for (int i = 0; i < TAU; ++i) c[i] = F::reduce(acc[i]); }
};

template <> struct SlotField<GoldilocksRing> {
    typedef Goldilocks F;
    static constexpr int TAU = 3;
    static LF_HD void mul(u64* c, const u64* a, const u64* b) {
        u64 b1n = F::mul_nu(b[1]), b2n = F::mul_nu(b[2]);
        Acc192 x; u64 c0, c1, c2;
        x.clear(); x.mac(a[0], b[0]); x.mac(a[1], b2n);  x.mac(a[2], b1n);  c0 = F::reduce192(x);
        x.clear(); x.mac(a[0], b[1]); x.mac(a[1], b[0]); x.mac(a[2], b2n);  c1 = F::reduce192(x);
        x.clear(); x.mac(a[0], b[2]); x.mac(a[1], b[1]); x.mac(a[2], b[0]); c2 = F::reduce192(x);
        c[0] = c0; c[1] = c1; c[2] = c2;
    }
    static LF_HD void mul_inl(u64* c, const u64* a, const u64* b) { mul(c, a, b); }
    static LF_HD void sqr(u64* c, const u64* a) {
        u64 a1n = F::mul_nu(a[1]), a2n = F::mul_nu(a[2]);
        u64 d1 = F::add(a[1], a[1]), d2 = F::add(a[2], a[2]);
        Acc192 x; u64 c0, c1, c2;
        x.clear(); x.mac(a[0], a[0]); x.mac(d1, a2n);                      c0 = F::reduce192(x);
        x.clear(); x.mac(a[0], d1);   x.mac(a[2], a2n);                    c1 = F::reduce192(x);
        x.clear(); x.mac(a[0], d2);   x.mac(a[1], a[1]);                   c2 = F::reduce192(x);
        c[0] = c0; c[1] = c1; c[2] = c2;
    }
    // lazy multiply-accumulate: acc[l] += (a*b)[l] with bn = (b0, nu*b1, nu*b2 precomputed by prep())
    struct Prepped { u64 b0, b1, b2, b1n, b2n; };
    static LF_HD Prepped prep(const u64* b) { Prepped p; p.b0 = b[0]; p.b1 = b[1]; p.b2 = b[2]; p.b1n = F::mul_nu(b[1]); p.b2n = F::mul_nu(b[2]); return p; }
    // c = a * b with b prepared: one accumulator live at a time
    static LF_HD void mul_prepped(u64* c, const u64* a, const Prepped& p) {
        Acc192 x; u64 c0, c1, c2;
        x.clear(); x.mac(a[0], p.b0); x.mac(a[1], p.b2n); x.mac(a[2], p.b1n); c0 = F::reduce192(x);
        x.clear(); x.mac(a[0], p.b1); x.mac(a[1], p.b0);  x.mac(a[2], p.b2n); c1 = F::reduce192(x);
        x.clear(); x.mac(a[0], p.b2); x.mac(a[1], p.b1);  x.mac(a[2], p.b0);  c2 = F::reduce192(x);
        c[0] = c0; c[1] = c1; c[2] = c2;
    }
    static LF_HD void mac(Acc192* acc, const u64* a, const Prepped& p) {
        acc[0].mac(a[0], p.b0); acc[0].mac(a[1], p.b2n); acc[0].mac(a[2], p.b1n);
        acc[1].mac(a[0], p.b1); acc[1].mac(a[1], p.b0);  acc[1].mac(a[2], p.b2n);
        acc[2].mac(a[0], p.b2); acc[2].mac(a[1], p.b1);  acc[2].mac(a[2], p.b0);
    }
    static LF_HD void add(u64* c, const u64* a, const u64* b) { for (int i = 0; i < 3; ++i) c[i] = F::add(a[i], b[i]); }
    static LF_HD void sub(u64* c, const u64* a, const u64* b) { for (int i = 0; i < 3; ++i) c[i] = F::sub(a[i], b[i]); }
    // Accumulating dot products (k_dot): three lazily reduced sums per output, fixed operand prepared once per x.
    // (A Karatsuba form with six sums -- 6 instead of 9 wide multiplies per slot-field product -- was measured on B200 and
    // lost 2.7x: the three extra modular additions per product and the doubled accumulator registers cost more than the
    // saved IMAD.WIDE issue slots; see tools/dot_ab.py.)
    static constexpr int NDOT = 3;
    typedef Prepped DotPrepped;
    static LF_HD DotPrepped dot_prep(const u64* x) { return prep(x); }
    static LF_HD void dot_mac(Acc192* acc, const u64* y, const DotPrepped& x) { mac(acc, y, x); }
    static LF_HD void dot_finish(u64* c, const Acc192* acc) { for (int i = 0; i < 3; ++i) c[i] = F::reduce192(acc[i]); }
};

// BabyBear slot field on BALANCED representatives (|a| <= (p-1)/2, signed 32-bit words).  (p/2)^2 < 2^59.82, so a sum of NINE
// products fits a signed 64-bit accumulator (9 (p-1)^2 / 4 = 0.9888 * 2^63): one multiply-accumulate is ONE IMAD.WIDE with no carry
// word, where the unsigned 96-bit accumulator of BabyBear::Acc needs a second instruction per product.  A product in Fq[Y]/(Y^9 - nu)
// is then 81 + 8 multiply-accumulates and 17 reductions: the 8 wrapped coefficient sums H_k = sum_{i>k} a_i b_{k+9-i} are reduced
// once and enter c_k = L_k + nu H_k as one more product; a fixed right operand (fix_variables: the round's challenge) carries its
// nu-multiples with it and needs only the 9 final reductions.  Exact integer arithmetic: any representative of the right residue
// class gives the same canonical limb, so results are bit-identical to BabyBear::mul chains.
// Reduction of |x| < 2^63: fold the upper word with 2^32 = 2^28 - 2 (mod p) -> |y| < 2^59 + 2^32; quotient estimate
// q = floor(((y >> 29) + 2) * round(2^61 / p) / 2^32), off from y / p by (-0.859, +0.659]; remainder in 32-bit wrap-around
// arithmetic, then one conditional +-p on either side.
struct BbBal {
    static constexpr int P = 2013265921, HALF = (P - 1) / 2;
    static constexpr int NUB = (int)(1398021245LL - 2013265921LL);                              // nu as a balanced representative
    static constexpr long long C32 = (1LL << 32) % P;
    static constexpr int M29 = (int)(((1LL << 61) + P / 2) / P);
    static LF_HD int bal(u32 a) { return (int)a - ((int)a > HALF ? P : 0); }                     // canonical [0, p) -> balanced
    static LF_HD u32 canon(int r) { return (u32)(r + (r < 0 ? P : 0)); }                         // (-p, p) -> canonical
    static LF_HD int fix(int r) { if (r > HALF) r -= P; if (r < -HALF) r += P; return r; }       // (-3p/2, 3p/2) -> balanced
    // signed 32 x 32 + 64 multiply-add: one IMAD.WIDE accumulating in place.  Written as a mad.lo.cc / madc.hi pair, which ptxas
    // fuses into one chained IMAD.WIDE; `mad.wide.s32` and the C expression are both re-associated by ptxas into independent wide
    // multiplies plus a tree of 64-bit three-input adds (two more instructions per product; measured on sm_100a, CUDA 12.9).
    static LF_HD long long madw(int a, int b, long long c) {
#if defined(__CUDA_ARCH__)
        int lo = (int)c, hi = (int)(c >> 32);
        asm("mad.lo.cc.s32 %0, %2, %3, %0;\n\tmadc.hi.s32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
        return (long long)(((u64)(u32)hi << 32) | (u32)lo);
#else
        return (long long)a * b + c;
#endif
    }
    static LF_HD int mulhi(int a, int b) {
#if defined(__CUDA_ARCH__)
        return __mulhi(a, b);
#else
        return (int)(((long long)a * b) >> 32);
#endif
    }
    // q = floor(((y >> 29) + 2) * M29 / 2^32): 2 * M29 / 2^32 = 0.533 is the rounding offset, so q - y / p lies in (-0.859, +0.659]
    // and the 32-bit remainder in [-0.659 p, 0.859 p)
    static LF_HD int red_small(long long y) {                                                     // |y| < 2^59 + 2^33
        const int q = mulhi((int)(y >> 29) + 2, M29);
        return fix((int)((u32)y - (u32)q * (u32)P));
    }
    static LF_HD int red(long long x) { return red_small(madw((int)(x >> 32), (int)C32, (long long)(u64)(u32)x)); }
    // c = a * b (all balanced, 9 limbs).  c may alias a or b.
    static LF_HD void mul(int* c, const int* a, const int* b) {
        int h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { long long x = 0;
#pragma unroll
            for (int i = k + 1; i < 9; ++i) x = madw(a[i], b[k + 9 - i], x);
            h[k] = red(x); }
        int r[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { long long x = 0;
            if (k < 8) x = madw(h[k < 8 ? k : 0], NUB, x);
#pragma unroll
            for (int i = 0; i <= k; ++i) x = madw(a[i], b[k - i], x);
            r[k] = red(x); }
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = r[k];
    }
    // fixed right operand: b and nu * b (balanced)
    struct Fixed { int b[9], bn[9]; };
    static LF_HD Fixed fixed(const u64* b_canonical) { Fixed f;
#pragma unroll
        for (int i = 0; i < 9; ++i) { f.b[i] = bal((u32)b_canonical[i]); f.bn[i] = red((long long)f.b[i] * NUB); } return f; }
    // c = add + a * b: 81 multiply-accumulates, 9 reductions (add, a balanced)
    static LF_HD void mul_fixed_add(int* c, const int* a, const Fixed& f, const int* add) {
        int r[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { long long x = add[k];
#pragma unroll
            for (int i = 0; i < 9; ++i) x = madw(a[i], i <= k ? f.b[k - i] : f.bn[k + 9 - i], x);
            r[k] = red(x); }
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = r[k];
    }
};

// BabyBear ring: slot field Fq9 = Fq[Y]/(Y^9 - nu).  Same interface as the generic SlotField, but the 81 partial products of a
// multiplication are straight-line lazily reduced MACs (one IMAD.WIDE + one carry add each, 9 per output limb, one 96-bit accumulator
// live at a time) with the fixed operand's nu-multiples prepared once, and the accumulating dot products (k_dot) keep the 17
// coefficients of the unreduced product so that their inner loop has no multiplication by nu at all.  The functions stay real calls
// (LF_HD_CALL): the generic sumcheck kernel instantiates them at many call sites.
template <> struct SlotField<BabyBearRing> {
    typedef BabyBear F;
    static constexpr int TAU = 9;
    // element arrays may be u64 or packed u32 words (template parameters below): values are < 2^31 either way
    struct Prepped { u32 b[9], bn[9]; };      // b and nu * b (bn[0] unused)
    template <class TB> static LF_HD Prepped prep(const TB* b) { Prepped p;
#pragma unroll
        for (int i = 0; i < 9; ++i) { p.b[i] = (u32)b[i]; p.bn[i] = (u32)F::mul_nu((u64)b[i]); } return p; }
    template <class TA> static LF_HD void mac(F::Acc* acc, const TA* a, const Prepped& p) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[k].mac((u64)a[i], i <= k ? (u64)p.b[k - i] : (u64)p.bn[k + 9 - i]);
    }
    template <class TC, class TA> static LF_HD void mul_prepped_inl(TC* c, const TA* a, const Prepped& p) {
        u32 r[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { F::Acc x; x.clear();
#pragma unroll
            for (int i = 0; i < 9; ++i) x.mac((u64)a[i], i <= k ? (u64)p.b[k - i] : (u64)p.bn[k + 9 - i]);
            r[k] = (u32)F::reduce(x); }
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = (TC)r[k];
    }
    // general products: on the device through the balanced-representative form (BbBal: one IMAD.WIDE per multiply-accumulate,
    // 54 instructions of conversion around ~330 of product, against ~600 for the unsigned 96-bit accumulators below)
    template <class TC, class TA, class TB> static LF_HD void mul_inl(TC* c, const TA* a, const TB* b) {
#if defined(__CUDA_ARCH__)
        int x[9], y[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { x[i] = BbBal::bal((u32)a[i]); y[i] = BbBal::bal((u32)b[i]); }
        BbBal::mul(x, x, y);
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = (TC)BbBal::canon(x[i]);
#else
        const Prepped p = prep(b); mul_prepped_inl(c, a, p);
#endif
    }
    static LF_HD_CALL void mul_prepped(u64* c, const u64* a, const Prepped& p) { mul_prepped_inl(c, a, p); }
    static LF_HD_CALL void mul(u64* c, const u64* a, const u64* b) { mul_inl(c, a, b); }
    static LF_HD_CALL void sqr(u64* c, const u64* a) { mul_inl(c, a, a); }
    template <class TC, class TA, class TB> static LF_HD void add(TC* c, const TA* a, const TB* b) {
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = (TC)F::add((u64)a[i], (u64)b[i]); }
    template <class TC, class TA, class TB> static LF_HD void sub(TC* c, const TA* a, const TB* b) {
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = (TC)F::sub((u64)a[i], (u64)b[i]); }
    // accumulating dot products: the 17 coefficients of the unreduced product, folded with nu once at the end
    static constexpr int NDOT = 17;
    struct DotPrepped { u32 x[9]; };
    static LF_HD DotPrepped dot_prep(const u64* x) { DotPrepped p;
#pragma unroll
        for (int i = 0; i < 9; ++i) p.x[i] = (u32)x[i]; return p; }
    static LF_HD void dot_mac(F::Acc* acc, const u64* y, const DotPrepped& x) {
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i + j].mac(y[i], (u64)x.x[j]);
    }
    static LF_HD void dot_finish(u64* c, F::Acc* acc) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const u64 hi = F::reduce(acc[k + 9]); acc[k].mac(hi, F::NU); }
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = F::reduce(acc[k]);
    }
};

}  // namespace lf
