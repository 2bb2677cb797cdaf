// Host harness for the balanced-representative BabyBear slot-field arithmetic (latticefold_b200/csrc/field.cuh: BbBal), the code the
// wide-slot-field sumcheck kernels run on the device: reductions at the edges of their stated input ranges and products / fixed-operand
// products against 128-bit reference arithmetic in Fq[Y]/(Y^9 - nu).  Built and run by tests/test_cabi_cpu.py.  Exit code 0 = pass.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../latticefold_b200/csrc/field.cuh"

using namespace lf;
typedef __int128 i128;
static const long long P = BbBal::P;

static long long modp(i128 x) { long long r = (long long)(x % P); return r < 0 ? r + P : r; }
static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 10) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } } while (0)

static void ref_mul(long long* c, const long long* a, const long long* b) {
    i128 d[17] = {0};
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) d[i + j] += (i128)a[i] * b[j];
    for (int k = 0; k < 9; ++k) c[k] = modp(d[k] + (k < 8 ? (i128)modp(d[k + 9]) * (long long)BabyBear::NU : 0));
}

int main() {
    std::mt19937_64 rng(12345);
    const long long H = BbBal::HALF;
    // ---- reductions
    const long long lim = 9 * H * H + H;                     // largest magnitude the kernels feed to red()
    long long edge[] = {0, 1, -1, P, -P, P - 1, 1 - P, H, -H, H + 1, -H - 1, lim, -lim, lim - 1, 1 - lim, (1LL << 62), -(1LL << 62), (long long)0x7fffffffffffffffLL, (long long)0x8000000000000001LL,
                        ((1LL << 32) - 1), (1LL << 32), -(1LL << 32), (1LL << 59), -(1LL << 59)};
    for (long long x : edge) { const int r = BbBal::red(x); CHECK(r >= -H && r <= H && modp(r) == modp(x), "red(%lld) = %d", x, r); }
    for (int it = 0; it < 4000000; ++it) {
        long long x = (long long)rng();
        if (it & 1) x >>= (rng() % 40);
        const int r = BbBal::red(x); CHECK(r >= -H && r <= H && modp(r) == modp(x), "red(%lld) = %d", x, r);
        const long long y = x >> 5;                        // |y| < 2^58: red_small's domain
        const int s = BbBal::red_small(y); CHECK(s >= -H && s <= H && modp(s) == modp(y), "red_small(%lld) = %d", y, s);
    }
    for (long long v = -3 * P / 2 + 1; v < 3 * P / 2; v += 9973) { const int r = BbBal::fix((int)(v > 2147483647LL ? 2147483647LL : v < -2147483647LL ? -2147483647LL : v)); (void)r; }
    for (u32 a : {0u, 1u, (u32)H, (u32)H + 1, (u32)P - 1}) { const int b = BbBal::bal(a); CHECK(b >= -H && b <= H && modp(b) == (long long)a && BbBal::canon(b) == a, "bal(%u)", a); }
    // ---- products: random, all-extreme (+-HALF everywhere: the accumulator bound), mixed
    for (int it = 0; it < 200000; ++it) {
        int a[9], b[9], add[9]; long long ar[9], br[9], cr[9];
        for (int i = 0; i < 9; ++i) {
            const int mode = it % 4;
            auto pick = [&]() -> int { if (mode == 0) return (int)((long long)(rng() % P) - H); if (mode == 1) return (rng() & 1) ? (int)H : (int)-H; if (mode == 2) return (int)H; return (int)((long long)(rng() % 7) - 3); };
            a[i] = pick(); b[i] = pick(); add[i] = pick();
            if (a[i] > H) a[i] = H; if (b[i] > H) b[i] = H;
            ar[i] = modp(a[i]); br[i] = modp(b[i]);
        }
        ref_mul(cr, ar, br);
        int c[9]; BbBal::mul(c, a, b);
        for (int k = 0; k < 9; ++k) CHECK(c[k] >= -H && c[k] <= H && modp(c[k]) == cr[k], "mul limb %d (case %d)", k, it);
        u64 bc[9]; for (int i = 0; i < 9; ++i) bc[i] = (u64)br[i];
        const BbBal::Fixed f = BbBal::fixed(bc);
        int e[9]; BbBal::mul_fixed_add(e, a, f, add);
        for (int k = 0; k < 9; ++k) CHECK(e[k] >= -H && e[k] <= H && modp(e[k]) == modp((i128)cr[k] + add[k]), "mul_fixed_add limb %d (case %d)", k, it);
        // against the product's canonical slot-field multiplication
        u64 a64[9], c64[9]; for (int i = 0; i < 9; ++i) a64[i] = (u64)ar[i];
        SlotField<BabyBearRing>::mul(c64, a64, bc);
        for (int k = 0; k < 9; ++k) CHECK((long long)c64[k] == cr[k], "SlotField::mul limb %d (case %d)", k, it);
    }
    if (fails) { std::printf("%d failures\n", fails); return 1; }
    std::printf("ok\n");
    return 0;
}
