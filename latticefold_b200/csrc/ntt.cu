// K12 kernels + plan + C ABI for the batched negacyclic NTT (definition and design notes: ntt.cuh; ABI: include/lf_b200.h).
#include "ntt.cuh"
#include "engine.cuh"
#include <cstring>

using namespace lf;
using namespace lf::ntt;

// omega_16^e for BabyBear, Montgomery form, e = 0..15 (filled at first plan creation)
__constant__ u32 c_bb_w16[16];
__device__ __forceinline__ u32 lf::ntt::BbF::mul_w16(u32 a, int e) { e &= 15; return e == 0 ? a : mul_tw(a, c_bb_w16[e]); }

namespace {

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
// shared-memory index of logical word a: one pad word in 16, so that the radix-16 scatter of pass 1 (16 consecutive words per
// thread) hits distinct banks.  (An XOR swizzle inside groups of 32 words was measured for the 4-byte field and lost 8 %.)
template <class T> __device__ __forceinline__ int sidx(int a) { return a + (a >> 4); }

// radix-R DFT (decimation in frequency) of the R registers x[g + j*G], j < R; natural order in and out.
// root = omega_16^(16/R) forward, its inverse for INV.  Everything is unrolled: exponents and register indices are constants.
template <class F, int R, int G, bool INV> __device__ __forceinline__ void dft_regs(typename F::T* x, int g) {
    typedef typename F::T T;
#pragma unroll
    for (int len = R / 2; len >= 1; len >>= 1) {
#pragma unroll
        for (int b0 = 0; b0 < R; b0 += 2 * len) {
#pragma unroll
            for (int i = 0; i < len; ++i) {
                T u = x[g + (b0 + i) * G], v = x[g + (b0 + i + len) * G];
                x[g + (b0 + i) * G] = F::add(u, v);
                int e = i * (R / (2 * len)) * (16 / R);               // exponent of omega_16
                if (INV) e = (16 - e) & 15;
                x[g + (b0 + i + len) * G] = F::mul_w16(F::sub(u, v), e);
            }
        }
    }
    if (R > 2) {                                                       // bit-reversal back to natural order (register renaming)
        T y[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { int r = 0;
#pragma unroll
            for (int bit = 1, rb = R >> 1; bit < R; bit <<= 1, rb >>= 1) if (k & bit) r |= rb;
            y[k] = x[g + r * G]; }
#pragma unroll
        for (int k = 0; k < R; ++k) x[g + k * G] = y[k];
    }
}

// One Stockham pass over the 16 registers of a thread.  Butterfly b = J * LPREV + K (K < LPREV) reads words j*(N/R) + b,
// multiplies input j by tw[(j-1)*LPREV + K] (forward: psi^((N/L)(2K+1)j); inverse: omega^(-(N/L)Kj), absent in the first pass),
// and writes output k to J*L + k*LPREV + K.
template <class F, int N, int R, int LPREV, bool INV, bool LAST>
__device__ __forceinline__ void pass(typename F::T* x, int t, typename F::T* s, const typename F::T* __restrict__ tw,
                                     const typename F::T* __restrict__ post, typename F::T* __restrict__ out, size_t ostride, bool valid) {
    typedef typename F::T T;
    constexpr int G = 16 / R, TP = N / 16, L = LPREV * R;
    if (LPREV > 1) {
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = s[sidx<T>(q * TP + t)];
    }
    if (!(INV && LPREV == 1)) {
#pragma unroll
        for (int q = G; q < 16; ++q) { int j = q / G, g = q % G, K = (t + g * TP) & (LPREV - 1); x[q] = F::mul_tw(x[q], tw[(j - 1) * LPREV + K]); }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) dft_regs<F, R, G, INV>(x, g);
    if (LAST) {
        if (valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) { int k = q / G, g = q % G, idx = k * (N / R) + t + g * TP;
                T v = x[q]; if (INV) v = F::mul_tw(v, post[idx]); out[(size_t)idx * ostride] = v; }
        }
    } else {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) { int k = q / G, g = q % G, b = t + g * TP, J = b / LPREV, K = b & (LPREV - 1); s[sidx<T>(J * L + k * LPREV + K)] = x[q]; }
        __syncthreads();
    }
}

// rows = batch * stride sub-polynomials of N points; row r = (poly, lo) with poly = r / stride, lo = r % stride.
//   forward: row input  = in[poly*N*stride + lo + stride*j]   (stride 1: contiguous, staged by one TMA bulk copy)
//            row output = out[r*N + k]
//   inverse: row input  = in[r*N + k]                          (always contiguous: TMA)
//            row output = out[poly*N*stride + lo + stride*j]
// tw: per-pass twiddle tables back to back; post: inverse only, scale * psi^-k.
// STRIDED (four-step rows, stride > 1): four adjacent rows per CTA with the row index in the low two bits of the thread id, so
// that a warp's strided accesses fall on whole 32-byte sectors (4 x 8 B) instead of one word per sector.
template <class F, int LOGN, bool INV, bool STRIDED>
__global__ void __launch_bounds__(Geo<LOGN, STRIDED>::THREADS, Geo<LOGN, STRIDED>::MINB)
k_ntt_cta(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, const typename F::T* __restrict__ tw,
          const typename F::T* __restrict__ post, size_t rows, int stride) {
    typedef typename F::T T; typedef Geo<LOGN, STRIDED> Gm;
    constexpr int N = Gm::N, TP = Gm::TP, PP = Gm::PP, P = Gm::P, R1 = Gm::R1;
    constexpr bool TMA_IN = !STRIDED || INV;
    constexpr int RS = Gm::PADN + (STRIDED ? 32 / (int)sizeof(T) : 0);     // words between the rows of a CTA; the stagger spreads
                                                                           // the 4 interleaved rows of a warp over distinct banks
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, pl = STRIDED ? tid % PP : tid / TP, t = STRIDED ? tid / PP : tid % TP;
    const size_t row0 = (size_t)blockIdx.x * PP, row = row0 + pl;
    const bool valid = row < rows;
    T x[16];
    if (TMA_IN) {
        const int nrows = (int)(rows - row0 < (size_t)PP ? rows - row0 : (size_t)PP);
        const u32 bar = smem_u32(&mbar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)(nrows * N * sizeof(T))) : "memory");
            for (int r = 0; r < nrows; ++r)      // one bulk copy per row: rows land RS words apart (bank stagger, see Geo::RS)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + r * RS)), "l"(in + (row0 + r) * N), "r"((u32)(N * sizeof(T))), "r"(bar) : "memory");
        }
        __syncthreads();                       // the barrier is initialised before anyone polls it
        u32 done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
        if (valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = smem[pl * RS + q * TP + t];
        }
        __syncthreads();                       // raw (unpadded) image fully consumed before the padded layout overwrites it
    } else {
        const T* src = in + (row / stride) * (size_t)N * stride + (row % stride);
        if (valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = src[(size_t)(q * TP + t) * stride];
        }
    }
    if (!valid) {
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = 0;
    }
    T* s = smem + pl * RS;
    T* dst; size_t ostride;
    if (INV) { dst = out + (row / stride) * (size_t)N * stride + (row % stride); ostride = stride; }
    else { dst = out + row * N; ostride = 1; }
    constexpr int TW1 = (R1 - 1), TW2 = 15 * R1, TW3 = 15 * R1 * 16;      // table sizes of passes 1..3
    pass<F, N, R1, 1, INV, P == 1>(x, t, s, tw, post, dst, ostride, valid);
    if constexpr (P >= 2) pass<F, N, 16, R1, INV, P == 2>(x, t, s, tw + TW1, post, dst, ostride, valid);
    if constexpr (P >= 3) pass<F, N, 16, R1 * 16, INV, P == 3>(x, t, s, tw + TW1 + TW2, post, dst, ostride, valid);
    if constexpr (P >= 4) pass<F, N, 16, R1 * 256, INV, P == 4>(x, t, s, tw + TW1 + TW2 + TW3, post, dst, ostride, valid);
}

// four-step cross pass, N = R * N2: forward  out[k*N2 + c] = sum_j omega_R^(jk) ctw[(j-1)*N2 + c] * in[j*N2 + c]
//                                   inverse  out[j*N2 + c] = ctw[j*N2 + c] * sum_k omega_R^(-jk) in[k*N2 + c]   (ctw carries 1/R)
template <class F, int R, bool INV>
__global__ void __launch_bounds__(256) k_ntt_cross(const typename F::T* __restrict__ in, typename F::T* __restrict__ out,
                                                   const typename F::T* __restrict__ ctw, int n2, size_t total) {
    typedef typename F::T T;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    size_t poly = idx / n2; int c = (int)(idx % n2);
    const T* src = in + poly * (size_t)R * n2 + c; T* dst = out + poly * (size_t)R * n2 + c;
    T x[R];
#pragma unroll
    for (int j = 0; j < R; ++j) x[j] = src[(size_t)j * n2];
    if (!INV) {
#pragma unroll
        for (int j = 1; j < R; ++j) x[j] = F::mul_tw(x[j], ctw[(size_t)(j - 1) * n2 + c]);
    }
    dft_regs<F, R, 1, INV>(x, 0);
    if (INV) {
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = F::mul_tw(x[j], ctw[(size_t)j * n2 + c]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) dst[(size_t)j * n2] = x[j];
}

template <class F> __global__ void k_ntt_pointwise(const typename F::T* __restrict__ a, const typename F::T* __restrict__ b,
                                                   typename F::T* __restrict__ out, typename F::T rsq, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // Goldilocks: rsq = 1 (canonical product).  BabyBear: mul_tw(a, b) = a*b/2^32, corrected by rsq = Montgomery form of 2^32
    typename F::T v = F::mul_tw(a[i], b[i]);
    if (F::ID == 1) v = F::mul_tw(v, rsq);
    out[i] = v;
}

// ------------------------------------------------------------------------------------------------ host: roots and tables
template <class F> u64 hmul(u64 a, u64 b) { return (u64)(((u128)a * b) % F::P); }
template <class F> u64 hpow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = hmul<F>(r, a); a = hmul<F>(a, a); e >>= 1; } return r; }
template <class F> u64 rho() {      // the 2^A-th root all psi_N derive from (rule in ntt.cuh)
    u64 r0 = hpow<F>(F::GEN, (F::P - 1) >> F::TWO_ADICITY);
    if (F::ID != 0) return r0;
    for (u64 u = 1; u < 64; u += 2) { u64 r = hpow<F>(r0, u); if (hpow<F>(r, (u64)1 << (F::TWO_ADICITY - 5)) == 64) return r; }
    throw LfException(LF_ERR_UNSUPPORTED, "no 2-adic root with rho^(2^27) = 2^6");
}
template <class F> u64 psi_of(int log_n) { return hpow<F>(rho<F>(), (u64)1 << (F::TWO_ADICITY - 1 - log_n)); }

}  // namespace

struct lf_ntt_plan {
    int field = 0, log_n = 0, sub_log = 0, n1 = 1;       // n1 > 1: four-step with 2^sub_log-point sub-transforms
    void *tw_f = nullptr, *tw_i = nullptr, *post_i = nullptr, *ctw_f = nullptr, *ctw_i = nullptr, *scratch = nullptr;
    size_t scratch_polys = 0; u64 psi = 0;
};

namespace {

template <class F> typename F::T* upload_table(lf_ctx* c, const std::vector<u64>& v) {
    typedef typename F::T T; std::vector<T> h(v.size()); for (size_t i = 0; i < v.size(); ++i) h[i] = (T)F::to_tw(v[i]);
    T* d = nullptr; LF_CUDA(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    LF_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream)); LF_CUDA(cudaStreamSynchronize(c->stream));
    return d;
}

template <class F> void build_plan(lf_ctx* c, lf_ntt_plan* pl) {
    const int log_n = pl->log_n;
    if (log_n > 14) { pl->sub_log = 12; pl->n1 = 1 << (log_n - 12); } else { pl->sub_log = log_n; pl->n1 = 1; }
    const int ls = pl->sub_log; const u64 n = (u64)1 << ls, N = (u64)1 << log_n;
    const u64 psi = psi_of<F>(ls), psi_inv = hpow<F>(psi, F::P - 2);          // psi_{N2} = psi_N^{N1}
    pl->psi = psi_of<F>(log_n);
    std::vector<u64> pw(2 * n), pwi(2 * n); pw[0] = pwi[0] = 1;
    for (u64 i = 1; i < 2 * n; ++i) { pw[i] = hmul<F>(pw[i - 1], psi); pwi[i] = hmul<F>(pwi[i - 1], psi_inv); }
    const int P = (ls + 3) / 4, R1 = 1 << (ls - 4 * (P - 1));
    std::vector<u64> tf, ti;
    u64 lprev = 1;
    for (int i = 0; i < P; ++i) {
        u64 R = i == 0 ? R1 : 16, L = lprev * R;
        for (u64 j = 1; j < R; ++j) for (u64 K = 0; K < lprev; ++K) {
            tf.push_back(pw[((n / L) * (2 * K + 1) * j) % (2 * n)]);
            ti.push_back(pwi[(2 * (n / L) * K * j) % (2 * n)]);
        }
        lprev = L;
    }
    const u64 n_inv = hpow<F>(n % F::P, F::P - 2);
    std::vector<u64> post(n); for (u64 k = 0; k < n; ++k) post[k] = hmul<F>(n_inv, pwi[k]);
    pl->tw_f = upload_table<F>(c, tf); pl->tw_i = upload_table<F>(c, ti); pl->post_i = upload_table<F>(c, post);
    if (pl->n1 > 1) {
        const u64 R = pl->n1, PSI = pl->psi, PSI_inv = hpow<F>(PSI, F::P - 2), r_inv = hpow<F>(R, F::P - 2);
        std::vector<u64> cf((R - 1) * n), ci(R * n);
        for (u64 j = 0; j < R; ++j) {
            u64 bf = hpow<F>(PSI, j), bi = hpow<F>(PSI_inv, j), sf = hmul<F>(bf, bf), si = hmul<F>(bi, bi);   // psi^(j(2c+1)) = bf * sf^c
            u64 vf = bf, vi = hmul<F>(bi, r_inv);
            for (u64 cidx = 0; cidx < n; ++cidx) { if (j) cf[(j - 1) * n + cidx] = vf; ci[j * n + cidx] = vi; vf = hmul<F>(vf, sf); vi = hmul<F>(vi, si); }
        }
        pl->ctw_f = upload_table<F>(c, cf); pl->ctw_i = upload_table<F>(c, ci);
        (void)N;      // the row scratch is allocated by the first transform, sized to its batch (capped at 4 GiB per launch pair)
    }
    if (F::ID == 1) {
        u64 w16 = hpow<F>(psi_of<F>(4), 2); u32 tab[16]; u64 v = 1;
        for (int e = 0; e < 16; ++e) { tab[e] = (u32)F::to_tw(v); v = hmul<F>(v, w16); }
        LF_CUDA(cudaMemcpyToSymbol(c_bb_w16, tab, sizeof(tab)));
    }
}

template <class Fn> void launch(lf_ctx* c, const char* name, Fn&& fn) {
    if (c->profiling) {
        cudaEvent_t a, b; LF_CUDA(cudaEventCreate(&a)); LF_CUDA(cudaEventCreate(&b));
        LF_CUDA(cudaEventRecord(a, c->stream)); fn(); LF_CUDA(cudaEventRecord(b, c->stream));
        c->prof.push_back({name, a, b});
    } else fn();
    ++c->launches; LF_CUDA(cudaGetLastError());
}

template <class F, int LOGN, bool INV, bool STRIDED> void launch_cta(lf_ctx* c, const lf_ntt_plan* pl, const void* in, void* out, size_t rows, int stride) {
    typedef typename F::T T; typedef Geo<LOGN, STRIDED> Gm;
    auto kern = k_ntt_cta<F, LOGN, INV, STRIDED>;
    const size_t smem = (size_t)Gm::PP * (Gm::PADN + (STRIDED ? 32 / sizeof(T) : 0)) * sizeof(T);
    static bool attr_set[8] = {false};      // per device
    if (!attr_set[c->device & 7]) { LF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set[c->device & 7] = true; }
    const unsigned grid = (unsigned)((rows + Gm::PP - 1) / Gm::PP);
    launch(c, INV ? "k_ntt_cta_inv" : "k_ntt_cta_fwd", [&] {
        kern<<<grid, Gm::THREADS, smem, c->stream>>>((const T*)in, (T*)out, (const T*)(INV ? pl->tw_i : pl->tw_f), (const T*)pl->post_i, rows, stride);
    });
}

template <class F, bool INV> void dispatch_cta(lf_ctx* c, const lf_ntt_plan* pl, const void* in, void* out, size_t rows, int stride) {
    switch (pl->sub_log) {
#define LF_NTT_CASE(L) case L: launch_cta<F, L, INV, false>(c, pl, in, out, rows, stride); break;
        LF_NTT_CASE(8) LF_NTT_CASE(9) LF_NTT_CASE(10) LF_NTT_CASE(11) LF_NTT_CASE(12) LF_NTT_CASE(13) LF_NTT_CASE(14)
#undef LF_NTT_CASE
        default: throw LfException(LF_ERR_UNSUPPORTED, "NTT size outside 2^8 .. 2^16");
    }
}
// strided sub-transforms exist for the 4096-point rows of the four-step path only
template <class F, bool INV> void dispatch_sub(lf_ctx* c, const lf_ntt_plan* pl, const void* in, void* out, size_t rows, int stride) {
    launch_cta<F, 12, INV, true>(c, pl, in, out, rows, stride);
}

template <class F, int R, bool INV> void launch_cross(lf_ctx* c, const lf_ntt_plan* pl, const void* in, void* out, size_t polys) {
    typedef typename F::T T; const int n2 = 1 << pl->sub_log; const size_t total = polys * n2;
    launch(c, INV ? "k_ntt_cross_inv" : "k_ntt_cross_fwd", [&] {
        k_ntt_cross<F, R, INV><<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>((const T*)in, (T*)out, (const T*)(INV ? pl->ctw_i : pl->ctw_f), n2, total);
    });
}
template <class F, bool INV> void dispatch_cross(lf_ctx* c, const lf_ntt_plan* pl, const void* in, void* out, size_t polys) {
    switch (pl->n1) {
        case 8: launch_cross<F, 8, INV>(c, pl, in, out, polys); break;
        case 16: launch_cross<F, 16, INV>(c, pl, in, out, polys); break;
        default: throw LfException(LF_ERR_UNSUPPORTED, "four-step radix");
    }
}

template <class F, bool INV> void exec(lf_ctx* c, lf_ntt_plan* pl, const void* in, void* out, size_t batch) {
    typedef typename F::T T;
    if (!batch) return;
    if (pl->n1 == 1) { dispatch_cta<F, INV>(c, pl, in, out, batch, 1); return; }
    const size_t N = (size_t)1 << pl->log_n;
    // The sub-transform rows go through a scratch of the batch's size.  These kernels are bound by the integer pipes, not by HBM
    // (ncu: ALU pipe ~90 % busy at 25 % of the HBM peak), so a second trip through memory costs less than the launch tails of
    // L2-sized chunks did (measured: 86 launches per transform at 2 GiB, 11 % of peak).
    const size_t want = std::min(batch, std::max<size_t>(1, ((size_t)4 << 30) / (N * sizeof(T))));
    if (pl->scratch_polys < want) {
        LF_CUDA(cudaStreamSynchronize(c->stream)); if (pl->scratch) cudaFree(pl->scratch); pl->scratch = nullptr; pl->scratch_polys = 0;
        LF_CUDA(cudaMalloc(&pl->scratch, want * N * sizeof(T))); pl->scratch_polys = want;
    }
    for (size_t p0 = 0; p0 < batch; p0 += pl->scratch_polys) {
        size_t np = std::min(pl->scratch_polys, batch - p0);
        const T* ci = (const T*)in + p0 * N; T* co = (T*)out + p0 * N;
        if (!INV) { dispatch_sub<F, false>(c, pl, ci, pl->scratch, np * pl->n1, pl->n1); dispatch_cross<F, false>(c, pl, pl->scratch, co, np); }
        else { dispatch_cross<F, true>(c, pl, ci, pl->scratch, np); dispatch_sub<F, true>(c, pl, pl->scratch, co, np * pl->n1, pl->n1); }
    }
}

thread_local std::string g_ntt_err;
template <class Fn> lf_status guard(lf_ctx* ctx, Fn&& fn) {
    try { fn(); return LF_OK; }
    catch (const LfException& e) { if (ctx) ctx->err = e.what(); else g_ntt_err = e.what(); return e.code; }
    catch (const std::exception& e) { if (ctx) ctx->err = e.what(); else g_ntt_err = e.what(); return LF_ERR_INVALID_ARG; }
}
size_t esize(int field) { return field == LF_FIELD_GOLDILOCKS ? 8 : 4; }
void check_plan(const lf_ctx* c, const lf_ntt_plan* pl) { if (!c || !pl) throw LfException(LF_ERR_INVALID_ARG, "null context or plan"); }

void run(lf_ctx* c, const lf_ntt_plan* cpl, bool inv, const void* in, void* out, size_t batch) {
    LF_CUDA(cudaSetDevice(c->device));
    lf_ntt_plan* pl = const_cast<lf_ntt_plan*>(cpl);      // the scratch grows on demand; a plan is used from one context at a time
    if (pl->field == LF_FIELD_GOLDILOCKS) { if (inv) exec<GlF, true>(c, pl, in, out, batch); else exec<GlF, false>(c, pl, in, out, batch); }
    else { if (inv) exec<BbF, true>(c, pl, in, out, batch); else exec<BbF, false>(c, pl, in, out, batch); }
}
void run_host(lf_ctx* c, const lf_ntt_plan* pl, bool inv, const void* h_in, void* h_out, size_t batch) {
    check_plan(c, pl);
    size_t bytes = batch * ((size_t)esize(pl->field) << pl->log_n); if (!bytes) return;
    LF_CUDA(cudaSetDevice(c->device));
    void* d = nullptr; LF_CUDA(cudaMalloc(&d, bytes));
    try {
        LF_CUDA(cudaMemcpyAsync(d, h_in, bytes, cudaMemcpyHostToDevice, c->stream));
        run(c, pl, inv, d, d, batch);
        LF_CUDA(cudaMemcpyAsync(h_out, d, bytes, cudaMemcpyDeviceToHost, c->stream)); LF_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) { cudaFree(d); throw; }
    cudaFree(d);
}

}  // namespace

extern "C" {

lf_status lf_ntt_root(int32_t field, int32_t log_n, uint64_t* psi_out) {
    return guard(nullptr, [&] {
        if (log_n < 1 || log_n > 26 || (field == LF_FIELD_GOLDILOCKS && log_n > 31)) throw LfException(LF_ERR_UNSUPPORTED, "log_n out of range");
        if (field == LF_FIELD_GOLDILOCKS) *psi_out = psi_of<GlF>(log_n); else if (field == LF_FIELD_BABYBEAR) *psi_out = psi_of<BbF>(log_n);
        else throw LfException(LF_ERR_UNSUPPORTED, "unknown field id");
    });
}

lf_status lf_ntt_plan_create(lf_ctx* ctx, int32_t field, int32_t log_n, lf_ntt_plan** out) {
    *out = nullptr;
    return guard(ctx, [&] {
        if (!ctx) throw LfException(LF_ERR_INVALID_ARG, "null context");
        if (field != LF_FIELD_GOLDILOCKS && field != LF_FIELD_BABYBEAR) throw LfException(LF_ERR_UNSUPPORTED, "unknown field id");
        if (log_n < 8 || log_n > 16) throw LfException(LF_ERR_UNSUPPORTED, "NTT size outside 2^8 .. 2^16");
        LF_CUDA(cudaSetDevice(ctx->device));
        std::unique_ptr<lf_ntt_plan> pl(new lf_ntt_plan); pl->field = field; pl->log_n = log_n;
        if (field == LF_FIELD_GOLDILOCKS) build_plan<GlF>(ctx, pl.get()); else build_plan<BbF>(ctx, pl.get());
        *out = pl.release();
    });
}
void lf_ntt_plan_free(lf_ctx* ctx, lf_ntt_plan* pl) {
    if (!pl) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(pl->tw_f); cudaFree(pl->tw_i); cudaFree(pl->post_i); cudaFree(pl->ctw_f); cudaFree(pl->ctw_i); cudaFree(pl->scratch);
    delete pl;
}
lf_status lf_ntt_forward_device(lf_ctx* ctx, const lf_ntt_plan* pl, const void* d_in, void* d_out, size_t batch) {
    return guard(ctx, [&] { check_plan(ctx, pl); run(ctx, pl, false, d_in, d_out, batch); });
}
lf_status lf_ntt_inverse_device(lf_ctx* ctx, const lf_ntt_plan* pl, const void* d_in, void* d_out, size_t batch) {
    return guard(ctx, [&] { check_plan(ctx, pl); run(ctx, pl, true, d_in, d_out, batch); });
}
lf_status lf_ntt_forward_host(lf_ctx* ctx, const lf_ntt_plan* pl, const void* h_in, void* h_out, size_t batch) {
    return guard(ctx, [&] { run_host(ctx, pl, false, h_in, h_out, batch); });
}
lf_status lf_ntt_inverse_host(lf_ctx* ctx, const lf_ntt_plan* pl, const void* h_in, void* h_out, size_t batch) {
    return guard(ctx, [&] { run_host(ctx, pl, true, h_in, h_out, batch); });
}
lf_status lf_ntt_pointwise_mul_device(lf_ctx* ctx, const lf_ntt_plan* pl, const void* d_a, const void* d_b, void* d_out, size_t batch) {
    return guard(ctx, [&] {
        check_plan(ctx, pl); size_t total = batch << pl->log_n; if (!total) return;
        LF_CUDA(cudaSetDevice(ctx->device));
        launch(ctx, "k_ntt_pointwise", [&] {
            unsigned grid = (unsigned)((total + 255) / 256);
            if (pl->field == LF_FIELD_GOLDILOCKS) k_ntt_pointwise<GlF><<<grid, 256, 0, ctx->stream>>>((const u64*)d_a, (const u64*)d_b, (u64*)d_out, 1, total);
            else k_ntt_pointwise<BbF><<<grid, 256, 0, ctx->stream>>>((const u32*)d_a, (const u32*)d_b, (u32*)d_out, (u32)BbF::to_tw(BbF::to_tw(1)), total);
        });
    });
}
/* negacyclic product of two batches of coefficient vectors: INTT(NTT(a) . NTT(b)), host buffers */
lf_status lf_ntt_negacyclic_mul_host(lf_ctx* ctx, const lf_ntt_plan* pl, const void* h_a, const void* h_b, void* h_out, size_t batch) {
    return guard(ctx, [&] {
        check_plan(ctx, pl); size_t bytes = batch * ((size_t)esize(pl->field) << pl->log_n); if (!bytes) return;
        LF_CUDA(cudaSetDevice(ctx->device));
        void *da = nullptr, *db = nullptr; LF_CUDA(cudaMalloc(&da, bytes)); if (cudaMalloc(&db, bytes) != cudaSuccess) { cudaFree(da); throw LfException(LF_ERR_CUDA, "cudaMalloc"); }
        try {
            LF_CUDA(cudaMemcpyAsync(da, h_a, bytes, cudaMemcpyHostToDevice, ctx->stream)); LF_CUDA(cudaMemcpyAsync(db, h_b, bytes, cudaMemcpyHostToDevice, ctx->stream));
            run(ctx, pl, false, da, da, batch); run(ctx, pl, false, db, db, batch);
            lf_status rc = lf_ntt_pointwise_mul_device(ctx, pl, da, db, da, batch); if (rc != LF_OK) throw LfException(rc, ctx->err);
            run(ctx, pl, true, da, da, batch);
            LF_CUDA(cudaMemcpyAsync(h_out, da, bytes, cudaMemcpyDeviceToHost, ctx->stream)); LF_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) { cudaFree(da); cudaFree(db); throw; }
        cudaFree(da); cudaFree(db);
    });
}

}  // extern "C"
