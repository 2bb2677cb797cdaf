// Integer-pipe ceilings on B200 for the kernels of this repo (DESIGN.md "what bounds the hot kernels"):
//   (1) IMAD.WIDE.U32 issue rate with independent accumulator chains,
//   (2) the lazily reduced 64x64 multiply-accumulate (Acc192::mac) rate, registers only,
//   (3) full Goldilocks mulmod (mul + reduce128) rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../latticefold_b200/csrc -o imad_peak imad_peak.cu
#include "field.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace lf;

template <int NACC> __global__ void k_mac(u64* out, u64 seed, int iters) {
    Acc192 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i].clear();
    u64 a = seed + threadIdx.x * 0x9E3779B97F4A7C15ULL, b = seed ^ (blockIdx.x * 0xD1342543DE82EF95ULL + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i].mac(a + i, b);
        a = a * 3 + 1; b ^= a;
    }
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r ^= Goldilocks::reduce192(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NCH> __global__ void k_wide(u64* out, u32 seed, int iters) {
    u64 acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = i;
    u32 a = seed + threadIdx.x, b = seed * 7 + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"(b));
        a += 3; b ^= a;
    }
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NCH> __global__ void k_mulmod(u64* out, u64 seed, int iters) {
    u64 x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = seed + i + threadIdx.x;
    u64 m = seed | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = Goldilocks::mul(x[i], m);
    }
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) r ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64* out; cudaMalloc(&out, 8ull * sms * 8 * 1024);
    const int iters = 4096;
    for (int bpsm : {1, 2, 4, 8}) for (int threads : {128, 256}) {
        const int blocks = sms * bpsm; const double thr = (double)blocks * threads;
        float ms = time_it([&] { k_wide<8><<<blocks, threads>>>(out, 12345u, iters); });
        double wide = thr * iters * 8 / (ms * 1e-3);
        ms = time_it([&] { k_mac<6><<<blocks, threads>>>(out, 12345ull, iters); });
        double mac = thr * iters * 6 / (ms * 1e-3);
        ms = time_it([&] { k_mulmod<6><<<blocks, threads>>>(out, 12345ull, iters); });
        double mm = thr * iters * 6 / (ms * 1e-3);
        printf("blocks/SM %d threads %d : IMAD.WIDE %.3e /s (%.1f lanes/clk/SM @1.9GHz) | Acc192::mac %.3e /s | mulmod %.3e /s\n", bpsm, threads, wide, wide / sms / 1.9e9, mac, mm);
    }
    return 0;
}
