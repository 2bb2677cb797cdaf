// Device Poseidon (width 24, 8 + 22 rounds, x^7, Goldilocks) as a SERIAL chain of permutations -- the shape of the Fiat-Shamir sponge
// of the prover step (transcript/poseidon.rs:29-75): how fast can the GPU hash one transcript?  SURVEY 8(f) rank 1 asks for an
// on-device transcript; this measurement is why the transcript stays on the host.
//   variant W: one warp per sponge, lane i < 24 holds state[i], the dense MDS layer is 24 shuffles + 24 multiply-accumulates per lane
//   variant B: one block of 24 warps... not built: every round needs a block barrier (30 x ~50 ns) on top of the same arithmetic chain
//   variant T: one thread per sponge (all 24 lanes in registers): the throughput form, for many independent sponges
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../latticefold_b200/csrc -o poseidon_dev poseidon_dev.cu ; run: ./poseidon_dev
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "field.cuh"
#include "poseidon_w24_tables.inc"
using lf::u64; typedef lf::Goldilocks F;
constexpr int W = 24, RF = 8, RP = 22, NR = RF + RP;
__constant__ u64 c_ark[NR * W]; __constant__ u64 c_mds[W * W];
__device__ __forceinline__ u64 pow7(u64 x) { u64 x2 = F::mul(x, x), x4 = F::mul(x2, x2); return F::mul(F::mul(x4, x2), x); }
// one warp = one sponge; `chain` dependent permutations
__global__ void k_warp(u64* io, int chain) {
    const int lane = threadIdx.x & 31; u64 s = lane < W ? io[blockIdx.x * W + lane] : 0;
    u64 mrow[W];
#pragma unroll
    for (int j = 0; j < W; ++j) mrow[j] = lane < W ? c_mds[lane * W + j] : 0;
    for (int it = 0; it < chain; ++it)
        for (int r = 0; r < NR; ++r) {
            if (lane < W) s = F::add(s, c_ark[r * W + lane]);
            const bool full = r < RF / 2 || r >= RF / 2 + RP;
            if (full || lane == 0) s = pow7(s);
            lf::Acc192 acc; acc.clear();
#pragma unroll
            for (int j = 0; j < W; ++j) { const u64 v = __shfl_sync(0xffffffffu, s, j); acc.mac(mrow[j], v); }
            s = F::reduce192(acc);
        }
    if (lane < W) io[blockIdx.x * W + lane] = s;
}
// one thread = one sponge
__global__ void k_thread(u64* io, int chain, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
    u64 s[W];
    for (int i = 0; i < W; ++i) s[i] = io[(size_t)t * W + i];
    for (int it = 0; it < chain; ++it)
        for (int r = 0; r < NR; ++r) {
            for (int i = 0; i < W; ++i) s[i] = F::add(s[i], c_ark[r * W + i]);
            const bool full = r < RF / 2 || r >= RF / 2 + RP;
            if (full) { for (int i = 0; i < W; ++i) s[i] = pow7(s[i]); } else s[0] = pow7(s[0]);
            u64 o[W];
            for (int i = 0; i < W; ++i) { lf::Acc192 acc; acc.clear(); for (int j = 0; j < W; ++j) acc.mac(c_mds[i * W + j], s[j]); o[i] = F::reduce192(acc); }
            for (int i = 0; i < W; ++i) s[i] = o[i];
        }
    for (int i = 0; i < W; ++i) io[(size_t)t * W + i] = s[i];
}
int main() {
    std::vector<u64> ark(NR * W), mds(W * W);
    for (int i = 0; i < NR * W; ++i) ark[i] = POSEIDON_W24_ARK[i] % F::P;
    for (int i = 0; i < W * W; ++i) mds[i] = POSEIDON_W24_MDS[i] % F::P;
    cudaMemcpyToSymbol(c_ark, ark.data(), ark.size() * 8); cudaMemcpyToSymbol(c_mds, mds.data(), mds.size() * 8);
    const int chain = 2000; u64* d; cudaMalloc(&d, (size_t)148 * 1024 * W * 8); cudaMemset(d, 1, (size_t)148 * 1024 * W * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
    for (int sponges : {1, 148, 148 * 8}) {
        k_warp<<<sponges, 32>>>(d, 10); cudaDeviceSynchronize();
        cudaEventRecord(a); k_warp<<<sponges, 32>>>(d, chain); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("warp-per-sponge  : %5d sponge(s), %d chained permutations: %.3f us per permutation (serial latency), %.1f M perm/s aggregate\n", sponges, chain, ms * 1e3 / chain, sponges * (double)chain / ms / 1e3);
    }
    for (int sponges : {1, 148 * 128, 148 * 1024}) {
        k_thread<<<(sponges + 127) / 128, 128>>>(d, 2, sponges); cudaDeviceSynchronize();
        cudaEventRecord(a); k_thread<<<(sponges + 127) / 128, 128>>>(d, 50, sponges); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("thread-per-sponge: %6d sponge(s), 50 chained permutations: %.3f us per permutation (serial latency), %.1f M perm/s aggregate\n", sponges, ms * 1e3 / 50, sponges * 50.0 / ms / 1e3);
    }
    printf("host reference: 2.5-3.5 us per permutation on one Xeon core with AVX-512 IFMA (tools/poseidon_bench.py); one prover step hashes ~2100 permutations in ONE serial chain\n");
    return 0;
}
