// tools/ladder_lab2.cu -- round-2 ladder experiments: wave quantisation (CTA size x resident warps per SM, limited by a
// dynamic shared-memory request) and arithmetic variants selected at compile time:
//   -DC25519_LADDER_CSWAP      round-1 form: physical conditional swap of the slots (default now: select folded into the doubling)
//   -DC25519_ALU_SMALL_MUL     x19 / x38 on the ALU pipe instead of IMAD
// Build one binary per variant:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-D...] -o tools/ladder_lab2_<tag> tools/ladder_lab2.cu
// Prints the FNV-1a hash of the canonical results of a 4096-op batch (must be identical for every variant) and the time of
// the production-shaped kernel (projective result to a 96-byte scratch record) at n = 2^20 for each launch shape.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../curve25519_b200/csrc/x25519.cuh"
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<int T, bool FULL>
__global__ void __launch_bounds__(T) k_ladder(uint8_t* __restrict__ out, const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ sk32, size_t n)
{
    extern __shared__ u32 dyn[];                    // [8][T] scalar words (+ padding that limits residency)
    u32 (*ks)[T] = reinterpret_cast<u32 (*)[T]>(dyn);
    const size_t i = (size_t)blockIdx.x * T + threadIdx.x; if (i >= n) return;
    fe k; fe_load(k, sk32 + 32 * i); k.v[0] &= 0xfffffff8u; k.v[7] = (k.v[7] | 0x40000000u) & 0x7fffffffu;
#pragma unroll
    for (int w = 0; w < 8; w++) ks[w][threadIdx.x] = k.v[w];
    fe u; fe_load(u, pk32 + 32 * i); const int t = threadIdx.x;
    if (FULL) { fe r; x25519_ladder(r, u, [&](int w) { return ks[w][t]; }); fe_store(out + 32 * i, r); }
    else { fe PX, PZ; x25519_ladder_projective(PX, PZ, u, [&](int w) { return ks[w][t]; }); fe_store(out + 96 * i, PX); fe_store(out + 96 * i + 32, PZ); }
}

struct Ctx { uint8_t *out, *pk, *sk; size_t n; };

template<int T> int shape(Ctx& c, int warps_per_sm)
{
    auto kern = k_ladder<T, false>;
    const int ctas = warps_per_sm * 32 / T;
    if (ctas * T != warps_per_sm * 32 || ctas > 32) return 0;
    // request enough dynamic shared memory that exactly `ctas` CTAs fit in 227 KB (1 KB reserved per CTA by the runtime)
    size_t smem = (size_t)(227 * 1024) / ctas - 1024; smem &= ~(size_t)127; if (smem < sizeof(u32) * 8 * T) smem = sizeof(u32) * 8 * T;
    CHK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    unsigned grid = (unsigned)((c.n + T - 1) / T);
    kern<<<grid, T, smem>>>(c.out, c.pk, c.sk, c.n); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e9;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); kern<<<grid, T, smem>>>(c.out, c.pk, c.sk, c.n); cudaEventRecord(e1); CHK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double waves = (double)c.n / (148.0 * occ * T);
    printf("T=%3d want %2d warps/SM: occ=%2d CTAs (%2d warps/SM) regs=%3d spill=%zuB waves=%6.2f  %8.3f ms  %7.2f Mops/s\n",
           T, warps_per_sm, occ, occ * T / 32, fa.numRegs, (size_t)fa.localSizeBytes, waves, best, c.n / (best * 1e3));
    return 0;
}

int main(int argc, char** argv)
{
    Ctx c; c.n = (argc > 1) ? (size_t)atol(argv[1]) : (size_t)1 << 20;
    CHK(cudaMalloc(&c.out, 96 * c.n)); CHK(cudaMalloc(&c.pk, 32 * c.n)); CHK(cudaMalloc(&c.sk, 32 * c.n));
    std::vector<uint8_t> h(32 * c.n); uint64_t s = 0x9e3779b97f4a7c15ull;
    for (size_t i = 0; i < 32 * c.n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint8_t)(s >> 24); } cudaMemcpy(c.pk, h.data(), 32 * c.n, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < 32 * c.n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint8_t)(s >> 24); } cudaMemcpy(c.sk, h.data(), 32 * c.n, cudaMemcpyHostToDevice);
    {   // correctness fingerprint: canonical affine results of the first 4096 operations
        const size_t m = 4096; CHK(cudaFuncSetAttribute(k_ladder<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
        k_ladder<128, true><<<(unsigned)(m / 128), 128, 4096>>>(c.out, c.pk, c.sk, m); CHK(cudaDeviceSynchronize());
        std::vector<uint8_t> r(32 * m); cudaMemcpy(r.data(), c.out, 32 * m, cudaMemcpyDeviceToHost);
        unsigned long long hsh = 1469598103934665603ull; for (size_t i = 0; i < r.size(); i++) { hsh ^= r[i]; hsh *= 1099511628211ull; }
        printf("fingerprint(4096 ops) = %016llx\n", hsh);
    }
    for (int w : {16, 17, 18, 19, 20, 21, 22, 24}) shape<32>(c, w);
    for (int w : {16, 18, 20, 22}) shape<64>(c, w);
    for (int w : {16, 20}) shape<128>(c, w);
    return 0;
}
