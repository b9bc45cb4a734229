// tools/coop_lab.cu -- the north_star's kernel mapping, measured: one field element per 8-lane sub-group, ONE 32-bit limb per
// lane, operands broadcast with __shfl_sync, carries resolved across lanes -- against the shipped mapping (one element per
// thread, all limbs in registers, fe25519.cuh).  Both compute z <- z * y mod 2^255-19 in a dependent chain; results are
// compared (canonical form) before anything is timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I curve25519_b200/csrc -o tools/coop_lab tools/coop_lab.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "fe25519.cuh"
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
typedef unsigned long long ull;

// ---- limb-per-lane arithmetic: lane l (0..7 of its sub-group) holds limb l -------------------------------------------
__device__ __forceinline__ u32 grp_get(u32 v, int src, int base) { return __shfl_sync(0xffffffffu, v, base + src); }

// resolve carries of eight 64-bit column values V_l (weight 2^(32 l)), folding the carry out of limb 7 back with x38
__device__ __forceinline__ u32 coop_carry(ull V, int l, int base)
{
    // three parallel passes bring every column below 2^32 + small; then a short data-dependent tail (rare)
#pragma unroll 1
    for (int pass = 0; pass < 12; pass++) {
        const u32 lo = (u32)V, hi_lo = (u32)(V >> 32);
        u32 cin = __shfl_sync(0xffffffffu, hi_lo, base + ((l + 7) & 7));      // carry of the limb below (lane 7 -> lane 0 wraps)
        const ull add = (l == 0) ? (ull)cin * 38ull : (ull)cin;                // 2^256 = 38 mod p
        V = (ull)lo + add;
        const unsigned pending = __ballot_sync(0xffffffffu, (V >> 32) != 0);
        if (pending == 0) break;
    }
    return (u32)V;
}

// z = x * y mod (2^256 - 38), limb l of each in lane l
__device__ __forceinline__ u32 coop_mul(u32 x, u32 y, int l, int base)
{
    ull lo_l = 0, lo_h = 0, hi_l = 0, hi_h = 0;     // column l (low half) and column l + 8 (high half), 32-bit halves summed apart
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const u32 yj = grp_get(y, j, base);
        const u32 xr = grp_get(x, (l - j) & 7, base);                          // x_{l-j}: its product with y_j lands in column l (or l + 8)
        const ull p = (ull)xr * yj;
        if (l >= j) { lo_l += (u32)p; lo_h += p >> 32; } else { hi_l += (u32)p; hi_h += p >> 32; }
    }
    // word k of the 512-bit product = low halves of column k + high halves of column k - 1
    const ull lo_h_dn = __shfl_sync(0xffffffffu, lo_h, base + ((l + 7) & 7));  // from the lane below (lane 7's for lane 0)
    const ull hi_h_dn = __shfl_sync(0xffffffffu, hi_h, base + ((l + 7) & 7));
    const ull W_lo = lo_l + (l ? lo_h_dn : 0ull);                              // word l         (< 2^36)
    const ull W_hi = hi_l + (l ? hi_h_dn : lo_h_dn);                           // word l + 8; word 8 takes column 7's high halves
    const ull W_16 = (l == 0) ? hi_h_dn : 0ull;                                // word 16 = column 15's high halves (lane 7's hi_h)
    // fold: 2^256 = 38, 2^512 = 1444
    const ull V = W_lo + 38ull * W_hi + 1444ull * W_16;                        // < 2^36 + 2^42 + 2^47
    return coop_carry(V, l, base);
}

template <int COOP>
__global__ void __launch_bounds__(256) k_chain(u32* out, const u32* a, const u32* b, int iters, size_t n_elems, ull* cyc)
{
    ull t0 = 0, t1 = 0;
    if (COOP) {
        const size_t lane_global = (size_t)blockIdx.x * 256 + threadIdx.x;
        const size_t e = lane_global >> 3; const int l = threadIdx.x & 7; const int base = threadIdx.x & 24;
        const size_t ee = e < n_elems ? e : n_elems - 1;
        u32 z = a[8 * ee + l]; const u32 y = b[8 * ee + l];
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; it++) z = coop_mul(z, y, l, base);
        t1 = clock64();
        if (e < n_elems) out[8 * e + l] = z;
    } else {
        const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
        const size_t ee = e < n_elems ? e : n_elems - 1;
        fe z, y;
#pragma unroll
        for (int i = 0; i < 8; i++) { z.v[i] = a[8 * ee + i]; y.v[i] = b[8 * ee + i]; }
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; it++) fe_mul(z, z, y);
        t1 = clock64();
        if (e < n_elems)
#pragma unroll
            for (int i = 0; i < 8; i++) out[8 * e + i] = z.v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

static ull modp_canon_cmp(const u32* x, const u32* y)      // compare two loosely reduced values modulo p (host, test only)
{
    auto canon = [](const u32* v, unsigned __int128 out[2]) {
        // value mod p with p = 2^255 - 19: fold bit 255.. then conditional subtract (two 128-bit halves)
        unsigned __int128 lo = 0, hi = 0;
        for (int i = 3; i >= 0; i--) lo = (lo << 32) | v[i];
        for (int i = 7; i >= 4; i--) hi = (hi << 32) | v[i];
        for (int r = 0; r < 3; r++) {
            unsigned __int128 top = hi >> 127; hi &= (((unsigned __int128)1 << 127) - 1);
            unsigned __int128 add = top * 19; unsigned __int128 nlo = lo + add; if (nlo < lo) hi++; lo = nlo;
        }
        // if value >= p subtract p
        const unsigned __int128 plo = ~(unsigned __int128)0 - 18, phi = (((unsigned __int128)1 << 127) - 1);
        if (hi > phi || (hi == phi && lo >= plo)) { unsigned __int128 nlo = lo - plo; if (lo < plo) hi--; lo = nlo; hi -= phi; }
        out[0] = lo; out[1] = hi;
    };
    unsigned __int128 a[2], b[2]; canon(x, a); canon(y, b);
    return (a[0] != b[0]) || (a[1] != b[1]);
}

int main()
{
    const size_t n = (size_t)148 * 2048 * 4;          // elements
    u32 *a, *b, *o1, *o2; ull* cyc;
    CHK(cudaMalloc(&a, 32 * n)); CHK(cudaMalloc(&b, 32 * n)); CHK(cudaMalloc(&o1, 32 * n)); CHK(cudaMalloc(&o2, 32 * n)); CHK(cudaMalloc(&cyc, 64));
    std::vector<u32> h(8 * n); ull s = 0x9e3779b97f4a7c15ull;
    for (auto& w : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = (u32)(s >> 16); }
    for (int i = 0; i < 8; i++) { h[i] = 0xffffffffu; h[8 + i] = i == 7 ? 0x7fffffffu : (i ? 0xffffffffu : 0xffffffecu); }   // edge operands: 2^256-1, p-1
    CHK(cudaMemcpy(a, h.data(), 32 * n, cudaMemcpyHostToDevice));
    for (auto& w : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = (u32)(s >> 16); }
    for (int i = 0; i < 8; i++) h[i] = 0xffffffffu;
    CHK(cudaMemcpy(b, h.data(), 32 * n, cudaMemcpyHostToDevice));
    // ---- correctness: 37 chained multiplications, both mappings, compared modulo p
    const size_t m = 1 << 16;
    k_chain<0><<<(unsigned)((m + 255) / 256), 256>>>(o1, a, b, 37, m, cyc);
    k_chain<1><<<(unsigned)((m * 8 + 255) / 256), 256>>>(o2, a, b, 37, m, cyc);
    CHK(cudaDeviceSynchronize());
    std::vector<u32> r1(8 * m), r2(8 * m);
    cudaMemcpy(r1.data(), o1, 32 * m, cudaMemcpyDeviceToHost); cudaMemcpy(r2.data(), o2, 32 * m, cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t e = 0; e < m; e++) bad += modp_canon_cmp(&r1[8 * e], &r2[8 * e]);
    printf("correctness: %zu of %zu chained products differ between the two mappings\n", bad, m);
    // ---- throughput at saturation and latency of a lone warp
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int coop = 0; coop < 2; coop++) {
        const int iters = 2000;
        const size_t elems = coop ? n / 4 : n;         // keep the run short: the cooperative mapping needs 8 lanes per element
        const unsigned grid = (unsigned)(((coop ? elems * 8 : elems) + 255) / 256);
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (coop) k_chain<1><<<grid, 256>>>(o2, a, b, iters, elems, cyc); else k_chain<0><<<grid, 256>>>(o1, a, b, iters, elems, cyc);
            cudaEventRecord(e1); CHK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double muls = (double)elems * iters / (ms * 1e-3);
        if (coop) k_chain<1><<<1, 32>>>(o2, a, b, iters, 4, cyc); else k_chain<0><<<1, 32>>>(o1, a, b, iters, 32, cyc);
        CHK(cudaDeviceSynchronize()); ull c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        cudaFuncAttributes fa; if (coop) cudaFuncGetAttributes(&fa, k_chain<1>); else cudaFuncGetAttributes(&fa, k_chain<0>);
        printf("%-58s regs=%3d  %8.2f G fe_mul/s at saturation   lone-warp latency %6.1f cycles per fe_mul\n",
               coop ? "limb per lane, 8 lanes per element, __shfl_sync (north_star)" : "element per thread, limbs in registers (shipped)",
               fa.numRegs, muls / 1e9, (double)c / iters);
    }
    return 0;
}
