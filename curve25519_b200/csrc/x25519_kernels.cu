// x25519_kernels.cu -- batched X25519 kernels (sm_100a).
//
//   k_x25519_ladder : curve25519_dh_CreateSharedKey (curve25519_dh.c:201) and the ladder flavour of
//                     curve25519_dh_CalculatePublicKey (:192); one operation per thread.
//
// Data layout in HBM: three arrays of 32-byte records (scalars in/out, peer points in, results out),
// record i belongs to thread i.  Each thread moves its record with two 16-byte vector accesses; a warp
// therefore touches one contiguous, 128-byte-aligned 1 KB span per array.  Algorithmic HBM traffic is
// 128 B per operation (32 sk in, 32 sk clamped out, 32 pk in, 32 out) against ~150 k integer
// multiply-adds, i.e. the kernel is bound by the integer pipes, not by HBM.
#include "kernels.h"
#include "x25519.cuh"

namespace c25519 {

constexpr int kLadderThreads = 128;

__global__ void __launch_bounds__(kLadderThreads)
k_x25519_ladder(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, uint8_t* __restrict__ sk32, size_t n)
{
    __shared__ u32 ks[8][kLadderThreads];          // scalar words, one column per thread (conflict-free)
    const size_t i = (size_t)blockIdx.x * kLadderThreads + threadIdx.x;
    if (i >= n) return;
    fe k;
    {   // ecp_TrimSecretKey (curve25519_utils.c:28-32), written back in place like the reference does
        uint4* q = reinterpret_cast<uint4*>(sk32 + 32 * i);
        uint4 a = q[0], b = q[1];
        a.x &= 0xfffffff8u;
        b.w = (b.w | 0x40000000u) & 0x7fffffffu;
        q[0] = a; q[1] = b;
        k.v[0] = a.x; k.v[1] = a.y; k.v[2] = a.z; k.v[3] = a.w; k.v[4] = b.x; k.v[5] = b.y; k.v[6] = b.z; k.v[7] = b.w;
    }
#pragma unroll
    for (int w = 0; w < 8; w++) ks[w][threadIdx.x] = k.v[w];
    fe u;
    if (pk32) fe_load(u, pk32 + 32 * i);           // all 256 bits (curve25519_dh.c:104)
    else fe_set_u32(u, 9);                         // ecp_BasePoint (curve25519_dh.c:37)
    fe r;
    const int t = threadIdx.x;
    x25519_ladder(r, u, [&](int w) { return ks[w][t]; });
    fe_store(out32 + 32 * i, r);
}

cudaError_t launch_x25519_ladder(uint8_t* out32, const uint8_t* pk32_or_null, uint8_t* sk32_inout, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + kLadderThreads - 1) / kLadderThreads);
    k_x25519_ladder<<<grid, kLadderThreads, 0, s>>>(out32, pk32_or_null, sk32_inout, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace c25519
