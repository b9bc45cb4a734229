// x25519_kernels.cu -- batched X25519 kernels (sm_100a).
//
//   k_x25519_ladder : curve25519_dh_CreateSharedKey (curve25519_dh.c:201) and the ladder flavour of
//                     curve25519_dh_CalculatePublicKey (:192); one operation per thread.
//
// Data layout in HBM: three arrays of 32-byte records (scalars in/out, peer points in, results out),
// record i belongs to thread i.  Each thread moves its record with ONE 256-bit access (LDG/STG.E.ENL2.256); a warp
// therefore touches one contiguous 1 KB span per array.  Algorithmic HBM traffic is
// 128 B per operation (32 sk in, 32 sk clamped out, 32 pk in, 32 out) against ~150 k integer
// multiply-adds, i.e. the kernel is bound by the integer pipes, not by HBM.
#include "kernels.h"
#include "normalize.cuh"
#include "sha512.cuh"
#include "x25519.cuh"
#include "x25519_warp.cuh"

namespace c25519 {

constexpr int kLadderThreads = 128;

// DEFER = true : stop at the projective result, write (X, Z) to the 96-byte scratch record i; the affine
//                result is produced for all operations by k_normalize (one shared inversion per 16 operations).
// DEFER = false: finish in place with a private inversion (tiny, latency-bound batches; the n = 1 legacy calls).
// ptxas settles on 98 registers (4 resident CTAs = 16 warps per SM).  Forcing 5 CTAs (<= 96 registers) makes it spill 36 bytes
// inside the loop; launch shapes between 16 and 21 warps per SM all land within 1 % anyway (profiles/r2_ladder_lab2.txt).
template <bool DEFER>
__global__ void __launch_bounds__(kLadderThreads)
k_x25519_ladder(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, uint8_t* __restrict__ sk32, size_t n,
                uint8_t* __restrict__ scratch)
{
    __shared__ u32 ks[8][kLadderThreads];          // scalar words, one column per thread (conflict-free)
    const size_t i = (size_t)blockIdx.x * kLadderThreads + threadIdx.x;
    if (i >= n) return;
    fe k;
    // ecp_TrimSecretKey (curve25519_utils.c:28-32), written back in place like the reference does
    fe_load_plain(k, sk32 + 32 * i);
    k.v[0] &= 0xfffffff8u;
    k.v[7] = (k.v[7] | 0x40000000u) & 0x7fffffffu;
    fe_store(sk32 + 32 * i, k);
#pragma unroll
    for (int w = 0; w < 8; w++) ks[w][threadIdx.x] = k.v[w];
    fe u;
    if (pk32) fe_load(u, pk32 + 32 * i);           // all 256 bits (curve25519_dh.c:104)
    else fe_set_u32(u, 9);                         // ecp_BasePoint (curve25519_dh.c:37)
    const int t = threadIdx.x;
    if (DEFER) {
        fe PX, PZ;
        x25519_ladder_projective(PX, PZ, u, [&](int w) { return ks[w][t]; });
        fe_store(scratch + kScratchXZ * i, PX);
        fe_store(scratch + kScratchXZ * i + 32, PZ);
    } else {
        fe r;
        x25519_ladder(r, u, [&](int w) { return ks[w][t]; });
        fe_store(out32 + 32 * i, r);
    }
}

// Small batches (n < kDeferThreshold): one operation per 4-lane group, 32 operations per 128-thread CTA (x25519.cuh:
// mont_step_quad).  The four lanes of a group load the same records, run the cooperative ladder, lane 0 stores.
__global__ void __launch_bounds__(kLadderThreads)
k_x25519_ladder_quad(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, uint8_t* __restrict__ sk32, size_t n)
{
    const int role = threadIdx.x & 3;
    size_t i = (size_t)blockIdx.x * (kLadderThreads / 4) + (threadIdx.x >> 2);
    const bool live = i < n;
    if (!live) i = n - 1;                              // keep whole warps converged for the shuffles; no stores from dead groups
    fe k;
    fe_load_plain(k, sk32 + 32 * i);
    k.v[0] &= 0xfffffff8u;
    k.v[7] = (k.v[7] | 0x40000000u) & 0x7fffffffu;
    fe u;
    if (pk32) fe_load(u, pk32 + 32 * i); else fe_set_u32(u, 9);
    __syncwarp();                                      // every lane of the group has read sk before lane 0 overwrites it
    fe r;
    x25519_ladder_quad(r, u, [&](int w) { return k.v[w]; }, role);
    if (live && role == 0) { fe_store(sk32 + 32 * i, k); fe_store(out32 + 32 * i, r); }
}

// Tiny batches (n <= kWarpThreshold): ONE operation per WARP -- limb per lane, four role groups, shuffles (x25519_warp.cuh).
// 128-thread CTAs = 4 operations; with at most one warp per SM sub-partition the latency of the ladder is all that matters.
__global__ void __launch_bounds__(kLadderThreads)
k_x25519_ladder_warp(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, uint8_t* __restrict__ sk32, size_t n)
{
    const size_t i = (size_t)blockIdx.x * (kLadderThreads / 32) + (threadIdx.x >> 5);
    if (i >= n) return;                                // warp-uniform
    const wlane c = w_lane();
    fe k;
    fe_load_plain(k, sk32 + 32 * i);
    k.v[0] &= 0xfffffff8u;
    k.v[7] = (k.v[7] | 0x40000000u) & 0x7fffffffu;
    u32 u;
    if (pk32) u = reinterpret_cast<const u32*>(pk32 + 32 * i)[c.limb]; else u = c.limb == 0 ? 9u : 0u;
    __syncwarp();                                      // every lane has read sk before lane 0 rewrites it
    const u32 res = w_x25519(u, [&](int w) { return k.v[w]; }, c);
    fe r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = __shfl_sync(0xffffffffu, res, j);     // group 0 holds the result, limb j in lane j
    fe_canon(r);
    if ((threadIdx.x & 31) == 0) { fe_store(sk32 + 32 * i, k); fe_store(out32 + 32 * i, r); }
}

// Generic scalar multiplication, no clamping, scalar NOT modified (ecp_PointMultiply, curve25519_dh.c:94).
__global__ void __launch_bounds__(kLadderThreads)
k_x25519_ladder_raw(const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ k32, size_t n, uint8_t* __restrict__ scratch)
{
    __shared__ u32 ks[8][kLadderThreads];
    const size_t i = (size_t)blockIdx.x * kLadderThreads + threadIdx.x;
    if (i >= n) return;
    fe k, u;
    fe_load(k, k32 + 32 * i);
#pragma unroll
    for (int w = 0; w < 8; w++) ks[w][threadIdx.x] = k.v[w];
    fe_load(u, pk32 + 32 * i);
    const int t = threadIdx.x;
    fe PX, PZ;
    x25519_ladder_projective_raw(PX, PZ, u, [&](int w) { return ks[w][t]; });
    fe_store(scratch + kScratchXZ * i, PX);
    fe_store(scratch + kScratchXZ * i + 32, PZ);
}

cudaError_t launch_x25519_ladder_raw(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + kLadderThreads - 1) / kLadderThreads);
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, n * kScratchXZ, s);
    if (e != cudaSuccess) return e;
    k_x25519_ladder_raw<<<grid, kLadderThreads, 0, s>>>(point32, scalar32, n, scratch);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = launch_normalize(kNormX, scratch, kScratchXZ, n, out32, 32, nullptr, 0, nullptr, 0, nullptr, s);
    cudaError_t e2 = wipe_and_free(scratch, n * kScratchXZ, s);
    return e != cudaSuccess ? e : e2;
}

// Shared KEY = first key_size bytes of SHA-512(shared secret): what the reference's C++ wrapper derives
// (X25519Private::CreateSharedKey, C++/x25519.cpp:75-95).  One 32-byte secret per thread, one compression.
__global__ void __launch_bounds__(128)
k_x25519_kdf_sha512(uint8_t* __restrict__ key_out, unsigned key_size, const uint8_t* __restrict__ secret32, size_t n)
{
    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    fe sct; fe_load_plain(sct, secret32 + 32 * i);
    u64 pre[4], dg[8]; u32 w[16];
    le_limbs_to_be64(pre, sct.v);
    sha512_prefixed<4>(dg, pre, nullptr, 0);
    sha512_digest_to_le_words(w, dg);
    uint8_t* o = key_out + (size_t)key_size * i;
    for (unsigned b = 0; b < key_size; b++) o[b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
}

cudaError_t launch_x25519_shared_kdf(uint8_t* key_out, unsigned key_size, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    uint8_t* secret = nullptr;
    cudaError_t e = cudaMallocAsync(&secret, n * 32, s);
    if (e != cudaSuccess) return e;
    e = launch_x25519_ladder(secret, pk32, sk32_inout, n, s);
    if (e == cudaSuccess) {
        k_x25519_kdf_sha512<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(key_out, key_size, secret, n);
        count_launch();
        e = cudaGetLastError();
    }
    cudaError_t e2 = wipe_and_free(secret, n * 32, s);
    return e != cudaSuccess ? e : e2;
}

// ---- batched normalisation kernel (shared by every operation of the path) ---------------------------
template <int MODE>
__global__ void __launch_bounds__(128)
k_normalize(uint8_t* __restrict__ scratch, size_t rec_stride, size_t n, size_t nthreads, int K, uint8_t* __restrict__ out, size_t out_stride,
            uint8_t* __restrict__ out2, size_t out2_stride, const uint8_t* __restrict__ cmp, size_t cmp_stride, int32_t* __restrict__ ok)
{
    const size_t tid = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (tid >= nthreads) return;
    normalize_walk<MODE>(scratch, rec_stride, n, tid, nthreads, K, out, out_stride, out2, out2_stride, cmp, cmp_stride, ok);
}

// X25519 results scattered to all ranks' gathered arrays by the normalisation kernel itself (peer stores).
__global__ void __launch_bounds__(128)
k_normalize_scatter(uint8_t* __restrict__ scratch, size_t rec_stride, size_t n, size_t nthreads, int K, PeerScatter sc)
{
    const size_t tid = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (tid >= nthreads) return;
    normalize_walk<kNormX>(scratch, rec_stride, n, tid, nthreads, K, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, &sc);
}

cudaError_t launch_x25519_ladder_scatter(uint8_t* const* out_ptrs, int world, int rank, const uint8_t* pk32, uint8_t* sk32_inout,
                                         size_t n_local, cudaStream_t s)
{
    if (n_local == 0) return cudaSuccess;
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return cudaErrorInvalidValue;
    PeerScatter sc; sc.world = world; sc.row_offset = (size_t)rank * n_local;
    for (int g = 0; g < 8; g++) sc.dst[g] = g < world ? out_ptrs[g] : nullptr;
    const unsigned grid = (unsigned)((n_local + kLadderThreads - 1) / kLadderThreads);
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, n_local * kScratchXZ, s);
    if (e != cudaSuccess) return e;
    k_x25519_ladder<true><<<grid, kLadderThreads, 0, s>>>(nullptr, pk32, sk32_inout, n_local, scratch);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) {
        size_t k = n_local / 8192; if (k < 1) k = 1; if (k > 16) k = 16;
        const size_t nthreads = (n_local + k - 1) / k;
        k_normalize_scatter<<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(scratch, kScratchXZ, n_local, nthreads, (int)k, sc);
        count_launch();
        e = cudaGetLastError();
    }
    cudaError_t e2 = wipe_and_free(scratch, n_local * kScratchXZ, s);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_normalize(int mode, uint8_t* scratch, size_t rec_stride, size_t n, uint8_t* out, size_t out_stride,
                             uint8_t* out2, size_t out2_stride, const uint8_t* cmp, size_t cmp_stride, int32_t* ok, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    // records walked per thread: 16 (one inversion per 16 operations) from 2^17 operations up -- the slice size of the host
    // pipeline, where the few resulting CTAs run underneath the other slices' ladders -- fewer for smaller batches (latency)
    size_t k = n / 8192; if (k < 1) k = 1; if (k > 16) k = 16;
    const size_t nthreads = (n + k - 1) / k;
    const unsigned grid = (unsigned)((nthreads + 127) / 128);
    switch (mode) {
    case kNormX: k_normalize<kNormX><<<grid, 128, 0, s>>>(scratch, rec_stride, n, nthreads, (int)k, out, out_stride, out2, out2_stride, cmp, cmp_stride, ok); break;
    case kNormEncode: k_normalize<kNormEncode><<<grid, 128, 0, s>>>(scratch, rec_stride, n, nthreads, (int)k, out, out_stride, out2, out2_stride, cmp, cmp_stride, ok); break;
    default: k_normalize<kNormCompare><<<grid, 128, 0, s>>>(scratch, rec_stride, n, nthreads, (int)k, out, out_stride, out2, out2_stride, cmp, cmp_stride, ok); break;
    }
    count_launch();
    return cudaGetLastError();
}

// The two halves of a large X25519 batch, separately launchable so that a caller can run the (latency-bound) batched
// inversion of one slice on a side stream underneath the ladder of the next slice (engine.cu: x25519_pipelined):
//   launch_x25519_projective : allocate the scratch on `s`, run the ladder, leave (X : Z) records in *scratch_out
//   launch_x25519_finish     : normalise into out32 on `s_finish` (the caller orders it after the ladder), wipe and free
cudaError_t launch_x25519_projective(uint8_t** scratch_out, const uint8_t* pk32_or_null, uint8_t* sk32_inout, size_t n, cudaStream_t s)
{
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, n * kScratchXZ, s);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((n + kLadderThreads - 1) / kLadderThreads);
    k_x25519_ladder<true><<<grid, kLadderThreads, 0, s>>>(nullptr, pk32_or_null, sk32_inout, n, scratch);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { wipe_and_free(scratch, n * kScratchXZ, s); return e; }
    *scratch_out = scratch;
    return cudaSuccess;
}
cudaError_t launch_x25519_finish(uint8_t* scratch, uint8_t* out32, size_t n, cudaStream_t s_finish)
{
    cudaError_t e = launch_normalize(kNormX, scratch, kScratchXZ, n, out32, 32, nullptr, 0, nullptr, 0, nullptr, s_finish);
    cudaError_t e2 = wipe_and_free(scratch, n * kScratchXZ, s_finish);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_x25519_ladder(uint8_t* out32, const uint8_t* pk32_or_null, uint8_t* sk32_inout, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + kLadderThreads - 1) / kLadderThreads);
    if (n <= kWarpThreshold) {                         // one warp per SM sub-partition at most: one operation per warp
        k_x25519_ladder_warp<<<(unsigned)((n + 3) / 4), kLadderThreads, 0, s>>>(out32, pk32_or_null, sk32_inout, n);
        count_launch();
        return cudaGetLastError();
    }
    if (n < kQuadThreshold) {                          // latency-bound: four lanes per operation
        k_x25519_ladder_quad<<<(unsigned)((n + kLadderThreads / 4 - 1) / (kLadderThreads / 4)), kLadderThreads, 0, s>>>(out32, pk32_or_null, sk32_inout, n);
        count_launch();
        return cudaGetLastError();
    }
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, n * kScratchXZ, s);
    if (e != cudaSuccess) return e;
    k_x25519_ladder<true><<<grid, kLadderThreads, 0, s>>>(out32, pk32_or_null, sk32_inout, n, scratch);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = launch_normalize(kNormX, scratch, kScratchXZ, n, out32, 32, nullptr, 0, nullptr, 0, nullptr, s);
    cudaError_t e2 = wipe_and_free(scratch, n * kScratchXZ, s);
    return e != cudaSuccess ? e : e2;
}

}  // namespace c25519
