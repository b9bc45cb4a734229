// test_kernels.cu -- unit-test and calibration kernels exported through c25519_test_primitive() and
// c25519_imad_peak_kernel().  They let tests/ differential-fuzz every device primitive against the
// reference's exported ecp_* / eco_* symbols, and let bench.py measure the
// IMAD.WIDE roofline denominator on the very device it is timing.
#include "kernels.h"
#include "fe25519.cuh"
#include "sc25519.cuh"
#include "sha512.cuh"

namespace c25519 {

__global__ void __launch_bounds__(128)
k_test_fe(int op, uint8_t* __restrict__ out, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y, z;
    fe_load(x, a + 32 * i);
    if (b) fe_load(y, b + 32 * i); else fe_set_u32(y, 0);
    switch (op) {
    case 0: fe_mul(z, x, y); break;
    case 1: fe_sqr(z, x); break;
    case 2: fe_add(z, x, y); break;
    case 3: fe_sub(z, x, y); break;
    case 4: fe_invert(z, x); break;
    case 5: fe_pow22523(z, x); break;
    case 6: fe_mul_small_add(z, x, 121665u, y); break;
    case 10: { fe p, q; fe_mul(p, x, x); fe_mul(q, y, y); fe_add_nn(z, p, q); } break;   // lazy add of two N values
    default: fe_copy(z, x); break;
    }
    fe_canon(z);
    fe_store(out + 32 * i, z);
}

// op 8: (a || b) as a 512-bit little-endian integer mod L -> 32 bytes;  op 9: SHA-512(a || b) -> 64 bytes
__global__ void __launch_bounds__(128)
k_test_sc_sha(int op, uint8_t* __restrict__ out, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
    fe_load(x, a + 32 * i); fe_load(y, b + 32 * i);
    if (op == 8) {
        u32 w[16], r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { w[k] = x.v[k]; w[8 + k] = y.v[k]; }
        sc_reduce512(r, w);
        fe z;
#pragma unroll
        for (int k = 0; k < 8; k++) z.v[k] = r[k];
        fe_store(out + 32 * i, z);
    } else {
        u64 pre[8], dg[8]; u32 w[16];
        le_limbs_to_be64(pre, x.v); le_limbs_to_be64(pre + 4, y.v);
        sha512_prefixed<8>(dg, pre, nullptr, 0);
        sha512_digest_to_le_words(w, dg);
        fe lo, hi;
#pragma unroll
        for (int k = 0; k < 8; k++) { lo.v[k] = w[k]; hi.v[k] = w[8 + k]; }
        fe_store(out + 64 * i, lo); fe_store(out + 64 * i + 32, hi);
    }
}

cudaError_t launch_test_primitive(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (op == 8 || op == 9) k_test_sc_sha<<<grid, 128, 0, s>>>(op, out, a, b, n);
    else k_test_fe<<<grid, 128, 0, s>>>(op, out, a, b, n);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// IMAD.WIDE.U32 peak: 8 independent 64-bit accumulator chains per thread, 32 multiply-accumulates per
// loop trip, no memory traffic.  148 SMs x 4 resident CTAs x 256 threads.
constexpr int kPeakBlocks = 148 * 4;
constexpr int kPeakThreads = 256;

__global__ void __launch_bounds__(kPeakThreads)
k_imad_peak(uint32_t* sink, int iters)
{
    unsigned long long a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    u32 x = 0x9e3779b9u ^ threadIdx.x, y = 0x85ebca6bu + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\t"
                         "mad.wide.u32 %1, %8, %9, %1;\n\t"
                         "mad.wide.u32 %2, %8, %9, %2;\n\t"
                         "mad.wide.u32 %3, %8, %9, %3;\n\t"
                         "mad.wide.u32 %4, %8, %9, %4;\n\t"
                         "mad.wide.u32 %5, %8, %9, %5;\n\t"
                         "mad.wide.u32 %6, %8, %9, %6;\n\t"
                         "mad.wide.u32 %7, %8, %9, %7;"
                         : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(a3), "+l"(a4), "+l"(a5), "+l"(a6), "+l"(a7)
                         : "r"(x), "r"(y));
        }
    }
    unsigned long long r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x1234567ull) sink[0] = (u32)r;      // never true in practice; keeps the chains alive
}

cudaError_t launch_imad_peak(uint64_t* mac_per_launch, uint32_t* sink, int iters, cudaStream_t s)
{
    if (mac_per_launch) *mac_per_launch = (uint64_t)kPeakBlocks * kPeakThreads * (uint64_t)iters * 32ull;
    k_imad_peak<<<kPeakBlocks, kPeakThreads, 0, s>>>(sink, iters);
    count_launch();
    return cudaGetLastError();
}

}  // namespace c25519
