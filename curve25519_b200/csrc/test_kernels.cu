// test_kernels.cu -- unit-test and calibration kernels exported through c25519_test_primitive() and
// c25519_imad_peak_kernel().  They let tests/ differential-fuzz every device primitive against the
// reference's exported ecp_* / eco_* symbols, and let bench.py measure the
// IMAD.WIDE roofline denominator on the very device it is timing.
#include "kernels.h"
#include "fe25519.cuh"
#include "ge25519.cuh"
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

// op 8: (a || b) as a 512-bit little-endian integer mod L -> 32 bytes;  op 9: SHA-512(a || b) -> 64 bytes;
// op 11: (a*b + a) mod L -> 32 bytes
__global__ void __launch_bounds__(128)
k_test_sc_sha(int op, uint8_t* __restrict__ out, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
    fe_load(x, a + 32 * i);
    if (b) fe_load(y, b + 32 * i); else fe_set_u32(y, 0);
    if (op == 12 || op == 13) { // comb index extraction: ecp_8Folds (32 bytes out) / ecp_4Folds (64 bytes out) of a
        const int cnt = op == 12 ? 32 : 64;
        uint8_t* o = out + (size_t)cnt * i;
        for (int j = 0; j < cnt; j++) o[j] = (uint8_t)(op == 12 ? comb8_index(x.v, j) : comb4_index(x.v, j));
    } else if (op == 11) {     // (a*b + a) mod L through sc_muladd (eco_MulReduce + eco_AddReduce + eco_Mod)
        u32 r[8];
        sc_muladd(r, x.v, y.v, x.v);
        fe z;
#pragma unroll
        for (int k = 0; k < 8; k++) z.v[k] = r[k];
        fe_store(out + 32 * i, z);
    } else if (op == 8) {
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
    if (op == 8 || op == 9 || op >= 11) k_test_sc_sha<<<grid, 128, 0, s>>>(op, out, a, b, n);
    else k_test_fe<<<grid, 128, 0, s>>>(op, out, a, b, n);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// IMAD.WIDE.U32 rate of this device, measured two ways.  tests/test_sass.py disassembles the built object and
// asserts that each loop body really holds eight IMAD.WIDE.U32 of the stated form (round 1's "fresh" loop let ptxas
// hoist seven of its eight products):
//   form 0 "fresh"      : IMAD.WIDE.U32 Rd, Ra, Rb, RZ   -- every product fed back into its own multiplicand
//   form 1 "accumulate" : IMAD.WIDE.U32 Rd, Ra, Rb, Rd   -- 64-bit register accumulator, the form a multi-precision
//                         multiply-accumulate uses for most of its products
// 8 independent chains per loop trip, no memory traffic, 148 SMs x 4 resident CTAs x 256 threads.
constexpr int kPeakBlocks = 148 * 4;
constexpr int kPeakThreads = 256;

template <int FORM>
__global__ void __launch_bounds__(kPeakThreads)
k_imad_peak(uint32_t* sink, const uint32_t* src, int iters)
{
    const int t = blockIdx.x * kPeakThreads + threadIdx.x;
    unsigned long long a0 = src[(t + 0) & 63], a1 = src[(t + 1) & 63], a2 = src[(t + 2) & 63], a3 = src[(t + 3) & 63],
                       a4 = src[(t + 4) & 63], a5 = src[(t + 5) & 63], a6 = src[(t + 6) & 63], a7 = src[(t + 7) & 63];
    u32 x0 = src[(t + 8) & 63] | 1, x1 = src[(t + 9) & 63] | 1, x2 = src[(t + 10) & 63] | 1, x3 = src[(t + 11) & 63] | 1,
        x4 = src[(t + 12) & 63] | 1, x5 = src[(t + 13) & 63] | 1, x6 = src[(t + 14) & 63] | 1, x7 = src[(t + 15) & 63] | 1;
    u32 y0 = src[(t + 16) & 63] | 1, y1 = src[(t + 17) & 63] | 1, y2 = src[(t + 18) & 63] | 1, y3 = src[(t + 19) & 63] | 1,
        y4 = src[(t + 20) & 63] | 1, y5 = src[(t + 21) & 63] | 1, y6 = src[(t + 22) & 63] | 1, y7 = src[(t + 23) & 63] | 1;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        unsigned long long p0, p1, p2, p3, p4, p5, p6, p7;
        asm volatile("mul.wide.u32 %0,%8,%16; mul.wide.u32 %1,%9,%17; mul.wide.u32 %2,%10,%18; mul.wide.u32 %3,%11,%19; "
                     "mul.wide.u32 %4,%12,%20; mul.wide.u32 %5,%13,%21; mul.wide.u32 %6,%14,%22; mul.wide.u32 %7,%15,%23;"
                     : "=l"(p0), "=l"(p1), "=l"(p2), "=l"(p3), "=l"(p4), "=l"(p5), "=l"(p6), "=l"(p7)
                     : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(x6), "r"(x7),
                       "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(y4), "r"(y5), "r"(y6), "r"(y7));
        if (FORM == 0) {        // every chain feeds its product back into its own multiplicand (one 3-input XOR on the ALU pipe
                                // each), so all eight products are loop-variant: nothing can be hoisted out of the loop
            x0 ^= (u32)p0 ^ (u32)(p0 >> 32); x1 ^= (u32)p1 ^ (u32)(p1 >> 32); x2 ^= (u32)p2 ^ (u32)(p2 >> 32); x3 ^= (u32)p3 ^ (u32)(p3 >> 32);
            x4 ^= (u32)p4 ^ (u32)(p4 >> 32); x5 ^= (u32)p5 ^ (u32)(p5 >> 32); x6 ^= (u32)p6 ^ (u32)(p6 >> 32); x7 ^= (u32)p7 ^ (u32)(p7 >> 32);
        } else {                // ptxas folds these 64-bit adds into the multiply: IMAD.WIDE.U32 Rd, Ra, Rb, Rd.  The accumulators
                                // change every trip, so the multiply-adds cannot be hoisted even where x_k, y_k are loop-invariant.
            a0 += p0; a1 += p1; a2 += p2; a3 += p3; a4 += p4; a5 += p5; a6 += p6; a7 += p7;
        }
        if (FORM != 0) x0 += 1;
    }
    unsigned long long r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    u32 q = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if ((u32)r == 0x12345u && q == 77u) sink[0] = (u32)(r >> 32);      // never true in practice; keeps the chains alive
}

cudaError_t launch_imad_peak(uint64_t* mac_per_launch, uint32_t* sink, int iters, cudaStream_t s)
{
    // iters < 0 selects the accumulate form (|iters| trips); sink must hold >= 64 words of arbitrary data
    const int form = iters < 0 ? 1 : 0;
    const int n = iters < 0 ? -iters : iters;
    if (mac_per_launch) *mac_per_launch = (uint64_t)kPeakBlocks * kPeakThreads * (uint64_t)n * 8ull;
    if (form) k_imad_peak<1><<<kPeakBlocks, kPeakThreads, 0, s>>>(sink, sink, n);
    else k_imad_peak<0><<<kPeakBlocks, kPeakThreads, 0, s>>>(sink, sink, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace c25519
