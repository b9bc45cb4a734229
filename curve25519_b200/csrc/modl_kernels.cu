// modl_kernels.cu -- batched arithmetic modulo the group order L ("BPO" in the reference), one operation per thread.
//
// The reference's self-test carries a small mod-L toolbox for its split-key feature (k1 * k2 == 1 mod L, so that
// k2 * (k1 * P) == P): eco_MontMul / eco_ToMont / eco_FromMont (test/curve25519_selftest.c:206-254), eco_ExpModBPO
// (:258-276), eco_InvModBPO (:279-282), eco_MulMod / eco_AddMod (:160-175) on top of the library's eco_MulReduce,
// eco_AddReduce and eco_Mod (source/curve25519_order.c:110-136).  These kernels are their batched equivalents.  Every
// result is canonical (in [0, L)), which is what the reference produces once eco_Mod has run.
#include "kernels.h"
#include "sc25519.cuh"
#include "../../include/c25519_b200.h"

namespace c25519 {

// (2^256)^-1 mod L and L - 2, little-endian limbs
__device__ __constant__ const u32 kScRinv[8] = {0x3d5f0d24u, 0xc766cca4u, 0x973f754cu, 0xb7f5c66au, 0x8ffa36beu, 0x614e7543u, 0x26fe9183u, 0x09db6c6fu};
__device__ __constant__ const u32 kScLm2[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u};

C25519_DEV void sc_mul(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8])
{
    u32 z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) z[i] = 0;
    sc_muladd(r, a, b, z);
}

// y = x^e mod L, left-to-right square-and-multiply over all 256 exponent bits with a branch-free select (what
// eco_ExpModBPO does byte by byte with a data-dependent multiply)
C25519_DEV void sc_pow(u32 (&y)[8], const u32 (&x)[8], const u32 (&e)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = i == 0 ? 1u : 0u;
#pragma unroll 1
    for (int bit = 255; bit >= 0; --bit) {
        u32 t[8];
        sc_mul(y, y, y);
        sc_mul(t, y, x);
        const bool on = (e[bit >> 5] >> (bit & 31)) & 1u;
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = on ? t[i] : y[i];
    }
}

__global__ void __launch_bounds__(128)
k_modl(int op, uint8_t* __restrict__ out, const uint8_t* __restrict__ a32, const uint8_t* __restrict__ b32, size_t n)
{
    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    fe A, B;
    fe_load(A, a32 + 32 * i);
    if (b32) fe_load(B, b32 + 32 * i); else fe_set_u32(B, 0);
    u32 r[8];
    switch (op) {
    case C25519_MODL_MULMOD: sc_mul(r, A.v, B.v); break;                                    // eco_MulMod
    case C25519_MODL_ADDMOD: { u32 one[8] = {1, 0, 0, 0, 0, 0, 0, 0}; sc_muladd(r, A.v, one, B.v); } break;   // eco_AddMod
    case C25519_MODL_MONTMUL: { u32 t[8], ri[8];                                            // eco_MontMul: a*b/R
#pragma unroll
        for (int k = 0; k < 8; k++) ri[k] = kScRinv[k];
        sc_mul(t, A.v, B.v); sc_mul(r, t, ri); } break;
    case C25519_MODL_EXPMOD: sc_pow(r, A.v, B.v); break;                                    // eco_ExpModBPO
    default: { u32 e[8];                                                                    // eco_InvModBPO: a^(L-2)
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = kScLm2[k];
        sc_pow(r, A.v, e); } break;
    }
    fe z;
#pragma unroll
    for (int k = 0; k < 8; k++) z.v[k] = r[k];
    fe_store(out + 32 * i, z);
}

cudaError_t launch_modl(int op, uint8_t* out32, const uint8_t* a32, const uint8_t* b32, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_modl<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, out32, a32, b32, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace c25519
