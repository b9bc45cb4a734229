// sc25519.cuh -- arithmetic modulo the group order L = 2^252 + 27742317777372353535851937790883648493,
// one scalar per thread.  Replaces source/curve25519_order.c of the reference:
//   eco_DigestToWords :139 + eco_Mod :125   -> sc_reduce512  (512-bit little-endian value -> canonical [0,L))
//   eco_MulReduce :110 + eco_AddReduce :132 + eco_Mod :125 -> sc_muladd  ((h*a + r) mod L, canonical)
// The reference folds one 32-bit word at a time through eco_ReduceHiWord (:80); every value that reaches
// an output is canonicalised by eco_Mod, so any exact reduction gives identical bytes.  Here: three folds
// at 2^256 with c = 2^256 mod L = -16*delta (delta = L - 2^252, so 2^256 = 16 L - 16 delta), a signed
// recombination, and one conditional add.  This is < 1 % of any Ed25519 operation; plain C with 64-bit
// accumulators (IMAD.WIDE) is ample.
#pragma once
#include "fe25519.cuh"

namespace c25519 {

// 16*delta = 2^256 mod L negated, 129 bits, 5 limbs
__device__ __constant__ const u32 kSc16Delta[5] = {0xcf5d3ed0u, 0x812631a5u, 0x2f79cd65u, 0x4def9deau, 0x00000001u};
__device__ __constant__ const u32 kScL[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};

// out[NA+NB] = a[NA] * b[NB]
template <int NA, int NB>
C25519_DEV void bn_mul(u32* out, const u32* a, const u32* b)
{
#pragma unroll
    for (int i = 0; i < NA + NB; i++) out[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        u64 carry = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            u64 t = (u64)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (u32)t; carry = t >> 32;
        }
        out[i + NB] = (u32)carry;
    }
}

// r (canonical, 8 limbs) = x (16 limbs, little-endian 512-bit) mod L
C25519_DEV void sc_reduce512(u32 (&r)[8], const u32 (&x)[16])
{
    u32 c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) c[i] = kSc16Delta[i];
    // x = lo + 2^256 hi  ==  lo - hi*c
    u32 t[13]; bn_mul<8, 5>(t, &x[8], c);                // t = hi*c  < 2^385
    // t = t_lo + 2^256 t_hi == t_lo - t_hi*c
    u32 u[10]; bn_mul<5, 5>(u, &t[8], c);                // u = t_hi*c < 2^258
    // u = u_lo + 2^256 u_hi == u_lo - u_hi*c  (u_hi < 4)
    u32 v[6]; u32 uhi = u[8];                            // u[9] == 0
    { u64 carry = 0;
#pragma unroll
      for (int j = 0; j < 5; j++) { u64 q = (u64)uhi * c[j] + carry; v[j] = (u32)q; carry = q >> 32; }
      v[5] = (u32)carry; }
    // s = lo - t_lo + u_lo - v  + 32 L        (|lo - t_lo + u_lo - v| < 2^257 < 32 L, so s > 0 and s < 2^259)
    // 32 L = 2^257 + 32 delta ; 32*delta = 2 * (16 delta) is 130 bits
    long long acc = 0; u32 s[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        long long term = 0;
        if (i < 8) term += (long long)x[i] - (long long)t[i] + (long long)u[i];
        if (i < 6) term -= (long long)v[i];
        // 32 L limbs: low 5 limbs = 2 * c (130 bits -> limb 4 holds the top bits), limb 8 = 2 (2^257)
        if (i < 5) term += (long long)(((u64)c[i] << 1) & 0xffffffffull) + (i ? (long long)(c[i - 1] >> 31) : 0);
        if (i == 8) term += 2;
        acc += term;
        s[i] = (u32)acc;
        acc >>= 32;                                       // arithmetic shift keeps the borrow
    }
    // q = s >> 252 (< 2^7);  s - q L = (s mod 2^252) - q*delta, add L back if negative
    u32 q = (s[7] >> 28) | (s[8] << 4);
    s[7] &= 0x0fffffffu;
    u32 qd[5];                                            // q * delta, delta = 125 bits = 4 limbs
    { u64 carry = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) { u64 p = (u64)q * kScL[j] + carry; qd[j] = (u32)p; carry = p >> 32; }
      qd[4] = (u32)carry; }
    long long b = 0; u32 d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        b += (long long)s[i] - (long long)(i < 5 ? qd[i] : 0u);
        d[i] = (u32)b; b >>= 32;
    }
    const u32 neg = (u32)b;                               // 0 or 0xffffffff
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 p = (u64)d[i] + (kScL[i] & neg) + carry;
        r[i] = (u32)p; carry = p >> 32;
    }
}

// s = (h*a + r) mod L, canonical.  h, a, r: 8 limbs each (any 256-bit values).
C25519_DEV void sc_muladd(u32 (&s)[8], const u32 (&h)[8], const u32 (&a)[8], const u32 (&r)[8])
{
    u32 p[16]; bn_mul<8, 8>(p, h, a);
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { u64 t = (u64)p[i] + (i < 8 ? r[i] : 0u) + carry; p[i] = (u32)t; carry = t >> 32; }
    // h*a + r < 2^512 always (max (2^256-1)^2 + 2^256-1), carry == 0
    sc_reduce512(s, p);
}

}  // namespace c25519
