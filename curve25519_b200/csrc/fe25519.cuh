// fe25519.cuh -- GF(2^255-19) arithmetic for sm_100a, one field element per thread, limbs in registers.
//
// Replaces (on the GPU) the reference's field layer source/curve25519_mehdi.c:
//   ecp_MulReduce :278   -> fe_mul      ecp_SqrReduce :310 -> fe_sqr     ecp_WordMulAddReduce :243 -> fe_mul_small_add
//   ecp_AddReduce :134   -> fe_add      ecp_SubReduce :161 -> fe_sub     ecp_Mod :185 -> fe_canon
//   ecp_Inverse   :340   -> fe_invert   ecp_ModExp2523 (ed25519_verify.c:116) -> fe_pow22523
//
// Representation: 8 x 32-bit saturated little-endian limbs (the same radix as the reference's portable-C
// build, so byte<->limb codecs are free), values are ANY 256-bit integer interpreted mod p and kept
// only loosely reduced; fe_canon() produces the unique representative in [0,p) that every public output
// of the reference passes through (ecp_Mod), which is what makes bit-exact parity possible with a
// completely different instruction schedule.
//
// Bounds vocabulary used in the comments:
//   W ("wide")   : any value < 2^256
//   N ("narrow") : value < 2^255 + 2^25   (output of fe_mul / fe_sqr / fe_mul_small_add)
//
// Instruction budget (sm_100a SASS, checked with cuobjdump): a 32x32->64 multiply-accumulate with carry
// in/out is ONE IMAD.WIDE.U32[.X]; fe_mul = 64 (product) + 8 (x38 fold) + 1 IMAD, fe_sqr = 36 + 8 + 1.
// The product is accumulated in two interleaved register files ("even" and "odd" aligned 64-bit
// columns) so that every partial product is a single IMAD.WIDE.U32.X chained through carry predicates.
#pragma once
#include <cstdint>

namespace c25519 {

typedef uint32_t u32;
typedef uint64_t u64;
#define C25519_DEV __device__ __forceinline__

struct fe { u32 v[8]; };

// ------------------------------------------------------------------ tiny PTX helpers
C25519_DEV void mul_wide(u32& lo, u32& hi, u32 a, u32 b)
{ asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=&r"(hi) : "r"(a), "r"(b)); }

// One schoolbook row "acc += x * y" split into the two carry chains described above.
//   lo[0..7]  += x_e[0..3] * y   (four 64-bit columns, all previously written; carry-out -> hi[7])
//   hi[0..7]  += x_o[0..3] * y   (top column hi[6..7] is fresh: set, not accumulated)
// `hi` starts one 32-bit word above `lo`.
C25519_DEV void mad_row(u32* lo, u32* hi, u32 xe0, u32 xe1, u32 xe2, u32 xe3, u32 xo0, u32 xo1, u32 xo2, u32 xo3, u32 y)
{
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %16, %24, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %24, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %24, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %24, 0;\n\t"
        "madc.hi.u32    %7, %19, %24, 0;\n\t"
        "mad.lo.cc.u32  %8,  %20, %24, %8;\n\t"
        "madc.hi.cc.u32 %9,  %20, %24, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %24, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %24, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %24, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %24, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %24, %14;\n\t"
        "madc.hi.cc.u32 %15, %23, %24, %15;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        "}"
        : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "=&r"(hi[6]), "=&r"(hi[7]),
          "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
        : "r"(xo0), "r"(xo1), "r"(xo2), "r"(xo3), "r"(xe0), "r"(xe1), "r"(xe2), "r"(xe3), "r"(y));
}

// x * 19 for small x.  The multiply pipe is the bound of every kernel here, so with C25519_ALU_SMALL_MUL the product is
// built from two shift-adds on the ALU pipe instead of an IMAD.
C25519_DEV u32 mul19(u32 x)
{
#ifdef C25519_ALU_SMALL_MUL
    u32 t, r;
    asm("{ .reg .u32 s; shl.b32 s, %2, 1; add.u32 %0, s, %2; shl.b32 s, %2, 4; add.u32 %1, s, %0; }" : "=&r"(t), "=r"(r) : "r"(x));
    return r;
#else
    return x * 19u;
#endif
}
// c * 38 for c in {0, 1}
C25519_DEV u32 bit38(u32 c)
{
#ifdef C25519_ALU_SMALL_MUL
    return (0u - c) & 38u;
#else
    return c * 38u;
#endif
}

// Final fold of a 9-word value Z[0..7] + w8 * 2^256 (w8 < 2^20) at bit 255:
//   hi = (w8 << 1) | (Z7 >> 31);  Z = (Z mod 2^255) + 19 * hi          -> N  (no carry out possible)
C25519_DEV void fold9(u32* z, u32 w8)
{
    u32 hi = (w8 << 1) | (z[7] >> 31);
    u32 m = mul19(hi);
    z[7] &= 0x7fffffffu;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3]), "+r"(z[4]), "+r"(z[5]), "+r"(z[6]), "+r"(z[7])
        : "r"(m));
}

// Reduce a 16-word product T (T[0..15]) to 8 words: Z = T_lo + 38 * T_hi, then fold9.   Output N.
C25519_DEV void reduce16(fe& z, u32* t)
{
    u32 u[8], w8;
#ifdef C25519_FRESH_FOLD
    {   // EXPERIMENT: all eight x38 products with a zero accumulator, merged by one more add chain
        u32 e[8];
        mul_wide(e[0], e[1], t[8], 38u); mul_wide(e[2], e[3], t[10], 38u);
        mul_wide(e[4], e[5], t[12], 38u); mul_wide(e[6], e[7], t[14], 38u);
        asm("add.cc.u32  %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %1, %10;\n\t"
            "addc.cc.u32 %2, %2, %11;\n\t"
            "addc.cc.u32 %3, %3, %12;\n\t"
            "addc.cc.u32 %4, %4, %13;\n\t"
            "addc.cc.u32 %5, %5, %14;\n\t"
            "addc.cc.u32 %6, %6, %15;\n\t"
            "addc.cc.u32 %7, %7, %16;\n\t"
            "addc.u32 %8, 0, 0;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=&r"(w8)
            : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]));
    }
#else
    // even columns of 38*T_hi accumulate straight into T_lo (carry chain); odd columns are fresh
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %9, %17, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %13, %17, %4;\n\t"
        "madc.hi.cc.u32 %5, %13, %17, %5;\n\t"
        "madc.lo.cc.u32 %6, %15, %17, %6;\n\t"
        "madc.hi.cc.u32 %7, %15, %17, %7;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        "}"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=&r"(w8)
        : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(38u));
#endif
    mul_wide(u[0], u[1], t[9], 38u);
    mul_wide(u[2], u[3], t[11], 38u);
    mul_wide(u[4], u[5], t[13], 38u);
    mul_wide(u[6], u[7], t[15], 38u);
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(w8)
        : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]));
    fold9(t, w8);                       // w8 <= 37 + 1 + 1
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = t[i];
}

// merge the two column files: T[k] = A[k] + B[k-1] (k = 1..n-1), carry into T[n]
// A: positions 0..15, B: positions 1..14 (B[0..13]).
C25519_DEV void merge_cols(u32* a, const u32* b)
{
    asm("add.cc.u32  %0, %0, %15;\n\t"
        "addc.cc.u32 %1, %1, %16;\n\t"
        "addc.cc.u32 %2, %2, %17;\n\t"
        "addc.cc.u32 %3, %3, %18;\n\t"
        "addc.cc.u32 %4, %4, %19;\n\t"
        "addc.cc.u32 %5, %5, %20;\n\t"
        "addc.cc.u32 %6, %6, %21;\n\t"
        "addc.cc.u32 %7, %7, %22;\n\t"
        "addc.cc.u32 %8, %8, %23;\n\t"
        "addc.cc.u32 %9, %9, %24;\n\t"
        "addc.cc.u32 %10, %10, %25;\n\t"
        "addc.cc.u32 %11, %11, %26;\n\t"
        "addc.cc.u32 %12, %12, %27;\n\t"
        "addc.cc.u32 %13, %13, %28;\n\t"
        "addc.u32    %14, %14, 0;"
        : "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
          "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]),
          "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]));
}

// ------------------------------------------------------------------ multiplication (ecp_MulReduce)
// Schoolbook 8x8: 64 products.  z = x*y mod (2^256-38), folded once more at bit 255.  Inputs W, output N.
C25519_DEV void fe_mul_schoolbook(fe& z, const fe& x, const fe& y)
{
    const u32* a = x.v; const u32* b = y.v;
    u32 A[16], B[14];
    // row 0: all columns fresh
    mul_wide(A[0], A[1], a[0], b[0]); mul_wide(A[2], A[3], a[2], b[0]);
    mul_wide(A[4], A[5], a[4], b[0]); mul_wide(A[6], A[7], a[6], b[0]);
    mul_wide(B[0], B[1], a[1], b[0]); mul_wide(B[2], B[3], a[3], b[0]);
    mul_wide(B[4], B[5], a[5], b[0]); mul_wide(B[6], B[7], a[7], b[0]);
    // row j: even-index limbs of x land on column parity j, odd-index limbs on parity j+1
    mad_row(B + 0, A + 2, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[1]);
    mad_row(A + 2, B + 2, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[2]);
    mad_row(B + 2, A + 4, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[3]);
    mad_row(A + 4, B + 4, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[4]);
    mad_row(B + 4, A + 6, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[5]);
    mad_row(A + 6, B + 6, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[6]);
    mad_row(B + 6, A + 8, a[0], a[2], a[4], a[6], a[1], a[3], a[5], a[7], b[7]);
    merge_cols(A, B);
    reduce16(z, A);
}

// 4x4-limb product (128 x 128 -> 256 bits), same two-column-file scheme: 7 fresh + 9 accumulating IMAD.WIDE.
//   out[0..7] = a[0..3] * b[0..3]
C25519_DEV void mul4x4(u32* out, const u32* a, const u32* b)
{
    u32 B[6];                  // B[k] sits at word position k+1 (positions 1..6); `out` doubles as file A (0..7)
    mul_wide(out[0], out[1], a[0], b[0]); mul_wide(out[2], out[3], a[2], b[0]);
    mul_wide(B[0], B[1], a[1], b[0]);     mul_wide(B[2], B[3], a[3], b[0]);
    // row 1: even limbs -> B cols 1,3 (accumulate, carry -> out[5]); odd limbs -> A col 2 (acc), col 4 (fresh)
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %9, %12, %0;\n\t"  "madc.hi.cc.u32 %1, %9, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, 0;\n\t"  "madc.hi.u32    %3, %10, %12, 0;\n\t"
        "mad.lo.cc.u32  %4, %8, %12, %4;\n\t"  "madc.hi.cc.u32 %5, %8, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t" "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(out[2]), "+r"(out[3]), "=&r"(out[4]), "=&r"(out[5]), "+r"(B[0]), "+r"(B[1]), "+r"(B[2]), "+r"(B[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[3]), "r"(a[2]), "r"(b[1]));
    // row 2: even limbs -> A cols 2,4 (acc, carry -> B[5]); odd limbs -> B col 3 (acc), col 5 (fresh)
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %9, %12, %0;\n\t"  "madc.hi.cc.u32 %1, %9, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, 0;\n\t"  "madc.hi.u32    %3, %10, %12, 0;\n\t"
        "mad.lo.cc.u32  %4, %8, %12, %4;\n\t"  "madc.hi.cc.u32 %5, %8, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t" "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(B[2]), "+r"(B[3]), "=&r"(B[4]), "=&r"(B[5]), "+r"(out[2]), "+r"(out[3]), "+r"(out[4]), "+r"(out[5])
        : "r"(a[0]), "r"(a[1]), "r"(a[3]), "r"(a[2]), "r"(b[2]));
    // row 3: even limbs -> B cols 3,5 (acc, carry -> out[7]); odd limbs -> A col 4 (acc), col 6 (fresh)
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %9, %12, %0;\n\t"  "madc.hi.cc.u32 %1, %9, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, 0;\n\t"  "madc.hi.u32    %3, %10, %12, 0;\n\t"
        "mad.lo.cc.u32  %4, %8, %12, %4;\n\t"  "madc.hi.cc.u32 %5, %8, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t" "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(out[4]), "+r"(out[5]), "=&r"(out[6]), "=&r"(out[7]), "+r"(B[2]), "+r"(B[3]), "+r"(B[4]), "+r"(B[5])
        : "r"(a[0]), "r"(a[1]), "r"(a[3]), "r"(a[2]), "r"(b[3]));
    // out[k] += B[k-1], k = 1..6, carry into out[7]
    asm("add.cc.u32  %0, %0, %7;\n\t"
        "addc.cc.u32 %1, %1, %8;\n\t"
        "addc.cc.u32 %2, %2, %9;\n\t"
        "addc.cc.u32 %3, %3, %10;\n\t"
        "addc.cc.u32 %4, %4, %11;\n\t"
        "addc.cc.u32 %5, %5, %12;\n\t"
        "addc.u32    %6, %6, 0;"
        : "+r"(out[1]), "+r"(out[2]), "+r"(out[3]), "+r"(out[4]), "+r"(out[5]), "+r"(out[6]), "+r"(out[7])
        : "r"(B[0]), "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]));
}

// One level of Karatsuba over 128-bit halves (EXPERIMENT, not the default -- see the note at fe_mul below):
// 48 products instead of 64, and 21 fresh + 27 accumulating multiplies instead of 15 + 49 (on B200 an IMAD.WIDE
// with a 64-bit register accumulator issues at ~0.56x the rate of one with a zero accumulator,
// profiles/r1_ubench4.txt), at the price of ~75 extra add/sub instructions.
//   x = xl + 2^128 xh, y = yl + 2^128 yh
//   z0 = xl yl, z2 = xh yh, zm = (xl + xh)(yl + yh)  (129-bit sums: 4-limb product + carry-bit corrections)
//   x y = z0 + 2^128 (zm - z0 - z2) + 2^256 z2
C25519_DEV void fe_mul_karatsuba(fe& z, const fe& x, const fe& y)
{
    u32 T[16];                          // T[0..7] = z0, T[8..15] = z2
    mul4x4(T, x.v, y.v);
    mul4x4(T + 8, x.v + 4, y.v + 4);
    u32 xs[4], ys[4], cx, cy;
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, 0, 0;"
        : "=&r"(xs[0]), "=&r"(xs[1]), "=&r"(xs[2]), "=&r"(xs[3]), "=&r"(cx)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]));
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, 0, 0;"
        : "=&r"(ys[0]), "=&r"(ys[1]), "=&r"(ys[2]), "=&r"(ys[3]), "=&r"(cy)
        : "r"(y.v[0]), "r"(y.v[1]), "r"(y.v[2]), "r"(y.v[3]), "r"(y.v[4]), "r"(y.v[5]), "r"(y.v[6]), "r"(y.v[7]));
    u32 M[9];                           // zm = (xl+xh)(yl+yh) < 2^258: 9 words
    mul4x4(M, xs, ys);
    {   // carry-bit corrections: + cx * ys * 2^128 + cy * xs * 2^128 + cx cy 2^256
        const u32 mx = 0u - cx, my = 0u - cy;
        u32 p0 = ys[0] & mx, p1 = ys[1] & mx, p2 = ys[2] & mx, p3 = ys[3] & mx;
        u32 q0 = xs[0] & my, q1 = xs[1] & my, q2 = xs[2] & my, q3 = xs[3] & my;
        M[8] = cx & cy;
        asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t addc.cc.u32 %3, %3, %8;\n\t addc.u32 %4, %4, 0;"
            : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8]) : "r"(p0), "r"(p1), "r"(p2), "r"(p3));
        asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t addc.cc.u32 %3, %3, %8;\n\t addc.u32 %4, %4, 0;"
            : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8]) : "r"(q0), "r"(q1), "r"(q2), "r"(q3));
    }
    // M -= z0; M -= z2    (the cross term xl yh + xh yl, 0 <= M < 2^257)
    asm("sub.cc.u32 %0, %0, %9;\n\t subc.cc.u32 %1, %1, %10;\n\t subc.cc.u32 %2, %2, %11;\n\t subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t subc.cc.u32 %5, %5, %14;\n\t subc.cc.u32 %6, %6, %15;\n\t subc.cc.u32 %7, %7, %16;\n\t subc.u32 %8, %8, 0;"
        : "+r"(M[0]), "+r"(M[1]), "+r"(M[2]), "+r"(M[3]), "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(T[0]), "r"(T[1]), "r"(T[2]), "r"(T[3]), "r"(T[4]), "r"(T[5]), "r"(T[6]), "r"(T[7]));
    asm("sub.cc.u32 %0, %0, %9;\n\t subc.cc.u32 %1, %1, %10;\n\t subc.cc.u32 %2, %2, %11;\n\t subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t subc.cc.u32 %5, %5, %14;\n\t subc.cc.u32 %6, %6, %15;\n\t subc.cc.u32 %7, %7, %16;\n\t subc.u32 %8, %8, 0;"
        : "+r"(M[0]), "+r"(M[1]), "+r"(M[2]), "+r"(M[3]), "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(T[8]), "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
    // T[4..12] += M, ripple into T[13..15]
    asm("add.cc.u32 %0, %0, %12;\n\t addc.cc.u32 %1, %1, %13;\n\t addc.cc.u32 %2, %2, %14;\n\t addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t addc.cc.u32 %5, %5, %17;\n\t addc.cc.u32 %6, %6, %18;\n\t addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t addc.cc.u32 %9, %9, 0;\n\t addc.cc.u32 %10, %10, 0;\n\t addc.u32 %11, %11, 0;"
        : "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(M[0]), "r"(M[1]), "r"(M[2]), "r"(M[3]), "r"(M[4]), "r"(M[5]), "r"(M[6]), "r"(M[7]), "r"(M[8]));
    reduce16(z, T);
}

// Measured on B200 (profiles/r1_ladder_lab_karatsuba_vs_schoolbook.txt): the Karatsuba variant is 3-4 % SLOWER in
// the ladder (48.7-49.1 vs 50.7-50.9 M ops/s): ptxas places a good part of the extra additions and register
// moves on the multiply pipe (IMAD.X / IMAD.MOV / IMAD.IADD), which eats the 16 saved products.  Schoolbook
// is therefore the default; -DC25519_MUL_KARATSUBA selects the other (same tests, same results).
#ifdef C25519_MUL_KARATSUBA
C25519_DEV void fe_mul(fe& z, const fe& x, const fe& y) { fe_mul_karatsuba(z, x, y); }
#else
C25519_DEV void fe_mul(fe& z, const fe& x, const fe& y) { fe_mul_schoolbook(z, x, y); }
#endif

// ------------------------------------------------------------------ squaring (ecp_SqrReduce)
// 28 off-diagonal products (two column files as in fe_mul), doubled, plus 8 diagonal squares.  Output N.
C25519_DEV void fe_sqr(fe& z, const fe& x)
{
    const u32* a = x.v;
    u32 A[16], B[14];       // A: even positions 0..15 ; B[k]: position k+1 (1..14)
    A[0] = 0; A[1] = 0; A[14] = 0; A[15] = 0;
    // row 0: x0 * {x1..x7}   -> positions 1..7 (+hi)
    mul_wide(B[0], B[1], a[0], a[1]); mul_wide(B[2], B[3], a[0], a[3]);
    mul_wide(B[4], B[5], a[0], a[5]); mul_wide(B[6], B[7], a[0], a[7]);
    mul_wide(A[2], A[3], a[0], a[2]); mul_wide(A[4], A[5], a[0], a[4]); mul_wide(A[6], A[7], a[0], a[6]);
    // row 1: x1 * {x2..x7}: even j -> B cols 3,5,7 (accumulate, carry -> A[9]); odd j -> A cols 4,6 (acc), 8 (fresh)
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %12, %13, %0;\n\t"  "madc.hi.cc.u32 %1, %12, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %12, %14, %2;\n\t"  "madc.hi.cc.u32 %3, %12, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %15, 0;\n\t"   "madc.hi.u32    %5, %12, %15, 0;\n\t"
        "mad.lo.cc.u32  %6, %12, %16, %6;\n\t"  "madc.hi.cc.u32 %7, %12, %16, %7;\n\t"
        "madc.lo.cc.u32 %8, %12, %17, %8;\n\t"  "madc.hi.cc.u32 %9, %12, %17, %9;\n\t"
        "madc.lo.cc.u32 %10, %12, %18, %10;\n\t" "madc.hi.cc.u32 %11, %12, %18, %11;\n\t"
        "addc.u32 %5, %5, 0;\n\t"
        "}"
        : "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7]), "=&r"(A[8]), "=&r"(A[9]),
          "+r"(B[2]), "+r"(B[3]), "+r"(B[4]), "+r"(B[5]), "+r"(B[6]), "+r"(B[7])
        : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[2]), "r"(a[4]), "r"(a[6]));
    // row 2: x2 * {x3..x7}: odd j -> B cols 5,7 (acc), 9 (fresh); even j -> A cols 6,8 (acc, carry -> B col 9 hi = B[9])
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %10, %11, %0;\n\t"  "madc.hi.cc.u32 %1, %10, %11, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, %2;\n\t"  "madc.hi.cc.u32 %3, %10, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %13, 0;\n\t"   "madc.hi.u32    %5, %10, %13, 0;\n\t"
        "mad.lo.cc.u32  %6, %10, %14, %6;\n\t"  "madc.hi.cc.u32 %7, %10, %14, %7;\n\t"
        "madc.lo.cc.u32 %8, %10, %15, %8;\n\t"  "madc.hi.cc.u32 %9, %10, %15, %9;\n\t"
        "addc.u32 %5, %5, 0;\n\t"
        "}"
        : "+r"(B[4]), "+r"(B[5]), "+r"(B[6]), "+r"(B[7]), "=&r"(B[8]), "=&r"(B[9]),
          "+r"(A[6]), "+r"(A[7]), "+r"(A[8]), "+r"(A[9])
        : "r"(a[2]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[4]), "r"(a[6]));
    // row 3: x3 * {x4..x7}: odd j -> A col 8 (acc), 10 (fresh); even j -> B cols 7,9 (acc, carry -> A[11])
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %8, %9, %0;\n\t"    "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, 0;\n\t"    "madc.hi.u32    %3, %8, %10, 0;\n\t"
        "mad.lo.cc.u32  %4, %8, %11, %4;\n\t"   "madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, %12, %6;\n\t"   "madc.hi.cc.u32 %7, %8, %12, %7;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(A[8]), "+r"(A[9]), "=&r"(A[10]), "=&r"(A[11]),
          "+r"(B[6]), "+r"(B[7]), "+r"(B[8]), "+r"(B[9])
        : "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[4]), "r"(a[6]));
    // row 4: x4 * {x5,x6,x7}: odd j -> B col 9 (acc), 11 (fresh); even j -> A col 10 (acc, carry -> B[11])
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %6, %7, %0;\n\t"    "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, 0;\n\t"     "madc.hi.u32    %3, %6, %8, 0;\n\t"
        "mad.lo.cc.u32  %4, %6, %9, %4;\n\t"    "madc.hi.cc.u32 %5, %6, %9, %5;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(B[8]), "+r"(B[9]), "=&r"(B[10]), "=&r"(B[11]),
          "+r"(A[10]), "+r"(A[11])
        : "r"(a[4]), "r"(a[5]), "r"(a[7]), "r"(a[6]));
    // row 5: x5 * {x6,x7}: odd j (7) -> A col 12 fresh; even j (6) -> B col 11 (acc, carry -> A[13])
    asm("{\n\t"
        "mul.lo.u32     %0, %4, %5;\n\t"        "mul.hi.u32     %1, %4, %5;\n\t"
        "mad.lo.cc.u32  %2, %4, %6, %2;\n\t"    "madc.hi.cc.u32 %3, %4, %6, %3;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(A[12]), "=&r"(A[13]), "+r"(B[10]), "+r"(B[11])
        : "r"(a[5]), "r"(a[7]), "r"(a[6]));
    // row 6: x6 * x7 -> B col 13 fresh
    mul_wide(B[12], B[13], a[6], a[7]);
    // T = A + (B << 32)
    merge_cols(A, B);
    // T = 2T  (positions 1..15): one add chain.  (-DC25519_SQR_DOUBLE_BY_SHIFT: fifteen independent funnel shifts instead of the
    // fifteen-long carry chain -- measured 0.5 % SLOWER in the ladder, 19.44 vs 19.34 ms: the chain's latency is hidden.)
#ifndef C25519_SQR_DOUBLE_BY_SHIFT
    asm("add.cc.u32  %0, %0, %0;\n\t"
        "addc.cc.u32 %1, %1, %1;\n\t"
        "addc.cc.u32 %2, %2, %2;\n\t"
        "addc.cc.u32 %3, %3, %3;\n\t"
        "addc.cc.u32 %4, %4, %4;\n\t"
        "addc.cc.u32 %5, %5, %5;\n\t"
        "addc.cc.u32 %6, %6, %6;\n\t"
        "addc.cc.u32 %7, %7, %7;\n\t"
        "addc.cc.u32 %8, %8, %8;\n\t"
        "addc.cc.u32 %9, %9, %9;\n\t"
        "addc.cc.u32 %10, %10, %10;\n\t"
        "addc.cc.u32 %11, %11, %11;\n\t"
        "addc.cc.u32 %12, %12, %12;\n\t"
        "addc.cc.u32 %13, %13, %13;\n\t"
        "addc.u32    %14, %14, %14;"
        : "+r"(A[1]), "+r"(A[2]), "+r"(A[3]), "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7]), "+r"(A[8]),
          "+r"(A[9]), "+r"(A[10]), "+r"(A[11]), "+r"(A[12]), "+r"(A[13]), "+r"(A[14]), "+r"(A[15]));
#else
#pragma unroll
    for (int k = 15; k >= 2; k--) A[k] = __funnelshift_l(A[k - 1], A[k], 1);      // (A[k] : A[k-1]) << 1, high word
    A[1] <<= 1;                                                                    // A[0] == 0 here
#endif
#ifdef C25519_FRESH_DIAG
    {   // EXPERIMENT: diagonal squares with a zero accumulator + one 16-word add chain
        u32 d[16];
#pragma unroll
        for (int i = 0; i < 8; i++) mul_wide(d[2 * i], d[2 * i + 1], a[i], a[i]);
        asm("add.cc.u32  %0, %0, %16;\n\t"  "addc.cc.u32 %1, %1, %17;\n\t"  "addc.cc.u32 %2, %2, %18;\n\t"  "addc.cc.u32 %3, %3, %19;\n\t"
            "addc.cc.u32 %4, %4, %20;\n\t"  "addc.cc.u32 %5, %5, %21;\n\t"  "addc.cc.u32 %6, %6, %22;\n\t"  "addc.cc.u32 %7, %7, %23;\n\t"
            "addc.cc.u32 %8, %8, %24;\n\t"  "addc.cc.u32 %9, %9, %25;\n\t"  "addc.cc.u32 %10, %10, %26;\n\t" "addc.cc.u32 %11, %11, %27;\n\t"
            "addc.cc.u32 %12, %12, %28;\n\t" "addc.cc.u32 %13, %13, %29;\n\t" "addc.cc.u32 %14, %14, %30;\n\t" "addc.u32 %15, %15, %31;"
            : "+r"(A[0]), "+r"(A[1]), "+r"(A[2]), "+r"(A[3]), "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7]),
              "+r"(A[8]), "+r"(A[9]), "+r"(A[10]), "+r"(A[11]), "+r"(A[12]), "+r"(A[13]), "+r"(A[14]), "+r"(A[15])
            : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
              "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]));
    }
#else
    // T += sum x_i^2 * 2^(64 i)
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %16, %16, %0;\n\t"   "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"   "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"   "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"   "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"   "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t" "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t" "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t" "madc.hi.u32    %15, %23, %23, %15;\n\t"
        "}"
        : "+r"(A[0]), "+r"(A[1]), "+r"(A[2]), "+r"(A[3]), "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7]),
          "+r"(A[8]), "+r"(A[9]), "+r"(A[10]), "+r"(A[11]), "+r"(A[12]), "+r"(A[13]), "+r"(A[14]), "+r"(A[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#endif
    reduce16(z, A);
}

// ------------------------------------------------------------------ z = y + b*x  (ecp_WordMulAddReduce)
// b < 2^18.  Inputs W, output N.
C25519_DEV void fe_mul_small_add(fe& z, const fe& y, u32 b, const fe& x)
{
    u32 t[8], u[8], w8;
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = y.v[i];
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %9, %13, %0;\n\t"   "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"  "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"  "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"  "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        "}"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=&r"(w8)
        : "r"(x.v[0]), "r"(x.v[2]), "r"(x.v[4]), "r"(x.v[6]), "r"(b));
    mul_wide(u[0], u[1], x.v[1], b); mul_wide(u[2], u[3], x.v[3], b);
    mul_wide(u[4], u[5], x.v[5], b); mul_wide(u[6], u[7], x.v[7], b);
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(w8)
        : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]));
    fold9(t, w8);                       // w8 < 2^18 + 2  ->  19*hi < 2^24
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = t[i];
}

// ------------------------------------------------------------------ add / sub
// z = x + y mod (2^256-38).  Inputs W, output W.   (ecp_AddReduce)
C25519_DEV void fe_add(fe& z, const fe& x, const fe& y)
{
    u32 t[8], c;
    asm("add.cc.u32  %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(c)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(y.v[0]), "r"(y.v[1]), "r"(y.v[2]), "r"(y.v[3]), "r"(y.v[4]), "r"(y.v[5]), "r"(y.v[6]), "r"(y.v[7]));
    u32 m = bit38(c), c2;
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, 0, 0;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=&r"(c2)
        : "r"(m));
    t[0] += bit38(c2);                  // a second wrap leaves t < 38, so this cannot carry
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = t[i];
}

// z = x + y for N inputs: the sum exceeds 2^256 only when both are >= 2^255 - 2^25, and then the wrapped
// sum is < 2^26, so the fix-up touches limb 0 only.  Output W.
C25519_DEV void fe_add_nn(fe& z, const fe& x, const fe& y)
{
    u32 t[8], c;
    asm("add.cc.u32  %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(c)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(y.v[0]), "r"(y.v[1]), "r"(y.v[2]), "r"(y.v[3]), "r"(y.v[4]), "r"(y.v[5]), "r"(y.v[6]), "r"(y.v[7]));
    t[0] += bit38(c);
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = t[i];
}

// z = x - y mod (2^256-38).  Inputs W, output W.   (ecp_SubReduce)
C25519_DEV void fe_sub(fe& z, const fe& x, const fe& y)
{
    u32 t[8], b;
    asm("sub.cc.u32  %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(b)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(y.v[0]), "r"(y.v[1]), "r"(y.v[2]), "r"(y.v[3]), "r"(y.v[4]), "r"(y.v[5]), "r"(y.v[6]), "r"(y.v[7]));
    u32 m = b & 38u, b2;                // b = 0 or 0xffffffff
    asm("sub.cc.u32  %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32    %8, 0, 0;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=&r"(b2)
        : "r"(m));
    t[0] -= b2 & 38u;                   // second wrap: t >= 2^256 - 38, limb 0 only
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = t[i];
}

// ------------------------------------------------------------------ canonical form (ecp_Mod)
// z in [0,p).  Input W.
C25519_DEV void fe_canon(fe& z)
{
    u32 t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = z.v[i];
    fold9(t, 0);                        // t < 2^255 + 19
    // s = t + 19; if bit 255 of s is set then t >= p and the answer is s mod 2^255
    u32 s[8];
    asm("add.cc.u32  %0, %8, 19;\n\t"
        "addc.cc.u32 %1, %9, 0;\n\t"
        "addc.cc.u32 %2, %10, 0;\n\t"
        "addc.cc.u32 %3, %11, 0;\n\t"
        "addc.cc.u32 %4, %12, 0;\n\t"
        "addc.cc.u32 %5, %13, 0;\n\t"
        "addc.cc.u32 %6, %14, 0;\n\t"
        "addc.u32    %7, %15, 0;"
        : "=&r"(s[0]), "=&r"(s[1]), "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
    bool ge = (s[7] >> 31) != 0;
    s[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = ge ? s[i] : t[i];
}

// z <- the same field element with a narrow representative (< 2^255 + 19): one fold at bit 255, no multiplication.
C25519_DEV void fe_narrow(fe& z) { fold9(z.v, 0); }

C25519_DEV void fe_set_u32(fe& z, u32 v)
{ z.v[0] = v;
#pragma unroll
  for (int i = 1; i < 8; i++) z.v[i] = 0; }

C25519_DEV void fe_copy(fe& z, const fe& x)
{
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = x.v[i]; }

// branch-free select / swap (lane-uniform control flow: secret bits never steer a branch, which on a GPU
// is a divergence question, not a side-channel one)
C25519_DEV void fe_cswap(fe& a, fe& b, bool s)
{
#pragma unroll
    for (int i = 0; i < 8; i++) { u32 x = a.v[i], y = b.v[i]; a.v[i] = s ? y : x; b.v[i] = s ? x : y; }
}
C25519_DEV void fe_select(fe& z, const fe& a, const fe& b, bool pick_b)
{
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = pick_b ? b.v[i] : a.v[i];
}
C25519_DEV bool fe_is_zero_canon(const fe& a)   // a must be canonical
{ u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r |= a.v[i];
  return r == 0; }

// z = -x mod p.   Input W, output W.
C25519_DEV void fe_neg(fe& z, const fe& x)
{ fe zero; fe_set_u32(zero, 0); fe_sub(z, zero, x); }

// n successive squarings (n >= 1), not unrolled: the body is one fe_sqr
C25519_DEV void fe_sqr_n(fe& z, const fe& x, int n)
{
    fe_sqr(z, x);
#pragma unroll 1
    for (int i = 1; i < n; i++) fe_sqr(z, z);
}

// z^(2^250-1) and z^11: the shared head of the two fixed addition chains
// (ecp_Inverse curve25519_mehdi.c:340-409, ecp_ModExp2523 ed25519_verify.c:116-135)
C25519_DEV void fe_pow_2_250_1(fe& out, fe& z11, const fe& z)
{
    fe z2, z9, t, a5, a10, a50;
    fe_sqr(z2, z);
    fe_sqr_n(t, z2, 2);
    fe_mul(z9, t, z);
    fe_mul(z11, z9, z2);
    fe_sqr(t, z11);
    fe_mul(a5, t, z9);                 // 2^5 - 1
    fe_sqr_n(t, a5, 5);   fe_mul(a10, t, a5);      // 2^10 - 1
    fe_sqr_n(t, a10, 10); fe_mul(z2, t, a10);      // 2^20 - 1   (z2 reused)
    fe_sqr_n(t, z2, 20);  fe_mul(t, t, z2);        // 2^40 - 1
    fe_sqr_n(t, t, 10);   fe_mul(a50, t, a10);     // 2^50 - 1
    fe_sqr_n(t, a50, 50); fe_mul(z9, t, a50);      // 2^100 - 1  (z9 reused)
    fe_sqr_n(t, z9, 100); fe_mul(t, t, z9);        // 2^200 - 1
    fe_sqr_n(t, t, 50);   fe_mul(out, t, a50);     // 2^250 - 1
}

// z^(p-2); 0 -> 0 like the reference.   (ecp_Inverse)
C25519_DEV void fe_invert(fe& out, const fe& z)
{ fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqr_n(t, t, 5); fe_mul(out, t, z11); }

// z^((p-5)/8) = z^(2^252-3).   (ecp_ModExp2523)
C25519_DEV void fe_pow22523(fe& out, const fe& z)
{ fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqr_n(t, t, 2); fe_mul(out, t, z); }

// ------------------------------------------------------------------ byte codecs (little-endian host == limb order)
// ecp_BytesToWords curve25519_utils.c:43 / ecp_WordsToBytes :61: a 32-byte record IS the limb array, and
// sm_100 moves it with ONE 256-bit access per thread (PTX ld/st.global.v8.b32, SASS LDG/STG.E.ENL2.256).
// Records must be 32-byte aligned (checked at the C ABI).
C25519_DEV void fe_load(fe& z, const uint8_t* p)        // read-only input (non-coherent path)
{
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(z.v[0]), "=r"(z.v[1]), "=r"(z.v[2]), "=r"(z.v[3]), "=r"(z.v[4]), "=r"(z.v[5]), "=r"(z.v[6]), "=r"(z.v[7])
                 : "l"(p));
}
C25519_DEV void fe_load_plain(fe& z, const uint8_t* p)  // coherent load (data written earlier by this grid or stream)
{
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(z.v[0]), "=r"(z.v[1]), "=r"(z.v[2]), "=r"(z.v[3]), "=r"(z.v[4]), "=r"(z.v[5]), "=r"(z.v[6]), "=r"(z.v[7])
                 : "l"(p) : "memory");
}
C25519_DEV void fe_store(uint8_t* p, const fe& z)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(z.v[0]), "r"(z.v[1]), "r"(z.v[2]), "r"(z.v[3]), "r"(z.v[4]), "r"(z.v[5]), "r"(z.v[6]), "r"(z.v[7])
                 : "memory");
}

}  // namespace c25519
