// x25519.cuh -- Montgomery-ladder X25519 (variable base), one scalar multiplication per thread.
//
// Replaces source/curve25519_dh.c of the reference on the GPU:
//   ecp_MontDouble :40-54 -> mont_double     ecp_Mont :57-84 -> mont_step     ecp_PointMultiply :94-157 -> x25519_ladder
//
// Same algorithm shape as the reference (start from the top set bit with P = (u:1), Q = 2P, then one
// "P <- P+Q, Q <- 2Q" step per remaining bit with the roles of P and Q chosen by the bit), with the
// reference's pointer-table operand selection (ECP_MONT :89) replaced by a branch-free select of the point
// that gets doubled (the differential addition is symmetric), and its projective-Z randomisation (:123)
// dropped (result-neutral).
// All 256 bits of u are used (no bit-255 masking, :104); Z = 0 at the end gives 32 zero bytes (:148-150).
#pragma once
#include "fe25519.cuh"

namespace c25519 {

// (X2:Z2) = 2 (X:Z)                                   2S + 2M + 1W + 3A
C25519_DEV void mont_double(fe& X2, fe& Z2, const fe& X, const fe& Z)
{
    fe A, B;
    fe_add(A, X, Z);
    fe_sub(B, X, Z);
    fe_sqr(A, A);
    fe_sqr(B, B);
    fe_mul(X2, A, B);
    fe_sub(B, A, B);
    fe_mul_small_add(A, A, 121665u, B);
    fe_mul(Z2, A, B);
}

// (S,D) <- (S + D, 2 D) where S - D = (base : 1).      5M + 4S + 1W + 7A     -- the hot loop body
// SX,SZ,DX,DZ are N (outputs of fe_mul/fe_sqr) on entry and on exit.
template <typename BaseLoader>
C25519_DEV void mont_step_with(fe& SX, fe& SZ, fe& DX, fe& DZ, BaseLoader load_base)
{
    fe A, B, C, D;
    fe_sub(A, SX, SZ);
    fe_add_nn(B, SX, SZ);
    fe_sub(C, DX, DZ);
    fe_add_nn(D, DX, DZ);
    fe_mul(A, A, D);
    fe_mul(B, B, C);
    fe_add_nn(SX, A, B);
    fe_sub(B, A, B);
    fe_sqr(SX, SX);
    fe_sqr(A, B);
    { fe base; load_base(base); fe_mul(SZ, A, base); }
    fe_sqr(A, D);
    fe_sqr(B, C);
    fe_mul(DX, A, B);
    fe_sub(B, A, B);
    fe_mul_small_add(A, A, 121665u, B);
    fe_mul(DZ, A, B);
}

// The same step with the conditional swap folded in: the differential addition is symmetric in its two inputs, so only
// the DOUBLING has to know which point it doubles -- `dbl_s` selects the S slot instead of the D slot (what a full
// cswap(S, D) before mont_step_with would achieve) with 16 limb selects on the sum/difference pair instead of 32 swaps.
template <typename BaseLoader>
C25519_DEV void mont_step_sel(fe& SX, fe& SZ, fe& DX, fe& DZ, bool dbl_s, BaseLoader load_base)
{
    fe A, B, C, D, P, M;
    fe_sub(A, SX, SZ);
    fe_add_nn(B, SX, SZ);
    fe_sub(C, DX, DZ);
    fe_add_nn(D, DX, DZ);
    fe_select(P, D, B, dbl_s);          // x + z of the point to double
    fe_select(M, C, A, dbl_s);          // x - z
    fe_mul(A, A, D);
    fe_mul(B, B, C);
    fe_add_nn(SX, A, B);
    fe_sub(B, A, B);
    fe_sqr(SX, SX);
    fe_sqr(A, B);
    { fe base; load_base(base); fe_mul(SZ, A, base); }
    fe_sqr(A, P);
    fe_sqr(B, M);
    fe_mul(DX, A, B);
    fe_sub(B, A, B);
    fe_mul_small_add(A, A, 121665u, B);
    fe_mul(DZ, A, B);
}

C25519_DEV void mont_step(fe& SX, fe& SZ, fe& DX, fe& DZ, const fe& base)
{ mont_step_with(SX, SZ, DX, DZ, [&](fe& b) { fe_copy(b, base); }); }

// (PX : PZ) = projective x-coordinate of [k]u.   k must have bit 254 set and bit 255 clear (clamped);
// kw(w) returns 32-bit word w of k.  The affine result PX/PZ is produced later, for many operations at once,
// by the batched inversion in normalize.cuh (PZ == 0 there yields 32 zero bytes like the reference).
template <typename KeyWord>
C25519_DEV void x25519_ladder_projective(fe& PX, fe& PZ, const fe& u, KeyWord kw)
{
    fe R0X, R0Z, R1X, R1Z;          // R0 = P, R1 = Q while `cur` is true
    fe_copy(R0X, u);
    fe_set_u32(R0Z, 1);
    mont_double(R1X, R1Z, R0X, R0Z);
    // make R0X narrow: u is an arbitrary 256-bit value, the step wants N inputs for its lazy additions
    fe_narrow(R0X);
    bool cur = true;
#ifndef C25519_LADDER_UNROLL
#define C25519_LADDER_UNROLL 1
#endif
    constexpr int kLadderUnroll = C25519_LADDER_UNROLL;     // > 1 was measured: no gain (profiles/), code size doubles
#pragma unroll kLadderUnroll
    for (int bit = 253; bit >= 0; --bit) {
        bool b = (kw(bit >> 5) >> (bit & 31)) & 1u;
        bool s = (b != cur);
        cur = b;
#ifdef C25519_LADDER_CSWAP           // round-1 form: physical swap of the two slots, 0.75 % slower (profiles/r2_ladder_lab2.txt)
        fe_cswap(R0X, R1X, s);
        fe_cswap(R0Z, R1Z, s);
        mont_step(R0X, R0Z, R1X, R1Z, u);    // bit = 1: P += Q, Q = 2Q ; bit = 0: Q += P, P = 2P
#else
        mont_step_sel(R0X, R0Z, R1X, R1Z, s, [&](fe& bb) { fe_copy(bb, u); });
#endif
    }
    fe_select(PX, R1X, R0X, cur);
    fe_select(PZ, R1Z, R0Z, cur);
}

// Generic k*P for ANY 256-bit scalar (no clamping): ecp_PointMultiply's general contract (curve25519_dh.c:94-157:
// "MSB-first from the top set bit; K = 0 gives 32 zero bytes").  Instead of searching for the top set bit -- which
// would make lanes diverge -- the ladder starts one level higher, at (P, Q) = (O, (u:1)) with O = (1:0), and walks
// all 256 bits: leading zero bits map O to O (doubling) and (u:1) to itself up to a projective factor (differential
// addition with O), so when the top set bit arrives the state equals the reference's start (P, Q) = ((u:1), 2(u:1))
// projectively, and K = 0 ends with Z = 0, i.e. zeros.  Two ladder steps dearer than the clamped fast path.
template <typename KeyWord>
C25519_DEV void x25519_ladder_projective_raw(fe& PX, fe& PZ, const fe& u, KeyWord kw)
{
    fe R0X, R0Z, R1X, R1Z;
    fe_set_u32(R0X, 1); fe_set_u32(R0Z, 0);          // P = O
    fe_set_u32(R1Z, 1);
    fe_copy(R1X, u); fe_narrow(R1X);                  // Q = (u : 1), narrow representative
    bool cur = true;
#pragma unroll 1
    for (int bit = 255; bit >= 0; --bit) {
        bool b = (kw(bit >> 5) >> (bit & 31)) & 1u;
        bool s = (b != cur);
        cur = b;
        mont_step_sel(R0X, R0Z, R1X, R1Z, s, [&](fe& bb) { fe_copy(bb, u); });
    }
    fe_select(PX, R1X, R0X, cur);
    fe_select(PZ, R1Z, R0Z, cur);
}

// ---- warp-cooperative ladder for SMALL batches: one scalar multiplication per 4-lane group --------------------------------
// With fewer operations than the machine has thread slots, throughput is irrelevant and the latency of one dependent chain
// (254 steps x 10 field operations back to back) is all that counts.  A ladder step has only three multiplication LEVELS:
//     level 1:  DA = A.D      CB = B.C      PP = P^2       MM = M^2            (P, M = x+z, x-z of the point being doubled)
//     level 2:  (DA+CB)^2     (DA-CB)^2     PP.MM          121665.(PP-MM)
//     level 3:               u.(DA-CB)^2                   (PP-MM).(PP + 121665 (PP-MM))
// so four lanes, each executing ONE field multiplication per level on its own operands and trading results through
// __shfl_sync, walk a step in three multiplication latencies instead of ten.  Every lane keeps a full copy of the state
// (SX, SZ, DX, DZ); the additions are replicated (cheap).  Same rational maps as mont_step_sel, hence the same results.
C25519_DEV void fe_shfl_xor(fe& out, const fe& in, int mask)
{
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = __shfl_xor_sync(0xffffffffu, in.v[i], mask);
}
C25519_DEV void fe_bcast4(fe& out, const fe& in, int role)         // value held by lane `role` of each 4-lane group
{
    const int src = (threadIdx.x & 28) | role;
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = __shfl_sync(0xffffffffu, in.v[i], src);
}
C25519_DEV void fe_pick4(fe& out, int role, const fe& r0, const fe& r1, const fe& r2, const fe& r3)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 a = (role & 1) ? r1.v[i] : r0.v[i], b = (role & 1) ? r3.v[i] : r2.v[i];
        out.v[i] = (role & 2) ? b : a;
    }
}
// one step; on entry and exit all four lanes hold identical narrow (SX, SZ, DX, DZ)
C25519_DEV void mont_step_quad(fe& SX, fe& SZ, fe& DX, fe& DZ, bool dbl_s, const fe& u, int role)
{
    fe A, B, C, D, P, M, op1, op2, r, other;
    fe_sub(A, SX, SZ); fe_add_nn(B, SX, SZ); fe_sub(C, DX, DZ); fe_add_nn(D, DX, DZ);
    fe_select(P, D, B, dbl_s); fe_select(M, C, A, dbl_s);
    // level 1: lane 0: A.D   lane 1: B.C   lane 2: P.P   lane 3: M.M
    fe_pick4(op1, role, A, B, P, M); fe_pick4(op2, role, D, C, P, M);
    fe_mul(r, op1, op2);
    fe_shfl_xor(other, r, 1);                      // lanes 0 <-> 1 trade DA / CB, lanes 2 <-> 3 trade PP / MM
    // level 2 operands: lane 0: (DA+CB)^2   lane 1: (DA-CB)^2   lane 2: PP.MM   lane 3: (PP-MM) . 121665
    fe sum, dif, k;
    fe_add_nn(sum, r, other);
    if (role & 1) fe_sub(dif, other, r); else fe_sub(dif, r, other);     // lane 1: DA - CB, lane 3: PP - MM (lanes 0, 2: unused sign)
    fe_set_u32(k, 121665u);
    fe_pick4(op1, role, sum, dif, r, dif); fe_pick4(op2, role, sum, dif, other, k);
    fe r2;
    fe_mul(r2, op1, op2);
    // level 3: lane 1: u . (DA-CB)^2    lane 3: E . (PP + 121665 E), E = PP - MM (lane 3 holds MM in r, PP in other)
    fe F; fe_add_nn(F, other, r2);                 // meaningful on lane 3 only: PP + 121665 E
    fe_pick4(op1, role, r2, r2, r2, dif); fe_pick4(op2, role, r2, u, r2, F);
    fe r3;
    fe_mul(r3, op1, op2);
    // results: SX' = lane 0's r2, SZ' = lane 1's r3, DX' = lane 2's r2, DZ' = lane 3's r3
    fe_bcast4(SX, r2, 0); fe_bcast4(SZ, r3, 1); fe_bcast4(DX, r2, 2); fe_bcast4(DZ, r3, 3);
}

// canonical x-coordinate of [k]u by the four-lane ladder (k clamped: bit 254 set, bit 255 clear); all four lanes return it
template <typename KeyWord>
C25519_DEV void x25519_ladder_quad(fe& out, const fe& u, KeyWord kw, int role)
{
    fe R0X, R0Z, R1X, R1Z;
    fe_copy(R0X, u); fe_set_u32(R0Z, 1);
    mont_double(R1X, R1Z, R0X, R0Z);
    fe_narrow(R0X);
    bool cur = true;
#pragma unroll 1
    for (int bit = 253; bit >= 0; --bit) {
        bool b = (kw(bit >> 5) >> (bit & 31)) & 1u;
        bool s = (b != cur);
        cur = b;
        mont_step_quad(R0X, R0Z, R1X, R1Z, s, u, role);
    }
    fe PX, PZ, zi;
    fe_select(PX, R1X, R0X, cur); fe_select(PZ, R1Z, R0Z, cur);
    fe_invert(zi, PZ);
    fe_mul(out, PX, zi);
    fe_canon(out);
}

// out = canonical x-coordinate of [k]u (single-operation form: own inversion).
template <typename KeyWord>
C25519_DEV void x25519_ladder(fe& out, const fe& u, KeyWord kw)
{
    fe PX, PZ, zi;
    x25519_ladder_projective(PX, PZ, u, kw);
    fe_invert(zi, PZ);
    fe_mul(out, PX, zi);
    fe_canon(out);
}

}  // namespace c25519
