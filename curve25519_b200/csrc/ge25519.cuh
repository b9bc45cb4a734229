// ge25519.cuh -- twisted-Edwards point arithmetic for Ed25519 / the fixed-base comb, one point per thread.
//
// Replaces the reference's point layer:
//   edp_DoublePoint    ed25519_sign.c:122-143   -> ge_double
//   edp_AddAffinePoint ed25519_sign.c:97-115    -> ge_add_affine   (precomputed affine operand, Z2 = 1)
//   edp_AddPoint       ed25519_verify.c:142-161 -> ge_add_pe       (precomputed projective operand)
//   edp_ExtPoint2PE    ed25519_sign.c:270-276   -> ge_to_pe
//   ecp_8Folds / ecp_4Folds curve25519_utils.c:144 / :125 -> comb8_index / comb4_index
//   edp_BasePointMult  ed25519_sign.c:215-244   -> ge_base_comb
//   ed25519_CalculateX ed25519_verify.c:66-100  -> ge_recover_x
//
// Every function performs the same sequence of rational maps on projective points as its reference
// counterpart (so even off-curve "garbage" inputs, which the reference never rejects, follow the same
// trajectory); internal representatives may differ by a common projective factor (e.g. ge_double returns
// (-1)x the reference's coordinates), which cancels in every output.
#pragma once
#include "curve_constants.cuh"
#include "fe25519.cuh"

namespace c25519 {

struct ge_ext { fe x, y, z, t; };          // Ext_POINT  (curve25519_mehdi.h:60-65), T = XY/Z
struct ge_pa  { fe ypx, ymx, t2d; };       // PA_POINT   (curve25519_mehdi.h:77-82)
struct ge_pe  { fe ypx, ymx, t2d, z2; };   // PE_POINT   (curve25519_mehdi.h:68-74)

C25519_DEV void fe_from_const(fe& z, const u32* c)
{
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = c[i];
}

// P <- 2P.   4S + 4M (3M when WITH_T is false).   Coordinates of P are N on entry and on exit.
// The doubling does not read T, so inside a run of consecutive doublings (the 3 x 64 doublings of the per-key
// table, ed25519_verify.c:206,212,222) only the last one has to produce it: X, Y, Z -- hence every later
// point -- are unchanged.
template <bool WITH_T = true>
C25519_DEV void ge_double(ge_ext& p)
{
    fe xx, yy, zz2, s, g, f, e;
    fe_add_nn(e, p.x, p.y);
    fe_sqr(xx, p.x);
    fe_sqr(yy, p.y);
    fe_sqr(zz2, p.z);
    fe_sqr(e, e);                   // (X+Y)^2
    fe_add_nn(zz2, zz2, zz2);       // 2 Z^2                 W
    fe_add_nn(s, yy, xx);           // S = Y^2 + X^2  (= -H) W
    fe_sub(g, yy, xx);              // G = Y^2 - X^2         W
    fe_sub(f, zz2, g);              // F' = 2Z^2 - G  (= -F) W
    fe_sub(e, e, s);                // E = (X+Y)^2 - X^2 - Y^2
    fe_mul(p.x, e, f);
    fe_mul(p.y, s, g);
    fe_mul(p.z, g, f);
    if (WITH_T) fe_mul(p.t, e, s);
}

// shared tail of the two additions: given A, B, C (N) and D = 2 Z1 Z2 (W or N).
// WITH_T = false skips T3 = E H: an addition whose result is only ever doubled next (the doubling does not read T) or
// encoded does not need it; X, Y, Z -- and therefore every later point -- are unchanged.
template <bool WITH_T = true>
C25519_DEV void ge_add_tail(ge_ext& r, const fe& a, const fe& b, const fe& c, const fe& d, bool d_is_narrow)
{
    fe e, h, f, g;
    fe_sub(e, b, a);                // E = B - A
    fe_add_nn(h, b, a);             // H = B + A
    fe_sub(f, d, c);                // F = D - C
    if (d_is_narrow) fe_add_nn(g, d, c); else fe_add(g, d, c);   // G = D + C
    fe_mul(r.x, e, f);
    fe_mul(r.y, h, g);
    if (WITH_T) fe_mul(r.t, e, h);
    fe_mul(r.z, g, f);
}

// P <- P + Q, Q precomputed affine (Z2 = 1).   7M (6M without T).
template <bool WITH_T = true>
C25519_DEV void ge_add_affine(ge_ext& p, const ge_pa& q)
{
    fe a, b, c, d;
    fe_sub(a, p.y, p.x);
    fe_add_nn(b, p.y, p.x);
    fe_mul(a, a, q.ymx);
    fe_mul(b, b, q.ypx);
    fe_mul(c, p.t, q.t2d);
    fe_add_nn(d, p.z, p.z);         // D = 2 Z1      W
    ge_add_tail<WITH_T>(p, a, b, c, d, false);
}

// R <- P + Q, Q precomputed projective.   8M (7M without T).   (R may alias P)
template <bool WITH_T = true>
C25519_DEV void ge_add_pe(ge_ext& r, const ge_ext& p, const ge_pe& q)
{
    fe a, b, c, d;
    fe_sub(a, p.y, p.x);
    fe_add_nn(b, p.y, p.x);
    fe_mul(a, a, q.ymx);
    fe_mul(b, b, q.ypx);
    fe_mul(c, p.t, q.t2d);
    fe_mul(d, p.z, q.z2);           // D = Z1 * 2 Z2  N
    ge_add_tail<WITH_T>(r, a, b, c, d, true);
}

// Ext -> PE.   1M.   p coordinates N.
C25519_DEV void ge_to_pe(ge_pe& r, const ge_ext& p)
{
    fe k2d; fe_from_const(k2d, k2D);
    fe_add_nn(r.ypx, p.y, p.x);
    fe_sub(r.ymx, p.y, p.x);
    fe_mul(r.t2d, p.t, k2d);
    fe_add_nn(r.z2, p.z, p.z);
}

// S <- the point a PE entry represents, scaled by 2: (2X, 2Y, 2Z, 2T)   (ed25519_verify.c:258-263)
C25519_DEV void ge_from_pe(ge_ext& s, const ge_pe& q)
{
    fe di; fe_from_const(di, kDInv);
    fe_sub(s.x, q.ypx, q.ymx);
    fe_add(s.y, q.ypx, q.ymx);
    fe_mul(s.t, q.t2d, di);
    fe_copy(s.z, q.z2);
    // make x, y, z narrow for the lazy additions downstream
    fe_narrow(s.x); fe_narrow(s.y); fe_narrow(s.z);
}

// comb indices ------------------------------------------------------------------------------------
// ecp_8Folds: index i (0..31) has bit k = bit (31 - i) of 32-bit word k of the scalar
C25519_DEV u32 comb8_index(const u32 (&s)[8], int i)
{
    const int sh = 31 - i;
    u32 r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r |= ((s[k] >> sh) & 1u) << k;
    return r;
}
// ecp_4Folds: index i (0..63) has bit k = bit (63 - i) of 64-bit limb k
C25519_DEV u32 comb4_index(const u32 (&s)[8], int i)
{
    const int bit = 63 - i;
    const int sh = bit & 31;
    u32 r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 w = (bit >= 32) ? s[2 * k + 1] : s[2 * k];
        r |= ((w >> sh) & 1u) << k;
    }
    return r;
}

// shared-memory comb table ------------------------------------------------------------------------
// 256 entries, kCombStrideWords (28) words apart: 24 data words + 4 pad words so that 16-byte vector reads
// by lanes with unrelated indices spread over the banks (entry stride of 7 quad-banks is odd mod 8).
constexpr int kCombStrideWords = 28;

C25519_DEV void comb_load(ge_pa& q, const u32* __restrict__ table_smem, u32 idx)
{
    const uint4* e = reinterpret_cast<const uint4*>(table_smem + idx * kCombStrideWords);
    uint4 a0 = e[0], a1 = e[1], b0 = e[2], b1 = e[3], c0 = e[4], c1 = e[5];
    q.ypx.v[0] = a0.x; q.ypx.v[1] = a0.y; q.ypx.v[2] = a0.z; q.ypx.v[3] = a0.w;
    q.ypx.v[4] = a1.x; q.ypx.v[5] = a1.y; q.ypx.v[6] = a1.z; q.ypx.v[7] = a1.w;
    q.ymx.v[0] = b0.x; q.ymx.v[1] = b0.y; q.ymx.v[2] = b0.z; q.ymx.v[3] = b0.w;
    q.ymx.v[4] = b1.x; q.ymx.v[5] = b1.y; q.ymx.v[6] = b1.z; q.ymx.v[7] = b1.w;
    q.t2d.v[0] = c0.x; q.t2d.v[1] = c0.y; q.t2d.v[2] = c0.z; q.t2d.v[3] = c0.w;
    q.t2d.v[4] = c1.x; q.t2d.v[5] = c1.y; q.t2d.v[6] = c1.z; q.t2d.v[7] = c1.w;
}

// S <- a * B by the 8-fold comb: 32 table look-ups, 31 doublings, 31 affine additions
// (edp_BasePointMult with the Z-randomiser taken as 1: the start point is (2x, 2y, 2, 2xy)).
// On return S.t is stale (never needed: every caller encodes or normalises X, Y, Z).
C25519_DEV void ge_base_comb(ge_ext& S, const u32 (&a)[8], const u32* __restrict__ table_smem)
{
    ge_pa q;
    comb_load(q, table_smem, comb8_index(a, 0));
    {
        fe di; fe_from_const(di, kDInv);
        fe_sub(S.x, q.ypx, q.ymx);          // 2x   (canonical inputs: result W)
        fe_add_nn(S.y, q.ypx, q.ymx);       // 2y
        fe_mul(S.t, q.t2d, di);             // 2xy
        fe_set_u32(S.z, 2);
        fe_narrow(S.x); fe_narrow(S.y);
    }
#pragma unroll 1
    for (int i = 1; i < 32; i++) {
        ge_double(S);
        comb_load(q, table_smem, comb8_index(a, i));
        ge_add_affine<false>(S, q);         // next: a doubling or the end -- T is never read again
    }
}

// (x, y) = (X/Z, Y/Z) canonical, and the 32-byte point encoding  y | (x & 1) << 255
// (ecp_Inverse + 2 x ecp_MulMod, ed25519_sign.c:265-267; ed25519_PackPoint curve25519_mehdi.h:130)
C25519_DEV void ge_encode(u32 (&enc)[8], const ge_ext& p)
{
    fe zi, x, y;
    fe_invert(zi, p.z);
    fe_mul(x, p.x, zi); fe_mul(y, p.y, zi);
    fe_canon(x); fe_canon(y);
#pragma unroll
    for (int i = 0; i < 8; i++) enc[i] = y.v[i];
    enc[7] |= (x.v[0] & 1u) << 31;
}

// x with the requested parity such that (x, y) satisfies the curve equation when possible; when no square
// root exists the same deterministic value as the reference's formula is produced (never fails).
C25519_DEV void ge_recover_x(fe& x, const fe& y, u32 parity)
{
    fe u, v, a, b, one, cd;
    fe_set_u32(one, 1); fe_from_const(cd, kD);
    fe_sqr(u, y);
    fe_mul(v, u, cd);
    fe_sub(u, u, one);              // u = y^2 - 1
    fe_add(v, v, one);              // v = d y^2 + 1
    fe_sqr(b, v);
    fe_mul(a, u, b);
    fe_mul(a, a, v);                // a = u v^3
    fe_sqr(b, b);                   // v^4
    fe_mul(b, a, b);                // u v^7
    fe_pow22523(b, b);
    fe_mul(x, b, a);
    fe_sqr(b, x);
    fe_mul(b, b, v);
    fe_sub(b, b, u);
    fe_canon(b);
    if (!fe_is_zero_canon(b)) { fe si; fe_from_const(si, kSqrtM1); fe_mul(x, x, si); }
    fe_canon(x);
    if ((x.v[0] ^ parity) & 1u) fe_neg(x, x);     // p - x (x = 0 gives p, the same field element)
}

}  // namespace c25519
