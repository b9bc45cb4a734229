// x25519_warp.cuh -- ONE X25519 operation per WARP: the north_star's kernel mapping, used where latency is the metric.
//
// Lane layout: lane = 8 * role + limb.  Each of the four 8-lane groups holds a field element with ONE 32-bit limb per
// lane; operands are broadcast with __shfl_sync and carries resolved across lanes (w_carry).  The four groups ("roles")
// execute the four independent field multiplications of a ladder level at the same time and trade results with single
// shuffles, so a ladder step costs three limb-parallel multiplications end to end instead of ten serial ones:
//     level 1:  DA = A.D      CB = B.C      PP = P^2       MM = M^2            (P, M = x+z, x-z of the point being doubled)
//     level 2:  (DA+CB)^2     (DA-CB)^2     PP.MM          121665.(PP-MM)
//     level 3:               u.(DA-CB)^2                   (PP-MM).(PP + 121665 (PP-MM))
// Same rational maps as mont_step_sel (x25519.cuh), i.e. as ecp_Mont / ecp_MontDouble of the reference
// (source/curve25519_dh.c:40-84), hence the same canonical results.  Throughput is a small fraction of the
// thread-per-operation kernel's (profiles/r2_coop_lab.txt: 4.2x fewer multiplications per second for the limb-per-lane
// mapping alone); with fewer operations than the machine has SM sub-partitions (592) nothing else is competing for them.
// Values are kept as fully carried 32-bit limbs of ANY representative below 2^256 (the reference's loose reduction).
#pragma once
#include "fe25519.cuh"

namespace c25519 {

struct wlane { int limb, base, role; };          // base = first lane of this lane's 8-lane group

C25519_DEV wlane w_lane()
{
    wlane c; const int lane = threadIdx.x & 31;
    c.limb = lane & 7; c.base = lane & 24; c.role = lane >> 3;
    return c;
}

// limbs of a 256-bit value held one per lane <- eight 64-bit column sums V_l of weight 2^(32 l); the carry out of limb 7 wraps
// around with x38 (2^256 = 38 mod p).  All four groups iterate together until no lane has a carry left: two passes in the
// common case (after the first, carries are single bits), more only when a carry ripples through all-ones limbs.  It always
// terminates: a pass without wrap-around moves every carry one limb up (at most 8 times), a wrap-around lowers the integer
// value by at least 2^255.  (A ballot-based carry look-ahead was tried: three ballots cost more than the passes they save.)
C25519_DEV u32 w_carry(u64 V, const wlane& c)
{
    while (true) {
        const u32 lo = (u32)V, hi = (u32)(V >> 32);
        const u32 cin = __shfl_sync(0xffffffffu, hi, c.base + ((c.limb + 7) & 7));
        V = (u64)lo + (c.limb == 0 ? (u64)cin * 38ull : (u64)cin);
        if (__ballot_sync(0xffffffffu, (V >> 32) != 0) == 0) break;
    }
    return (u32)V;
}

C25519_DEV u32 w_add(u32 a, u32 b, const wlane& c) { return w_carry((u64)a + b, c); }
// a - b + 2 (2^256 - 38): the constant's limbs (2^33 - 76, then 2^33 - 2) dominate any b limb, so no column goes negative
C25519_DEV u32 w_sub(u32 a, u32 b, const wlane& c)
{
    const u64 K = c.limb == 0 ? 0x1ffffffb4ull : 0x1fffffffeull;
    return w_carry((u64)a + K - b, c);
}

// z = x * y mod (2^256 - 38)          (tools/coop_lab.cu measures exactly this routine against fe_mul)
C25519_DEV u32 w_mul(u32 x, u32 y, const wlane& c)
{
    u64 lo_l = 0, lo_h = 0, hi_l = 0, hi_h = 0;      // columns l and l + 8, 32-bit halves of the products summed apart
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const u32 yj = __shfl_sync(0xffffffffu, y, c.base + j);
        const u32 xr = __shfl_sync(0xffffffffu, x, c.base + ((c.limb - j) & 7));   // x_{l-j}: x_{l-j} y_j lands in column l or l + 8
        const u64 p = (u64)xr * yj;
        if (c.limb >= j) { lo_l += (u32)p; lo_h += p >> 32; } else { hi_l += (u32)p; hi_h += p >> 32; }
    }
    const int below = c.base + ((c.limb + 7) & 7);
    const u64 lo_h_dn = __shfl_sync(0xffffffffu, lo_h, below);
    const u64 hi_h_dn = __shfl_sync(0xffffffffu, hi_h, below);
    const u64 W_lo = lo_l + (c.limb ? lo_h_dn : 0ull);          // word l of the 512-bit product (before carries)
    const u64 W_hi = hi_l + (c.limb ? hi_h_dn : lo_h_dn);       // word l + 8
    return w_carry(W_lo + 38ull * W_hi, c);                     // fold at 2^256; column 15 has no high half (word 16 = 0)
}

C25519_DEV u32 w_pick4(int role, u32 r0, u32 r1, u32 r2, u32 r3)
{
    const u32 a = (role & 1) ? r1 : r0, b = (role & 1) ? r3 : r2;
    return (role & 2) ? b : a;
}
C25519_DEV u32 w_from_role(u32 v, int role, const wlane& c) { return __shfl_sync(0xffffffffu, v, 8 * role + c.limb); }

// one ladder step; every group holds the same (SX, SZ, DX, DZ) on entry and on exit
C25519_DEV void w_mont_step(u32& SX, u32& SZ, u32& DX, u32& DZ, bool dbl_s, u32 u, const wlane& c)
{
    const int r = c.role;
    // level 0: role 0 needs A = SX-SZ and D = DX+DZ, role 1 B = SX+SZ and C = DX-DZ, role 2 P = x+z, role 3 M = x-z of the doubled point
    const u32 px = dbl_s ? SX : DX, pz = dbl_s ? SZ : DZ;
    const u32 a1 = r < 2 ? SX : px, b1 = r < 2 ? SZ : pz;
    const bool sub1 = (r == 0) || (r == 3);                 // role 0: A, role 3: M are differences
    const u64 K = c.limb == 0 ? 0x1ffffffb4ull : 0x1fffffffeull;
    const u32 v1 = w_carry(sub1 ? (u64)a1 + K - b1 : (u64)a1 + b1, c);
    const u32 v2 = w_carry(r == 0 ? (u64)DX + DZ : (u64)DX + K - DZ, c);      // role 0: D, role 1: C (roles 2, 3: unused)
    // level 1: A.D | B.C | P.P | M.M
    const u32 m1 = w_mul(v1, r < 2 ? v2 : v1, c);
    const u32 other = __shfl_xor_sync(0xffffffffu, m1, 8);  // roles 0 <-> 1 trade DA / CB, roles 2 <-> 3 trade PP / MM
    // level 2 operands: role 0: DA+CB, role 1: DA-CB, role 3: E = PP-MM (role 2 multiplies PP.MM as they are)
    const u32 t = w_carry(r == 0 ? (u64)m1 + other : (u64)other + K - m1, c);
    const u32 k121665 = c.limb == 0 ? 121665u : 0u;
    const u32 m2 = w_mul(w_pick4(r, t, t, m1, t), w_pick4(r, t, t, other, k121665), c);
    // level 3: role 1: u.(DA-CB)^2, role 3: E.(PP + 121665 E)   (role 3 holds PP in `other`)
    const u32 F = w_carry((u64)other + m2, c);
    const u32 m3 = w_mul(r == 3 ? t : m2, r == 3 ? F : u, c);
    SX = w_from_role(m2, 0, c); SZ = w_from_role(m3, 1, c); DX = w_from_role(m2, 2, c); DZ = w_from_role(m3, 3, c);
}

// z^(p-2), every group redundantly (254 squarings + 11 multiplications, the chain of fe_invert / ecp_Inverse); 0 -> 0
C25519_DEV u32 w_sqr_n(u32 x, int n, const wlane& c)
{
#pragma unroll 1
    for (int i = 0; i < n; i++) x = w_mul(x, x, c);
    return x;
}
C25519_DEV u32 w_invert(u32 z, const wlane& c)
{
    const u32 z2 = w_mul(z, z, c);
    u32 t = w_sqr_n(z2, 2, c);
    const u32 z9 = w_mul(t, z, c);
    const u32 z11 = w_mul(z9, z2, c);
    t = w_mul(z11, z11, c);
    const u32 a5 = w_mul(t, z9, c);
    t = w_sqr_n(a5, 5, c);   const u32 a10 = w_mul(t, a5, c);
    t = w_sqr_n(a10, 10, c); const u32 a20 = w_mul(t, a10, c);
    t = w_sqr_n(a20, 20, c); t = w_mul(t, a20, c);
    t = w_sqr_n(t, 10, c);   const u32 a50 = w_mul(t, a10, c);
    t = w_sqr_n(a50, 50, c); const u32 a100 = w_mul(t, a50, c);
    t = w_sqr_n(a100, 100, c); t = w_mul(t, a100, c);
    t = w_sqr_n(t, 50, c);   t = w_mul(t, a50, c);            // 2^250 - 1
    t = w_sqr_n(t, 5, c);
    return w_mul(t, z11, c);
}

// canonical x-coordinate of [k]u (k clamped), limb-per-lane; kw(w) = word w of k (every lane holds the scalar)
template <typename KeyWord>
C25519_DEV u32 w_x25519(u32 u, KeyWord kw, const wlane& c)
{
    // P = (u : 1), Q = 2P (ecp_MontDouble: 2S + 2M + 1W), all groups redundantly
    u32 R0X = u, R0Z = c.limb == 0 ? 1u : 0u, R1X, R1Z;
    {
        const u32 a = w_add(R0X, R0Z, c), b = w_sub(R0X, R0Z, c);
        const u32 aa = w_mul(a, a, c), bb = w_mul(b, b, c);
        R1X = w_mul(aa, bb, c);
        const u32 e = w_sub(aa, bb, c);
        const u32 f = w_add(aa, w_mul(e, c.limb == 0 ? 121665u : 0u, c), c);
        R1Z = w_mul(f, e, c);
    }
    bool cur = true;
#pragma unroll 1
    for (int bit = 253; bit >= 0; --bit) {
        const bool b = (kw(bit >> 5) >> (bit & 31)) & 1u;
        const bool s = (b != cur);
        cur = b;
        w_mont_step(R0X, R0Z, R1X, R1Z, s, u, c);
    }
    const u32 PX = cur ? R0X : R1X, PZ = cur ? R0Z : R1Z;
    return w_mul(PX, w_invert(PZ, c), c);
}

}  // namespace c25519
