// normalize.cuh -- batched projective -> affine conversion (Montgomery's simultaneous inversion).
//
// Every operation of the path ends with "invert Z, multiply, canonicalise": ecp_Inverse + ecp_MulMod at
// curve25519_dh.c:148-149 and :176-177, ed25519_sign.c:265-267, ed25519_verify.c:277-279.  One inversion is
// 254 squarings + 11 multiplications, i.e. 9 % of an X25519 ladder and 36 % of a fixed-base operation.  The
// scalar-multiplication kernels therefore stop at the projective result and write it to a scratch record;
// this kernel lets each thread walk K records (strided by the thread count, so every access is coalesced
// across the warp), multiply the Z's together, invert ONCE, and unwind:  4 multiplications + 265/K per
// operation instead of 265.  Z == 0 (low-order inputs; only reachable for X25519 and undecodable Ed25519
// keys) is excluded from the product and yields all-zero coordinates, exactly what inverse(0) = 0 gives in
// the reference.  Outputs are canonical (fe_canon), so they are bit-identical to per-operation inversion.
#pragma once
#include "fe25519.cuh"

namespace c25519 {

// scratch record layouts (8-limb fields, 32 B each)
//   XZ  record (96 B):  [0] X   [1] Z   [2] prefix
//   XYZ record (128 B): [0] X   [1] Y   [2] Z   [3] prefix
constexpr int kScratchXZ = 96;
constexpr int kScratchXYZ = 128;

// Fused gather: the X25519 result of operation i is stored straight into EVERY rank's copy of the gathered result
// array (peer-mapped pointers over NVLink / NVSwitch), at row rank * n_local + i, instead of being all-gathered by a
// separate collective afterwards.
struct PeerScatter {
    uint8_t* dst[8];        // device pointers to the [world * n_local, 32] result arrays of ranks 0..world-1 (own included)
    int world;
    size_t row_offset;      // rank * n_local
};

enum NormalizeMode {
    kNormX = 0,        // out32[i] = X/Z                                   (X25519 ladder; comb public key with X = Z+Y, Z = Z-Y)
    kNormEncode = 1,   // out[i * out_stride .. +32) = encode(X/Z, Y/Z); optionally mirrored into out2
    kNormCompare = 2,  // ok[i] = (encode(X/Z, Y/Z) == cmp[i * cmp_stride .. +32))
};

template <int MODE>
C25519_DEV void normalize_walk(uint8_t* __restrict__ scratch, size_t rec_stride, size_t n, size_t tid, size_t nthreads, int K,
                               uint8_t* __restrict__ out, size_t out_stride, uint8_t* __restrict__ out2, size_t out2_stride,
                               const uint8_t* __restrict__ cmp, size_t cmp_stride, int32_t* __restrict__ ok,
                               const PeerScatter* __restrict__ scatter = nullptr)
{
    constexpr int ZF = (MODE == kNormX) ? 1 : 2;          // field index of Z
    constexpr int PF = ZF + 1;                            // field index of the prefix slot
    fe acc; fe_set_u32(acc, 1);
    int cnt = 0;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const size_t i = tid + (size_t)k * nthreads;
        if (i >= n) break;
        uint8_t* rec = scratch + rec_stride * i;
        fe z, zc;
        fe_load_plain(z, rec + 32 * ZF);
        fe_copy(zc, z); fe_canon(zc);
        if (fe_is_zero_canon(zc)) fe_set_u32(z, 1);
        fe_store(rec + 32 * PF, acc);
        fe_mul(acc, acc, z);
        cnt++;
    }
    if (cnt == 0) return;
    fe inv;
    fe_invert(inv, acc);
#pragma unroll 1
    for (int k = cnt - 1; k >= 0; k--) {
        const size_t i = tid + (size_t)k * nthreads;
        uint8_t* rec = scratch + rec_stride * i;
        fe z, zc, pre, zi, x;
        fe_load_plain(z, rec + 32 * ZF);
        fe_load_plain(pre, rec + 32 * PF);
        fe_load_plain(x, rec);
        fe_copy(zc, z); fe_canon(zc);
        const bool zero = fe_is_zero_canon(zc);
        if (zero) fe_set_u32(z, 1);
        fe_mul(zi, inv, pre);                 // 1 / Z_i
        fe_mul(inv, inv, z);                  // drop Z_i from the running inverse
        if (zero) fe_set_u32(zi, 0);          // inverse(0) = 0, like ecp_Inverse
        fe_mul(x, x, zi);
        fe_canon(x);
        if (MODE == kNormX) {
            if (scatter) {
                for (int g = 0; g < scatter->world; g++) fe_store(scatter->dst[g] + 32 * (scatter->row_offset + i), x);
            } else {
                fe_store(out + out_stride * i, x);
            }
        } else {
            fe y;
            fe_load_plain(y, rec + 32);
            fe_mul(y, y, zi);
            fe_canon(y);
            y.v[7] |= (x.v[0] & 1u) << 31;    // ed25519_PackPoint (curve25519_mehdi.h:130)
            if (MODE == kNormEncode) {
                fe_store(out + out_stride * i, y);
                if (out2) fe_store(out2 + out2_stride * i, y);
            } else {
                fe r; fe_load(r, cmp + cmp_stride * i);
                u32 diff = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) diff |= y.v[w] ^ r.v[w];
                ok[i] = diff == 0 ? 1 : 0;
            }
        }
    }
}

}  // namespace c25519
