// sha512.cuh -- SHA-512 for one message per thread (FIPS 180-4), the hash inside Ed25519 keygen / sign /
// verify.  Replaces source/sha512.c of the reference (SHA512_Init :50, SHA512_Update :118, SHA512_Final :67,
// SHA512_Transform :226-294) with a formulation that suits a register machine without byte-addressable
// state: the input is "a short prefix held in registers, followed by message bytes in global memory", which
// is exactly the three shapes Ed25519 needs
//     H(seed)                      prefix = 4 words, no message          (ed25519_sign.c:355-357, :385-387)
//     H(prefix32 || msg)           prefix = 4 words                      (ed25519_sign.c:392-395)
//     H(R || pk || msg)            prefix = 8 words                      (ed25519_sign.c:404-408, ed25519_verify.c:299-303)
// 64-bit words are kept as native u64; ptxas lowers rotates to funnel shifts (SHF.L.W / SHF.R.W).
#pragma once
#include <cstdint>
#include "fe25519.cuh"

namespace c25519 {

__device__ __constant__ const u64 kSha512K[80] = {
    0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull,
    0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull,
    0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull,
    0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,
    0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull, 0x983e5152ee66dfabull,
    0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
    0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull,
    0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
    0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull,
    0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull, 0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull,
    0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull,
    0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
    0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull,
    0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull,
    0x113f9804bef90daeull, 0x1b710b35131c471bull, 0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull,
    0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};

C25519_DEV u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
C25519_DEV u64 bswap64(u64 x)
{
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
}

// one compression: state += F(state, w);  w[16] is consumed as the rolling message schedule
C25519_DEV void sha512_compress(u64 (&st)[8], u64 (&w)[16])
{
    u64 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int r = 0; r < 80; r += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (r) {
                u64 w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                u64 s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
                u64 s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
                w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
            }
            u64 t1 = h + (rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41)) + ((e & f) ^ (~e & g)) + kSha512K[r + j] + w[j];
            u64 t2 = (rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// Big-endian 64-bit word at byte offset `pos` of the padded tail  msg[0..len) || 0x80 || 0x00...
// (the 128-bit length field is patched in by the caller).
C25519_DEV u64 sha512_msg_word(const uint8_t* __restrict__ msg, u64 len, u64 pos, bool aligned8)
{
    if (pos + 8 <= len) {
        if (aligned8) return bswap64(__ldg(reinterpret_cast<const unsigned long long*>(msg + pos)));
        u64 v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) v = (v << 8) | msg[pos + k];
        return v;
    }
    if (pos > len) return 0;
    u64 v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u64 p = pos + k;
        u64 byte = p < len ? (u64)msg[p] : (p == len ? 0x80ull : 0ull);
        v = (v << 8) | byte;
    }
    return v;
}

// digest[8] (big-endian words as u64) = SHA-512( prefix bytes || msg[0..len) )
// PW = number of 64-bit prefix words (4 or 8); prefix[] holds them already big-endian.
template <int PW>
C25519_DEV void sha512_prefixed(u64 (&digest)[8], const u64 (&prefix)[PW], const uint8_t* __restrict__ msg, u64 len)
{
    u64 st[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const u64 total = (u64)PW * 8 + len;
    const u64 nblocks = (total + 17 + 127) / 128;
    const bool aligned8 = ((reinterpret_cast<uintptr_t>(msg) & 7) == 0);     // message words then sit at 8-byte multiples
#pragma unroll 1
    for (u64 blk = 0; blk < nblocks; blk++) {
        u64 w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            u64 v;
            if (j < PW && blk == 0) v = prefix[j];
            else v = sha512_msg_word(msg, len, blk * 128 + (u64)j * 8 - (u64)PW * 8, aligned8);
            if (blk == nblocks - 1) {
                if (j == 14) v = 0;                 // message length < 2^61 bytes
                if (j == 15) v = total << 3;
            }
            w[j] = v;
        }
        sha512_compress(st, w);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) digest[i] = st[i];
}

// the digest as 16 little-endian 32-bit words of its byte string (what ecp_BytesToWords would produce)
C25519_DEV void sha512_digest_to_le_words(u32 (&out)[16], const u64 (&digest)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 le = bswap64(digest[i]);            // bytes of digest word i, first byte in the low bits
        out[2 * i] = (u32)le; out[2 * i + 1] = (u32)(le >> 32);
    }
}

// 32 bytes held as 8 little-endian u32 limbs -> 4 big-endian u64 message words
C25519_DEV void le_limbs_to_be64(u64* out4, const u32* limbs8)
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        out4[i] = ((u64)__byte_perm(limbs8[2 * i], 0, 0x0123) << 32) | (u64)__byte_perm(limbs8[2 * i + 1], 0, 0x0123);
}

}  // namespace c25519
