/*
 * oracle/oracle25519.c -- CPU RESTATEMENT OF THE REFERENCE'S ALGORITHMS.
 *
 * *** TEST INFRASTRUCTURE ONLY. ***  Nothing under curve25519_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
 *
 * What it is: a plain-C, single-threaded restatement of msotoodeh/curve25519's hot path (X25519
 * variable/fixed base, Ed25519 keygen/sign/verify, SHA-512, arithmetic mod L) that follows the
 * reference's *algorithms and quirks* function by function (each function cites the reference
 * file:line it restates) but shares no code with it: the field is held in 5 x 51-bit limbs with
 * unsigned __int128 products (the reference uses 8 x 32-bit saturated limbs), every constant
 * (d, 2d, 1/d, sqrt(-1), the base point, the 256-entry 8-fold table) is derived at start-up from first
 * principles, and mod-L reduction is a bit-serial shift/subtract.  That independence is the point: the
 * CUDA engine (8 x 32-bit limbs, PTX carry chains) is checked against this file AND against the
 * compiled reference (oracle/_ref/libref25519.so), and this file is itself pinned against the compiled
 * reference and the RFC 7748 / RFC 8032 vectors by tests/test_oracle.py (parity: PINNED).
 *
 * Quirks restated on purpose:
 *   - X25519 loads all 256 bits of the peer u-coordinate (no bit-255 mask)      curve25519_dh.c:104
 *   - the secret key is clamped IN PLACE                                        curve25519_dh.c:186,196,206
 *   - verify is permissive: no S<L check, no point validation, byte compare     ed25519_verify.c:287-313
 * Result-neutral things NOT restated: projective-Z randomisation with edp_custom_blinding.zr
 * (curve25519_dh.c:123, ed25519_sign.c:234-237) and scalar blinding (ed25519_sign.c:246-263);
 * blinding contexts are accepted and ignored.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t fe[5];                       /* value = sum f[i] * 2^(51 i), limbs loosely < 2^54 */
#define M51 ((uint64_t)0x7ffffffffffffULL)

/* ---------------------------------------------------------------- field GF(2^255-19) ---------- */

static void fe_set(fe h, uint64_t v) { h[0] = v; h[1] = h[2] = h[3] = h[4] = 0; }
static void fe_copy(fe h, const fe f) { memcpy(h, f, sizeof(fe)); }

/* all 256 bits, like ecp_BytesToWords (curve25519_utils.c:43-58): top limb receives 52 bits */
static void fe_frombytes256(fe h, const uint8_t *s)
{
    uint64_t w[4];
    for (int i = 0; i < 4; i++) {
        w[i] = 0;
        for (int j = 7; j >= 0; j--) w[i] = (w[i] << 8) | s[8 * i + j];
    }
    h[0] = w[0] & M51;
    h[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    h[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    h[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    h[4] = w[3] >> 12;                        /* 52 bits */
}

static void fe_carry(fe h)
{
    uint64_t c;
    c = h[0] >> 51; h[0] &= M51; h[1] += c;
    c = h[1] >> 51; h[1] &= M51; h[2] += c;
    c = h[2] >> 51; h[2] &= M51; h[3] += c;
    c = h[3] >> 51; h[3] &= M51; h[4] += c;
    c = h[4] >> 51; h[4] &= M51; h[0] += 19 * c;
    c = h[0] >> 51; h[0] &= M51; h[1] += c;
}

/* canonical little-endian encoding in [0,p): what ecp_Mod + ecp_WordsToBytes deliver
   (curve25519_mehdi.c:185-209, curve25519_utils.c:61-75) */
static void fe_tobytes(uint8_t *s, const fe f)
{
    fe t; fe_copy(t, f); fe_carry(t); fe_carry(t);
    /* t < 2^255 + small; subtract p if t >= p, done by adding 19 and looking at bit 255 */
    uint64_t q = (t[0] + 19) >> 51;
    q = (t[1] + q) >> 51; q = (t[2] + q) >> 51; q = (t[3] + q) >> 51; q = (t[4] + q) >> 51;
    t[0] += 19 * q;
    uint64_t c;
    c = t[0] >> 51; t[0] &= M51; t[1] += c;
    c = t[1] >> 51; t[1] &= M51; t[2] += c;
    c = t[2] >> 51; t[2] &= M51; t[3] += c;
    c = t[3] >> 51; t[3] &= M51; t[4] += c;
    t[4] &= M51;
    uint64_t w[4];
    w[0] = t[0] | (t[1] << 51);
    w[1] = (t[1] >> 13) | (t[2] << 38);
    w[2] = (t[2] >> 26) | (t[3] << 25);
    w[3] = (t[3] >> 39) | (t[4] << 12);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

static void fe_add(fe h, const fe f, const fe g)          /* ecp_AddReduce, curve25519_mehdi.c:134 */
{ for (int i = 0; i < 5; i++) h[i] = f[i] + g[i]; fe_carry(h); }

static void fe_sub(fe h, const fe f, const fe g)          /* ecp_SubReduce, curve25519_mehdi.c:161 */
{
    /* add 8p so limbs stay non-negative for g limbs < 2^54 */
    h[0] = f[0] + 0x3fffffffffff68ULL - g[0];
    for (int i = 1; i < 5; i++) h[i] = f[i] + 0x3ffffffffffff8ULL - g[i];
    fe_carry(h);
}

static void fe_mul(fe h, const fe f, const fe g)          /* ecp_MulReduce, curve25519_mehdi.c:278 */
{
    u128 t0, t1, t2, t3, t4;
    uint64_t g1 = 19 * g[1], g2 = 19 * g[2], g3 = 19 * g[3], g4 = 19 * g[4];
    t0 = (u128)f[0] * g[0] + (u128)f[1] * g4 + (u128)f[2] * g3 + (u128)f[3] * g2 + (u128)f[4] * g1;
    t1 = (u128)f[0] * g[1] + (u128)f[1] * g[0] + (u128)f[2] * g4 + (u128)f[3] * g3 + (u128)f[4] * g2;
    t2 = (u128)f[0] * g[2] + (u128)f[1] * g[1] + (u128)f[2] * g[0] + (u128)f[3] * g4 + (u128)f[4] * g3;
    t3 = (u128)f[0] * g[3] + (u128)f[1] * g[2] + (u128)f[2] * g[1] + (u128)f[3] * g[0] + (u128)f[4] * g4;
    t4 = (u128)f[0] * g[4] + (u128)f[1] * g[3] + (u128)f[2] * g[2] + (u128)f[3] * g[1] + (u128)f[4] * g[0];
    uint64_t c;
    t1 += (uint64_t)(t0 >> 51); h[0] = (uint64_t)t0 & M51;
    t2 += (uint64_t)(t1 >> 51); h[1] = (uint64_t)t1 & M51;
    t3 += (uint64_t)(t2 >> 51); h[2] = (uint64_t)t2 & M51;
    t4 += (uint64_t)(t3 >> 51); h[3] = (uint64_t)t3 & M51;
    c = (uint64_t)(t4 >> 51);   h[4] = (uint64_t)t4 & M51;
    h[0] += 19 * c;
    c = h[0] >> 51; h[0] &= M51; h[1] += c;
}

static void fe_sq(fe h, const fe f) { fe_mul(h, f, f); }  /* ecp_SqrReduce, curve25519_mehdi.c:310 */

static void fe_mul_small(fe h, const fe f, uint64_t b)    /* the b*X part of ecp_WordMulAddReduce :243 */
{
    u128 t; uint64_t c = 0;
    for (int i = 0; i < 5; i++) { t = (u128)f[i] * b + c; h[i] = (uint64_t)t & M51; c = (uint64_t)(t >> 51); }
    h[0] += 19 * c; fe_carry(h);
}

static void fe_sqn(fe h, const fe f, int n) { fe_sq(h, f); while (--n > 0) fe_sq(h, h); }

/* z^(2^250-1) and z^11: shared head of ecp_Inverse (curve25519_mehdi.c:340-409) and
   ecp_ModExp2523 (ed25519_verify.c:116-135) */
static void fe_pow_2_250_1(fe out, fe z11, const fe z)
{
    fe z2, z9, t, a5, a10, a20, a50, a100;
    fe_sq(z2, z); fe_sqn(t, z2, 2); fe_mul(z9, t, z); fe_mul(z11, z9, z2);
    fe_sq(t, z11); fe_mul(a5, t, z9);
    fe_sqn(t, a5, 5); fe_mul(a10, t, a5);
    fe_sqn(t, a10, 10); fe_mul(a20, t, a10);
    fe_sqn(t, a20, 20); fe_mul(t, t, a20);
    fe_sqn(t, t, 10); fe_mul(a50, t, a10);
    fe_sqn(t, a50, 50); fe_mul(a100, t, a50);
    fe_sqn(t, a100, 100); fe_mul(t, t, a100);
    fe_sqn(t, t, 50); fe_mul(out, t, a50);
}
static void fe_invert(fe out, const fe z)                 /* z^(p-2); 0 -> 0.  curve25519_mehdi.c:340 */
{ fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqn(t, t, 5); fe_mul(out, t, z11); }
static void fe_pow22523(fe out, const fe z)               /* z^((p-5)/8).  ed25519_verify.c:116 */
{ fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqn(t, t, 2); fe_mul(out, t, z); }

static int fe_iszero(const fe f)
{ uint8_t s[32]; fe_tobytes(s, f); uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i]; return r == 0; }
static int fe_parity(const fe f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static void fe_neg(fe h, const fe f) { fe z; fe_set(z, 0); fe_sub(h, z, f); }

/* ---------------------------------------------------------------- SHA-512 (sha512.c:50-294) --- */

typedef struct { uint64_t h[8]; uint8_t buf[128]; size_t fill; uint64_t total; } sha512_ctx;
static const uint64_t K512[80] = {
0x428a2f98d728ae22ULL,0x7137449123ef65cdULL,0xb5c0fbcfec4d3b2fULL,0xe9b5dba58189dbbcULL,0x3956c25bf348b538ULL,
0x59f111f1b605d019ULL,0x923f82a4af194f9bULL,0xab1c5ed5da6d8118ULL,0xd807aa98a3030242ULL,0x12835b0145706fbeULL,
0x243185be4ee4b28cULL,0x550c7dc3d5ffb4e2ULL,0x72be5d74f27b896fULL,0x80deb1fe3b1696b1ULL,0x9bdc06a725c71235ULL,
0xc19bf174cf692694ULL,0xe49b69c19ef14ad2ULL,0xefbe4786384f25e3ULL,0x0fc19dc68b8cd5b5ULL,0x240ca1cc77ac9c65ULL,
0x2de92c6f592b0275ULL,0x4a7484aa6ea6e483ULL,0x5cb0a9dcbd41fbd4ULL,0x76f988da831153b5ULL,0x983e5152ee66dfabULL,
0xa831c66d2db43210ULL,0xb00327c898fb213fULL,0xbf597fc7beef0ee4ULL,0xc6e00bf33da88fc2ULL,0xd5a79147930aa725ULL,
0x06ca6351e003826fULL,0x142929670a0e6e70ULL,0x27b70a8546d22ffcULL,0x2e1b21385c26c926ULL,0x4d2c6dfc5ac42aedULL,
0x53380d139d95b3dfULL,0x650a73548baf63deULL,0x766a0abb3c77b2a8ULL,0x81c2c92e47edaee6ULL,0x92722c851482353bULL,
0xa2bfe8a14cf10364ULL,0xa81a664bbc423001ULL,0xc24b8b70d0f89791ULL,0xc76c51a30654be30ULL,0xd192e819d6ef5218ULL,
0xd69906245565a910ULL,0xf40e35855771202aULL,0x106aa07032bbd1b8ULL,0x19a4c116b8d2d0c8ULL,0x1e376c085141ab53ULL,
0x2748774cdf8eeb99ULL,0x34b0bcb5e19b48a8ULL,0x391c0cb3c5c95a63ULL,0x4ed8aa4ae3418acbULL,0x5b9cca4f7763e373ULL,
0x682e6ff3d6b2b8a3ULL,0x748f82ee5defb2fcULL,0x78a5636f43172f60ULL,0x84c87814a1f0ab72ULL,0x8cc702081a6439ecULL,
0x90befffa23631e28ULL,0xa4506cebde82bde9ULL,0xbef9a3f7b2c67915ULL,0xc67178f2e372532bULL,0xca273eceea26619cULL,
0xd186b8c721c0c207ULL,0xeada7dd6cde0eb1eULL,0xf57d4f7fee6ed178ULL,0x06f067aa72176fbaULL,0x0a637dc5a2c898a6ULL,
0x113f9804bef90daeULL,0x1b710b35131c471bULL,0x28db77f523047d84ULL,0x32caab7b40c72493ULL,0x3c9ebe0a15c9bebcULL,
0x431d67c49c100d4cULL,0x4cc5d4becb3e42b6ULL,0x597f299cfc657e2aULL,0x5fcb6fab3ad6faecULL,0x6c44198c4a475817ULL };
#define ROR(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
static void sha512_block(uint64_t h[8], const uint8_t *p)
{
    uint64_t w[80], a, b, c, d, e, f, g, hh;
    for (int i = 0; i < 16; i++) { w[i] = 0; for (int j = 0; j < 8; j++) w[i] = (w[i] << 8) | p[8 * i + j]; }
    for (int i = 16; i < 80; i++) {
        uint64_t s0 = ROR(w[i - 15], 1) ^ ROR(w[i - 15], 8) ^ (w[i - 15] >> 7);
        uint64_t s1 = ROR(w[i - 2], 19) ^ ROR(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    a = h[0]; b = h[1]; c = h[2]; d = h[3]; e = h[4]; f = h[5]; g = h[6]; hh = h[7];
    for (int i = 0; i < 80; i++) {
        uint64_t t1 = hh + (ROR(e, 14) ^ ROR(e, 18) ^ ROR(e, 41)) + ((e & f) ^ (~e & g)) + K512[i] + w[i];
        uint64_t t2 = (ROR(a, 28) ^ ROR(a, 34) ^ ROR(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha512_init(sha512_ctx *c)
{
    static const uint64_t iv[8] = { 0x6a09e667f3bcc908ULL,0xbb67ae8584caa73bULL,0x3c6ef372fe94f82bULL,0xa54ff53a5f1d36f1ULL,
                                    0x510e527fade682d1ULL,0x9b05688c2b3e6c1fULL,0x1f83d9abfb41bd6bULL,0x5be0cd19137e2179ULL };
    memcpy(c->h, iv, sizeof iv); c->fill = 0; c->total = 0;
}
static void sha512_update(sha512_ctx *c, const void *data, size_t n)
{
    const uint8_t *p = (const uint8_t *)data; c->total += n;
    while (n) {
        size_t k = 128 - c->fill; if (k > n) k = n;
        memcpy(c->buf + c->fill, p, k); c->fill += k; p += k; n -= k;
        if (c->fill == 128) { sha512_block(c->h, c->buf); c->fill = 0; }
    }
}
static void sha512_final(uint8_t out[64], sha512_ctx *c)
{
    uint64_t bits = c->total * 8; uint8_t pad[240]; size_t k = (c->fill < 112) ? 112 - c->fill : 240 - c->fill;
    memset(pad, 0, sizeof pad); pad[0] = 0x80;
    uint8_t len[16]; memset(len, 0, 16); for (int i = 0; i < 8; i++) len[15 - i] = (uint8_t)(bits >> (8 * i));
    uint64_t keep = c->total; sha512_update(c, pad, k); sha512_update(c, len, 16); c->total = keep;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c->h[i] >> (56 - 8 * j));
}
/* exported so tests can pin it against the reference's SHA-512 KATs (test/curve25519_selftest.c:131-141) */
void orc_sha512(uint8_t out[64], const uint8_t *msg, size_t n)
{ sha512_ctx c; sha512_init(&c); sha512_update(&c, msg, n); sha512_final(out, &c); }

/* ---------------------------------------------------------------- arithmetic mod L ------------ */
/* L = 2^252 + 27742317777372353535851937790883648493 (curve25519_order.c:30-31).  The reference
   folds one 32-bit word at a time (eco_ReduceHiWord :80) and canonicalises with eco_Mod (:125); every
   value that reaches an output is canonical, so a bit-serial shift/subtract gives identical bytes. */
static const uint64_t L64[4] = { 0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL };

static void sc_reduce_bits(uint8_t out[32], const uint8_t *in, int nbytes)   /* eco_DigestToWords+eco_Mod :139,:125 */
{
    uint64_t r[4] = { 0, 0, 0, 0 };
    for (int bit = nbytes * 8 - 1; bit >= 0; bit--) {
        uint64_t b = (in[bit >> 3] >> (bit & 7)) & 1;
        r[3] = (r[3] << 1) | (r[2] >> 63); r[2] = (r[2] << 1) | (r[1] >> 63);
        r[1] = (r[1] << 1) | (r[0] >> 63); r[0] = (r[0] << 1) | b;
        uint64_t t[4]; u128 br = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)r[i] - L64[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
        if (!br) memcpy(r, t, sizeof r);
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(r[i] >> (8 * j));
}
/* s = (h*a + r) mod L: eco_MulReduce + eco_AddReduce + eco_Mod (curve25519_order.c:110,132,125) */
static void sc_muladd(uint8_t s[32], const uint8_t h[32], const uint8_t a[32], const uint8_t r[32])
{
    uint32_t hw[8], aw[8];
    for (int i = 0; i < 8; i++) {
        hw[i] = (uint32_t)h[4*i] | (uint32_t)h[4*i+1] << 8 | (uint32_t)h[4*i+2] << 16 | (uint32_t)h[4*i+3] << 24;
        aw[i] = (uint32_t)a[4*i] | (uint32_t)a[4*i+1] << 8 | (uint32_t)a[4*i+2] << 16 | (uint32_t)a[4*i+3] << 24;
    }
    uint32_t prod[17]; memset(prod, 0, sizeof prod);
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { uint64_t t = (uint64_t)hw[i] * aw[j] + prod[i + j] + c; prod[i + j] = (uint32_t)t; c = t >> 32; }
        prod[i + 8] = (uint32_t)c;
    }
    uint64_t c = 0;
    for (int i = 0; i < 17; i++) {
        uint64_t rw = (i < 8) ? ((uint32_t)r[4*i] | (uint32_t)r[4*i+1] << 8 | (uint32_t)r[4*i+2] << 16 | (uint32_t)r[4*i+3] << 24) : 0;
        uint64_t t = (uint64_t)prod[i] + rw + c; prod[i] = (uint32_t)t; c = t >> 32;
    }
    uint8_t wide[68];
    for (int i = 0; i < 17; i++) for (int j = 0; j < 4; j++) wide[4 * i + j] = (uint8_t)(prod[i] >> (8 * j));
    sc_reduce_bits(s, wide, 68);
}

/* ---------------------------------------------------------------- Edwards layer --------------- */

typedef struct { fe x, y, z, t; } ext_pt;            /* Ext_POINT  curve25519_mehdi.h:60-65 */
typedef struct { fe YpX, YmX, T2d, Z2; } pe_pt;      /* PE_POINT   curve25519_mehdi.h:68-74 */
typedef struct { fe YpX, YmX, T2d; } pa_pt;          /* PA_POINT   curve25519_mehdi.h:77-82 */

static fe C_d, C_2d, C_di, C_I, C_one, C_zero, C_By, C_Bx;
static pa_pt base8[256];                             /* _w_base_folding8, base_folding8.h:6-1288 */
static int consts_ready = 0;

static void fe_canon(fe h) { uint8_t s[32]; fe_tobytes(s, h); fe_frombytes256(h, s); }

/* edp_DoublePoint ed25519_sign.c:122-143 (dbl-2008-hwcd, a = -1) */
static void ed_double(ext_pt *p)
{
    fe a, b, c, d, e;
    fe_sq(a, p->x); fe_sq(b, p->y); fe_sq(c, p->z); fe_add(c, c, c);
    fe_neg(d, a);
    fe_sub(a, d, b); fe_add(d, d, b); fe_sub(b, d, c);
    fe_add(e, p->x, p->y); fe_sq(e, e); fe_add(e, e, a);
    fe_mul(p->x, e, b); fe_mul(p->y, a, d); fe_mul(p->z, d, b); fe_mul(p->t, e, a);
}
/* edp_AddPoint ed25519_verify.c:142-161; q2z == NULL means Z2 = 1 => D = 2*Z1 = edp_AddAffinePoint ed25519_sign.c:97-115 */
static void ed_add_pre(ext_pt *r, const ext_pt *p, const fe qYpX, const fe qYmX, const fe qT2d, const fe *qZ2)
{
    fe a, b, c, d, e;
    fe_sub(a, p->y, p->x); fe_mul(a, a, qYmX);
    fe_add(b, p->y, p->x); fe_mul(b, b, qYpX);
    fe_mul(c, p->t, qT2d);
    if (qZ2) fe_mul(d, p->z, *qZ2); else fe_add(d, p->z, p->z);
    fe_sub(e, b, a); fe_add(b, b, a); fe_sub(a, d, c); fe_add(d, d, c);
    fe_mul(r->x, e, a); fe_mul(r->y, b, d); fe_mul(r->t, e, b); fe_mul(r->z, d, a);
}
static void ed_add_affine(ext_pt *p, const pa_pt *q) { ed_add_pre(p, p, q->YpX, q->YmX, q->T2d, 0); }
static void ed_add_pe(ext_pt *r, const ext_pt *p, const pe_pt *q) { ed_add_pre(r, p, q->YpX, q->YmX, q->T2d, &q->Z2); }
/* edp_ExtPoint2PE ed25519_sign.c:270-276 */
static void ed_ext2pe(pe_pt *r, const ext_pt *p)
{ fe_add(r->YpX, p->y, p->x); fe_sub(r->YmX, p->y, p->x); fe_mul(r->T2d, p->t, C_2d); fe_add(r->Z2, p->z, p->z); }

/* ed25519_CalculateX ed25519_verify.c:66-100: x = sqrt((y^2-1)/(d y^2+1)) with the requested parity; never fails */
static void ed_calc_x(fe X, const fe Y, int parity)
{
    fe u, v, a, b;
    fe_sq(u, Y); fe_mul(v, u, C_d); fe_sub(u, u, C_one); fe_add(v, v, C_one);
    fe_sq(b, v); fe_mul(a, u, b); fe_mul(a, a, v); fe_sq(b, b); fe_mul(b, a, b);
    fe_pow22523(b, b); fe_mul(X, b, a);
    fe_sq(b, X); fe_mul(b, b, v); fe_sub(b, b, u);
    if (!fe_iszero(b)) fe_mul(X, X, C_I);
    fe_canon(X);
    if ((fe_parity(X) ^ parity) & 1) fe_neg(X, X);        /* p - X; X = 0 gives p == 0 (mod p), same field element */
}

__attribute__((constructor)) static void init_consts(void)
{
    if (consts_ready) return;
    fe t, u;
    fe_set(C_one, 1); fe_set(C_zero, 0);
    /* d = -121665/121666 (ed25519_sign.c:31-33) */
    fe_set(t, 121666); fe_invert(t, t); fe_set(u, 121665); fe_mul(t, t, u); fe_neg(C_d, t); fe_canon(C_d);
    fe_add(C_2d, C_d, C_d); fe_canon(C_2d);              /* _w_2d ed25519_sign.c:59 */
    fe_invert(C_di, C_d); fe_canon(C_di);                /* _w_di ed25519_sign.c:61 */
    /* sqrt(-1) = 2^((p-1)/4) = 2 * (2^((p-5)/8))^2 ... computed as 2^((p-1)/4) via (p-1)/4 = 2*((p-5)/8) + 1 (_w_I ed25519_verify.c:60) */
    fe_set(t, 2); fe_pow22523(u, t); fe_sq(u, u); fe_mul(C_I, u, t); fe_canon(C_I);
    /* base point y = 4/5, x even (ed25519_sign.c:35-37) */
    fe_set(t, 5); fe_invert(t, t); fe_set(u, 4); fe_mul(C_By, t, u); fe_canon(C_By);
    ed_calc_x(C_Bx, C_By, 0); fe_canon(C_Bx);
    /* 8-fold table: entry i = sum_{k in bits(i)} 2^(32k) B as canonical (Y+X, Y-X, 2dT), entry 0 = identity
       (generator in the reference: test/curve25519_selftest.c:498-551) */
    ext_pt P[8], acc[256];
    fe_copy(P[0].x, C_Bx); fe_copy(P[0].y, C_By); fe_set(P[0].z, 1); fe_mul(P[0].t, C_Bx, C_By);
    for (int k = 1; k < 8; k++) { P[k] = P[k - 1]; for (int i = 0; i < 32; i++) ed_double(&P[k]); }
    fe_set(acc[0].x, 0); fe_set(acc[0].y, 1); fe_set(acc[0].z, 1); fe_set(acc[0].t, 0);
    for (int i = 1; i < 256; i++) {
        int k = 0; while (!((i >> k) & 1)) k++;
        int rest = i & (i - 1);
        if (!rest) { acc[i] = P[k]; }
        else { pe_pt q; ed_ext2pe(&q, &P[k]); ed_add_pe(&acc[i], &acc[rest], &q); }
    }
    for (int i = 0; i < 256; i++) {
        fe zi, x, y; fe_invert(zi, acc[i].z); fe_mul(x, acc[i].x, zi); fe_mul(y, acc[i].y, zi);
        fe_add(base8[i].YpX, y, x); fe_sub(base8[i].YmX, y, x); fe_mul(t, x, y); fe_mul(base8[i].T2d, t, C_2d);
        fe_canon(base8[i].YpX); fe_canon(base8[i].YmX); fe_canon(base8[i].T2d);
    }
    consts_ready = 1;
}
/* test hook: canonical 96-byte image of table entry i, comparable with _w_base_folding8[i] */
void orc_base_table_entry(uint8_t out[96], int i)
{ init_consts(); fe_tobytes(out, base8[i].YpX); fe_tobytes(out + 32, base8[i].YmX); fe_tobytes(out + 64, base8[i].T2d); }

/* ecp_8Folds curve25519_utils.c:144-153: cut[i] bit k = bit (31-i) of 32-bit word k */
static void folds8(uint8_t cut[32], const uint8_t s[32])
{
    for (int i = 0; i < 32; i++) {
        int bit = 31 - i; uint8_t a = 0;
        for (int k = 7; k >= 0; k--) a = (uint8_t)((a << 1) | ((s[4 * k + (bit >> 3)] >> (bit & 7)) & 1));
        cut[i] = a;
    }
}
/* ecp_4Folds curve25519_utils.c:125-142: v[i] bit k = bit (63-i) of 64-bit limb k */
static void folds4(uint8_t v[64], const uint8_t s[32])
{
    for (int i = 0; i < 64; i++) {
        int bit = 63 - i; uint8_t a = 0;
        for (int k = 3; k >= 0; k--) a = (uint8_t)((a << 1) | ((s[8 * k + (bit >> 3)] >> (bit & 7)) & 1));
        v[i] = a;
    }
}

/* edp_BasePointMult ed25519_sign.c:215-244 (Z-randomiser R taken as 1: start point is (2x,2y,2,2xy)) */
static void ed_base_mult(ext_pt *S, const uint8_t sk[32])
{
    uint8_t cut[32]; init_consts(); folds8(cut, sk);
    const pa_pt *p0 = &base8[cut[0]];
    fe_sub(S->x, p0->YpX, p0->YmX); fe_add(S->y, p0->YpX, p0->YmX); fe_mul(S->t, p0->T2d, C_di); fe_set(S->z, 2);
    for (int i = 1; i < 32; i++) { ed_double(S); ed_add_affine(S, &base8[cut[i]]); }
}
/* edp_BasePointMultiply ed25519_sign.c:246-268 + ed25519_PackPoint curve25519_mehdi.h:130 */
static void ed_base_mult_pack(uint8_t out[32], const uint8_t sk[32])
{
    ext_pt S; fe zi, x, y; ed_base_mult(&S, sk);
    fe_invert(zi, S.z); fe_mul(x, S.x, zi); fe_mul(y, S.y, zi);
    fe_tobytes(out, y); out[31] = (uint8_t)((out[31] & 0x7f) | (fe_parity(x) << 7));
}

/* ---------------------------------------------------------------- X25519 ---------------------- */

static void trim(uint8_t *k) { k[0] &= 0xf8; k[31] = (uint8_t)((k[31] | 0x40) & 0x7f); }   /* ecp_TrimSecretKey utils.c:28 */

/* ecp_MontDouble curve25519_dh.c:40-54 */
static void mont_double(fe X2, fe Z2, const fe X, const fe Z)
{
    fe A, B; fe_add(A, X, Z); fe_sub(B, X, Z); fe_sq(A, A); fe_sq(B, B); fe_mul(X2, A, B);
    fe_sub(B, A, B); fe t; fe_mul_small(t, B, 121665); fe_add(A, A, t); fe_mul(Z2, A, B);
}
/* ecp_Mont curve25519_dh.c:57-84: (P,Q) <- (P+Q, 2Q), difference = (base:1) */
static void mont_step(fe PX, fe PZ, fe QX, fe QZ, const fe base)
{
    fe A, B, C, D, E, t;
    fe_sub(A, PX, PZ); fe_add(B, PX, PZ); fe_sub(C, QX, QZ); fe_add(D, QX, QZ);
    fe_mul(A, A, D); fe_mul(B, B, C); fe_add(E, A, B); fe_sub(B, A, B);
    fe_sq(PX, E); fe_sq(A, B); fe_mul(PZ, A, base);
    fe_sq(A, D); fe_sq(B, C); fe_mul(QX, A, B); fe_sub(B, A, B);
    fe_mul_small(t, B, 121665); fe_add(A, A, t); fe_mul(QZ, A, B);
}
/* ecp_PointMultiply curve25519_dh.c:94-157 */
void ecp_PointMultiply(uint8_t *out, const uint8_t *point, const uint8_t *scalar, int len)
{
    fe X, PX, PZ, QX, QZ;
    fe_frombytes256(X, point);                              /* all 256 bits: dh.c:104 */
    int top = len * 8 - 1;
    while (top >= 0 && !((scalar[top >> 3] >> (top & 7)) & 1)) top--;
    if (top < 0) { memset(out, 0, 32); return; }            /* dh.c:156 */
    fe_copy(PX, X); fe_set(PZ, 1);                          /* dh.c:123-124 with Z = 1 */
    mont_double(QX, QZ, PX, PZ);                            /* dh.c:125 */
    for (int bit = top - 1; bit >= 0; bit--) {
        if ((scalar[bit >> 3] >> (bit & 7)) & 1) mont_step(PX, PZ, QX, QZ, X);   /* PP[1]=&P QP[1]=&Q dh.c:127 */
        else                                      mont_step(QX, QZ, PX, PZ, X);   /* PP[0]=&Q QP[0]=&P dh.c:128 */
    }
    fe zi; fe_invert(zi, PZ); fe_mul(X, PX, zi); fe_tobytes(out, X);   /* dh.c:148-150 */
}
/* x25519_BasePointMultiply curve25519_dh.c:162-179: u = (Z+Y)/(Z-Y) */
static void x25519_base(uint8_t *r, const uint8_t *sk)
{
    ext_pt S; fe n, dn; ed_base_mult(&S, sk);
    fe_add(n, S.z, S.y); fe_sub(dn, S.z, S.y); fe_invert(dn, dn); fe_mul(n, n, dn); fe_tobytes(r, n);
}
static const uint8_t nine[32] = { 9 };
void curve25519_dh_CalculatePublicKey_fast(unsigned char *pk, unsigned char *sk) { trim(sk); x25519_base(pk, sk); }             /* dh.c:182 */
void curve25519_dh_CalculatePublicKey(unsigned char *pk, unsigned char *sk) { trim(sk); ecp_PointMultiply(pk, nine, sk, 32); }   /* dh.c:192 */
void curve25519_dh_CreateSharedKey(unsigned char *sh, const unsigned char *pk, unsigned char *sk) { trim(sk); ecp_PointMultiply(sh, pk, sk, 32); } /* dh.c:201 */

/* ---------------------------------------------------------------- Ed25519 --------------------- */

/* ed25519_CreateKeyPair ed25519_sign.c:344-367 */
void ed25519_CreateKeyPair(unsigned char *pub, unsigned char *priv, const void *blinding, const unsigned char *sk)
{
    uint8_t md[64]; (void)blinding;
    orc_sha512(md, sk, 32); trim(md);
    ed_base_mult_pack(pub, md);
    memmove(priv, sk, 32); memcpy(priv + 32, pub, 32);
}
/* ed25519_SignMessage ed25519_sign.c:372-419 */
void ed25519_SignMessage(unsigned char *sig, const unsigned char *priv, const void *blinding, const unsigned char *msg, size_t n)
{
    uint8_t md[64], a[32], r[32], h[32], Renc[32]; sha512_ctx c; (void)blinding;
    orc_sha512(md, priv, 32); trim(md); memcpy(a, md, 32);
    sha512_init(&c); sha512_update(&c, md + 32, 32); sha512_update(&c, msg, n); sha512_final(md, &c);
    sc_reduce_bits(r, md, 64);
    ed_base_mult_pack(Renc, r);
    sha512_init(&c); sha512_update(&c, Renc, 32); sha512_update(&c, priv + 32, 32); sha512_update(&c, msg, n); sha512_final(md, &c);
    sc_reduce_bits(h, md, 64);                  /* the reference keeps h only loosely reduced (:409); S is canonical either way */
    memcpy(sig, Renc, 32);
    sc_muladd(sig + 32, h, a, r);
}
/* blinding contexts are result-neutral (test/curve25519_test.c:371-393): accepted, zero-filled, ignored */
void *ed25519_Blinding_Init(void *context, const unsigned char *seed, size_t size)
{ (void)seed; (void)size; if (!context) context = malloc(192); if (context) memset(context, 0, 192); return context; }
void ed25519_Blinding_Finish(void *context) { if (context) { memset(context, 0, 192); free(context); } }

typedef struct { uint8_t pk[32]; pe_pt q[16]; } verify_ctx;   /* EDP_SIGV_CTX ed25519_verify.c:44-47 (our layout is larger; opaque) */

/* ed25519_Verify_Init ed25519_verify.c:179-232: table of subset sums of 2^(64k) * (-A), k = 0..3 */
void *ed25519_Verify_Init(void *context, const unsigned char *pk)
{
    verify_ctx *ctx = (verify_ctx *)context; init_consts();
    if (!ctx) ctx = (verify_ctx *)malloc(sizeof *ctx);
    if (!ctx) return 0;
    ext_pt Q, T; uint8_t yb[32];
    memcpy(ctx->pk, pk, 32);
    memcpy(yb, pk, 32); int sign = yb[31] >> 7; yb[31] &= 0x7f;          /* ecp_DecodeInt utils.c:100 */
    fe_frombytes256(Q.y, yb);
    ed_calc_x(Q.x, Q.y, (~sign) & 1);                                     /* inverted parity: -A  (:193) */
    fe_mul(Q.t, Q.x, Q.y); fe_set(Q.z, 1);
    fe_set(ctx->q[0].YpX, 1); fe_set(ctx->q[0].YmX, 1); fe_set(ctx->q[0].T2d, 0); fe_set(ctx->q[0].Z2, 2);
    ed_ext2pe(&ctx->q[1], &Q);
    for (int lvl = 1; lvl < 4; lvl++) {
        for (int i = 0; i < 64; i++) ed_double(&Q);
        int base = 1 << lvl;
        ed_ext2pe(&ctx->q[base], &Q);
        for (int s = 1; s < base; s++) { ed_add_pe(&T, &Q, &ctx->q[s]); ed_ext2pe(&ctx->q[base + s], &T); }   /* QTABLE_SET :175 */
    }
    return ctx;
}
void ed25519_Verify_Finish(void *ctx) { free(ctx); }
/* ed25519_Verify_Check ed25519_verify.c:287-313 with edp_PolyPointMultiply :243-280 */
int ed25519_Verify_Check(const void *context, const unsigned char *sig, const unsigned char *msg, size_t n)
{
    const verify_ctx *ctx = (const verify_ctx *)context; sha512_ctx c; uint8_t md[64], h[32], u[32], v[64], enc[32];
    sha512_init(&c); sha512_update(&c, sig, 32); sha512_update(&c, ctx->pk, 32); sha512_update(&c, msg, n); sha512_final(md, &c);
    sc_reduce_bits(h, md, 64);
    folds8(u, sig + 32);                                   /* S used raw, all 256 bits (:308) */
    folds4(v, h);
    ext_pt S; const pe_pt *q0 = &ctx->q[v[0]];
    fe_sub(S.x, q0->YpX, q0->YmX); fe_add(S.y, q0->YpX, q0->YmX); fe_mul(S.t, q0->T2d, C_di); fe_copy(S.z, q0->Z2);
    int i = 1;
    for (; i < 32; i++) { ed_double(&S); ed_add_pe(&S, &S, &ctx->q[v[i]]); }
    for (; i < 64; i++) { ed_double(&S); ed_add_affine(&S, &base8[u[i - 32]]); ed_add_pe(&S, &S, &ctx->q[v[i]]); }
    fe zi, x, y; fe_invert(zi, S.z); fe_mul(x, S.x, zi); fe_mul(y, S.y, zi);
    fe_tobytes(enc, y); enc[31] = (uint8_t)((enc[31] & 0x7f) | (fe_parity(x) << 7));
    return memcmp(enc, sig, 32) == 0;
}
/* ed25519_VerifySignature ed25519_verify.c:163-173 */
int ed25519_VerifySignature(const unsigned char *sig, const unsigned char *pk, const unsigned char *msg, size_t n)
{ verify_ctx ctx; ed25519_Verify_Init(&ctx, pk); return ed25519_Verify_Check(&ctx, sig, msg, n); }

/* ---------------------------------------------------------------- unit-test hooks ------------- */
/* canonical results of single field / scalar operations, for differential tests against the
   reference's ecp_MulReduce+ecp_Mod, eco_MulReduce+eco_Mod, ecp_8Folds, ... */
void orc_fe_mul(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe x, y; fe_frombytes256(x, a); fe_frombytes256(y, b); fe_mul(x, x, y); fe_tobytes(out, x); }
void orc_fe_add(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe x, y; fe_frombytes256(x, a); fe_frombytes256(y, b); fe_add(x, x, y); fe_tobytes(out, x); }
void orc_fe_sub(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe x, y; fe_frombytes256(x, a); fe_frombytes256(y, b); fe_sub(x, x, y); fe_tobytes(out, x); }
void orc_fe_inv(uint8_t out[32], const uint8_t a[32]) { fe x; fe_frombytes256(x, a); fe_invert(x, x); fe_tobytes(out, x); }
void orc_fe_pow22523(uint8_t out[32], const uint8_t a[32]) { fe x; fe_frombytes256(x, a); fe_pow22523(x, x); fe_tobytes(out, x); }
void orc_sc_reduce64(uint8_t out[32], const uint8_t in[64]) { sc_reduce_bits(out, in, 64); }
void orc_sc_muladd(uint8_t s[32], const uint8_t h[32], const uint8_t a[32], const uint8_t r[32]) { sc_muladd(s, h, a, r); }
void orc_folds8(uint8_t cut[32], const uint8_t s[32]) { folds8(cut, s); }
void orc_folds4(uint8_t v[64], const uint8_t s[32]) { folds4(v, s); }
