// legacy_internals.cu -- the reference library's INTERNAL symbols, re-exported on top of the GPU primitives.
//
// The reference's self-test (test/curve25519_selftest.c, built with -DECP_SELF_TEST) and its C++ wrappers (C++/x25519.cpp,
// C++/ed25519.cpp) link against functions and tables that are not part of the 11-function public API: the word-level field
// and scalar arithmetic (ecp_* / eco_*, source/curve25519_mehdi.h:93-160), the Edwards point operations (edp_*,
// source/ed25519_sign.c / ed25519_verify.c), the streaming SHA-512 API (source/sha512.h:74-87) and a few constant tables.
// Exporting them -- every arithmetic one executed on the GPU by the very device functions the batch kernels are made of --
// lets those programs run against this engine unchanged (tests/test_gpu_dropin.py).
// Each call is one tiny launch (one thread): correct, bit-compatible where the reference defines the bits (canonical values,
// digests, carries), and slow; throughput lives in the batched ABI.
//
// Representation note: where the reference returns a LOOSELY reduced value (ecp_MulReduce & co.: any representative below
// 2^256), the representative returned here may differ; it is congruent and below 2^256, which is all the reference's own
// callers rely on (they canonicalise with ecp_Mod / eco_Mod before comparing).  Only data plumbing (copies, compares, byte
// <-> word codecs, hex printing) runs on the host.
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ge25519.cuh"
#include "sc25519.cuh"
#include "sha512.cuh"

namespace c25519 {

enum LegacyOp {
    L_ADD = 1, L_SUB, L_ADDREDUCE, L_SUBREDUCE, L_MULREDUCE, L_SQRREDUCE, L_MOD, L_MULMOD, L_MUL, L_INVERSE,
    L_ECO_MULREDUCE, L_ECO_ADDREDUCE, L_ECO_MOD, L_ECO_REDUCEHIWORD, L_ECO_DIGEST,
    L_EDP_ADDAFFINE, L_EDP_ADDPOINT, L_EDP_DOUBLE, L_EDP_BASEMULT, L_ED_CALCX, L_SHA512_BLOCKS,
};

C25519_DEV void ld_fe(fe& z, const u32* p)
{
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = p[i];
}
C25519_DEV void st_fe(u32* p, const fe& z)
{
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = z.v[i];
}
C25519_DEV void ld_ext(ge_ext& p, const u32* w)        // Ext_POINT {x, y, z, t}; narrow representatives for the lazy additions
{
    ld_fe(p.x, w); ld_fe(p.y, w + 8); ld_fe(p.z, w + 16); ld_fe(p.t, w + 24);
    fe_narrow(p.x); fe_narrow(p.y); fe_narrow(p.z);
}
C25519_DEV void st_ext(u32* w, const ge_ext& p) { st_fe(w, p.x); st_fe(w + 8, p.y); st_fe(w + 16, p.z); st_fe(w + 24, p.t); }

// io: `nin` input words followed by the output words.  One thread does the work; the comb table (only L_EDP_BASEMULT reads
// it) is staged into shared memory by TMA like in the batch kernels.
__global__ void __launch_bounds__(32)
k_legacy_op(int op, u32* __restrict__ io, int nin, const u32* __restrict__ gtable)
{
    __shared__ __align__(128) u32 s_table[kCombEntries * kCombStrideWords];
    if (op == L_EDP_BASEMULT) {
        for (int i = threadIdx.x; i < kCombEntries * kCombStrideWords; i += 32) s_table[i] = gtable[i];
        __syncwarp();
    }
    if (threadIdx.x != 0) return;
    const u32* in = io;
    u32* out = io + nin;
    fe x, y, z;
    switch (op) {
    case L_ADD: case L_SUB: {                               // ecp_Add / ecp_Sub: plain 256-bit, carry / borrow returned
        u64 c = 0;
        for (int i = 0; i < 8; i++) {
            if (op == L_ADD) { c += (u64)in[i] + in[8 + i]; out[i] = (u32)c; c >>= 32; }
            else { u64 d = (u64)in[i] - in[8 + i] - c; out[i] = (u32)d; c = (d >> 63) & 1; }
        }
        out[8] = op == L_ADD ? (u32)c : (u32)(0u - (u32)c);
    } break;
    case L_ADDREDUCE: ld_fe(x, in); ld_fe(y, in + 8); fe_add(z, x, y); st_fe(out, z); break;
    case L_SUBREDUCE: ld_fe(x, in); ld_fe(y, in + 8); fe_sub(z, x, y); st_fe(out, z); break;
    case L_MULREDUCE: ld_fe(x, in); ld_fe(y, in + 8); fe_mul(z, x, y); st_fe(out, z); break;
    case L_SQRREDUCE: ld_fe(x, in); fe_sqr(z, x); st_fe(out, z); break;
    case L_MOD: ld_fe(z, in); fe_canon(z); st_fe(out, z); break;
    case L_MULMOD: ld_fe(x, in); ld_fe(y, in + 8); fe_mul(z, x, y); fe_canon(z); st_fe(out, z); break;
    case L_MUL: { u32 a[8], b[8], t[16];                    // ecp_Mul: the exact 512-bit product
        for (int i = 0; i < 8; i++) { a[i] = in[i]; b[i] = in[8 + i]; }
        bn_mul<8, 8>(t, a, b);
        for (int i = 0; i < 16; i++) out[i] = t[i]; } break;
    case L_INVERSE: ld_fe(x, in); fe_invert(z, x); st_fe(out, z); break;
    case L_ECO_MULREDUCE: case L_ECO_ADDREDUCE: {
        u32 a[8], b[8], r[8], one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 8; i++) { a[i] = in[i]; b[i] = in[8 + i]; }
        if (op == L_ECO_MULREDUCE) sc_muladd(r, a, b, zero); else sc_muladd(r, a, one, b);
        for (int i = 0; i < 8; i++) out[i] = r[i]; } break;
    case L_ECO_MOD: case L_ECO_REDUCEHIWORD: case L_ECO_DIGEST: {
        u32 w[16], r[8];                                    // value = in[0..7] (+ in[8] * 2^256) (+ in[8..15] * 2^256)
        const int words = op == L_ECO_MOD ? 8 : (op == L_ECO_REDUCEHIWORD ? 9 : 16);
        for (int i = 0; i < 16; i++) w[i] = i < words ? in[i] : 0u;
        sc_reduce512(r, w);
        for (int i = 0; i < 8; i++) out[i] = r[i]; } break;
    case L_EDP_ADDAFFINE: { ge_ext p; ge_pa q; ld_ext(p, in); ld_fe(q.ypx, in + 32); ld_fe(q.ymx, in + 40); ld_fe(q.t2d, in + 48);
        ge_add_affine(p, q); st_ext(out, p); } break;
    case L_EDP_ADDPOINT: { ge_ext p, r; ge_pe q; ld_ext(p, in); ld_fe(q.ypx, in + 32); ld_fe(q.ymx, in + 40); ld_fe(q.t2d, in + 48); ld_fe(q.z2, in + 56);
        ge_add_pe(r, p, q); st_ext(out, r); } break;
    case L_EDP_DOUBLE: { ge_ext p; ld_ext(p, in); ge_double(p); st_ext(out, p); } break;
    case L_EDP_BASEMULT: { u32 a[8]; ge_ext S; fe zi;       // edp_BasePointMultiply: affine (x, y), canonical
        for (int i = 0; i < 8; i++) a[i] = in[i];
        ge_base_comb(S, a, s_table);
        fe_invert(zi, S.z);
        fe_mul(x, S.x, zi); fe_mul(y, S.y, zi); fe_canon(x); fe_canon(y);
        st_fe(out, x); st_fe(out + 8, y); } break;
    case L_ED_CALCX: {                                      // ed25519_CalculateX: x in [0, p] with the requested parity
        ld_fe(y, in);
        ge_recover_x(x, y, in[8] & 1u);                     // may have returned -x loosely reduced: pin the exact integer
        fe_canon(x);
        if ((x.v[0] ^ in[8]) & 1u) {                        // only reachable for x == 0: the reference returns p - 0 = p
            const u32 P[8] = {0xffffffedu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu};
            u64 b = 0;
            for (int i = 0; i < 8; i++) { u64 d = (u64)P[i] - x.v[i] - b; x.v[i] = (u32)d; b = (d >> 63) & 1; }
        }
        st_fe(out, x); } break;
    case L_SHA512_BLOCKS: {                                 // in: state (8 x u64 as lo, hi), nblocks, then nblocks x 128 bytes
        u64 st[8];
        for (int i = 0; i < 8; i++) st[i] = (u64)in[2 * i] | ((u64)in[2 * i + 1] << 32);
        const u32 nb = in[16];
        const u32* data = in + 17;
        for (u32 blk = 0; blk < nb; blk++) {
            u64 w[16];
            for (int j = 0; j < 16; j++) {
                const u32 lo = data[blk * 32 + 2 * j], hi = data[blk * 32 + 2 * j + 1];    // message bytes 8j..8j+7, little-endian words
                w[j] = ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
            }
            sha512_compress(st, w);
        }
        for (int i = 0; i < 8; i++) { out[2 * i] = (u32)st[i]; out[2 * i + 1] = (u32)(st[i] >> 32); }
    } break;
    default: break;
    }
}

cudaError_t launch_legacy_op(int op, uint32_t* io, int nin, const uint32_t* table, cudaStream_t s)
{
    k_legacy_op<<<1, 32, 0, s>>>(op, io, nin, table);
    count_launch();
    return cudaGetLastError();
}

// engine.cu: stage `nin` words in, run one legacy op on the default device, bring `nout` words back; aborts on failure
void legacy_run(int op, const uint32_t* in, int nin, uint32_t* out, int nout);

extern const uint32_t (&kCombTableHost)[kCombEntries * kCombWordsPerEntry];

}  // namespace c25519

using namespace c25519;

namespace {
inline void run2(int op, uint32_t* z, const uint32_t* x, const uint32_t* y, int nout = 8)
{
    uint32_t in[16]; memcpy(in, x, 32); memcpy(in + 8, y, 32);
    uint32_t out[16]; legacy_run(op, in, 16, out, nout); memcpy(z, out, 4 * (size_t)nout);
}
inline void run1(int op, uint32_t* z, const uint32_t* x)
{
    uint32_t out[8]; legacy_run(op, x, 8, out, 8); memcpy(z, out, 32);
}
}  // namespace

extern "C" {

// ---- constant tables: _w_P, _w_2d, _w_I, _w_NxBPO and _w_base_folding8 are emitted by tools/gen_base_table.py into
// comb_table.cu (computed from first principles, checked against the reference's in tests/test_oracle.py)
extern const uint32_t _w_P[8], _w_2d[8], _w_I[8], _w_NxBPO[16][8];
extern const unsigned char ecp_BasePoint[32];
const unsigned char ecp_BasePoint[32] = {9};                 // source/curve25519_dh.c:37

// ---- word-level field arithmetic mod 2^255 - 19 (source/curve25519_mehdi.c)
uint32_t ecp_Add(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { uint32_t o[9]; run2(L_ADD, o, X, Y, 9); memcpy(Z, o, 32); return o[8]; }
int32_t ecp_Sub(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { uint32_t o[9]; run2(L_SUB, o, X, Y, 9); memcpy(Z, o, 32); return (int32_t)o[8]; }
void ecp_AddReduce(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_ADDREDUCE, Z, X, Y); }
void ecp_SubReduce(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_SUBREDUCE, Z, X, Y); }
void ecp_MulReduce(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_MULREDUCE, Z, X, Y); }
void ecp_SqrReduce(uint32_t* Y, const uint32_t* X) { run1(L_SQRREDUCE, Y, X); }
void ecp_Mod(uint32_t* X) { run1(L_MOD, X, X); }
void ecp_MulMod(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_MULMOD, Z, X, Y); }
void ecp_Mul(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_MUL, Z, X, Y, 16); }
void ecp_Inverse(uint32_t* out, const uint32_t* z) { run1(L_INVERSE, out, z); }
// data plumbing (no arithmetic): host side
void ecp_SetValue(uint32_t* X, uint32_t value) { memset(X, 0, 32); X[0] = value; }
void ecp_Copy(uint32_t* Y, const uint32_t* X) { memmove(Y, X, 32); }
int ecp_CmpNE(const uint32_t* X, const uint32_t* Y) { return memcmp(X, Y, 32) != 0; }
int ecp_CmpLT(const uint32_t* X, const uint32_t* Y)
{
    for (int i = 7; i >= 0; i--) if (X[i] != Y[i]) return X[i] < Y[i];
    return 0;
}
uint32_t* ecp_BytesToWords(uint32_t* Y, const unsigned char* X)
{
    for (int i = 0; i < 8; i++) Y[i] = (uint32_t)X[4 * i] | ((uint32_t)X[4 * i + 1] << 8) | ((uint32_t)X[4 * i + 2] << 16) | ((uint32_t)X[4 * i + 3] << 24);
    return Y;
}
unsigned char* ecp_WordsToBytes(unsigned char* Y, const uint32_t* X)
{
    for (int i = 0; i < 8; i++) { Y[4 * i] = (unsigned char)X[i]; Y[4 * i + 1] = (unsigned char)(X[i] >> 8); Y[4 * i + 2] = (unsigned char)(X[i] >> 16); Y[4 * i + 3] = (unsigned char)(X[i] >> 24); }
    return Y;
}
unsigned char* ecp_EncodeInt(unsigned char* Y, const uint32_t* X, unsigned char parity)
{
    ecp_WordsToBytes(Y, X);
    Y[31] = (unsigned char)((Y[31] & 0x7f) | (parity << 7));
    return Y;
}
unsigned char ecp_DecodeInt(uint32_t* Y, const unsigned char* X)
{
    ecp_BytesToWords(Y, X);
    Y[7] &= 0x7fffffffu;
    return (unsigned char)((X[31] >> 7) & 1);
}
void ecp_PrintHexBytes(const char* name, const unsigned char* data, uint32_t size)
{
    printf("%s = 0x", name);
    while (size > 0) printf("%02X", data[--size]);
    printf("\n");
}
void ecp_PrintHexWords(const char* name, const uint32_t* data, uint32_t size)
{
    printf("%s = 0x", name);
    while (size > 0) printf("%08X", data[--size]);
    printf("\n");
}

// ---- arithmetic modulo the group order (source/curve25519_order.c); results are canonical
void eco_MulReduce(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_ECO_MULREDUCE, Z, X, Y); }
void eco_AddReduce(uint32_t* Z, const uint32_t* X, const uint32_t* Y) { run2(L_ECO_ADDREDUCE, Z, X, Y); }
void eco_Mod(uint32_t* X) { run1(L_ECO_MOD, X, X); }
void eco_ReduceHiWord(uint32_t* Y, uint32_t b, const uint32_t* X)
{
    uint32_t in[9]; memcpy(in, X, 32); in[8] = b;
    uint32_t out[8]; legacy_run(L_ECO_REDUCEHIWORD, in, 9, out, 8); memcpy(Y, out, 32);
}
void eco_DigestToWords(uint32_t* Y, const unsigned char* md)
{
    uint32_t in[16]; ecp_BytesToWords(in, md); ecp_BytesToWords(in + 8, md + 32);
    legacy_run(L_ECO_DIGEST, in, 16, Y, 8);
}

// ---- Edwards point operations (source/ed25519_sign.c:71-143, ed25519_verify.c:66-161, ed25519_sign.c:246-268)
void edp_AddAffinePoint(uint32_t* p /* Ext_POINT */, const uint32_t* q /* PA_POINT */)
{
    uint32_t in[56]; memcpy(in, p, 128); memcpy(in + 32, q, 96);
    legacy_run(L_EDP_ADDAFFINE, in, 56, p, 32);
}
void edp_AddBasePoint(uint32_t* p) { edp_AddAffinePoint(p, kCombTableHost + kCombWordsPerEntry); }       // + 1 * B
void edp_AddPoint(uint32_t* r, const uint32_t* p, const uint32_t* q /* PE_POINT */)
{
    uint32_t in[64]; memcpy(in, p, 128); memcpy(in + 32, q, 128);
    uint32_t out[32]; legacy_run(L_EDP_ADDPOINT, in, 64, out, 32); memcpy(r, out, 128);
}
void edp_DoublePoint(uint32_t* p)
{
    uint32_t in[32]; memcpy(in, p, 128);
    legacy_run(L_EDP_DOUBLE, in, 32, p, 32);
}
void edp_BasePointMultiply(uint32_t* R /* Affine_POINT */, const uint32_t* sk, const void* blinding)
{
    (void)blinding;                                          // result-neutral (ed25519_sign.c:254-263)
    legacy_run(L_EDP_BASEMULT, sk, 8, R, 16);
}
void ed25519_CalculateX(uint32_t* X, const uint32_t* Y, uint32_t parity)
{
    uint32_t in[9]; memcpy(in, Y, 32); in[8] = parity;
    uint32_t out[8]; legacy_run(L_ED_CALCX, in, 9, out, 8); memcpy(X, out, 32);
}
void ed25519_UnpackPoint(uint32_t* r /* Affine_POINT */, const unsigned char* p)
{
    const unsigned char parity = ecp_DecodeInt(r + 8, p);
    ed25519_CalculateX(r, r + 8, parity);
}

// ---- streaming SHA-512 with the reference's context layout (source/sha512.h:74-83): the compression runs on the GPU
struct SHA512_CTX_ { unsigned long long h[8], Nl, Nh; union { unsigned long long d[8]; unsigned char p[128]; } u; unsigned int num, md_len; };
static void sha512_blocks(SHA512_CTX_* c, const unsigned char* data, size_t nblocks)
{
    while (nblocks) {
        const size_t nb = nblocks > 512 ? 512 : nblocks;       // <= 64 KB per launch
        static thread_local uint32_t in[17 + 512 * 32];
        for (int i = 0; i < 8; i++) { in[2 * i] = (uint32_t)c->h[i]; in[2 * i + 1] = (uint32_t)(c->h[i] >> 32); }
        in[16] = (uint32_t)nb;
        memcpy(in + 17, data, nb * 128);
        uint32_t out[16];
        legacy_run(L_SHA512_BLOCKS, in, (int)(17 + nb * 32), out, 16);
        for (int i = 0; i < 8; i++) c->h[i] = (unsigned long long)out[2 * i] | ((unsigned long long)out[2 * i + 1] << 32);
        data += nb * 128; nblocks -= nb;
    }
}
void SHA512_Init(void* ctx)
{
    SHA512_CTX_* c = static_cast<SHA512_CTX_*>(ctx);
    static const unsigned long long iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                             0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    memcpy(c->h, iv, sizeof iv);
    c->Nl = c->Nh = 0; c->num = 0; c->md_len = 64;
}
void SHA512_Update(void* ctx, const void* data_, size_t len)
{
    SHA512_CTX_* c = static_cast<SHA512_CTX_*>(ctx);
    const unsigned char* data = static_cast<const unsigned char*>(data_);
    if (len == 0) return;
    const unsigned long long bits = (unsigned long long)len << 3;
    c->Nl += bits; if (c->Nl < bits) c->Nh++;
    c->Nh += (unsigned long long)len >> 61;
    if (c->num) {
        const size_t take = len < 128 - c->num ? len : 128 - c->num;
        memcpy(c->u.p + c->num, data, take);
        c->num += (unsigned)take; data += take; len -= take;
        if (c->num < 128) return;
        sha512_blocks(c, c->u.p, 1);
        c->num = 0;
    }
    if (len >= 128) { sha512_blocks(c, data, len / 128); data += len & ~(size_t)127; len &= 127; }
    if (len) { memcpy(c->u.p, data, len); c->num = (unsigned)len; }
}
void SHA512_Final(unsigned char* md, void* ctx)
{
    SHA512_CTX_* c = static_cast<SHA512_CTX_*>(ctx);
    unsigned char tail[256]; memset(tail, 0, sizeof tail);
    memcpy(tail, c->u.p, c->num);
    tail[c->num] = 0x80;
    const size_t nb = c->num + 17 <= 128 ? 1 : 2;
    for (int i = 0; i < 8; i++) { tail[nb * 128 - 1 - i] = (unsigned char)(c->Nl >> (8 * i)); tail[nb * 128 - 9 - i] = (unsigned char)(c->Nh >> (8 * i)); }
    sha512_blocks(c, tail, nb);
    for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) md[8 * i + k] = (unsigned char)(c->h[i] >> (56 - 8 * k));
    memset(c, 0, sizeof *c);
}

}  // extern "C"
