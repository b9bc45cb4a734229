// ed25519_kernels.cu -- fixed-base (8-fold comb) and double-base kernels (sm_100a), one operation per thread.
//
//   k_x25519_comb          curve25519_dh_CalculatePublicKey_fast   curve25519_dh.c:162-189
//   k_ed25519_keypair      ed25519_CreateKeyPair                   ed25519_sign.c:344-367
//   k_ed25519_sign         ed25519_SignMessage                     ed25519_sign.c:372-419
//   k_ed25519_verify_init  ed25519_Verify_Init                     ed25519_verify.c:179-232
//   k_ed25519_verify_check ed25519_Verify_Check                    ed25519_verify.c:287-313 (+ edp_PolyPointMultiply :243)
//   (ed25519_VerifySignature :163 = init + check over a per-call workspace, exactly like the reference)
//
// The 256-entry comb table (28 KB padded) is staged from global to shared memory once per CTA with one TMA
// bulk copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) and then read with 16-byte LDS by lanes
// holding unrelated 8-bit indices.  Per-key verification tables (16 x 128 B + 32 B key = 2080 B, the same
// size as the reference's EDP_SIGV_CTX) live in HBM, one contiguous record per key.
#include "kernels.h"
#include "ge25519.cuh"
#include "normalize.cuh"
#include "sc25519.cuh"
#include "sha512.cuh"

namespace c25519 {

constexpr int kThreads = 128;
// resident CTAs per SM the verification kernels are compiled for (register budget = 65536 / (128 x this)); measured in
// tools/verify_lab.cu, profiles/r2_verify_lab.txt
#ifndef C25519_VERIFY_INIT_MINB
#define C25519_VERIFY_INIT_MINB 4
#endif
#ifndef C25519_VERIFY_CHECK_MINB
#define C25519_VERIFY_CHECK_MINB 4
#endif
constexpr int kCombSmemWords = kCombEntries * kCombStrideWords;       // 7168 words = 28 672 B
constexpr int kCtxBytes = 2080;
static_assert(kCombStrideWords == kCombStrideWordsHost, "comb table stride mismatch between host image and kernels");

// ---- TMA bulk copy of the comb table into shared memory -------------------------------------------
C25519_DEV void stage_comb_table(u32* smem_table, unsigned long long* bar, const u32* __restrict__ gtable)
{
    const u32 bar_a = (u32)__cvta_generic_to_shared(bar);
    const u32 dst_a = (u32)__cvta_generic_to_shared(smem_table);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((u32)(kCombSmemWords * 4)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(gtable), "r"((u32)(kCombSmemWords * 4)), "r"(bar_a) : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
}

#define C25519_COMB_SMEM                                              \
    __shared__ __align__(128) u32 s_table[kCombSmemWords];            \
    __shared__ __align__(8) unsigned long long s_bar;                 \
    stage_comb_table(s_table, &s_bar, gtable)

C25519_DEV void load8(u32 (&w)[8], const uint8_t* p)       // 32-byte record, one 256-bit load
{
    fe t; fe_load(t, p);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = t.v[i];
}
C25519_DEV void load8_plain(u32 (&w)[8], const uint8_t* p)
{
    fe t; fe_load_plain(t, p);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = t.v[i];
}
C25519_DEV void store8(uint8_t* p, const u32 (&w)[8])
{
    fe t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = w[i];
    fe_store(p, t);
}
C25519_DEV void clamp(u32 (&k)[8]) { k[0] &= 0xfffffff8u; k[7] = (k[7] | 0x40000000u) & 0x7fffffffu; }   // ecp_TrimSecretKey

// ---- X25519 public key through the comb -----------------------------------------------------------
C25519_DEV void store_xyz(uint8_t* rec, const ge_ext& S)
{ fe_store(rec, S.x); fe_store(rec + 32, S.y); fe_store(rec + 64, S.z); }

// DEFER (all comb kernels): stop at the projective result, write it to the scratch record, and let
// k_normalize finish the batch with one shared inversion per 16 operations (normalize.cuh).
template <bool DEFER>
__global__ void __launch_bounds__(kThreads)
k_x25519_comb(uint8_t* __restrict__ pk32, uint8_t* __restrict__ sk32, size_t n, const u32* __restrict__ gtable,
              uint8_t* __restrict__ scratch)
{
    C25519_COMB_SMEM;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    u32 k[8];
    load8_plain(k, sk32 + 32 * i);
    clamp(k);
    store8(sk32 + 32 * i, k);
    ge_ext S;
    ge_base_comb(S, k, s_table);
    // u = (Z + Y) / (Z - Y)      (curve25519_dh.c:174-178)
    fe num, den, u;
    fe_add_nn(num, S.z, S.y);
    fe_sub(den, S.z, S.y);
    if (DEFER) {
        fe_store(scratch + kScratchXZ * i, num);
        fe_store(scratch + kScratchXZ * i + 32, den);
    } else {
        fe_invert(den, den);
        fe_mul(u, num, den);
        fe_canon(u);
        fe_store(pk32 + 32 * i, u);
    }
}

// a = clamp(SHA512(seed)[0..31]) as limbs; also returns the upper half of the digest (the "prefix" b)
C25519_DEV void expand_seed(u32 (&a)[8], u32 (&b)[8], const u32 (&seed)[8])
{
    u64 pre[4], dg[8]; u32 w[16];
    le_limbs_to_be64(pre, seed);
    sha512_prefixed<4>(dg, pre, nullptr, 0);
    sha512_digest_to_le_words(w, dg);
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = w[i]; b[i] = w[8 + i]; }
    clamp(a);
}

template <bool DEFER>
__global__ void __launch_bounds__(kThreads)
k_ed25519_keypair(uint8_t* __restrict__ pub32, uint8_t* __restrict__ priv64, const uint8_t* __restrict__ seed32, size_t n,
                  const u32* __restrict__ gtable, uint8_t* __restrict__ scratch)
{
    C25519_COMB_SMEM;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    u32 seed[8], a[8], b[8], enc[8];
    load8(seed, seed32 + 32 * i);
    expand_seed(a, b, seed);
    ge_ext S;
    ge_base_comb(S, a, s_table);
    store8(priv64 + 64 * i, seed);
    if (DEFER) {
        store_xyz(scratch + kScratchXYZ * i, S);          // k_normalize writes pub32[i] and priv64[i][32..64)
    } else {
        ge_encode(enc, S);
        store8(pub32 + 32 * i, enc);
        store8(priv64 + 64 * i + 32, enc);
    }
}

C25519_DEV void msg_span(const uint8_t*& m, u64& len, const uint8_t* msgs, const uint64_t* off, size_t fixed_len, size_t i)
{
    if (off) { m = msgs + off[i]; len = off[i + 1] - off[i]; }
    else { m = msgs + i * fixed_len; len = fixed_len; }
}

// PHASE 0: whole signature in one kernel (private inversion; tiny batches)
// PHASE 1: [a:b] = H(sk), r = H(b || m) mod L, R = r B projective -> private scratch record (X, Y, Z, -, a, r)
// PHASE 2: (after k_normalize wrote enc(R) to sig[0..32))  h = H(enc(R) || pk || m) mod L,  S = (h a + r) mod L
// The secret nonce r and scalar a never touch a caller-visible buffer (the reference zeroises both, ed25519_sign.c:417-418);
// the launcher wipes the scratch before releasing it.
constexpr int kSignScratch = 192;                     // X, Y, Z, prefix, a, r
template <int PHASE>
__global__ void __launch_bounds__(kThreads)
k_ed25519_sign(uint8_t* __restrict__ sig64, const uint8_t* __restrict__ priv64, const uint8_t* __restrict__ msgs,
               const uint64_t* __restrict__ off, size_t fixed_len, size_t n, const u32* __restrict__ gtable,
               uint8_t* __restrict__ scratch)
{
    __shared__ __align__(128) u32 s_table[PHASE == 2 ? 4 : kCombSmemWords];
    __shared__ __align__(8) unsigned long long s_bar;
    if (PHASE != 2) stage_comb_table(s_table, &s_bar, gtable);
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint8_t* m; u64 mlen;
    msg_span(m, mlen, msgs, off, fixed_len, i);
    u32 a[8], r[8], enc[8];
    if (PHASE != 2) {
        u32 seed[8], b[8];
        load8(seed, priv64 + 64 * i);
        expand_seed(a, b, seed);                                   // [a:b] = H(sk)              (:385-389)
        {   // r = H(b || m) mod L                                                               (:392-397)
            u64 pre[4], dg[8]; u32 w[16];
            le_limbs_to_be64(pre, b);
            sha512_prefixed<4>(dg, pre, m, mlen);
            sha512_digest_to_le_words(w, dg);
            sc_reduce512(r, w);
        }
        ge_ext S;                                                  // R = r B                    (:400-401)
        ge_base_comb(S, r, s_table);
        if (PHASE == 1) {
            uint8_t* rec = scratch + (size_t)kSignScratch * i;
            store_xyz(rec, S);
            store8(rec + 128, a);
            store8(rec + 160, r);
            return;
        }
        ge_encode(enc, S);
    } else {
        load8_plain(enc, sig64 + 64 * i);                       // plain loads: written by earlier launches
        load8_plain(a, scratch + (size_t)kSignScratch * i + 128);
        load8_plain(r, scratch + (size_t)kSignScratch * i + 160);
    }
    u32 pk[8], h[8], s[8];
    load8(pk, priv64 + 64 * i + 32);
    {   // h = H(enc(R) || pk || m) mod L ; S = (h a + r) mod L                               (:404-414)
        u64 pre[8], dg[8]; u32 w[16];
        le_limbs_to_be64(pre, enc);
        le_limbs_to_be64(pre + 4, pk);
        sha512_prefixed<8>(dg, pre, m, mlen);
        sha512_digest_to_le_words(w, dg);
        sc_reduce512(h, w);
        sc_muladd(s, h, a, r);
    }
    if (PHASE == 0) store8(sig64 + 64 * i, enc);
    store8(sig64 + 64 * i + 32, s);
}

// ---- verification ----------------------------------------------------------------------------------
C25519_DEV void store_pe(uint8_t* p, const ge_pe& q)
{ fe_store(p, q.ypx); fe_store(p + 32, q.ymx); fe_store(p + 64, q.t2d); fe_store(p + 96, q.z2); }
C25519_DEV void load_pe(ge_pe& q, const uint8_t* p)        // 128-byte table entry: four 256-bit loads (coherent: the table
{                                                           // may have been written by this launch's predecessor)
    fe_load_plain(q.ypx, p); fe_load_plain(q.ymx, p + 32); fe_load_plain(q.t2d, p + 64); fe_load_plain(q.z2, p + 96);
}

// ctx record: [0,32) public key bytes, [32 + 128 j, 32 + 128 (j+1)) table entry j = sum_{k in bits(j)} 2^(64k) (-A)
__global__ void __launch_bounds__(kThreads, C25519_VERIFY_INIT_MINB)
k_ed25519_verify_init(uint8_t* __restrict__ ctx, const uint8_t* __restrict__ pk32, size_t n)
{
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    uint8_t* rec = ctx + (size_t)kCtxBytes * i;
    uint8_t* tab = rec + 32;
    u32 pk[8];
    load8(pk, pk32 + 32 * i);
    store8(rec, pk);
    ge_ext Q;
    {   // -A: decode y (bit 255 is the sign of x), recover x with the INVERTED parity      (:192-195)
        const u32 sign = pk[7] >> 31;
#pragma unroll
        for (int k = 0; k < 8; k++) Q.y.v[k] = pk[k];
        Q.y.v[7] &= 0x7fffffffu;
        ge_recover_x(Q.x, Q.y, sign ^ 1u);
        fe_mul(Q.t, Q.x, Q.y);
        fe_set_u32(Q.z, 1);
        fe_narrow(Q.x); fe_narrow(Q.y);                         // narrow representatives for the lazy adds
    }
    {   // entry 0 = neutral element (1, 1, 0, 2)                                            (:201-204)
        ge_pe id; fe_set_u32(id.ypx, 1); fe_set_u32(id.ymx, 1); fe_set_u32(id.t2d, 0); fe_set_u32(id.z2, 2);
        store_pe(tab, id);
    }
    ge_pe pe;
    ge_to_pe(pe, Q);
    store_pe(tab + 128, pe);                                    // entry 1 = -A
#pragma unroll 1
    for (int lvl = 1; lvl < 4; lvl++) {
#pragma unroll 1
        for (int d = 0; d < 63; d++) ge_double<false>(Q);       // T is not an input of a doubling: skip it 63 times
        ge_double<true>(Q);                                     // Q = 2^(64 lvl) (-A)
        const int base = 1 << lvl;
        ge_to_pe(pe, Q);
        store_pe(tab + 128 * base, pe);
#pragma unroll 1
        for (int s = 1; s < base; s++) {                        // QTABLE_SET(base + s, s)      (:175)
            ge_pe prev; ge_ext T;
            load_pe(prev, tab + 128 * s);
            ge_add_pe(T, Q, prev);
            ge_to_pe(pe, T);
            store_pe(tab + 128 * (base + s), pe);
        }
    }
}

template <bool DEFER>
__global__ void __launch_bounds__(kThreads, C25519_VERIFY_CHECK_MINB)
k_ed25519_verify_check(int32_t* __restrict__ ok, const uint8_t* __restrict__ ctx, const uint32_t* __restrict__ key_index,
                       const uint8_t* __restrict__ sig64, const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ off,
                       size_t fixed_len, size_t n, const u32* __restrict__ gtable, uint8_t* __restrict__ scratch)
{
    C25519_COMB_SMEM;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint8_t* rec = ctx + (size_t)kCtxBytes * (key_index ? (size_t)key_index[i] : i);
    const uint8_t* tab = rec + 32;
    const uint8_t* m; u64 mlen;
    msg_span(m, mlen, msgs, off, fixed_len, i);
    u32 R[8], s[8], h[8];
    load8(R, sig64 + 64 * i);
    load8(s, sig64 + 64 * i + 32);                              // S is used raw, all 256 bits     (:308)
    {   // h = H(enc(R) || pk || m) mod L                                                        (:298-304)
        u32 pk[8];
        load8_plain(pk, rec);
        u64 pre[8], dg[8]; u32 w[16];
        le_limbs_to_be64(pre, R);
        le_limbs_to_be64(pre + 4, pk);
        sha512_prefixed<8>(dg, pre, m, mlen);
        sha512_digest_to_le_words(w, dg);
        sc_reduce512(h, w);
    }
    // T = s B + h (-A): 4-fold comb over the per-key table interleaved with the 8-fold base comb  (:243-280)
    ge_ext S;
    {
        ge_pe q0;
        load_pe(q0, tab + 128 * comb4_index(h, 0));
        ge_from_pe(S, q0);
    }
#ifdef C25519_CHECK_SPLIT_LOOP
    // EXPERIMENT (tools/verify_lab.cu): two loops, so the first 31 iterations (no base-table addition) have a body that fits
    // the 32 KB instruction cache
#pragma unroll 1
    for (int j = 1; j < 32; j++) {
        ge_double(S);
        ge_pe q;
        load_pe(q, tab + 128 * comb4_index(h, j));
        ge_add_pe<false>(S, S, q);
    }
#pragma unroll 1
    for (int j = 32; j < 64; j++) {
        ge_double(S);
        ge_pa qb;
        comb_load(qb, s_table, comb8_index(s, j - 32));
        ge_add_affine(S, qb);
        ge_pe q;
        load_pe(q, tab + 128 * comb4_index(h, j));
        ge_add_pe<false>(S, S, q);
    }
#else
#pragma unroll 1
    for (int j = 1; j < 64; j++) {
        ge_double(S);
        if (j >= 32) {
            ge_pa qb;
            comb_load(qb, s_table, comb8_index(s, j - 32));
            ge_add_affine(S, qb);
        }
        ge_pe q;
        load_pe(q, tab + 128 * comb4_index(h, j));
        ge_add_pe<false>(S, S, q);          // followed by a doubling or the end: T is not read again
    }
#endif
    if (DEFER) { store_xyz(scratch + kScratchXYZ * i, S); return; }     // k_normalize encodes and compares with R
    u32 enc[8];
    ge_encode(enc, S);
    u32 diff = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) diff |= enc[k] ^ R[k];          // memcmp(md, signature, 32) == 0          (:312)
    ok[i] = diff == 0 ? 1 : 0;
}

// ---- launchers ----------------------------------------------------------------------------------
static inline unsigned grid_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// run `body(scratch)` with a stream-ordered scratch buffer of `bytes`; `secret` buffers (anything derived from a private
// key) are wiped before they go back to the pool -- also when a launch inside `body` failed
template <typename Body>
static cudaError_t with_scratch(size_t bytes, cudaStream_t s, bool secret, Body body)
{
    uint8_t* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, bytes, s);
    if (e != cudaSuccess) return e;
    e = body(scratch);
    cudaError_t e2 = secret ? wipe_and_free(scratch, bytes, s) : cudaFreeAsync(scratch, s);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_x25519_comb(uint8_t* pk32, uint8_t* sk32_inout, size_t n, const uint32_t* table, cudaStream_t s)
{
    if (!n) return cudaSuccess;
    if (n < kDeferThreshold) {
        k_x25519_comb<false><<<grid_for(n), kThreads, 0, s>>>(pk32, sk32_inout, n, table, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    return with_scratch(n * kScratchXZ, s, true, [&](uint8_t* scratch) {
        k_x25519_comb<true><<<grid_for(n), kThreads, 0, s>>>(pk32, sk32_inout, n, table, scratch);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return launch_normalize(kNormX, scratch, kScratchXZ, n, pk32, 32, nullptr, 0, nullptr, 0, nullptr, s);
    });
}
cudaError_t launch_ed25519_keypair(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n, const uint32_t* table, cudaStream_t s)
{
    if (!n) return cudaSuccess;
    if (n < kDeferThreshold) {
        k_ed25519_keypair<false><<<grid_for(n), kThreads, 0, s>>>(pub32, priv64, seed32, n, table, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    return with_scratch(n * kScratchXYZ, s, true, [&](uint8_t* scratch) {
        k_ed25519_keypair<true><<<grid_for(n), kThreads, 0, s>>>(pub32, priv64, seed32, n, table, scratch);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return launch_normalize(kNormEncode, scratch, kScratchXYZ, n, pub32, 32, priv64 + 32, 64, nullptr, 0, nullptr, s);
    });
}
cudaError_t launch_ed25519_sign(uint8_t* sig64, const uint8_t* priv64, const uint8_t* msgs, const uint64_t* off, size_t fixed_len,
                                size_t n, const uint32_t* table, cudaStream_t s)
{
    if (!n) return cudaSuccess;
    if (n < kDeferThreshold) {
        k_ed25519_sign<0><<<grid_for(n), kThreads, 0, s>>>(sig64, priv64, msgs, off, fixed_len, n, table, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    return with_scratch(n * kSignScratch, s, true, [&](uint8_t* scratch) {
        k_ed25519_sign<1><<<grid_for(n), kThreads, 0, s>>>(sig64, priv64, msgs, off, fixed_len, n, table, scratch);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch_normalize(kNormEncode, scratch, kSignScratch, n, sig64, 64, nullptr, 0, nullptr, 0, nullptr, s);
        if (e != cudaSuccess) return e;
        k_ed25519_sign<2><<<grid_for(n), kThreads, 0, s>>>(sig64, priv64, msgs, off, fixed_len, n, table, scratch);
        count_launch();
        return cudaGetLastError();
    });
}
cudaError_t launch_ed25519_verify_init(uint8_t* ctx, const uint8_t* pk32, size_t n_keys, cudaStream_t s)
{
    if (!n_keys) return cudaSuccess;
    k_ed25519_verify_init<<<grid_for(n_keys), kThreads, 0, s>>>(ctx, pk32, n_keys);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_ed25519_verify_check(int32_t* ok, const uint8_t* ctx, const uint32_t* key_index, const uint8_t* sig64,
                                        const uint8_t* msgs, const uint64_t* off, size_t fixed_len, size_t n,
                                        const uint32_t* table, cudaStream_t s)
{
    if (!n) return cudaSuccess;
    if (n < kDeferThreshold) {
        k_ed25519_verify_check<false><<<grid_for(n), kThreads, 0, s>>>(ok, ctx, key_index, sig64, msgs, off, fixed_len, n, table, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    return with_scratch(n * kScratchXYZ, s, false, [&](uint8_t* scratch) {
        k_ed25519_verify_check<true><<<grid_for(n), kThreads, 0, s>>>(ok, ctx, key_index, sig64, msgs, off, fixed_len, n, table, scratch);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return launch_normalize(kNormCompare, scratch, kScratchXYZ, n, nullptr, 0, nullptr, 0, sig64, 64, ok, s);
    });
}

// Single-phase verification = init + check over a stream-ordered workspace of 2080 B per item, processed in
// slices so the workspace stays bounded (2^20 items -> 2.2 GB of the 180 GB) however large the batch is.
cudaError_t launch_ed25519_verify(int32_t* ok, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msgs, const uint64_t* off,
                                  size_t fixed_len, size_t n, const uint32_t* table, cudaStream_t s)
{
    if (!n) return cudaSuccess;
    constexpr size_t kSlice = (size_t)1 << 20;
    const size_t slice = n < kSlice ? n : kSlice;
    uint8_t* ws = nullptr;
    cudaError_t e = cudaMallocAsync(&ws, slice * kCtxBytes, s);
    if (e != cudaSuccess) return e;
    for (size_t base = 0; base < n && e == cudaSuccess; base += slice) {
        const size_t cnt = (n - base < slice) ? n - base : slice;
        e = launch_ed25519_verify_init(ws, pk32 + 32 * base, cnt, s);
        if (e != cudaSuccess) break;
        e = launch_ed25519_verify_check(ok + base, ws, nullptr, sig64 + 64 * base, off ? msgs : msgs + base * fixed_len,
                                        off ? off + base : nullptr, fixed_len, cnt, table, s);
    }
    cudaError_t e2 = cudaFreeAsync(ws, s);
    return e != cudaSuccess ? e : e2;
}

}  // namespace c25519
