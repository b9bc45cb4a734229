/*
 * c25519_legacy.h -- the reference's 11-function C API, re-exported by libcurve25519_b200.so.
 *
 * Same symbol names, argument order and buffer contracts as msotoodeh/curve25519's public headers
 * (include/curve25519_dh.h:34-48 and include/ed25519_signature.h:40-93 in the reference tree), so a
 * program written against the reference links against this library unchanged; include/curve25519_dh.h
 * and include/ed25519_signature.h in this directory simply forward to this file.  Each call is an n = 1
 * batch through c25519_*_host() (stage in, one kernel, stage out) -- correct and bit-exact, but the
 * point of the engine is the batched ABI in c25519_b200.h.
 *
 * Buffer sizes: secret key 32, public key 32, shared secret 32, Ed25519 private key 64 (seed || pk),
 * signature 64 (R || S).  The three DH calls clamp `sk` in place, like the reference.
 * Errors: the reference has no error channel.  If the GPU is unusable these wrappers abort() after
 * printing c25519_last_error() rather than return garbage; verify returns 0.
 */
#ifndef C25519_LEGACY_H
#define C25519_LEGACY_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ed25519_public_key_size   32
#define ed25519_secret_key_size   32
#define ed25519_private_key_size  64
#define ed25519_signature_size    64

/* Host-side key clamp exported by the reference's library and used by its own test program
 * (source/curve25519_utils.c:28-32, declared in source/curve25519_mehdi.h:96): two byte masks, no curve work. */
void ecp_TrimSecretKey(unsigned char *sk);
/* Generic scalar multiplication exported by the reference's library (source/curve25519_mehdi.h:93): Q = K * P on the
 * Montgomery x-line, K = `len` little-endian bytes (len <= 32), NOT clamped, not modified; K = 0 gives 32 zero bytes. */
void ecp_PointMultiply(unsigned char *Q, const unsigned char *P, const unsigned char *K, int len);

void curve25519_dh_CalculatePublicKey(unsigned char *pk, unsigned char *sk);
void curve25519_dh_CalculatePublicKey_fast(unsigned char *pk, unsigned char *sk);
void curve25519_dh_CreateSharedKey(unsigned char *shared, const unsigned char *pk, unsigned char *sk);

void ed25519_CreateKeyPair(unsigned char *pubKey, unsigned char *privKey, const void *blinding,
                           const unsigned char *sk);
void ed25519_SignMessage(unsigned char *signature, const unsigned char *privKey, const void *blinding,
                         const unsigned char *msg, size_t msg_size);
/* Blinding contexts are result-neutral in the reference (test/curve25519_test.c:371-393 expects identical
 * signatures with and without one); they are accepted and ignored here. */
void *ed25519_Blinding_Init(void *context, const unsigned char *seed, size_t size);
void ed25519_Blinding_Finish(void *context);

int ed25519_VerifySignature(const unsigned char *signature, const unsigned char *publicKey,
                            const unsigned char *msg, size_t msg_size);
/* context == NULL allocates; caller storage must hold 2080 bytes (sizeof(EDP_SIGV_CTX), ed25519_verify.c:44-47) */
void *ed25519_Verify_Init(void *context, const unsigned char *publicKey);
int ed25519_Verify_Check(const void *context, const unsigned char *signature, const unsigned char *msg,
                         size_t msg_size);
void ed25519_Verify_Finish(void *ctx);

/* ------------------------------------------------------------------------------------------------------------------
 * INTERNAL symbols of the reference's library, exported so that its self-test (test/curve25519_selftest.c built with
 * -DECP_SELF_TEST) and its C++ wrappers (C++/x25519.cpp, C++/ed25519.cpp) link against this engine unchanged
 * (curve25519_b200/csrc/legacy_internals.cu).  Word arrays are 8 x 32-bit little-endian limbs (the reference's
 * portable-C U_WORD, source/curve25519_mehdi.h:36-51); points use the reference's struct layouts as flat word arrays:
 * Affine_POINT {x,y} = 16 words, Ext_POINT {x,y,z,t} = 32, PA_POINT {Y+X,Y-X,2dT} = 24, PE_POINT {Y+X,Y-X,2dT,2Z} = 32
 * (source/curve25519_mehdi.h:54-82).  Every arithmetic function runs on the GPU, one launch per call; values the
 * reference leaves loosely reduced come back as a congruent representative below 2^256 (callers canonicalise with
 * ecp_Mod / eco_Mod, as the reference's own code does); eco_* results are canonical.
 * Prototypes: source/curve25519_mehdi.h:93-160, source/sha512.h:85-87.                                              */
extern const unsigned char ecp_BasePoint[32];
extern const uint32_t _w_P[8], _w_2d[8], _w_I[8], _w_NxBPO[16][8];
extern const uint32_t _w_base_folding8[256 * 24];            /* PA_POINT[256], source/base_folding8.h */
uint32_t ecp_Add(uint32_t *Z, const uint32_t *X, const uint32_t *Y);          /* plain 256-bit add, returns the carry */
int32_t  ecp_Sub(uint32_t *Z, const uint32_t *X, const uint32_t *Y);          /* plain 256-bit subtract, returns 0 / -1 */
void ecp_AddReduce(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void ecp_SubReduce(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void ecp_MulReduce(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void ecp_SqrReduce(uint32_t *Y, const uint32_t *X);
void ecp_Mod(uint32_t *X);
void ecp_MulMod(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void ecp_Mul(uint32_t *Z16, const uint32_t *X, const uint32_t *Y);            /* exact 512-bit product */
void ecp_Inverse(uint32_t *out, const uint32_t *z);
void ecp_SetValue(uint32_t *X, uint32_t value);
void ecp_Copy(uint32_t *Y, const uint32_t *X);
int  ecp_CmpNE(const uint32_t *X, const uint32_t *Y);
int  ecp_CmpLT(const uint32_t *X, const uint32_t *Y);
uint32_t *ecp_BytesToWords(uint32_t *Y, const unsigned char *X);
unsigned char *ecp_WordsToBytes(unsigned char *Y, const uint32_t *X);
unsigned char *ecp_EncodeInt(unsigned char *Y, const uint32_t *X, unsigned char parity);
unsigned char ecp_DecodeInt(uint32_t *Y, const unsigned char *X);
void ecp_PrintHexBytes(const char *name, const unsigned char *data, uint32_t size);
void ecp_PrintHexWords(const char *name, const uint32_t *data, uint32_t size);
void eco_MulReduce(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void eco_AddReduce(uint32_t *Z, const uint32_t *X, const uint32_t *Y);
void eco_Mod(uint32_t *X);
void eco_ReduceHiWord(uint32_t *Y, uint32_t b, const uint32_t *X);            /* (X + b 2^256) mod L */
void eco_DigestToWords(uint32_t *Y, const unsigned char *md64);
void edp_AddAffinePoint(uint32_t *p_ext, const uint32_t *q_pa);
void edp_AddBasePoint(uint32_t *p_ext);
void edp_AddPoint(uint32_t *r_ext, const uint32_t *p_ext, const uint32_t *q_pe);
void edp_DoublePoint(uint32_t *p_ext);
void edp_BasePointMultiply(uint32_t *r_affine, const uint32_t *sk, const void *blinding);
void ed25519_CalculateX(uint32_t *X, const uint32_t *Y, uint32_t parity);
void ed25519_UnpackPoint(uint32_t *r_affine, const unsigned char *p);
/* streaming SHA-512 on the reference's 216-byte SHA512_CTX (source/sha512.h:74-83); compression on the GPU */
void SHA512_Init(void *ctx);
void SHA512_Update(void *ctx, const void *data, size_t len);
void SHA512_Final(unsigned char *md, void *ctx);

#ifdef __cplusplus
}
#endif
#endif /* C25519_LEGACY_H */
