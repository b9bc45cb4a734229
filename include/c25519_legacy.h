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

#ifdef __cplusplus
}
#endif
#endif /* C25519_LEGACY_H */
