/*
 * oracle/batch_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thread-parallel loops over the reference's single-operation C API
 * (include/curve25519_dh.h:34-48, include/ed25519_signature.h:40-93 in the
 * reference tree).  The same file is linked into
 *   oracle/_ref/libref25519.so   (the reference's own portable-C sources) and
 *   oracle/liboracle25519.so     (our CPU restatement, oracle/oracle25519.c)
 * so that tests and bench.py's cpu_baseline / --impl reference arm can drive
 * either implementation over whole batches without Python call overhead.
 * Each worker walks a contiguous slice [lo,hi) of the batch and calls the
 * n=1 API exactly as a user of the reference would.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

/* the 11-function legacy API, as exported by whichever library we are linked into */
void curve25519_dh_CalculatePublicKey(unsigned char *pk, unsigned char *sk);
void curve25519_dh_CalculatePublicKey_fast(unsigned char *pk, unsigned char *sk);
void curve25519_dh_CreateSharedKey(unsigned char *shared, const unsigned char *pk, unsigned char *sk);
void ed25519_CreateKeyPair(unsigned char *pubKey, unsigned char *privKey, const void *blinding, const unsigned char *sk);
void ed25519_SignMessage(unsigned char *signature, const unsigned char *privKey, const void *blinding, const unsigned char *msg, size_t msg_size);
int  ed25519_VerifySignature(const unsigned char *signature, const unsigned char *publicKey, const unsigned char *msg, size_t msg_size);

enum { K_SHARED = 0, K_PUBLIC = 1, K_PUBLIC_FAST = 2, K_KEYPAIR = 3, K_SIGN = 4, K_VERIFY = 5 };

typedef struct {
    int kind;
    size_t lo, hi;
    uint8_t *a;            /* primary output                        */
    uint8_t *b;            /* second in/out buffer                  */
    const uint8_t *c;      /* input                                 */
    const uint8_t *msgs;
    const uint64_t *off;   /* n+1 offsets or NULL                   */
    size_t fixed_len;
    int32_t *ok;
} job_t;

static void *worker(void *p)
{
    job_t *j = (job_t *)p;
    for (size_t i = j->lo; i < j->hi; i++) {
        const uint8_t *m = 0; size_t ml = 0;
        if (j->kind >= K_SIGN) {
            if (j->off) { m = j->msgs + j->off[i]; ml = (size_t)(j->off[i + 1] - j->off[i]); }
            else        { m = j->msgs + i * j->fixed_len; ml = j->fixed_len; }
        }
        switch (j->kind) {
        case K_SHARED:      curve25519_dh_CreateSharedKey(j->a + 32 * i, j->c + 32 * i, j->b + 32 * i); break;
        case K_PUBLIC:      curve25519_dh_CalculatePublicKey(j->a + 32 * i, j->b + 32 * i); break;
        case K_PUBLIC_FAST: curve25519_dh_CalculatePublicKey_fast(j->a + 32 * i, j->b + 32 * i); break;
        case K_KEYPAIR:     ed25519_CreateKeyPair(j->a + 32 * i, j->b + 64 * i, 0, j->c + 32 * i); break;
        case K_SIGN:        ed25519_SignMessage(j->a + 64 * i, j->c + 64 * i, 0, m, ml); break;
        case K_VERIFY:      j->ok[i] = ed25519_VerifySignature(j->c + 64 * i, j->b + 32 * i, m, ml); break;
        }
    }
    return 0;
}

static double run(job_t proto, size_t n, int threads)
{
    struct timespec t0, t1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((size_t)threads > n && n > 0) threads = (int)n;
    pthread_t th[256];
    job_t jobs[256];
    clock_gettime(CLOCK_MONOTONIC, &t0);
    size_t per = (n + (size_t)threads - 1) / (size_t)threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].lo = (size_t)t * per;
        jobs[t].hi = jobs[t].lo + per > n ? n : jobs[t].lo + per;
        if (jobs[t].lo >= jobs[t].hi) break;
        if (threads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], 0, worker, &jobs[t]);
        started++;
    }
    if (threads > 1) for (int t = 0; t < started; t++) pthread_join(th[t], 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* All return the wall-clock seconds of the parallel region. */
double drv_x25519_shared(uint8_t *out, const uint8_t *pk, uint8_t *sk_inout, size_t n, int threads)
{ job_t j; memset(&j, 0, sizeof j); j.kind = K_SHARED; j.a = out; j.b = sk_inout; j.c = pk; return run(j, n, threads); }

double drv_x25519_public(uint8_t *pk, uint8_t *sk_inout, size_t n, int fast, int threads)
{ job_t j; memset(&j, 0, sizeof j); j.kind = fast ? K_PUBLIC_FAST : K_PUBLIC; j.a = pk; j.b = sk_inout; return run(j, n, threads); }

double drv_ed25519_keypair(uint8_t *pub, uint8_t *priv, const uint8_t *seed, size_t n, int threads)
{ job_t j; memset(&j, 0, sizeof j); j.kind = K_KEYPAIR; j.a = pub; j.b = priv; j.c = seed; return run(j, n, threads); }

double drv_ed25519_sign(uint8_t *sig, const uint8_t *priv, const uint8_t *msgs, const uint64_t *off,
                        size_t fixed_len, size_t n, int threads)
{ job_t j; memset(&j, 0, sizeof j); j.kind = K_SIGN; j.a = sig; j.c = priv; j.msgs = msgs; j.off = off; j.fixed_len = fixed_len; return run(j, n, threads); }

double drv_ed25519_verify(int32_t *ok, const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs,
                          const uint64_t *off, size_t fixed_len, size_t n, int threads)
{ job_t j; memset(&j, 0, sizeof j); j.kind = K_VERIFY; j.ok = ok; j.c = sig; j.b = (uint8_t *)pk; j.msgs = msgs; j.off = off; j.fixed_len = fixed_len; return run(j, n, threads); }
