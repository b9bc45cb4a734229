/*
 * oracle/dropin_main.c -- TEST INFRASTRUCTURE.  A main() for the reference's own test translation unit
 * (test/curve25519_test.c, compiled unmodified with -Dmain=reference_test_main) that runs its dh_test() and
 * signature_test() (RFC 8032 TEST 2, with and without blinding) and then the one correctness check that
 * lives at the top of its speed_test(): Google's curve25519-donna (test/curve25519_donna.c, also compiled
 * unmodified) must produce the same public key as our curve25519_dh_CalculatePublicKey[_fast] for the key
 * 0x42.. (test/curve25519_test.c:136-154).  speed_test()'s rdtsc timing loops themselves are not run: its
 * inline asm (".byte 0x0f,0x31" with an "=A" constraint, test/curve25519_test.c:46-51) clobbers rdx behind the
 * compiler's back on x86-64 and 10 000 single-operation GPU round trips add nothing to a correctness check.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>

extern unsigned char sk1[32], pk1[32], msg1[1], msg1_sig[64];      /* test/curve25519_test.c:412-424 */
int dh_test(void);
int signature_test(const unsigned char *sk, const unsigned char *expected_pk, const unsigned char *msg, size_t size,
                   const unsigned char *expected_sig);
void curve25519_donna(unsigned char *mypublic, const unsigned char *secret, const unsigned char *basepoint);
void ecp_TrimSecretKey(unsigned char *X);
void curve25519_dh_CalculatePublicKey(unsigned char *pk, unsigned char *sk);
void curve25519_dh_CalculatePublicKey_fast(unsigned char *pk, unsigned char *sk);

int main(void)
{
    int rc = dh_test();
    rc += signature_test(sk1, pk1, msg1, 1, msg1_sig);
    {
        static const unsigned char base[32] = { 9 };
        unsigned char sk[32], donna[32], ours[32], ours_fast[32];
        memset(sk, 0x42, 32);
        ecp_TrimSecretKey(sk);
        curve25519_donna(donna, sk, base);
        curve25519_dh_CalculatePublicKey(ours, sk);
        curve25519_dh_CalculatePublicKey_fast(ours_fast, sk);
        if (memcmp(donna, ours, 32) || memcmp(donna, ours_fast, 32)) { printf("donna cross-check FAILED\n"); rc++; }
        else printf("donna cross-check: public keys match\n");
    }
    printf("\ndropin: failures = %d\n", rc);
    return rc;
}
