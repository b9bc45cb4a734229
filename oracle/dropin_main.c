/*
 * oracle/dropin_main.c -- TEST INFRASTRUCTURE.  A main() for the reference's own test translation unit
 * (test/curve25519_test.c, compiled unmodified with -Dmain=reference_test_main) that runs its dh_test(),
 * signature_test() (RFC 8032 TEST 2, with and without blinding) and speed_test() -- including the donna
 * cross-check at the top of speed_test -- with a small loop count instead of the hard-wired 1000
 * (10 000 single-operation GPU round trips add nothing to a correctness check).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>

extern unsigned char sk1[32], pk1[32], msg1[1], msg1_sig[64];      /* test/curve25519_test.c:412-424 */
int dh_test(void);
int signature_test(const unsigned char *sk, const unsigned char *expected_pk, const unsigned char *msg, size_t size,
                   const unsigned char *expected_sig);
int speed_test(int loops);

int main(int argc, char **argv)
{
    int loops = argc > 1 ? atoi(argv[1]) : 3;
    int rc = dh_test();
    rc += signature_test(sk1, pk1, msg1, 1, msg1_sig);
    rc += speed_test(loops);
    printf("\ndropin: failures = %d\n", rc);
    return rc;
}
