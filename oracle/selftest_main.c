/*
 * oracle/selftest_main.c -- TEST INFRASTRUCTURE.  A main() for the reference's own self-test translation unit
 * (test/curve25519_selftest.c, compiled where it lies with -DECP_SELF_TEST).  That file drives the library's INTERNAL
 * symbols -- word-level field / scalar arithmetic, Edwards point operations, the streaming SHA-512 API, the constant tables
 * -- so linking it against libcurve25519_b200.so runs the whole ECP_SELF_TEST suite on the GPU engine
 * (curve25519_b200/csrc/legacy_internals.cu): curve25519_SelfTest (test/curve25519_selftest.c:603-817: mod-L folds, SHA-512
 * KATs incl. one million 'a', field identities, I*D mod BPO three ways, 1000 Montgomery products vs plain reduction, inverse,
 * sqrt(-1), y(B), (l-1)B == B, lB == O, the pk1/pk2 ECDH KAT, the k1*k2 split-key round trip, 1/k1 == k2) and
 * ed25519_selftest (:914-983: unpack(B), 7B / 11B / 127B through the comb vs double-and-add, v*B + u*A == O).
 * The same program linked against the compiled reference (oracle/_ref/libref25519.so) must print the same verdict.
 */
#include <stdio.h>

int curve25519_SelfTest(int level);
int ed25519_selftest(void);

int main(void)
{
    int a = curve25519_SelfTest(0);
    printf("curve25519_SelfTest(0): %d failure(s)\n", a);
    int b = ed25519_selftest();
    printf("ed25519_selftest(): %d failure(s)\n", b);
    printf("selftest: %s\n", (a | b) ? "FAILED" : "all checks passed");
    return (a | b) ? 1 : 0;
}
