/* oracle/selftest_print.c -- TEST INFRASTRUCTURE.  Hex printers the reference keeps in its test program
 * (test/curve25519_test.c:55-84) and its self-test calls; needed only when the self-test is linked against the compiled
 * reference library (our library exports them itself). */
#include <stdio.h>
void ecp_PrintHexBytes(const char *name, const unsigned char *data, unsigned size)
{ printf("%s = 0x", name); while (size > 0) printf("%02X", data[--size]); printf("\n"); }
void ecp_PrintHexWords(const char *name, const unsigned *data, unsigned size)
{ printf("%s = 0x", name); while (size > 0) printf("%08X", data[--size]); printf("\n"); }
