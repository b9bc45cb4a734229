// oracle/cxx_dropin_main.cpp -- TEST INFRASTRUCTURE.  Drives the reference's C++ convenience classes (C++/x25519.cpp,
// C++/ed25519.cpp, compiled where they lie) and prints every result in hex.  Built twice by oracle/Makefile: against the
// compiled reference (oracle/_ref/libref25519.so) and against libcurve25519_b200.so; tests/test_gpu_dropin.py requires
// byte-identical output.  The classes call curve25519_dh_*, ed25519_* (with the build-time blinding contexts of
// C++/custom_blinds.h), and -- for X25519Private::CreateSharedKey -- SHA512_Init/Update/Final (C++/x25519.cpp:75-95).
#include <cstdio>
#include <cstring>
#include "x25519.h"
#include "ed25519.h"

static void hex(const char* name, const unsigned char* p, int n)
{
    printf("%s=", name);
    for (int i = 0; i < n; i++) printf("%02x", p[i]);
    printf("\n");
}

int main()
{
    unsigned char a_sk[32], b_sk[32], buf[64], buf2[64];
    for (int i = 0; i < 32; i++) { a_sk[i] = (unsigned char)(0x11 + 7 * i); b_sk[i] = (unsigned char)(0xc3 ^ (5 * i)); }
    int rc = 0;
    {
        X25519Private alice(a_sk), bob(b_sk);
        hex("alice_sk_clamped", alice.GetPrivateKey(0), 32);
        hex("alice_pk", alice.GetPublicKey(0), 32);
        hex("bob_pk", bob.GetPublicKey(0), 32);
        alice.CreateShare(bob.GetPublicKey(0), buf); bob.CreateShare(alice.GetPublicKey(0), buf2);
        hex("share", buf, 32);
        if (memcmp(buf, buf2, 32)) { printf("X25519 shares differ\n"); rc++; }
        alice.CreateSharedKey(bob.GetPublicKey(0), buf, 64); bob.CreateSharedKey(alice.GetPublicKey(0), buf2, 64);
        hex("shared_key64", buf, 64);
        if (memcmp(buf, buf2, 64)) { printf("X25519 shared keys differ\n"); rc++; }
        alice.CreateSharedKey(bob.GetPublicKey(0), buf, 20);
        hex("shared_key20", buf, 20);
    }
    {
        ED25519Private signer(a_sk, 32);
        hex("ed_priv", signer.GetPrivateKey(), 64);
        hex("ed_pub", signer.GetPublicKey(), 32);
        const unsigned char msg[] = "drop-in check of the C++ wrappers";
        unsigned char sig[64];
        signer.SignMessage(msg, sizeof msg, sig);
        hex("sig", sig, 64);
        ED25519Public pub(signer.GetPublicKey());
        bool ok = pub.VeifySignature(msg, sizeof msg, sig);
        sig[10] ^= 4;
        bool bad = pub.VeifySignature(msg, sizeof msg, sig);
        printf("verify=%d tampered=%d\n", (int)ok, (int)bad);
        if (!ok || bad) rc++;
        ED25519Private reload(signer.GetPrivateKey(), 64);
        sig[10] ^= 4;
        unsigned char sig2[64];
        reload.SignMessage(msg, sizeof msg, sig2);
        if (memcmp(sig, sig2, 64)) { printf("reloaded key signs differently\n"); rc++; }
        ED25519Private rnd(0, 0);                       // random key pair: sign / verify round trip only (not printed)
        rnd.SignMessage(msg, 5, sig2);
        if (!ED25519Public(rnd.GetPublicKey()).VeifySignature(msg, 5, sig2)) { printf("random key round trip failed\n"); rc++; }
    }
    printf("cxx dropin: failures = %d\n", rc);
    return rc;
}
