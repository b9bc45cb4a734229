"""Public known-answer vectors (RFC 7748, RFC 8032) and the reference's own test constants, with their
status against the reference recorded in SURVEY.md section 8c / Appendix A.  Hex strings are in memory
byte order."""

# (scalar, u, expected CreateSharedKey output from the REFERENCE, note)
X25519_KAT = [
    ("a546e36bf0527c9d3b16154b82465edd62144c0ac1fc5a18506a2244ba449ac4",
     "e6db6867583030db3594c1a424b15f7c726624ec26b3353b10a903a6d0ab1c4c",
     "c3da55379de9c6908e94ea4df28d084f32eccf03491c71f754b4075577a28552", "RFC 7748 5.2 #1"),
    # RFC 7748 5.2 #2 has bit 255 of u set; the reference does not mask it (curve25519_dh.c:104) and returns
    # d5f3..., not the RFC's 95cb...
    ("4b66e9d4d1b4673c5ad22691957d6af5c11b6421e0ea01d42ca4169e7918ba0d",
     "e5210f12786811d3f4b7959d0538ae2c31dbe7106fc03c3efc4cd549c715a493",
     "d5f33573c9f6b8129483acce1e2534e95d3c41af6b00d0d30437b87cada57e4a", "RFC 7748 5.2 #2, unmasked u (reference quirk)"),
    ("4b66e9d4d1b4673c5ad22691957d6af5c11b6421e0ea01d42ca4169e7918ba0d",
     "e5210f12786811d3f4b7959d0538ae2c31dbe7106fc03c3efc4cd549c715a413",
     "95cbde9476e8907d7aade45cb4b873f88b595a68799fa152e6f8f7647aac7957", "RFC 7748 5.2 #2 with the caller masking bit 255"),
    ("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
     "de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f",
     "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742", "RFC 7748 6.1 shared secret (Alice)"),
    ("5dab087e624a8a4b79e17f8b83800ee66f3bb1292618b6fd1c2f8b27ff88e0eb",
     "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a",
     "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742", "RFC 7748 6.1 shared secret (Bob)"),
    ("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
     "ff" * 32,
     "f3ee9fd0e4a41b68c7a5d5d9ba98f14aaecca6fa8a7cd9df0750df6288fcd616", "u = 2^256-1: all 256 bits used"),
    ("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
     "09" + "00" * 30 + "80",
     "154b259daae67a5b0d49f13d09bdb4da14197e812a111867da5e358b0e2d4055", "u = 9 with bit 255 set"),
]

# low-order / degenerate u-coordinates: the reference returns 32 zero bytes (Z = 0 -> inverse(0) = 0)
X25519_LOW_ORDER_U = [
    "00" * 32,
    "01" + "00" * 31,
    "ec" + "ff" * 30 + "7f",      # p - 1
    "ed" + "ff" * 30 + "7f",      # p
    "ee" + "ff" * 30 + "7f",      # p + 1
    "e0eb7a7c3b41b8ae1656e3faf19fc46ada098deb9c32b1fd866205165f49b800",
    "5f9c95bca3508c24b1d0b1559c83ef5b04445cc4581c8e86d8224eddd09f1157",
]

# (secret key, public key) for curve25519_dh_CalculatePublicKey and _fast
X25519_PUBLIC_KAT = [
    ("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
     "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a", "RFC 7748 6.1 Alice"),
    ("5dab087e624a8a4b79e17f8b83800ee66f3bb1292618b6fd1c2f8b27ff88e0eb",
     "de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f", "RFC 7748 6.1 Bob"),
]
X25519_ITER_1 = "422c8e7a6227d7bca1350b3e2bb7279f7897b87bb6854b783c60e80311ae3079"
X25519_ITER_1000 = "684cf59ba83309552800ef566f2f4d3c1c3887c49360e3875f2eb94d99532c51"

# RFC 8032 7.1: (seed, public key, message, signature)
ED25519_KAT = [
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
     "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a", "",
     "e5564300c360ac729086e2cc806e828a84877f1eb8e5d974d873e065224901555fb8821590a33bacc61e39701cf9b46bd25bf5f0595bbe24655141438e7a100b"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
     "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c", "72",
     "92a009a9f0d4cab8720e820b5f642540a2b27b5416503f8fb3762223ebdb69da085ac1e43e15996e458f3613d0f11d8c387b2eaeb4302aeeb00d291612bb0c00"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7",
     "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025", "af82",
     "6291d657deec24024827e69c3abe01a30ce548a284743a445e3680d7db5ac3ac18ff9b538d16f290ae67f760984dc6594a7c15e9716ed28dc027beceea1ec40a"),
]
# TEST 2 is the vector the reference itself carries (test/curve25519_test.c:412-424).

# the reference's dh_test keys (test/curve25519_test.c:435-445) -> public keys and the agreed secret (Appendix A)
DH_TEST = {
    "alice_pk": "fddcda69eeca58e5d783ad1032d080d2758a4e427881b6a4a6fe43d9e7f4ac34",
    "bruce_pk": "6f8f72fc509f40933a025d278b565f8ea873db7417746764da53085cafde536e",
    "shared": "3517fe6814813323e59336af7db9e8aebbf26ccdef572cb71ee3c12d97043c23",
}

SHA512_ABC = ("ddaf35a193617abacc417349ae20413112e6fa4e89a97ea20a9eeee64b55d39a"
              "2192992a274fc1a836ba3c23a3feebbd454d4423643ce80e2a9ac94fa54ca49f")

L_ORDER = 2**252 + 27742317777372353535851937790883648493
P_FIELD = 2**255 - 19
# dh_test secret keys = SHA-256("1234"), SHA-256("abcd") (test/curve25519_test.c:435-445)
DH_TEST["alice_sk"] = "03ac674216f3e15c761ee1a5e255f067953623c8b388b4459e13f978d7c846f4"
DH_TEST["bruce_sk"] = "88d4266fd4e6338d13b845fcf289579d209c897823b9217da3e161936f031589"

# ---- the reference self-test's own constants (test/curve25519_selftest.c), memory byte order
SELFTEST_PK1 = "46f9d209c75369ac5f97a328a1667ac7f86d5ec9b20d515c1139a2563b101360"     # :109-111 (used as a raw scalar)
SELFTEST_PK2 = "5a266ad3d08d9e9b8bd92acccd87d5b996d1dbbab6bcc9756276d761f9375fa7"     # :113-115
SELFTEST_K1 = "0be3be63bc016aaac9e5279fb790fb44372b2d4da1735b5bb01ac0318d892103"      # :101-103, k1 * k2 == 1 mod L
SELFTEST_K2 = "3903e3277e4193612d3d40193d606821602 5ef90b98b24f250609421d4743605".replace(" ", "")   # :105-107
# I * D mod BPO (:128-129), little-endian bytes of W256(0xFDC0315D, ..., 0x00A63CC5)
SELFTEST_IXD_MOD_BPO = "5d31c0fd60f48e59f44916e17ceeeb2db4ef7802fe771833ce3ee0fbc53ca600"
CONST_I = 0x2b8324804fc1df0b2b4d00993dfbd7a72f431806ad2fe478c4ee1b274a0ea0b0      # sqrt(-1) mod p
CONST_D = 0x52036cee2b6ffe738cc740797779e89800700a4d4141d8ab75eb4dca135978a3      # Edwards d
