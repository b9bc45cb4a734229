/* drop-in name of the reference header include/ed25519_signature.h; see c25519_legacy.h */
#include "c25519_legacy.h"
