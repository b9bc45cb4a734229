/* drop-in name of the reference header include/curve25519_dh.h; see c25519_legacy.h */
#include "c25519_legacy.h"
