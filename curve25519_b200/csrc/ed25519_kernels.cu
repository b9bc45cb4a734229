// placeholder -- replaced by the real kernels in the next milestone
#include "kernels.h"
namespace c25519 {
cudaError_t launch_x25519_comb(uint8_t*, uint8_t*, size_t, const uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_ed25519_keypair(uint8_t*, uint8_t*, const uint8_t*, size_t, const uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_ed25519_sign(uint8_t*, const uint8_t*, const uint8_t*, const uint64_t*, size_t, size_t, const uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_ed25519_verify(int32_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint64_t*, size_t, size_t, const uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_ed25519_verify_init(uint8_t*, const uint8_t*, size_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_ed25519_verify_check(int32_t*, const uint8_t*, const uint32_t*, const uint8_t*, const uint8_t*, const uint64_t*, size_t, size_t, const uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
}
