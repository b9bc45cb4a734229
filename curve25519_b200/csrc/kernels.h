// kernels.h -- internal launcher interface between the host-side C ABI (engine.cu) and the kernel
// translation units.  Not installed; the public boundary is include/c25519_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace c25519 {

// one 8-fold base-point table entry = (Y+X, Y-X, 2dT), 3 x 8 limbs (PA_POINT, curve25519_mehdi.h:77-82)
constexpr int kCombEntries = 256;
constexpr int kCombWordsPerEntry = 24;
constexpr int kCombTableBytes = kCombEntries * kCombWordsPerEntry * 4;   // 24 576
constexpr int kCombStrideWordsHost = 28;   // padded device/shared-memory stride (== kCombStrideWords in ge25519.cuh)

cudaError_t launch_x25519_ladder(uint8_t* out32, const uint8_t* pk32_or_null, uint8_t* sk32_inout, size_t n, cudaStream_t s);
cudaError_t launch_x25519_projective(uint8_t** scratch_out, const uint8_t* pk32_or_null, uint8_t* sk32_inout, size_t n, cudaStream_t s);
cudaError_t launch_x25519_finish(uint8_t* scratch, uint8_t* out32, size_t n, cudaStream_t s_finish);
cudaError_t launch_x25519_ladder_raw(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n, cudaStream_t s);
cudaError_t launch_x25519_ladder_scatter(uint8_t* const* out_ptrs, int world, int rank, const uint8_t* pk32, uint8_t* sk32_inout,
                                         size_t n_local, cudaStream_t s);
cudaError_t launch_x25519_shared_kdf(uint8_t* key_out, unsigned key_size, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, cudaStream_t s);
cudaError_t launch_x25519_comb(uint8_t* pk32, uint8_t* sk32_inout, size_t n, const uint32_t* table, cudaStream_t s);
cudaError_t launch_ed25519_keypair(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n, const uint32_t* table, cudaStream_t s);
cudaError_t launch_ed25519_sign(uint8_t* sig64, const uint8_t* priv64, const uint8_t* msgs, const uint64_t* off, size_t fixed_len,
                                size_t n, const uint32_t* table, cudaStream_t s);
cudaError_t launch_ed25519_verify(int32_t* ok, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msgs, const uint64_t* off,
                                  size_t fixed_len, size_t n, const uint32_t* table, cudaStream_t s);
cudaError_t launch_ed25519_verify_init(uint8_t* ctx, const uint8_t* pk32, size_t n_keys, cudaStream_t s);
cudaError_t launch_ed25519_verify_check(int32_t* ok, const uint8_t* ctx, const uint32_t* key_index, const uint8_t* sig64,
                                        const uint8_t* msgs, const uint64_t* off, size_t fixed_len, size_t n,
                                        const uint32_t* table, cudaStream_t s);
// Batched projective->affine conversion (normalize.cuh).  mode: 0 = X/Z -> out (stride out_stride),
// 1 = encode(X/Z, Y/Z) -> out (and out2 if non-null), 2 = ok[i] = (encode == cmp[i]).  rec_stride = bytes per scratch record.
cudaError_t launch_normalize(int mode, uint8_t* scratch, size_t rec_stride, size_t n, uint8_t* out, size_t out_stride,
                             uint8_t* out2, size_t out2_stride, const uint8_t* cmp, size_t cmp_stride, int32_t* ok, cudaStream_t s);
// below this many operations a batch is latency-bound and each kernel does its own inversion (one launch)
constexpr size_t kDeferThreshold = 256;
// below this many operations the X25519 ladder runs warp-cooperatively, four lanes per operation (x25519.cuh:
// mont_step_quad): ~370 us per batch instead of ~620 us, at a quarter of the throughput -- which a batch this small cannot use
constexpr size_t kQuadThreshold = 8192;
// up to this many operations (one warp per SM sub-partition: 148 x 4) the ladder runs ONE operation per WARP -- limb per lane,
// four role groups, __shfl_sync (x25519_warp.cuh), the north_star's mapping
constexpr size_t kWarpThreshold = 592;

cudaError_t launch_modl(int op, uint8_t* out32, const uint8_t* a32, const uint8_t* b32, size_t n, cudaStream_t s);
cudaError_t launch_legacy_op(int op, uint32_t* io, int nin, const uint32_t* table, cudaStream_t s);
cudaError_t launch_test_primitive(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, cudaStream_t s);
cudaError_t launch_imad_peak(uint64_t* mac_per_launch, uint32_t* sink, int iters, cudaStream_t s);

void count_launch();

// zeroise a stream-ordered scratch buffer that held secret-derived data, then return it to the pool
inline cudaError_t wipe_and_free(void* p, size_t bytes, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, s);
    cudaError_t e2 = cudaFreeAsync(p, s);
    return e != cudaSuccess ? e : e2;
}

}  // namespace c25519
