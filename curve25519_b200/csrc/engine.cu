// engine.cu -- host side of the C ABI declared in include/c25519_b200.h and include/c25519_legacy.h.
//
// Responsibilities: device selection and one-time upload of the comb table, argument checking, the
// device-pointer batch entry points (thin: one kernel launch each, asynchronous on the caller's
// stream), the host-pointer entry points (sliced H2D -> kernel -> D2H pipeline rotating over four
// private streams, copying straight from/to the caller's buffers), and the reference's 11-function API as n = 1 batches.
// There is no CPU implementation of any operation in this library: if CUDA is unusable every call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/c25519_b200.h"
#include "../../include/c25519_legacy.h"
#include "kernels.h"

namespace c25519 {

extern const uint32_t kCombTableHost[kCombEntries * kCombWordsPerEntry];   // comb_table.cu (generated)

const uint32_t* g_comb_table_dev = nullptr;

namespace {

std::mutex g_mu;                       // guards init/shutdown and the host-pointer pipeline
bool g_ready = false;
int g_device = -1;
int g_requested_device = -1;           // set by c25519_init(); -1 = take $C25519_DEVICE, default 0
std::atomic<uint64_t> g_launches{0};
thread_local char t_err[256] = "";

// host-pointer pipeline resources (grow-only): kStages stages, each a private stream + device scratch
constexpr size_t kChunkOps = 1u << 17;             // operations per pipeline slice
constexpr int kStages = 4;                          // slices in flight (one private stream each)
struct Stage {
    cudaStream_t stream = nullptr;
    uint8_t* dev = nullptr;  size_t dev_cap = 0;
};
Stage g_stage[kStages];

int fail(int code, const char* what)
{
    if (code > 0) snprintf(t_err, sizeof t_err, "%s: %s", what, cudaGetErrorString((cudaError_t)code));
    else snprintf(t_err, sizeof t_err, "%s", what);
    return code;
}
#define CK(expr)                                                        \
    do {                                                                \
        cudaError_t e__ = (expr);                                       \
        if (e__ != cudaSuccess) return fail((int)e__, #expr);           \
    } while (0)

int ensure_init_locked()
{
    if (g_ready) return 0;
    int dev = 0;
    if (g_requested_device >= 0) dev = g_requested_device;
    else if (const char* e = getenv("C25519_DEVICE")) dev = atoi(e);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(C25519_E_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)");
    if (dev < 0 || dev >= count) return fail(C25519_E_BAD_ARGUMENT, "device ordinal out of range");
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) return fail(C25519_E_NO_DEVICE, "device is not compute capability 10.x (kernels are built for sm_100a only)");
    CK(cudaSetDevice(dev));
    // device image of the comb table: entries padded from 24 to kCombStrideWords (28) words so one TMA bulk
    // copy drops it into shared memory in its bank-conflict-avoiding layout (see ge25519.cuh)
    static uint32_t padded[kCombEntries * kCombStrideWordsHost];
    for (int e = 0; e < kCombEntries; e++)
        for (int w = 0; w < kCombStrideWordsHost; w++)
            padded[e * kCombStrideWordsHost + w] = w < kCombWordsPerEntry ? kCombTableHost[e * kCombWordsPerEntry + w] : 0u;
    uint32_t* t = nullptr;
    CK(cudaMalloc(&t, sizeof padded));
    CK(cudaMemcpy(t, padded, sizeof padded, cudaMemcpyHostToDevice));
    g_comb_table_dev = t;
    for (auto& st : g_stage) CK(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
    g_device = dev;
    g_ready = true;
    return 0;
}

int reserve(Stage& st, size_t bytes)
{
    if (st.dev_cap < bytes) {
        if (st.dev) { cudaStreamSynchronize(st.stream); cudaFree(st.dev); }
        st.dev = nullptr; st.dev_cap = 0;
        if (cudaMalloc(&st.dev, bytes) != cudaSuccess) return fail(C25519_E_OUT_OF_MEMORY, "cudaMalloc(stage)");
        st.dev_cap = bytes;
    }
    return 0;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A "field" of a staged batch: per-operation record size, direction, caller's host pointer.
struct Field { size_t rec; bool in, out; const uint8_t* src; uint8_t* dst; };

// Generic chunked pipeline.  For each chunk: H2D the `in` fields straight from the caller's buffers
// (truly asynchronous when they are pinned, driver-staged when pageable), run `launch` on the device
// copies, D2H the `out` fields straight into the caller's buffers.  kStages stages on as many streams rotate,
// so slice c+1's H2D overlaps slice c's kernel and slice c-1's D2H (both PCIe directions busy); stream order protects the reuse of a
// stage's device scratch.  Returns after both streams have drained.
template <int NF, typename Launch>
int run_host_pipeline(Field (&f)[NF], size_t n, Launch launch)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_init_locked()) return rc;
    CK(cudaSetDevice(g_device));
    if (n == 0) return 0;
    size_t rec_total = 0;
    for (int k = 0; k < NF; k++) rec_total += f[k].rec;
    // slice size: 2^17 operations, fewer when records are large (long messages) so a stage stays <= 256 MB
    const size_t by_bytes = std::max<size_t>(1, ((size_t)256 << 20) / rec_total);
    const size_t chunk = std::min(n, std::min(kChunkOps, by_bytes));
    size_t offs[NF + 1]; offs[0] = 0;
    for (int k = 0; k < NF; k++) offs[k + 1] = offs[k] + align_up(f[k].rec * chunk, 256);
    for (auto& st : g_stage) if (int rc = reserve(st, offs[NF])) return rc;
    int s = 0;
    for (size_t base = 0; base < n; base += chunk, s = (s + 1) % kStages) {
        const size_t cnt = std::min(chunk, n - base);
        Stage& st = g_stage[s];
        uint8_t* d[NF];
        for (int k = 0; k < NF; k++) {
            d[k] = st.dev + offs[k];
            if (f[k].in) CK(cudaMemcpyAsync(d[k], f[k].src + f[k].rec * base, f[k].rec * cnt, cudaMemcpyHostToDevice, st.stream));
        }
        CK(launch(d, cnt, st.stream));
        for (int k = 0; k < NF; k++)
            if (f[k].out) CK(cudaMemcpyAsync(f[k].dst + f[k].rec * base, d[k], f[k].rec * cnt, cudaMemcpyDeviceToHost, st.stream));
    }
    for (auto& st : g_stage) CK(cudaStreamSynchronize(st.stream));
    return 0;
}

inline bool misaligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) != 0; }

int check_ready()
{
    if (!g_ready) {
        std::lock_guard<std::mutex> lk(g_mu);
        return ensure_init_locked();
    }
    return 0;
}

}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace c25519

using namespace c25519;

extern "C" {

int c25519_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ready && device == g_device) return 0;
    if (g_ready) return fail(C25519_E_BAD_ARGUMENT, "already initialised on another device; call c25519_shutdown first");
    if (device < 0) return fail(C25519_E_BAD_ARGUMENT, "device ordinal must be >= 0");
    g_requested_device = device;
    return ensure_init_locked();
}

int c25519_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ready) return 0;
    cudaSetDevice(g_device);
    cudaDeviceSynchronize();
    for (auto& st : g_stage) {
        if (st.dev) cudaFree(st.dev);
        if (st.stream) cudaStreamDestroy(st.stream);
        st = Stage();
    }
    cudaFree(const_cast<uint32_t*>(g_comb_table_dev));
    g_comb_table_dev = nullptr;
    g_ready = false;
    g_requested_device = -1;
    return 0;
}

const char* c25519_last_error(void) { return t_err; }
uint64_t c25519_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------ device-pointer batch API
int c25519_x25519_shared_batch(uint8_t* out32, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!out32 || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out32) || misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_x25519_ladder(out32, pk32, sk32_inout, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_shared_batch_scatter(uint8_t* const* gathered_ptrs, int world, int rank, const uint8_t* pk32, uint8_t* sk32_inout,
                                       size_t n_local, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (!gathered_ptrs || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(C25519_E_BAD_ARGUMENT, "bad world / rank / pointer table");
    if (n_local && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    for (int g = 0; g < world; g++)
        if (!gathered_ptrs[g] || misaligned32(gathered_ptrs[g])) return fail(C25519_E_BAD_ARGUMENT, "gathered arrays must be non-null and 32-byte aligned");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_x25519_ladder_scatter(gathered_ptrs, world, rank, pk32, sk32_inout, n_local, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_shared_kdf_batch(uint8_t* key_out, size_t key_size, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!key_out || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (key_size == 0 || key_size > 64) return fail(C25519_E_BAD_ARGUMENT, "key_size must be 1..64 (bytes of the SHA-512 digest)");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_x25519_shared_kdf(key_out, (unsigned)key_size, pk32, sk32_inout, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_scalarmult_raw_batch(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!out32 || !point32 || !scalar32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out32) || misaligned32(point32) || misaligned32(scalar32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_x25519_ladder_raw(out32, point32, scalar32, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_public_batch(uint8_t* pk32, uint8_t* sk32_inout, size_t n, int ladder, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    if (ladder) CK(launch_x25519_ladder(pk32, nullptr, sk32_inout, n, (cudaStream_t)stream));
    else CK(launch_x25519_comb(pk32, sk32_inout, n, g_comb_table_dev, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_keypair_batch(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!pub32 || !priv64 || !seed32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(pub32) || misaligned32(priv64) || misaligned32(seed32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_ed25519_keypair(pub32, priv64, seed32, n, g_comb_table_dev, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_sign_batch(uint8_t* sig64, const uint8_t* priv64, const uint8_t* msgs, const uint64_t* msg_off,
                              size_t fixed_len, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!sig64 || !priv64)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig64) || misaligned32(priv64)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_ed25519_sign(sig64, priv64, msgs, msg_off, fixed_len, n, g_comb_table_dev, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_batch(int32_t* ok, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msgs,
                                const uint64_t* msg_off, size_t fixed_len, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!ok || !sig64 || !pk32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig64) || misaligned32(pk32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_ed25519_verify(ok, sig64, pk32, msgs, msg_off, fixed_len, n, g_comb_table_dev, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_init_batch(uint8_t* ctx, const uint8_t* pk32, size_t n_keys, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n_keys && (!ctx || !pk32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(ctx) || misaligned32(pk32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_ed25519_verify_init(ctx, pk32, n_keys, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_check_batch(int32_t* ok, const uint8_t* ctx, const uint32_t* key_index, const uint8_t* sig64,
                                      const uint8_t* msgs, const uint64_t* msg_off, size_t fixed_len, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    if (n && (!ok || !ctx || !sig64)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(ctx) || misaligned32(sig64)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    CK(launch_ed25519_verify_check(ok, ctx, key_index, sig64, msgs, msg_off, fixed_len, n, g_comb_table_dev, (cudaStream_t)stream));
    return 0;
}

int c25519_test_primitive(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, void* stream)
{
    if (int rc = check_ready()) return rc;
    CK(launch_test_primitive(op, out, a, b, n, (cudaStream_t)stream));
    return 0;
}

int c25519_imad_peak_kernel(uint64_t* mac_per_launch, uint32_t* sink, int iters, void* stream)
{
    if (int rc = check_ready()) return rc;
    CK(launch_imad_peak(mac_per_launch, sink, iters, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------ host-pointer batch API
int c25519_x25519_shared_host(uint8_t* out32, const uint8_t* pk32, uint8_t* sk32_inout, size_t n)
{
    if (n && (!out32 || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, out32}, {32, true, false, pk32, nullptr}, {32, true, true, sk32_inout, sk32_inout}};
    return run_host_pipeline(f, n, [](uint8_t** d, size_t cnt, cudaStream_t s) { return launch_x25519_ladder(d[0], d[1], d[2], cnt, s); });
}

int c25519_x25519_scalarmult_raw_host(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n)
{
    if (n && (!out32 || !point32 || !scalar32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, out32}, {32, true, false, point32, nullptr}, {32, true, false, scalar32, nullptr}};
    return run_host_pipeline(f, n, [](uint8_t** d, size_t cnt, cudaStream_t s) { return launch_x25519_ladder_raw(d[0], d[1], d[2], cnt, s); });
}

int c25519_x25519_public_host(uint8_t* pk32, uint8_t* sk32_inout, size_t n, int ladder)
{
    if (n && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[2] = {{32, false, true, nullptr, pk32}, {32, true, true, sk32_inout, sk32_inout}};
    return run_host_pipeline(f, n, [ladder](uint8_t** d, size_t cnt, cudaStream_t s) {
        return ladder ? launch_x25519_ladder(d[0], nullptr, d[1], cnt, s) : launch_x25519_comb(d[0], d[1], cnt, g_comb_table_dev, s);
    });
}

int c25519_ed25519_keypair_host(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n)
{
    if (n && (!pub32 || !priv64 || !seed32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, pub32}, {64, false, true, nullptr, priv64}, {32, true, false, seed32, nullptr}};
    return run_host_pipeline(f, n, [](uint8_t** d, size_t cnt, cudaStream_t s) {
        return launch_ed25519_keypair(d[0], d[1], d[2], cnt, g_comb_table_dev, s);
    });
}

// Ragged messages are staged as one extra blob per call (not chunked): the fixed-length fast path is the
// one the throughput configurations use.
static int ed25519_host_msgs(bool sign, uint8_t* out, const uint8_t* in64, const uint8_t* pk32, const uint8_t* msgs,
                             const uint64_t* msg_off, size_t fixed_len, size_t n)
{
    if (!msg_off) {
        const size_t ml = fixed_len ? fixed_len : 1;   // zero-length records still need a non-zero stride for staging
        if (sign) {
            Field f[3] = {{64, false, true, nullptr, out}, {64, true, false, in64, nullptr}, {ml, fixed_len != 0, false, msgs, nullptr}};
            return run_host_pipeline(f, n, [fixed_len](uint8_t** d, size_t cnt, cudaStream_t s) {
                return launch_ed25519_sign(d[0], d[1], d[2], nullptr, fixed_len, cnt, g_comb_table_dev, s);
            });
        }
        Field f[4] = {{4, false, true, nullptr, out}, {64, true, false, in64, nullptr}, {32, true, false, pk32, nullptr},
                      {ml, fixed_len != 0, false, msgs, nullptr}};
        return run_host_pipeline(f, n, [fixed_len](uint8_t** d, size_t cnt, cudaStream_t s) {
            return launch_ed25519_verify(reinterpret_cast<int32_t*>(d[0]), d[1], d[2], d[3], nullptr, fixed_len, cnt, g_comb_table_dev, s);
        });
    }
    // ragged: single shot on stage 0's stream with ad-hoc device buffers
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_init_locked()) return rc;
    CK(cudaSetDevice(g_device));
    if (n == 0) return 0;
    const size_t total = (size_t)msg_off[n];
    const size_t out_rec = sign ? 64 : 4;
    uint8_t *d_out = nullptr, *d_in = nullptr, *d_pk = nullptr, *d_msgs = nullptr; uint64_t* d_off = nullptr;
    cudaStream_t s = g_stage[0].stream;
    int rc = 0;
    auto cleanup = [&]() { cudaFree(d_out); cudaFree(d_in); cudaFree(d_pk); cudaFree(d_msgs); cudaFree(d_off); };
#define CKC(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { rc = fail((int)e__, #expr); cleanup(); return rc; } } while (0)
    CKC(cudaMalloc(&d_out, out_rec * n));
    CKC(cudaMalloc(&d_in, 64 * n));
    CKC(cudaMalloc(&d_msgs, total ? total : 1));
    CKC(cudaMalloc(&d_off, 8 * (n + 1)));
    CKC(cudaMemcpyAsync(d_in, in64, 64 * n, cudaMemcpyHostToDevice, s));
    if (total) CKC(cudaMemcpyAsync(d_msgs, msgs, total, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_off, msg_off, 8 * (n + 1), cudaMemcpyHostToDevice, s));
    if (sign) {
        CKC(launch_ed25519_sign(d_out, d_in, d_msgs, d_off, 0, n, g_comb_table_dev, s));
    } else {
        CKC(cudaMalloc(&d_pk, 32 * n));
        CKC(cudaMemcpyAsync(d_pk, pk32, 32 * n, cudaMemcpyHostToDevice, s));
        CKC(launch_ed25519_verify(reinterpret_cast<int32_t*>(d_out), d_in, d_pk, d_msgs, d_off, 0, n, g_comb_table_dev, s));
    }
    CKC(cudaMemcpyAsync(out, d_out, out_rec * n, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
#undef CKC
    cleanup();
    return 0;
}

int c25519_ed25519_sign_host(uint8_t* sig64, const uint8_t* priv64, const uint8_t* msgs, const uint64_t* msg_off, size_t fixed_len, size_t n)
{
    if (n && (!sig64 || !priv64)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    return ed25519_host_msgs(true, sig64, priv64, nullptr, msgs, msg_off, fixed_len, n);
}

int c25519_ed25519_verify_host(int32_t* ok, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msgs, const uint64_t* msg_off,
                               size_t fixed_len, size_t n)
{
    if (n && (!ok || !sig64 || !pk32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    return ed25519_host_msgs(false, reinterpret_cast<uint8_t*>(ok), sig64, pk32, msgs, msg_off, fixed_len, n);
}

// ------------------------------------------------------------------ the reference's 11-function API (n = 1)
static void die_if(int rc, const char* fn)
{
    if (rc == 0) return;
    fprintf(stderr, "libcurve25519_b200: %s failed (%d): %s -- there is no CPU fallback\n", fn, rc, c25519_last_error());
    abort();
}

// RFC 7748 clamp on a caller's host buffer (curve25519_utils.c:28-32); the batch kernels clamp on the device
void ecp_TrimSecretKey(unsigned char* sk) { sk[0] &= 0xf8; sk[31] = (unsigned char)((sk[31] | 0x40) & 0x7f); }

// Generic k*P exported by the reference's library (source/curve25519_mehdi.h:93) and used by its self-test:
// K is `len` little-endian bytes (len <= 32), not clamped, not modified.
void ecp_PointMultiply(unsigned char* Q, const unsigned char* P, const unsigned char* K, int len)
{
    unsigned char k[32] = {0};
    if (len > 32) len = 32;
    if (len > 0) memcpy(k, K, (size_t)len);
    die_if(c25519_x25519_scalarmult_raw_host(Q, P, k, 1), "ecp_PointMultiply");
}

void curve25519_dh_CalculatePublicKey(unsigned char* pk, unsigned char* sk)
{ die_if(c25519_x25519_public_host(pk, sk, 1, 1), "curve25519_dh_CalculatePublicKey"); }

void curve25519_dh_CalculatePublicKey_fast(unsigned char* pk, unsigned char* sk)
{ die_if(c25519_x25519_public_host(pk, sk, 1, 0), "curve25519_dh_CalculatePublicKey_fast"); }

void curve25519_dh_CreateSharedKey(unsigned char* shared, const unsigned char* pk, unsigned char* sk)
{ die_if(c25519_x25519_shared_host(shared, pk, sk, 1), "curve25519_dh_CreateSharedKey"); }

void ed25519_CreateKeyPair(unsigned char* pubKey, unsigned char* privKey, const void* blinding, const unsigned char* sk)
{
    (void)blinding;
    unsigned char seed[32]; memcpy(seed, sk, 32);     // privKey may alias sk in caller code
    die_if(c25519_ed25519_keypair_host(pubKey, privKey, seed, 1), "ed25519_CreateKeyPair");
}

void ed25519_SignMessage(unsigned char* signature, const unsigned char* privKey, const void* blinding, const unsigned char* msg, size_t msg_size)
{
    (void)blinding;
    die_if(c25519_ed25519_sign_host(signature, privKey, msg, nullptr, msg_size, 1), "ed25519_SignMessage");
}

// Blinding is result-neutral (ed25519_sign.c:246-263 only re-randomises the scalar and Z); the context is
// kept as an opaque 192-byte blob (sizeof(EDP_BLINDING_CTX), curve25519_mehdi.h:84-88) so caller-supplied
// storage of the reference's size is never overrun.
void* ed25519_Blinding_Init(void* context, const unsigned char* seed, size_t size)
{
    (void)seed; (void)size;
    if (!context) context = malloc(192);
    if (context) memset(context, 0, 192);
    return context;
}
void ed25519_Blinding_Finish(void* context)
{
    if (context) { memset(context, 0, 192); free(context); }
}

int ed25519_VerifySignature(const unsigned char* signature, const unsigned char* publicKey, const unsigned char* msg, size_t msg_size)
{
    int32_t ok = 0;
    die_if(c25519_ed25519_verify_host(&ok, signature, publicKey, msg, nullptr, msg_size, 1), "ed25519_VerifySignature");
    return ok;
}

// Two-phase verification, n = 1: the context is the 2080-byte device-format table copied back to the host.
void* ed25519_Verify_Init(void* context, const unsigned char* publicKey)
{
    uint8_t* ctx = static_cast<uint8_t*>(context);
    if (!ctx) ctx = static_cast<uint8_t*>(malloc(C25519_VERIFY_CTX_BYTES));
    if (!ctx) return nullptr;
    int rc;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        rc = ensure_init_locked();
        if (!rc) {
            cudaSetDevice(g_device);
            Stage& st = g_stage[0];
            rc = reserve(st, 4096);
            if (!rc) {
                cudaError_t e = cudaMemcpyAsync(st.dev + 2304, publicKey, 32, cudaMemcpyHostToDevice, st.stream);
                if (e == cudaSuccess) e = launch_ed25519_verify_init(st.dev, st.dev + 2304, 1, st.stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ctx, st.dev, C25519_VERIFY_CTX_BYTES, cudaMemcpyDeviceToHost, st.stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st.stream);
                if (e != cudaSuccess) rc = fail((int)e, "ed25519_Verify_Init");
            }
        }
    }
    die_if(rc, "ed25519_Verify_Init");
    return ctx;
}

int ed25519_Verify_Check(const void* context, const unsigned char* signature, const unsigned char* msg, size_t msg_size)
{
    int32_t ok = 0;
    int rc;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        rc = ensure_init_locked();
        if (!rc) {
            cudaSetDevice(g_device);
            Stage& st = g_stage[0];
            const size_t need = 4096 + align_up(msg_size + 1, 256);
            rc = reserve(st, need);
            if (!rc) {
                uint8_t* d_ctx = st.dev; uint8_t* d_sig = st.dev + 2304; int32_t* d_ok = reinterpret_cast<int32_t*>(st.dev + 2304 + 64);
                uint8_t* d_msg = st.dev + 4096;
                cudaError_t e = cudaMemcpyAsync(d_ctx, context, C25519_VERIFY_CTX_BYTES, cudaMemcpyHostToDevice, st.stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(d_sig, signature, 64, cudaMemcpyHostToDevice, st.stream);
                if (e == cudaSuccess && msg_size) e = cudaMemcpyAsync(d_msg, msg, msg_size, cudaMemcpyHostToDevice, st.stream);
                if (e == cudaSuccess) e = launch_ed25519_verify_check(d_ok, d_ctx, nullptr, d_sig, d_msg, nullptr, msg_size, 1, g_comb_table_dev, st.stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(&ok, d_ok, 4, cudaMemcpyDeviceToHost, st.stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st.stream);
                if (e != cudaSuccess) rc = fail((int)e, "ed25519_Verify_Check");
            }
        }
    }
    die_if(rc, "ed25519_Verify_Check");
    return ok;
}

void ed25519_Verify_Finish(void* ctx) { free(ctx); }

}  // extern "C"
