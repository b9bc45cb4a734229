// engine.cu -- host side of the C ABI declared in include/c25519_b200.h and include/c25519_legacy.h.
//
// Responsibilities: per-device state (comb table, staging pipelines, side stream), argument checking, the
// device-pointer batch entry points (thin: a few kernel launches each, asynchronous on the caller's stream), the
// host-pointer entry points (sliced H2D -> kernel -> D2H pipeline rotating over four private streams, copying straight
// from/to the caller's buffers), the multi-GPU entry points (local kernels + one NCCL exchange of result records,
// overlapped slice by slice), and the reference's C API as n = 1 batches.
// There is no CPU implementation of any operation in this library: if CUDA is unusable every call fails.
//
// Threading (the reference is fully reentrant): no global lock is held while work runs.  Each
// host-pointer call borrows one Pipeline (private streams + device staging buffers) from a small per-device pool, so
// concurrent callers overlap on the GPU; device-pointer calls touch no shared mutable state at all.  Every entry point
// saves and restores the calling thread's current CUDA device.  c25519_shutdown must not race with other calls.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/c25519_b200.h"
#include "../../include/c25519_legacy.h"
#include "kernels.h"

namespace c25519 {

extern const uint32_t (&kCombTableHost)[kCombEntries * kCombWordsPerEntry];   // comb_table.cu (generated)

namespace {

constexpr int kMaxDevices = 64;
constexpr size_t kChunkOps = 1u << 17;              // operations per pipeline slice
constexpr int kStages = 4;                          // slices in flight per pipeline (one private stream each)
constexpr int kMaxPipelines = 4;                    // concurrent host-pointer calls per device before callers queue
constexpr size_t kRaggedSliceBytes = (size_t)64 << 20;

struct Stage {
    cudaStream_t stream = nullptr;
    uint8_t* dev = nullptr;  size_t dev_cap = 0;  size_t used = 0;
};
struct Pipeline { Stage st[kStages]; bool busy = false; };

struct Device {
    std::atomic<bool> ready{false};
    const uint32_t* comb = nullptr;                 // device image of the comb table (padded stride)
    cudaStream_t side = nullptr;                    // finish stream: batched inversion + NCCL exchange of a slice
    cudaStream_t aux = nullptr;                     // second compute stream: odd slices' ladders (no drain bubble between slices)
    std::mutex mu;                                  // guards `pipes` and `regs`
    std::condition_variable cv;
    std::vector<Pipeline*> pipes;
    std::vector<struct Registration*> regs;         // result arrays registered for the peer-memory exchange
};
Device g_dev[kMaxDevices];
std::mutex g_init_mu;                               // serialises device initialisation and shutdown
std::atomic<int> g_default_device{-1};
std::atomic<uint64_t> g_launches{0};
thread_local char t_err[256] = "";

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

// Make `dev` current for the lifetime of the guard and put the caller's device back afterwards.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) { cudaSetDevice(dev); switched = true; }
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
};

int ensure_device(int dev)
{
    if (dev < 0 || dev >= kMaxDevices) return fail(C25519_E_BAD_ARGUMENT, "device ordinal out of range");
    Device& d = g_dev[dev];
    if (d.ready.load(std::memory_order_acquire)) return 0;
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (d.ready.load(std::memory_order_acquire)) return 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(C25519_E_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)");
    if (dev >= count) return fail(C25519_E_BAD_ARGUMENT, "device ordinal out of range");
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) return fail(C25519_E_NO_DEVICE, "device is not compute capability 10.x (kernels are built for sm_100a only)");
    DeviceGuard g(dev);
    // device image of the comb table: entries padded from 24 to kCombStrideWords (28) words so one TMA bulk
    // copy drops it into shared memory in its bank-conflict-avoiding layout (see ge25519.cuh)
    std::vector<uint32_t> padded((size_t)kCombEntries * kCombStrideWordsHost);
    for (int en = 0; en < kCombEntries; en++)
        for (int w = 0; w < kCombStrideWordsHost; w++)
            padded[(size_t)en * kCombStrideWordsHost + w] = w < kCombWordsPerEntry ? kCombTableHost[en * kCombWordsPerEntry + w] : 0u;
    uint32_t* t = nullptr;
    CK(cudaMalloc(&t, padded.size() * 4));
    CK(cudaMemcpy(t, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice));
    d.comb = t;
    {   // the finish stream outranks the compute streams: when ladder CTAs retire, the waiting inversion / NCCL CTAs of the
        // previous slice get the freed slots first -- otherwise they would only start once the next ladder's queue is empty
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&d.side, cudaStreamNonBlocking, greatest));
    }
    CK(cudaStreamCreateWithFlags(&d.aux, cudaStreamNonBlocking));
    {   // keep the stream-ordered scratch allocations of the launchers cached in the pool between calls
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    int expect = -1;
    g_default_device.compare_exchange_strong(expect, dev);
    d.ready.store(true, std::memory_order_release);
    return 0;
}

// device used by the host-pointer calls and the legacy wrappers: the first one initialised, else $C25519_DEVICE, else 0
int default_device(int* dev)
{
    int d = g_default_device.load();
    if (d < 0) { const char* e = getenv("C25519_DEVICE"); d = e ? atoi(e) : 0; }
    if (int rc = ensure_device(d)) return rc;
    *dev = d;
    return 0;
}

// Device that owns the memory behind `p` (device-pointer entry points).  A pointer that is not device memory of a
// compute-capability-10 GPU is refused; so are batches whose arrays live on different devices.
int device_of(const void* p, int* dev)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(C25519_E_NO_DEVICE, "cudaPointerGetAttributes failed: no usable CUDA device (this engine has no CPU fallback)"); }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
        return fail(C25519_E_BAD_ARGUMENT, "*_batch entry points take DEVICE pointers (use the *_host entry points for host memory)");
    *dev = a.device;
    return 0;
}
int batch_device(int* dev, std::initializer_list<const void*> ptrs)
{
    int found = -1;
    for (const void* p : ptrs) {
        if (!p) continue;
        int d = -1;
        if (int rc = device_of(p, &d)) return rc;
        if (found >= 0 && d != found) return fail(C25519_E_BAD_ARGUMENT, "record arrays live on different devices");
        found = d;
    }
    if (found < 0) {                                // n == 0 with all-null pointers: nothing will be launched
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); return fail(C25519_E_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)"); }
        if (cudaGetDevice(&found) != cudaSuccess) found = 0;
    }
    if (int rc = ensure_device(found)) return rc;
    *dev = found;
    return 0;
}

int reserve(Stage& st, size_t bytes)
{
    if (st.dev_cap < bytes) {
        if (st.dev) { cudaStreamSynchronize(st.stream); cudaFree(st.dev); }
        st.dev = nullptr; st.dev_cap = 0;
        if (cudaMalloc(&st.dev, bytes) != cudaSuccess) { cudaGetLastError(); return fail(C25519_E_OUT_OF_MEMORY, "cudaMalloc(stage)"); }
        st.dev_cap = bytes;
    }
    return 0;
}

// Borrow a pipeline of device `d` (the device must be current: new pipelines create streams).
Pipeline* acquire(Device& d)
{
    std::unique_lock<std::mutex> lk(d.mu);
    for (;;) {
        for (Pipeline* p : d.pipes) if (!p->busy) { p->busy = true; return p; }
        if ((int)d.pipes.size() < kMaxPipelines) {
            Pipeline* p = new Pipeline();
            bool ok = true;
            for (auto& st : p->st) ok = ok && cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) == cudaSuccess;
            if (!ok) { for (auto& st : p->st) if (st.stream) cudaStreamDestroy(st.stream); delete p; return nullptr; }
            p->busy = true;
            d.pipes.push_back(p);
            return p;
        }
        d.cv.wait(lk);
    }
}
void release(Device& d, Pipeline* p)
{
    { std::lock_guard<std::mutex> lk(d.mu); p->busy = false; }
    d.cv.notify_one();
}
// drain every stage (also on the error path: queued copies may still touch the caller's buffers), optionally wiping
// the staging memory that held secret keys / shared secrets first
int drain(Pipeline* p, bool wipe)
{
    int rc = 0;
    for (auto& st : p->st) {
        if (wipe && st.dev && st.used) cudaMemsetAsync(st.dev, 0, std::min(st.used, st.dev_cap), st.stream);
        cudaError_t e = cudaStreamSynchronize(st.stream);
        if (e != cudaSuccess && rc == 0) rc = fail((int)e, "cudaStreamSynchronize(stage)");
        st.used = 0;
    }
    return rc;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A "field" of a staged batch: per-operation record size, direction, caller's host pointer.
struct Field { size_t rec; bool in, out; const uint8_t* src; uint8_t* dst; };

// Generic chunked pipeline.  For each chunk: H2D the `in` fields straight from the caller's buffers
// (truly asynchronous when they are pinned, driver-staged when pageable), run `launch` on the device
// copies, D2H the `out` fields straight into the caller's buffers.  kStages stages on as many streams rotate,
// so slice c+1's H2D overlaps slice c's kernel and slice c-1's D2H (both PCIe directions busy); stream order protects
// the reuse of a stage's device scratch.  Returns after every stream has drained -- on failure too.
template <int NF, typename Launch>
int run_host_pipeline(Field (&f)[NF], size_t n, bool secret, Launch launch)
{
    int dev = 0;
    if (int rc = default_device(&dev)) return rc;
    if (n == 0) return 0;
    DeviceGuard g(dev);
    Device& D = g_dev[dev];
    Pipeline* P = acquire(D);
    if (!P) return fail(C25519_E_OUT_OF_MEMORY, "cannot create pipeline streams");
    size_t rec_total = 0;
    for (int k = 0; k < NF; k++) rec_total += f[k].rec;
    // slice size: 2^17 operations, fewer when records are large (long messages) so a stage stays <= 256 MB
    const size_t by_bytes = std::max<size_t>(1, ((size_t)256 << 20) / rec_total);
    const size_t chunk = std::min(n, std::min(kChunkOps, by_bytes));
    size_t offs[NF + 1]; offs[0] = 0;
    for (int k = 0; k < NF; k++) offs[k + 1] = offs[k] + align_up(f[k].rec * chunk, 256);
    int rc = 0;
    auto body = [&]() -> int {
        for (auto& st : P->st) if (int r = reserve(st, offs[NF])) return r;
        int s = 0;
        for (size_t base = 0; base < n; base += chunk, s = (s + 1) % kStages) {
            const size_t cnt = std::min(chunk, n - base);
            Stage& st = P->st[s];
            st.used = offs[NF];
            uint8_t* d[NF];
            for (int k = 0; k < NF; k++) {
                d[k] = st.dev + offs[k];
                if (f[k].in) CK(cudaMemcpyAsync(d[k], f[k].src + f[k].rec * base, f[k].rec * cnt, cudaMemcpyHostToDevice, st.stream));
            }
            CK(launch(d, cnt, st.stream, D));
            for (int k = 0; k < NF; k++)
                if (f[k].out) CK(cudaMemcpyAsync(f[k].dst + f[k].rec * base, d[k], f[k].rec * cnt, cudaMemcpyDeviceToHost, st.stream));
        }
        return 0;
    };
    rc = body();
    char saved[sizeof t_err]; memcpy(saved, t_err, sizeof saved);
    int rc2 = drain(P, secret);
    if (rc) memcpy(t_err, saved, sizeof saved);        // keep the first failure's message
    release(D, P);
    return rc ? rc : rc2;
}

inline bool misaligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) != 0; }

// ---- NCCL, bound at run time (dlopen) so the library loads on hosts without NCCL; the sharded calls then fail loudly ----
typedef struct { char internal[128]; } nccl_uid;
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*CommCount)(void*, int*) = nullptr;
    int (*CommUserRank)(void*, int*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Nccl g_nccl;
std::once_flag g_nccl_once;
constexpr int kNcclUint8 = 1;                        // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

int nccl_load()
{
    std::call_once(g_nccl_once, [] {
        const char* names[] = {getenv("C25519_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm) continue;
            g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.h) break;
        }
        if (!g_nccl.h) return;
        auto sym = [&](const char* s) { return dlsym(g_nccl.h, s); };
#define BIND(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(sym(name))
        BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
        BIND(CommCount, "ncclCommCount"); BIND(CommUserRank, "ncclCommUserRank"); BIND(AllGather, "ncclAllGather");
        BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd");
        BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.CommCount && g_nccl.CommUserRank &&
                    g_nccl.AllGather && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
    });
    if (!g_nccl.ok) return fail(C25519_E_NO_NCCL, "NCCL is not available (libnccl.so.2 not found; set C25519_NCCL_LIB)");
    return 0;
}
int nccl_fail(int code, const char* what)
{
    snprintf(t_err, sizeof t_err, "%s: NCCL error %d (%s)", what, code, g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "?");
    return C25519_E_NCCL;
}
#define NK(expr)                                                        \
    do {                                                                \
        int r__ = (expr);                                               \
        if (r__ != 0) return nccl_fail(r__, #expr);                     \
    } while (0)

int comm_shape(void* comm, int* world, int* rank)
{
    if (int rc = nccl_load()) return rc;
    if (!comm) return fail(C25519_E_BAD_ARGUMENT, "null NCCL communicator");
    NK(g_nccl.CommCount(comm, world));
    NK(g_nccl.CommUserRank(comm, rank));
    return 0;
}

// ---- peer-memory exchange: result arrays registered with c25519_sharded_register -----------------------------------------
// A registered result array is mapped into every rank of the communicator (CUDA IPC).  The exchange of the result records
// then needs NO SM at all: each rank PUSHES its rows into every peer's array with copy-engine transfers over NVLink
// (cudaMemcpyAsync on peer-mapped memory) and the ranks synchronise with stream memory operations on peer-mapped flags
// (cuStreamWriteValue32 / cuStreamWaitValue32).  Unlike NCCL's kernels -- which cannot get SM slots while a ladder grid
// fills the machine and therefore never overlap it -- the pushes of one slice run underneath the next slice's ladder.
struct Registration {
    uint8_t* base = nullptr; size_t bytes = 0; void* comm = nullptr; int world = 0, rank = 0;
    uint8_t* peer[8] = {};                          // every rank's array in THIS process's address space (peer[rank] = base)
    uint32_t* flags = nullptr;                      // local flags: entered[8], pushed[8]
    uint32_t* peer_flags[8] = {};
    void* opened[16] = {}; int n_opened = 0;        // cudaIpcOpenMemHandle results to close
    uint32_t epoch = 0;                             // one per exchange call; ranks advance in lock step (collective calls)
    // deferred mode (c25519_sharded_set_deferred): the exchange runs on `xs` and is NOT joined into the caller's stream;
    // `done` marks its completion (c25519_sharded_sync, and the next call's write-after-read guard)
    bool deferred = false, has_done = false;
    cudaStream_t xs = nullptr;
    cudaEvent_t done = nullptr;
};
typedef int (*StreamMemOp)(cudaStream_t, unsigned long long, unsigned, unsigned);
StreamMemOp g_write32 = nullptr, g_wait32 = nullptr;
typedef int (*MemGetAddressRange)(unsigned long long*, size_t*, unsigned long long);
MemGetAddressRange g_addr_range = nullptr;
std::once_flag g_drv_once;
int driver_load()
{
    std::call_once(g_drv_once, [] {
        cudaDriverEntryPointQueryResult q;
        void* f = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) g_write32 = (StreamMemOp)f;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) g_wait32 = (StreamMemOp)f;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) g_addr_range = (MemGetAddressRange)f;
    });
    if (!g_write32 || !g_wait32 || !g_addr_range) return fail(C25519_E_NO_DEVICE, "CUDA driver lacks stream memory operations");
    return 0;
}
#define DK(expr)                                                        \
    do {                                                                \
        int r__ = (expr);                                               \
        if (r__ != 0) { snprintf(t_err, sizeof t_err, "%s: CUDA driver error %d", #expr, r__); return C25519_E_BAD_ARGUMENT; } \
    } while (0)

Registration* find_registration(Device& D, const void* p, void* comm)
{
    std::lock_guard<std::mutex> lk(D.mu);
    for (Registration* r : D.regs)
        if (r->comm == comm && (const uint8_t*)p >= r->base && (const uint8_t*)p < r->base + r->bytes) return r;
    return nullptr;
}

// "every rank has entered exchange `epoch`" towards the peers (side stream, no SMs)
int peer_announce(Registration* R, uint32_t epoch, int slot, cudaStream_t s)
{
    for (int p = 0; p < R->world; p++)
        if (p != R->rank) DK(g_write32(s, (unsigned long long)(uintptr_t)(R->peer_flags[p] + slot * 8 + R->rank), epoch, 0));
    return 0;
}
int peer_await(Registration* R, uint32_t epoch, int slot, cudaStream_t s)
{
    for (int p = 0; p < R->world; p++)
        if (p != R->rank) DK(g_wait32(s, (unsigned long long)(uintptr_t)(R->flags + slot * 8 + p), epoch, 0 /* GEQ */));
    return 0;
}
// push rows [row0, row0 + cnt) of this rank's block into every peer's array (copy engines over NVLink)
int peer_push(Registration* R, uint8_t* all, size_t rec, size_t n_local, size_t row0, size_t cnt, cudaStream_t s)
{
    const size_t off = (size_t)(all - R->base) + ((size_t)R->rank * n_local + row0) * rec;
    for (int k = 1; k < R->world; k++) {
        const int p = (R->rank + k) % R->world;      // every rank starts with a different peer
        CK(cudaMemcpyAsync(R->peer[p] + off, R->base + off, cnt * rec, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

// Exchange `cnt` records of `rec` bytes starting at row `row0` of every rank's block of `all` ([world][n_local] records), on
// stream `s`.  Registered arrays: copy-engine pushes bracketed by the flag protocol (`first` / `last` mark the first and
// last slice of one call; epoch from exchange_begin).  Otherwise NCCL: the in-place all-gather for whole blocks, grouped
// send / recv for a slice.
struct Exchange { Registration* R = nullptr; uint32_t epoch = 0; };
int exchange_begin(Exchange& x, Device& D, void* all, void* comm, cudaStream_t s)
{
    x.R = find_registration(D, all, comm);
    if (!x.R) return 0;
    if (int rc = driver_load()) return rc;
    if (x.R->has_done) CK(cudaStreamWaitEvent(s, x.R->done, 0));     // a deferred exchange still in flight on this region goes first
    x.epoch = ++x.R->epoch;
    return peer_announce(x.R, x.epoch, 0, s);        // this rank's array may be overwritten from now on (stream order)
}
int exchange_rows(Exchange& x, uint8_t* all, size_t rec, size_t n_local, size_t row0, size_t cnt, bool first, bool last, int world, int rank,
                  void* comm, cudaStream_t s)
{
    if (x.R) {
        if (first) if (int rc = peer_await(x.R, x.epoch, 0, s)) return rc;      // every peer has reached this call
        if (int rc = peer_push(x.R, all, rec, n_local, row0, cnt, s)) return rc;
        if (last) {
            if (int rc = peer_announce(x.R, x.epoch, 1, s)) return rc;          // my rows have landed everywhere
            return peer_await(x.R, x.epoch, 1, s);                              // everybody's rows have landed here
        }
        return 0;
    }
    if (row0 == 0 && cnt == n_local) {               // whole blocks: the plain in-place all-gather
        NK(g_nccl.AllGather(all + (size_t)rank * n_local * rec, all, n_local * rec, kNcclUint8, comm, s));
        return 0;
    }
    NK(g_nccl.GroupStart());
    for (int p = 0; p < world; p++) {
        if (p == rank) continue;
        NK(g_nccl.Send(all + ((size_t)rank * n_local + row0) * rec, cnt * rec, kNcclUint8, p, comm, s));
        NK(g_nccl.Recv(all + ((size_t)p * n_local + row0) * rec, cnt * rec, kNcclUint8, p, comm, s));
    }
    NK(g_nccl.GroupEnd());
    return 0;
}
// whole-block exchange on the caller's stream (sign / verify / public keys / generic records)
int exchange_all(Device& D, uint8_t* all, size_t rec, size_t n_local, int world, int rank, void* comm, cudaStream_t s)
{
    Exchange x;
    if (int rc = exchange_begin(x, D, all, comm, s)) return rc;
    return exchange_rows(x, all, rec, n_local, 0, n_local, true, true, world, rank, comm, s);
}

// X25519 over a large HBM-resident batch, software-pipelined: the batch is cut into `slices`.  Even slices' ladders run on the
// caller's stream, odd slices' on the device's aux stream (forked from the caller's stream), so the CTAs of slice i+1 start
// the moment slots free up while slice i drains -- no bubble at the slice boundary.  Slice i's batched inversion
// (latency-bound: one 265-operation chain per thread) and -- for the sharded entry point -- its NCCL exchange run on the
// finish stream underneath the later slices' ladders; only the last slice's finish is exposed.  Everything joins the
// caller's stream again before the call returns.  `before(finish_stream)` runs once after the fork, `after_slice(row0, cnt,
// first, last, finish_stream)` once per slice.
template <typename Before, typename AfterSlice>
int x25519_pipelined(Device& D, uint8_t* out32, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, cudaStream_t s,
                     const size_t* bounds, int slices, Before before, AfterSlice after_slice)
{
    cudaEvent_t ev[8] = {}, fork = nullptr, done = nullptr, aux_done = nullptr;   // per-call events: callers never share one
    int rc = 0, made = 0;
    auto body = [&]() -> int {
        CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CK(cudaEventRecord(fork, s));                                 // earlier work on the caller's stream (e.g. the inputs)
        CK(cudaStreamWaitEvent(D.aux, fork, 0));
        CK(cudaStreamWaitEvent(D.side, fork, 0));
        if (int r = before(D.side)) return r;
        for (int i = 0; i < slices; i++) {
            const size_t row0 = bounds[i], cnt = bounds[i + 1] - bounds[i];       // bounds[0] = 0 < ... < bounds[slices] = n
            if (cnt == 0) continue;
            cudaStream_t cs = (i & 1) ? D.aux : s;
            uint8_t* scratch = nullptr;
            CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); made = i + 1;
            CK(launch_x25519_projective(&scratch, pk32 ? pk32 + 32 * row0 : nullptr, sk32_inout + 32 * row0, cnt, cs));
            CK(cudaEventRecord(ev[i], cs));
            CK(cudaStreamWaitEvent(D.side, ev[i], 0));
            CK(launch_x25519_finish(scratch, out32 + 32 * row0, cnt, D.side));
            if (int r = after_slice(row0, cnt, i == 0, i == slices - 1, D.side)) return r;
        }
        CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        CK(cudaEventRecord(done, D.side));
        CK(cudaStreamWaitEvent(s, done, 0));
        CK(cudaEventCreateWithFlags(&aux_done, cudaEventDisableTiming));
        CK(cudaEventRecord(aux_done, D.aux));
        CK(cudaStreamWaitEvent(s, aux_done, 0));
        return 0;
    };
    rc = body();
    for (int i = 0; i < made; i++) cudaEventDestroy(ev[i]);      // released by the runtime once the recorded work completes
    for (cudaEvent_t e : {fork, done, aux_done}) if (e) cudaEventDestroy(e);
    return rc;
}
constexpr size_t kPipelineMinOps = (size_t)1 << 18;      // below this a batch is one launch pair (nothing to overlap)

}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// One word-level legacy operation (csrc/legacy_internals.cu) on the default device: `nin` words in, `nout` words out.
// Like the other n = 1 wrappers it has no error channel: a CUDA failure aborts with a message (no CPU fallback).
void legacy_run(int op, const uint32_t* in, int nin, uint32_t* out, int nout)
{
    int dev = 0, rc = default_device(&dev);
    if (!rc) {
        DeviceGuard g(dev);
        Device& D = g_dev[dev];
        Pipeline* P = acquire(D);
        if (!P) rc = fail(C25519_E_OUT_OF_MEMORY, "cannot create pipeline streams");
        else {
            Stage& st = P->st[0];
            rc = reserve(st, align_up(4 * (size_t)(nin + nout), 256));
            if (!rc) {
                uint32_t* io = reinterpret_cast<uint32_t*>(st.dev);
                cudaError_t e = cudaMemcpyAsync(io, in, 4 * (size_t)nin, cudaMemcpyHostToDevice, st.stream);
                if (e == cudaSuccess) e = launch_legacy_op(op, io, nin, D.comb, st.stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(out, io + nin, 4 * (size_t)nout, cudaMemcpyDeviceToHost, st.stream);
                if (e == cudaSuccess) e = cudaMemsetAsync(io, 0, 4 * (size_t)(nin + nout), st.stream);     // operands may be secrets
                cudaError_t e2 = cudaStreamSynchronize(st.stream);
                if (e == cudaSuccess) e = e2;
                if (e != cudaSuccess) rc = fail((int)e, "legacy word-level operation");
            }
            release(D, P);
        }
    }
    if (rc) {
        fprintf(stderr, "libcurve25519_b200: legacy operation %d failed (%d): %s -- there is no CPU fallback\n", op, rc, t_err);
        abort();
    }
}

}  // namespace c25519

using namespace c25519;

extern "C" {

int c25519_init(int device)
{
    if (device < 0) return fail(C25519_E_BAD_ARGUMENT, "device ordinal must be >= 0");
    return ensure_device(device);
}

int c25519_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (int dev = 0; dev < kMaxDevices; dev++) {
        Device& d = g_dev[dev];
        if (!d.ready.load()) continue;
        DeviceGuard g(dev);
        cudaDeviceSynchronize();
        {
            std::lock_guard<std::mutex> lk2(d.mu);
            for (Pipeline* p : d.pipes) {
                for (auto& st : p->st) { if (st.dev) cudaFree(st.dev); if (st.stream) cudaStreamDestroy(st.stream); }
                delete p;
            }
            d.pipes.clear();
            for (Registration* r : d.regs) {
                for (int i = 0; i < r->n_opened; i++) cudaIpcCloseMemHandle(r->opened[i]);
                cudaFree(r->flags);
                if (r->xs) cudaStreamDestroy(r->xs);
                if (r->done) cudaEventDestroy(r->done);
                delete r;
            }
            d.regs.clear();
        }
        if (d.side) { cudaStreamDestroy(d.side); d.side = nullptr; }
        if (d.aux) { cudaStreamDestroy(d.aux); d.aux = nullptr; }
        cudaFree(const_cast<uint32_t*>(d.comb));
        d.comb = nullptr;
        d.ready.store(false);
    }
    g_default_device.store(-1);
    return 0;
}

const char* c25519_last_error(void) { return t_err; }
uint64_t c25519_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------ device-pointer batch API
#define BATCH_PROLOGUE(...)                                             \
    int dev__ = 0;                                                      \
    if (int rc = batch_device(&dev__, {__VA_ARGS__})) return rc;        \
    DeviceGuard guard__(dev__);                                         \
    Device& D = g_dev[dev__];                                           \
    (void)D

int c25519_x25519_shared_batch(uint8_t* out32, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, void* stream)
{
    if (n && (!out32 || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out32) || misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(out32, pk32, sk32_inout);
    // one ladder launch + one batched inversion: slicing a single-GPU batch was measured 0.3 - 1.5 % slower
    CK(launch_x25519_ladder(out32, pk32, sk32_inout, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_shared_batch_scatter(uint8_t* const* gathered_ptrs, int world, int rank, const uint8_t* pk32, uint8_t* sk32_inout,
                                       size_t n_local, void* stream)
{
    if (!gathered_ptrs || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(C25519_E_BAD_ARGUMENT, "bad world / rank / pointer table");
    if (n_local && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    for (int g = 0; g < world; g++)
        if (!gathered_ptrs[g] || misaligned32(gathered_ptrs[g])) return fail(C25519_E_BAD_ARGUMENT, "gathered arrays must be non-null and 32-byte aligned");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(pk32, sk32_inout, gathered_ptrs[rank]);     // peers' arrays live on other devices by design
    CK(launch_x25519_ladder_scatter(gathered_ptrs, world, rank, pk32, sk32_inout, n_local, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_shared_kdf_batch(uint8_t* key_out, size_t key_size, const uint8_t* pk32, uint8_t* sk32_inout, size_t n, void* stream)
{
    if (n && (!key_out || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (key_size == 0 || key_size > 64) return fail(C25519_E_BAD_ARGUMENT, "key_size must be 1..64 (bytes of the SHA-512 digest)");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(key_out, pk32, sk32_inout);
    CK(launch_x25519_shared_kdf(key_out, (unsigned)key_size, pk32, sk32_inout, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_scalarmult_raw_batch(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n, void* stream)
{
    if (n && (!out32 || !point32 || !scalar32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out32) || misaligned32(point32) || misaligned32(scalar32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(out32, point32, scalar32);
    CK(launch_x25519_ladder_raw(out32, point32, scalar32, n, (cudaStream_t)stream));
    return 0;
}

int c25519_x25519_public_batch(uint8_t* pk32, uint8_t* sk32_inout, size_t n, int ladder, void* stream)
{
    if (n && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(pk32) || misaligned32(sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(pk32, sk32_inout);
    if (ladder) CK(launch_x25519_ladder(pk32, nullptr, sk32_inout, n, (cudaStream_t)stream));
    else CK(launch_x25519_comb(pk32, sk32_inout, n, D.comb, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_keypair_batch(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n, void* stream)
{
    if (n && (!pub32 || !priv64 || !seed32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(pub32) || misaligned32(priv64) || misaligned32(seed32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(pub32, priv64, seed32);
    CK(launch_ed25519_keypair(pub32, priv64, seed32, n, D.comb, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_sign_batch(uint8_t* sig64, const uint8_t* priv64, const uint8_t* msgs, const uint64_t* msg_off,
                              size_t fixed_len, size_t n, void* stream)
{
    if (n && (!sig64 || !priv64)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig64) || misaligned32(priv64)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(sig64, priv64, msg_off);
    CK(launch_ed25519_sign(sig64, priv64, msgs, msg_off, fixed_len, n, D.comb, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_batch(int32_t* ok, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msgs,
                                const uint64_t* msg_off, size_t fixed_len, size_t n, void* stream)
{
    if (n && (!ok || !sig64 || !pk32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig64) || misaligned32(pk32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(ok, sig64, pk32);
    CK(launch_ed25519_verify(ok, sig64, pk32, msgs, msg_off, fixed_len, n, D.comb, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_init_batch(uint8_t* ctx, const uint8_t* pk32, size_t n_keys, void* stream)
{
    if (n_keys && (!ctx || !pk32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(ctx) || misaligned32(pk32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(ctx, pk32);
    CK(launch_ed25519_verify_init(ctx, pk32, n_keys, (cudaStream_t)stream));
    return 0;
}

int c25519_ed25519_verify_check_batch(int32_t* ok, const uint8_t* ctx, const uint32_t* key_index, const uint8_t* sig64,
                                      const uint8_t* msgs, const uint64_t* msg_off, size_t fixed_len, size_t n, void* stream)
{
    if (n && (!ok || !ctx || !sig64)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(ctx) || misaligned32(sig64)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(ok, ctx, sig64);
    CK(launch_ed25519_verify_check(ok, ctx, key_index, sig64, msgs, msg_off, fixed_len, n, D.comb, (cudaStream_t)stream));
    return 0;
}

int c25519_modl_batch(int op, uint8_t* out32, const uint8_t* a32, const uint8_t* b32, size_t n, void* stream)
{
    if (op < C25519_MODL_MULMOD || op > C25519_MODL_INVMOD) return fail(C25519_E_BAD_ARGUMENT, "unknown mod-L operation");
    if (n && (!out32 || !a32 || (op != C25519_MODL_INVMOD && !b32))) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out32) || misaligned32(a32) || misaligned32(b32)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(out32, a32, b32);
    CK(launch_modl(op, out32, a32, b32, n, (cudaStream_t)stream));
    return 0;
}

int c25519_test_primitive(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, void* stream)
{
    BATCH_PROLOGUE(out, a, b);
    CK(launch_test_primitive(op, out, a, b, n, (cudaStream_t)stream));
    return 0;
}

int c25519_imad_peak_kernel(uint64_t* mac_per_launch, uint32_t* sink, int iters, void* stream)
{
    BATCH_PROLOGUE(sink);
    CK(launch_imad_peak(mac_per_launch, sink, iters, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------ multi-GPU: local kernels + ONE NCCL exchange of results
int c25519_nccl_unique_id(void* id128)
{
    if (int rc = nccl_load()) return rc;
    if (!id128) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    nccl_uid u;
    NK(g_nccl.GetUniqueId(&u));
    memcpy(id128, &u, sizeof u);
    return 0;
}

int c25519_nccl_comm_init(void** comm, int world, int rank, const void* id128, int device)
{
    if (int rc = nccl_load()) return rc;
    if (!comm || !id128 || world < 1 || rank < 0 || rank >= world) return fail(C25519_E_BAD_ARGUMENT, "bad communicator arguments");
    if (int rc = ensure_device(device)) return rc;
    DeviceGuard g(device);
    nccl_uid u; memcpy(&u, id128, sizeof u);
    NK(g_nccl.CommInitRank(comm, world, u, rank));
    return 0;
}

int c25519_nccl_comm_destroy(void* comm)
{
    if (int rc = nccl_load()) return rc;
    if (comm) NK(g_nccl.CommDestroy(comm));
    return 0;
}

// Map `base[0..bytes)` (device memory of this rank; same size on every rank) into every rank of the communicator.  COLLECTIVE
// and synchronous.  Afterwards every *_sharded / allgather call whose result array lies inside the region exchanges its
// records with copy-engine pushes into the peers' arrays + stream-memory-op flags instead of NCCL kernels (see above).
int c25519_sharded_register(void* base_, size_t bytes, void* nccl_comm)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (!base_ || bytes == 0) return fail(C25519_E_BAD_ARGUMENT, "null region");
    if (world > 8) return fail(C25519_E_BAD_ARGUMENT, "peer-memory exchange supports up to 8 ranks (one NVSwitch domain)");
    if (int rc = driver_load()) return rc;
    BATCH_PROLOGUE(base_);
    struct Rec { cudaIpcMemHandle_t data, flags; unsigned long long offset; int ok; };
    struct Holder {                                  // frees a half-built registration on any early error return
        Registration* r; bool keep = false;
        ~Holder() { if (!keep && r) { for (int i = 0; i < r->n_opened; i++) cudaIpcCloseMemHandle(r->opened[i]); if (r->flags) cudaFree(r->flags); delete r; } }
    } holder{new Registration()};
    Registration* R = holder.r;
    R->base = static_cast<uint8_t*>(base_); R->bytes = bytes; R->comm = nccl_comm; R->world = world; R->rank = rank;
    Rec mine; memset(&mine, 0, sizeof mine);
    unsigned long long abase = 0; size_t asize = 0;
    int ok = g_addr_range(&abase, &asize, (unsigned long long)(uintptr_t)base_) == 0;
    ok = ok && cudaMalloc(&R->flags, 16 * sizeof(uint32_t)) == cudaSuccess && cudaMemset(R->flags, 0, 16 * sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.data, reinterpret_cast<void*>((uintptr_t)abase)) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.flags, R->flags) == cudaSuccess;
    cudaGetLastError();
    mine.offset = (unsigned long long)(uintptr_t)base_ - abase; mine.ok = ok;
    // all-gather the handles through the communicator itself
    Rec* d_all = nullptr; Rec h_all[8];
    CK(cudaMalloc(&d_all, sizeof(Rec) * world));
    CK(cudaMemcpy(d_all + rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NK(g_nccl.AllGather(d_all + rank, d_all, sizeof(Rec), kNcclUint8, nccl_comm, (cudaStream_t)0));
    CK(cudaStreamSynchronize((cudaStream_t)0));
    CK(cudaMemcpy(h_all, d_all, sizeof(Rec) * world, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    for (int p = 0; p < world; p++) ok = ok && h_all[p].ok;
    for (int p = 0; ok && p < world; p++) {
        if (p == rank) { R->peer[p] = R->base; R->peer_flags[p] = R->flags; continue; }
        void *pd = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pd, h_all[p].data, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        R->opened[R->n_opened++] = pd;
        if (cudaIpcOpenMemHandle(&pf, h_all[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        R->opened[R->n_opened++] = pf;
        R->peer[p] = static_cast<uint8_t*>(pd) + h_all[p].offset;
        R->peer_flags[p] = static_cast<uint32_t*>(pf);
    }
    cudaGetLastError();
    // the decision must be the same on every rank: agree through one more tiny all-gather
    int* d_ok = nullptr; int h_ok[8];
    CK(cudaMalloc(&d_ok, sizeof(int) * world));
    CK(cudaMemcpy(d_ok + rank, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NK(g_nccl.AllGather(d_ok + rank, d_ok, sizeof(int), kNcclUint8, nccl_comm, (cudaStream_t)0));
    CK(cudaStreamSynchronize((cudaStream_t)0));
    CK(cudaMemcpy(h_ok, d_ok, sizeof(int) * world, cudaMemcpyDeviceToHost));
    cudaFree(d_ok);
    for (int p = 0; p < world; p++) ok = ok && h_ok[p];
    if (!ok) {
        cudaGetLastError();
        return fail(C25519_E_BAD_ARGUMENT, "the region cannot be shared through CUDA IPC on every rank (allocate it with cudaMalloc / the default "
                                           "PyTorch allocator); calls keep using the NCCL exchange");
    }
    holder.keep = true;
    std::lock_guard<std::mutex> lk(D.mu);
    D.regs.push_back(R);
    return 0;
}

int c25519_sharded_unregister(void* base)
{
    for (int dev = 0; dev < kMaxDevices; dev++) {
        Device& D = g_dev[dev];
        if (!D.ready.load()) continue;
        Registration* R = nullptr;
        {
            std::lock_guard<std::mutex> lk(D.mu);
            for (size_t i = 0; i < D.regs.size(); i++)
                if (D.regs[i]->base == base) { R = D.regs[i]; D.regs.erase(D.regs.begin() + i); break; }
        }
        if (!R) continue;
        DeviceGuard g(dev);
        cudaDeviceSynchronize();
        for (int i = 0; i < R->n_opened; i++) cudaIpcCloseMemHandle(R->opened[i]);
        cudaFree(R->flags);
        if (R->xs) cudaStreamDestroy(R->xs);
        if (R->done) cudaEventDestroy(R->done);
        delete R;
        return 0;
    }
    return fail(C25519_E_BAD_ARGUMENT, "region was not registered");
}

static Registration* registration_by_base(void* base, int* dev_out)
{
    for (int dev = 0; dev < kMaxDevices; dev++) {
        Device& D = g_dev[dev];
        if (!D.ready.load()) continue;
        std::lock_guard<std::mutex> lk(D.mu);
        for (Registration* r : D.regs) if (r->base == base) { *dev_out = dev; return r; }
    }
    return nullptr;
}

// Deferred exchange for a registered region: c25519_x25519_shared_sharded then returns (stream order) as soon as THIS rank's rows
// are in place; the rows travel to the peers by copy engine on a private stream, underneath whatever the caller enqueues next
// (typically the next batch's ladder -- copy engines need no SM).  The gathered array is complete after c25519_sharded_sync.
int c25519_sharded_set_deferred(void* base, int on)
{
    int dev = 0;
    Registration* R = registration_by_base(base, &dev);
    if (!R) return fail(C25519_E_BAD_ARGUMENT, "region was not registered");
    DeviceGuard g(dev);
    if (on && !R->xs) {
        CK(cudaStreamCreateWithFlags(&R->xs, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&R->done, cudaEventDisableTiming));
    }
    R->deferred = on != 0;
    return 0;
}

// make `stream` wait for the last deferred exchange on the region (no-op when none is pending)
int c25519_sharded_sync(void* base, void* stream)
{
    int dev = 0;
    Registration* R = registration_by_base(base, &dev);
    if (!R) return fail(C25519_E_BAD_ARGUMENT, "region was not registered");
    DeviceGuard g(dev);
    if (R->has_done) CK(cudaStreamWaitEvent((cudaStream_t)stream, R->done, 0));
    return 0;
}

int c25519_allgather_records(void* all, size_t rec_bytes, size_t n_local, void* nccl_comm, void* stream)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (n_local == 0 || rec_bytes == 0) return 0;
    if (!all) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    BATCH_PROLOGUE(all);
    return world > 1 ? exchange_all(D, static_cast<uint8_t*>(all), rec_bytes, n_local, world, rank, nccl_comm, (cudaStream_t)stream) : 0;
}

// X25519 shared keys, sharded: this rank computes rows [rank*n_local, (rank+1)*n_local) of out_all and receives the other
// ranks' rows.
int c25519_x25519_shared_sharded(uint8_t* out_all, const uint8_t* pk32_local, uint8_t* sk32_local_inout, size_t n_local,
                                 void* nccl_comm, void* stream)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (n_local == 0) return 0;
    if (!out_all || !pk32_local || !sk32_local_inout) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(out_all) || misaligned32(pk32_local) || misaligned32(sk32_local_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(out_all, pk32_local, sk32_local_inout);
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* mine = out_all + (size_t)rank * n_local * 32;
    // Registered result array: the fused path.  The batched-inversion kernel stores every result straight into ALL ranks'
    // arrays through the peer mappings (k_normalize_scatter: 32-byte stores over NVLink), bracketed by the stream-memory-op
    // flag protocol; no separate collective runs at all.  Measured best at every N.
    if (world > 1) {
        if (Registration* R = find_registration(D, out_all, nccl_comm)) {
            if (int rc = driver_load()) return rc;
            if (R->deferred && n_local >= kQuadThreshold) {
                // Deferred: ladder now; my rows are rewritten only after the previous exchange has finished reading them; the
                // push to the peers runs on the registration's own stream (copy engines + flag operations, no SM) and is not
                // joined here -- it overlaps the caller's next work.  c25519_sharded_sync joins it.
                const uint32_t epoch = ++R->epoch;
                uint8_t* scratch = nullptr;
                CK(launch_x25519_projective(&scratch, pk32_local, sk32_local_inout, n_local, s));
                if (R->has_done) CK(cudaStreamWaitEvent(s, R->done, 0));
                CK(launch_x25519_finish(scratch, mine, n_local, s));
                cudaEvent_t ev = nullptr;
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                cudaError_t e = cudaEventRecord(ev, s);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(R->xs, ev, 0);
                cudaEventDestroy(ev);
                if (e != cudaSuccess) return fail((int)e, "deferred exchange: event");
                if (int rc = peer_announce(R, epoch, 0, R->xs)) return rc;
                if (int rc = peer_await(R, epoch, 0, R->xs)) return rc;
                if (int rc = peer_push(R, out_all, 32, n_local, 0, n_local, R->xs)) return rc;
                if (int rc = peer_announce(R, epoch, 1, R->xs)) return rc;
                if (int rc = peer_await(R, epoch, 1, R->xs)) return rc;
                CK(cudaEventRecord(R->done, R->xs));
                R->has_done = true;
                return 0;
            }
            if (R->has_done) CK(cudaStreamWaitEvent(s, R->done, 0));         // a deferred exchange still in flight goes first
            const uint32_t epoch = ++R->epoch;
            uint8_t* ptrs[8] = {};
            for (int g = 0; g < world; g++) ptrs[g] = R->peer[g] + (out_all - R->base);
            if (int rc = peer_announce(R, epoch, 0, s)) return rc;          // my array may be overwritten from here on
            if (int rc = peer_await(R, epoch, 0, s)) return rc;             // ... and so may everybody else's
            CK(launch_x25519_ladder_scatter(ptrs, world, rank, pk32_local, sk32_local_inout, n_local, s));
            if (int rc = peer_announce(R, epoch, 1, s)) return rc;          // my rows have landed everywhere
            return peer_await(R, epoch, 1, s);                              // everybody's rows have landed here
        }
    }
    // Unregistered: local kernels, then the in-place NCCL all-gather.  C25519_SHARD_MODE=1 selects the two-slice pipeline
    // (first slice's inversion + exchange on a side stream under the second slice's ladder); it was measured SLOWER at
    // N = 2 and N = 8 (the finish kernels get no SM slots while a ladder grid fills the machine), so it is not the default.
    static const int shard_mode = [] { const char* e = getenv("C25519_SHARD_MODE"); return e ? atoi(e) : 0; }();
    if (n_local < kPipelineMinOps || shard_mode == 0) {
        CK(launch_x25519_ladder(mine, pk32_local, sk32_local_inout, n_local, s));
        if (world > 1) return exchange_all(D, out_all, 32, n_local, world, rank, nccl_comm, s);
        return 0;
    }
    static const int shard_den = [] { const char* e = getenv("C25519_SHARD_TAIL_DEN"); return e ? atoi(e) : 8; }();  // last slice = 1/den
    const size_t cut = (n_local - n_local / (size_t)shard_den + 127) & ~(size_t)127;
    const size_t bounds[3] = {0, cut, n_local};
    Exchange x;
    return x25519_pipelined(D, mine, pk32_local, sk32_local_inout, n_local, s, bounds, 2,
        [&](cudaStream_t side) -> int { return world > 1 ? exchange_begin(x, D, out_all, nccl_comm, side) : 0; },
        [&](size_t row0, size_t cnt, bool first, bool last, cudaStream_t side) -> int {
            return world > 1 ? exchange_rows(x, out_all, 32, n_local, row0, cnt, first, last, world, rank, nccl_comm, side) : 0;
        });
}

int c25519_x25519_public_sharded(uint8_t* pk_all, uint8_t* sk32_local_inout, size_t n_local, int ladder, void* nccl_comm, void* stream)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (n_local == 0) return 0;
    if (!pk_all || !sk32_local_inout) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(pk_all) || misaligned32(sk32_local_inout)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(pk_all, sk32_local_inout);
    uint8_t* mine = pk_all + (size_t)rank * n_local * 32;
    if (ladder) CK(launch_x25519_ladder(mine, nullptr, sk32_local_inout, n_local, (cudaStream_t)stream));
    else CK(launch_x25519_comb(mine, sk32_local_inout, n_local, D.comb, (cudaStream_t)stream));
    return world > 1 ? exchange_all(D, pk_all, 32, n_local, world, rank, nccl_comm, (cudaStream_t)stream) : 0;
}

int c25519_ed25519_sign_sharded(uint8_t* sig_all, const uint8_t* priv64_local, const uint8_t* msgs_local, const uint64_t* msg_off_local,
                                size_t fixed_len, size_t n_local, void* nccl_comm, void* stream)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (n_local == 0) return 0;
    if (!sig_all || !priv64_local) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig_all) || misaligned32(priv64_local)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(sig_all, priv64_local);
    CK(launch_ed25519_sign(sig_all + (size_t)rank * n_local * 64, priv64_local, msgs_local, msg_off_local, fixed_len, n_local, D.comb, (cudaStream_t)stream));
    return world > 1 ? exchange_all(D, sig_all, 64, n_local, world, rank, nccl_comm, (cudaStream_t)stream) : 0;
}

int c25519_ed25519_verify_sharded(int32_t* ok_all, const uint8_t* sig64_local, const uint8_t* pk32_local, const uint8_t* msgs_local,
                                  const uint64_t* msg_off_local, size_t fixed_len, size_t n_local, void* nccl_comm, void* stream)
{
    int world = 0, rank = 0;
    if (int rc = comm_shape(nccl_comm, &world, &rank)) return rc;
    if (n_local == 0) return 0;
    if (!ok_all || !sig64_local || !pk32_local) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (misaligned32(sig64_local) || misaligned32(pk32_local)) return fail(C25519_E_BAD_ARGUMENT, "record arrays must be 32-byte aligned");
    BATCH_PROLOGUE(ok_all, sig64_local, pk32_local);
    CK(launch_ed25519_verify(ok_all + (size_t)rank * n_local, sig64_local, pk32_local, msgs_local, msg_off_local, fixed_len, n_local, D.comb, (cudaStream_t)stream));
    return world > 1 ? exchange_all(D, reinterpret_cast<uint8_t*>(ok_all), 4, n_local, world, rank, nccl_comm, (cudaStream_t)stream) : 0;
}

// ------------------------------------------------------------------ host-pointer batch API
int c25519_x25519_shared_host(uint8_t* out32, const uint8_t* pk32, uint8_t* sk32_inout, size_t n)
{
    if (n && (!out32 || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, out32}, {32, true, false, pk32, nullptr}, {32, true, true, sk32_inout, sk32_inout}};
    return run_host_pipeline(f, n, true, [](uint8_t** d, size_t cnt, cudaStream_t s, Device&) { return launch_x25519_ladder(d[0], d[1], d[2], cnt, s); });
}

int c25519_x25519_scalarmult_raw_host(uint8_t* out32, const uint8_t* point32, const uint8_t* scalar32, size_t n)
{
    if (n && (!out32 || !point32 || !scalar32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, out32}, {32, true, false, point32, nullptr}, {32, true, false, scalar32, nullptr}};
    return run_host_pipeline(f, n, true, [](uint8_t** d, size_t cnt, cudaStream_t s, Device&) { return launch_x25519_ladder_raw(d[0], d[1], d[2], cnt, s); });
}

int c25519_x25519_public_host(uint8_t* pk32, uint8_t* sk32_inout, size_t n, int ladder)
{
    if (n && (!pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[2] = {{32, false, true, nullptr, pk32}, {32, true, true, sk32_inout, sk32_inout}};
    return run_host_pipeline(f, n, true, [ladder](uint8_t** d, size_t cnt, cudaStream_t s, Device& D) {
        return ladder ? launch_x25519_ladder(d[0], nullptr, d[1], cnt, s) : launch_x25519_comb(d[0], d[1], cnt, D.comb, s);
    });
}

int c25519_x25519_shared_kdf_host(uint8_t* key_out, size_t key_size, const uint8_t* pk32, uint8_t* sk32_inout, size_t n)
{
    if (n && (!key_out || !pk32 || !sk32_inout)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    if (key_size == 0 || key_size > 64) return fail(C25519_E_BAD_ARGUMENT, "key_size must be 1..64 (bytes of the SHA-512 digest)");
    Field f[3] = {{key_size, false, true, nullptr, key_out}, {32, true, false, pk32, nullptr}, {32, true, true, sk32_inout, sk32_inout}};
    return run_host_pipeline(f, n, true, [key_size](uint8_t** d, size_t cnt, cudaStream_t s, Device&) {
        return launch_x25519_shared_kdf(d[0], (unsigned)key_size, d[1], d[2], cnt, s);
    });
}

int c25519_ed25519_keypair_host(uint8_t* pub32, uint8_t* priv64, const uint8_t* seed32, size_t n)
{
    if (n && (!pub32 || !priv64 || !seed32)) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    Field f[3] = {{32, false, true, nullptr, pub32}, {64, false, true, nullptr, priv64}, {32, true, false, seed32, nullptr}};
    return run_host_pipeline(f, n, true, [](uint8_t** d, size_t cnt, cudaStream_t s, Device& D) {
        return launch_ed25519_keypair(d[0], d[1], d[2], cnt, D.comb, s);
    });
}

int c25519_modl_host(int op, uint8_t* out32, const uint8_t* a32, const uint8_t* b32, size_t n)
{
    if (op < C25519_MODL_MULMOD || op > C25519_MODL_INVMOD) return fail(C25519_E_BAD_ARGUMENT, "unknown mod-L operation");
    if (n && (!out32 || !a32 || (op != C25519_MODL_INVMOD && !b32))) return fail(C25519_E_BAD_ARGUMENT, "null pointer");
    const bool two = b32 != nullptr;
    Field f[3] = {{32, false, true, nullptr, out32}, {32, true, false, a32, nullptr}, {32, two, false, b32, nullptr}};
    return run_host_pipeline(f, n, true, [op, two](uint8_t** d, size_t cnt, cudaStream_t s, Device&) {
        return launch_modl(op, d[0], d[1], two ? d[2] : nullptr, cnt, s);
    });
}

// Ed25519 with messages.  Fixed-length messages ride the generic pipeline as one more record field.  Ragged messages are
// sliced by operation count AND message bytes (<= 64 MB of message data per slice unless a single message is larger), each
// slice staged through a pipeline stage: offsets are uploaded unmodified and the kernels get a message base pointer
// shifted back by the slice's first offset, so nothing is rebased on the host.
static int ed25519_host_msgs(bool sign, uint8_t* out, const uint8_t* in64, const uint8_t* pk32, const uint8_t* msgs,
                             const uint64_t* msg_off, size_t fixed_len, size_t n)
{
    if (!msg_off) {
        const size_t ml = fixed_len ? fixed_len : 1;   // zero-length records still need a non-zero stride for staging
        if (sign) {
            Field f[3] = {{64, false, true, nullptr, out}, {64, true, false, in64, nullptr}, {ml, fixed_len != 0, false, msgs, nullptr}};
            return run_host_pipeline(f, n, true, [fixed_len](uint8_t** d, size_t cnt, cudaStream_t s, Device& D) {
                return launch_ed25519_sign(d[0], d[1], d[2], nullptr, fixed_len, cnt, D.comb, s);
            });
        }
        Field f[4] = {{4, false, true, nullptr, out}, {64, true, false, in64, nullptr}, {32, true, false, pk32, nullptr},
                      {ml, fixed_len != 0, false, msgs, nullptr}};
        return run_host_pipeline(f, n, false, [fixed_len](uint8_t** d, size_t cnt, cudaStream_t s, Device& D) {
            return launch_ed25519_verify(reinterpret_cast<int32_t*>(d[0]), d[1], d[2], d[3], nullptr, fixed_len, cnt, D.comb, s);
        });
    }
    int dev = 0;
    if (int rc = default_device(&dev)) return rc;
    if (n == 0) return 0;
    DeviceGuard g(dev);
    Device& D = g_dev[dev];
    Pipeline* P = acquire(D);
    if (!P) return fail(C25519_E_OUT_OF_MEMORY, "cannot create pipeline streams");
    const size_t out_rec = sign ? 64 : 4;
    auto body = [&]() -> int {
        int s = 0;
        for (size_t base = 0; base < n; s = (s + 1) % kStages) {
            size_t cnt = 1;
            while (base + cnt < n && cnt < kChunkOps && msg_off[base + cnt + 1] - msg_off[base] <= kRaggedSliceBytes) cnt++;
            if (msg_off[base + cnt] < msg_off[base]) return fail(C25519_E_BAD_ARGUMENT, "msg_off must be non-decreasing");
            const size_t bytes = (size_t)(msg_off[base + cnt] - msg_off[base]);
            const size_t o_out = 0, o_in = o_out + align_up(out_rec * cnt, 256), o_pk = o_in + align_up(64 * cnt, 256),
                         o_off = o_pk + align_up(32 * cnt, 256), o_msg = o_off + align_up(8 * (cnt + 1), 256),
                         total = o_msg + align_up(bytes ? bytes : 1, 256);
            Stage& st = P->st[s];
            if (int rc = reserve(st, total)) return rc;
            st.used = std::max(st.used, total);
            uint8_t* b = st.dev;
            CK(cudaMemcpyAsync(b + o_in, in64 + 64 * base, 64 * cnt, cudaMemcpyHostToDevice, st.stream));
            if (!sign) CK(cudaMemcpyAsync(b + o_pk, pk32 + 32 * base, 32 * cnt, cudaMemcpyHostToDevice, st.stream));
            CK(cudaMemcpyAsync(b + o_off, msg_off + base, 8 * (cnt + 1), cudaMemcpyHostToDevice, st.stream));
            if (bytes) CK(cudaMemcpyAsync(b + o_msg, msgs + msg_off[base], bytes, cudaMemcpyHostToDevice, st.stream));
            const uint8_t* vmsgs = b + o_msg - msg_off[base];       // kernels add the absolute offsets back
            const uint64_t* d_off = reinterpret_cast<const uint64_t*>(b + o_off);
            if (sign) CK(launch_ed25519_sign(b + o_out, b + o_in, vmsgs, d_off, 0, cnt, D.comb, st.stream));
            else CK(launch_ed25519_verify(reinterpret_cast<int32_t*>(b + o_out), b + o_in, b + o_pk, vmsgs, d_off, 0, cnt, D.comb, st.stream));
            CK(cudaMemcpyAsync(out + out_rec * base, b + o_out, out_rec * cnt, cudaMemcpyDeviceToHost, st.stream));
            base += cnt;
        }
        return 0;
    };
    int rc = body();
    char saved[sizeof t_err]; memcpy(saved, t_err, sizeof saved);
    int rc2 = drain(P, sign);
    if (rc) memcpy(t_err, saved, sizeof saved);
    release(D, P);
    return rc ? rc : rc2;
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

// ------------------------------------------------------------------ the reference's C API (n = 1)
static void die_if(int rc, const char* fn)
{
    if (rc == 0) return;
    fprintf(stderr, "libcurve25519_b200: %s failed (%d): %s -- there is no CPU fallback\n", fn, rc, c25519_last_error());
    abort();
}

// RFC 7748 clamp on a caller's host buffer (curve25519_utils.c:28-32); the batch kernels clamp on the device
void ecp_TrimSecretKey(unsigned char* sk) { sk[0] &= 0xf8; sk[31] = (unsigned char)((sk[31] | 0x40) & 0x7f); }

// Generic k*P exported by the reference's library (source/curve25519_mehdi.h:93) and used by its self-test:
// K is `len` little-endian bytes, not clamped, not modified.  The reference walks all `len` bytes; scalars wider than
// 256 bits have no batched representation here, so non-zero bytes beyond the 32nd are refused loudly, not truncated.
void ecp_PointMultiply(unsigned char* Q, const unsigned char* P, const unsigned char* K, int len)
{
    unsigned char k[32] = {0};
    for (int i = 32; i < len; i++)
        if (K[i]) { snprintf(t_err, sizeof t_err, "scalar wider than 256 bits (len = %d)", len); die_if(C25519_E_BAD_ARGUMENT, "ecp_PointMultiply"); }
    if (len > 32) len = 32;
    if (len > 0) memcpy(k, K, (size_t)len);
    die_if(c25519_x25519_scalarmult_raw_host(Q, P, k, 1), "ecp_PointMultiply");
    memset(k, 0, sizeof k);
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
    memset(seed, 0, sizeof seed);
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
static int verify_init_one(uint8_t* ctx, const unsigned char* publicKey)
{
    int dev = 0;
    if (int rc = default_device(&dev)) return rc;
    DeviceGuard g(dev);
    Device& D = g_dev[dev];
    Pipeline* P = acquire(D);
    if (!P) return fail(C25519_E_OUT_OF_MEMORY, "cannot create pipeline streams");
    Stage& st = P->st[0];
    int rc = reserve(st, 4096);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(st.dev + 2304, publicKey, 32, cudaMemcpyHostToDevice, st.stream);
        if (e == cudaSuccess) e = launch_ed25519_verify_init(st.dev, st.dev + 2304, 1, st.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx, st.dev, C25519_VERIFY_CTX_BYTES, cudaMemcpyDeviceToHost, st.stream);
        cudaError_t e2 = cudaStreamSynchronize(st.stream);
        if (e == cudaSuccess) e = e2;
        if (e != cudaSuccess) rc = fail((int)e, "ed25519_Verify_Init");
    }
    release(D, P);
    return rc;
}
void* ed25519_Verify_Init(void* context, const unsigned char* publicKey)
{
    uint8_t* ctx = static_cast<uint8_t*>(context);
    if (!ctx) ctx = static_cast<uint8_t*>(malloc(C25519_VERIFY_CTX_BYTES));
    if (!ctx) return nullptr;
    die_if(verify_init_one(ctx, publicKey), "ed25519_Verify_Init");
    return ctx;
}

static int verify_check_one(int32_t* ok, const void* context, const unsigned char* signature, const unsigned char* msg, size_t msg_size)
{
    int dev = 0;
    if (int rc = default_device(&dev)) return rc;
    DeviceGuard g(dev);
    Device& D = g_dev[dev];
    Pipeline* P = acquire(D);
    if (!P) return fail(C25519_E_OUT_OF_MEMORY, "cannot create pipeline streams");
    Stage& st = P->st[0];
    int rc = reserve(st, 4096 + align_up(msg_size + 1, 256));
    if (!rc) {
        uint8_t* d_ctx = st.dev; uint8_t* d_sig = st.dev + 2304; int32_t* d_ok = reinterpret_cast<int32_t*>(st.dev + 2304 + 64);
        uint8_t* d_msg = st.dev + 4096;
        cudaError_t e = cudaMemcpyAsync(d_ctx, context, C25519_VERIFY_CTX_BYTES, cudaMemcpyHostToDevice, st.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_sig, signature, 64, cudaMemcpyHostToDevice, st.stream);
        if (e == cudaSuccess && msg_size) e = cudaMemcpyAsync(d_msg, msg, msg_size, cudaMemcpyHostToDevice, st.stream);
        if (e == cudaSuccess) e = launch_ed25519_verify_check(d_ok, d_ctx, nullptr, d_sig, d_msg, nullptr, msg_size, 1, D.comb, st.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ok, d_ok, 4, cudaMemcpyDeviceToHost, st.stream);
        cudaError_t e2 = cudaStreamSynchronize(st.stream);
        if (e == cudaSuccess) e = e2;
        if (e != cudaSuccess) rc = fail((int)e, "ed25519_Verify_Check");
    }
    release(D, P);
    return rc;
}
int ed25519_Verify_Check(const void* context, const unsigned char* signature, const unsigned char* msg, size_t msg_size)
{
    int32_t ok = 0;
    die_if(verify_check_one(&ok, context, signature, msg, msg_size), "ed25519_Verify_Check");
    return ok;
}

void ed25519_Verify_Finish(void* ctx) { free(ctx); }

}  // extern "C"
