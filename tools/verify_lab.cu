// tools/verify_lab.cu -- launch-bound (occupancy vs spills) variants of the two verification kernels, timed at n = 2^20.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -DC25519_VERIFY_INIT_MINB=<3|4|5> -DC25519_VERIFY_CHECK_MINB=<4|5|6> -o tools/verify_lab_<tag> tools/verify_lab.cu
// The kernels are the production translation units, included verbatim; inputs are random bytes (the work per item does
// not depend on the data, except one conditional multiplication in the decompression).
#include <cstdio>
#include <vector>
#include "../curve25519_b200/csrc/comb_table.cu"
#include "../curve25519_b200/csrc/x25519_kernels.cu"
#include "../curve25519_b200/csrc/ed25519_kernels.cu"
namespace c25519 { void count_launch() {} }
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
int main()
{
    const size_t n = (size_t)1 << 20;
    std::vector<uint32_t> padded((size_t)kCombEntries * kCombStrideWordsHost, 0u);
    for (int e = 0; e < kCombEntries; e++) for (int w = 0; w < kCombWordsPerEntry; w++) padded[(size_t)e * kCombStrideWordsHost + w] = kCombTableHost[e * kCombWordsPerEntry + w];
    uint32_t* table; CHK(cudaMalloc(&table, padded.size() * 4)); CHK(cudaMemcpy(table, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice));
    uint8_t *sig, *pk, *msgs, *ws; int32_t* ok;
    CHK(cudaMalloc(&sig, 64 * n)); CHK(cudaMalloc(&pk, 32 * n)); CHK(cudaMalloc(&msgs, 64 * n)); CHK(cudaMalloc(&ok, 4 * n)); CHK(cudaMalloc(&ws, 2080 * n));
    std::vector<uint8_t> h(64 * n); uint64_t s = 0x9e3779b97f4a7c15ull;
    for (auto& b : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b = (uint8_t)(s >> 24); }
    CHK(cudaMemcpy(sig, h.data(), 64 * n, cudaMemcpyHostToDevice)); CHK(cudaMemcpy(msgs, h.data(), 64 * n, cudaMemcpyHostToDevice)); CHK(cudaMemcpy(pk, h.data() + 7, 32 * n, cudaMemcpyHostToDevice));
    cudaFuncAttributes fi, fc; cudaFuncGetAttributes(&fi, k_ed25519_verify_init); cudaFuncGetAttributes(&fc, k_ed25519_verify_check<true>);
    int oi = 0, oc = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oi, k_ed25519_verify_init, 128, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, k_ed25519_verify_check<true>, 128, 0);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float bi = 1e9, bc = 1e9;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        CHK(launch_ed25519_verify_init(ws, pk, n, 0));
        cudaEventRecord(e1);
        CHK(launch_ed25519_verify_check(ok, ws, nullptr, sig, msgs, nullptr, 64, n, table, 0));
        cudaEventRecord(e2); CHK(cudaDeviceSynchronize());
        float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
        if (r) { if (a < bi) bi = a; if (b < bc) bc = b; }
    }
    std::vector<int32_t> hok(n); cudaMemcpy(hok.data(), ok, 4 * n, cudaMemcpyDeviceToHost); long sum = 0; for (auto v : hok) sum += v;
    printf("init: regs=%3d spill=%3zuB occ=%d CTAs/SM %7.3f ms | check(+normalize): regs=%3d spill=%3zuB occ=%d %7.3f ms | total %7.3f ms = %6.2f M verifies/s (ok sum %ld)\n",
           fi.numRegs, (size_t)fi.localSizeBytes, oi, bi, fc.numRegs, (size_t)fc.localSizeBytes, oc, bc, bi + bc, n / ((bi + bc) * 1e3), sum);
    return 0;
}
