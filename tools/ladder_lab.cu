// tools/ladder_lab.cu -- launch-shape / ILP variants of the X25519 ladder kernel, timed side by side.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ladder_lab tools/ladder_lab.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../curve25519_b200/csrc/x25519.cuh"
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_ladder(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ sk32, size_t n)
{
    __shared__ u32 ks[8][T];
    const size_t i=(size_t)blockIdx.x*T+threadIdx.x; if(i>=n) return;
    fe k; fe_load(k, sk32+32*i); k.v[0]&=0xfffffff8u; k.v[7]=(k.v[7]|0x40000000u)&0x7fffffffu;
    #pragma unroll
    for(int w=0;w<8;w++) ks[w][threadIdx.x]=k.v[w];
    fe u; fe_load(u, pk32+32*i); fe r; const int t=threadIdx.x;
    x25519_ladder(r,u,[&](int w){return ks[w][t];});
    fe_store(out32+32*i,r);
}

// base point kept in shared memory (one column per thread) instead of 8 registers
template<int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_ladder_sb(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ sk32, size_t n)
{
    __shared__ u32 ks[8][T]; __shared__ u32 us[8][T];
    const size_t i=(size_t)blockIdx.x*T+threadIdx.x; if(i>=n) return;
    const int t=threadIdx.x;
    fe k; fe_load(k, sk32+32*i); k.v[0]&=0xfffffff8u; k.v[7]=(k.v[7]|0x40000000u)&0x7fffffffu;
    #pragma unroll
    for(int w=0;w<8;w++) ks[w][t]=k.v[w];
    fe R0X,R0Z,R1X,R1Z;
    { fe u; fe_load(u, pk32+32*i);
      #pragma unroll
      for(int w=0;w<8;w++) us[w][t]=u.v[w];
      fe_copy(R0X,u); fe_set_u32(R0Z,1); mont_double(R1X,R1Z,R0X,R0Z); fe_narrow(R0X); }
    bool cur=true;
    #pragma unroll 1
    for(int bit=253;bit>=0;--bit){
        bool b=(ks[bit>>5][t]>>(bit&31))&1u; bool s=(b!=cur);
        fe_cswap(R0X,R1X,s); fe_cswap(R0Z,R1Z,s); cur=b;
        mont_step_with(R0X,R0Z,R1X,R1Z,[&](fe& bb){
            #pragma unroll
            for(int w=0;w<8;w++) bb.v[w]=us[w][t]; });
    }
    fe PX,PZ,zi,r; fe_select(PX,R1X,R0X,cur); fe_select(PZ,R1Z,R0Z,cur); fe_invert(zi,PZ); fe_mul(r,PX,zi); fe_canon(r); fe_store(out32+32*i,r);
}

// two operations per thread, steps interleaved in one loop body
template<int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_ladder2(uint8_t* __restrict__ out32, const uint8_t* __restrict__ pk32, const uint8_t* __restrict__ sk32, size_t n)
{
    __shared__ u32 ks[16][T];
    const size_t i0=((size_t)blockIdx.x*T+threadIdx.x); const size_t half=(n+1)/2; if(i0>=half) return;
    const size_t i1 = (i0+half<n)? i0+half : i0;
    const int t=threadIdx.x;
    fe ka,kb; fe_load(ka, sk32+32*i0); fe_load(kb, sk32+32*i1);
    ka.v[0]&=0xfffffff8u; ka.v[7]=(ka.v[7]|0x40000000u)&0x7fffffffu; kb.v[0]&=0xfffffff8u; kb.v[7]=(kb.v[7]|0x40000000u)&0x7fffffffu;
    #pragma unroll
    for(int w=0;w<8;w++){ ks[w][t]=ka.v[w]; ks[8+w][t]=kb.v[w]; }
    fe ua,ub; fe_load(ua, pk32+32*i0); fe_load(ub, pk32+32*i1);
    fe A0X,A0Z,A1X,A1Z,B0X,B0Z,B1X,B1Z;
    fe_copy(A0X,ua); fe_set_u32(A0Z,1); mont_double(A1X,A1Z,A0X,A0Z);
    fe_copy(B0X,ub); fe_set_u32(B0Z,1); mont_double(B1X,B1Z,B0X,B0Z);
    { fe one; fe_set_u32(one,1); fe_mul(A0X,A0X,one); fe_mul(B0X,B0X,one); }
    bool ca=true, cb=true;
    #pragma unroll 1
    for(int bit=253;bit>=0;--bit){
        bool ba=(ks[bit>>5][t]>>(bit&31))&1u, bb=(ks[8+(bit>>5)][t]>>(bit&31))&1u;
        bool sa=(ba!=ca), sb=(bb!=cb);
        fe_cswap(A0X,A1X,sa); fe_cswap(A0Z,A1Z,sa); fe_cswap(B0X,B1X,sb); fe_cswap(B0Z,B1Z,sb);
        ca=ba; cb=bb;
        mont_step(A0X,A0Z,A1X,A1Z,ua);
        mont_step(B0X,B0Z,B1X,B1Z,ub);
    }
    fe PX,PZ,zi,r;
    fe_select(PX,A1X,A0X,ca); fe_select(PZ,A1Z,A0Z,ca); fe_invert(zi,PZ); fe_mul(r,PX,zi); fe_canon(r); fe_store(out32+32*i0,r);
    fe_select(PX,B1X,B0X,cb); fe_select(PZ,B1Z,B0Z,cb); fe_invert(zi,PZ); fe_mul(r,PX,zi); fe_canon(r); fe_store(out32+32*i1,r);
}

struct Ctx { uint8_t *out,*ref,*pk,*sk; size_t n; std::vector<uint8_t> href; };
template<typename K> int timeit(const char* name, K kern, int T, size_t nthreads, Ctx& c, bool first)
{
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,kern); int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,kern,T,0);
    unsigned grid=(unsigned)((nthreads+T-1)/T);
    cudaMemset(c.out,0,32*c.n);
    kern<<<grid,T>>>(c.out,c.pk,c.sk,c.n); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); kern<<<grid,T>>>(c.out,c.pk,c.sk,c.n); cudaEventRecord(e1); CHK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    std::vector<uint8_t> h(32*c.n); cudaMemcpy(h.data(),c.out,32*c.n,cudaMemcpyDeviceToHost);
    if(first) c.href=h;
    bool same = memcmp(h.data(),c.href.data(),32*c.n)==0;
    printf("%-28s regs=%3d spill=%4zuB occ=%2d blk/SM (%2d warps/SM)  %8.3f ms  %7.2f Mops/s  %s\n", name, fa.numRegs, (size_t)fa.localSizeBytes, occ, occ*T/32, best, c.n/(best*1e3), same?"OK":"MISMATCH");
    return 0;
}
int main(int argc,char**argv){
    Ctx c; c.n = (argc>1)? (size_t)atol(argv[1]) : (size_t)1<<20;
    CHK(cudaMalloc(&c.out,32*c.n)); CHK(cudaMalloc(&c.pk,32*c.n)); CHK(cudaMalloc(&c.sk,32*c.n));
    std::vector<uint8_t> h(32*c.n); uint64_t s=0x9e3779b97f4a7c15ull; 
    for(size_t i=0;i<32*c.n;i++){ s^=s<<13; s^=s>>7; s^=s<<17; h[i]=(uint8_t)(s>>24);} cudaMemcpy(c.pk,h.data(),32*c.n,cudaMemcpyHostToDevice);
    for(size_t i=0;i<32*c.n;i++){ s^=s<<13; s^=s>>7; s^=s<<17; h[i]=(uint8_t)(s>>24);} cudaMemcpy(c.sk,h.data(),32*c.n,cudaMemcpyHostToDevice);
    timeit("T128 min1 (ptxas free)", k_ladder<128,1>,128,c.n,c,true);
    timeit("smem-base T128 min5",    k_ladder_sb<128,5>,128,c.n,c,false);
    timeit("smem-base T128 min6",    k_ladder_sb<128,6>,128,c.n,c,false);
    timeit("smem-base T128 min7",    k_ladder_sb<128,7>,128,c.n,c,false);
    timeit("T128 min4 (<=128 regs)", k_ladder<128,4>,128,c.n,c,false);
    timeit("T128 min5 (<=96 regs)",  k_ladder<128,5>,128,c.n,c,false);
    timeit("T128 min6 (<=80 regs)",  k_ladder<128,6>,128,c.n,c,false);
    timeit("T64  min8",              k_ladder<64,8>,64,c.n,c,false);
    timeit("T256 min2",              k_ladder<256,2>,256,c.n,c,false);
    { unsigned long long hsh=1469598103934665603ull; for(size_t i=0;i<c.href.size();i++){ hsh^=c.href[i]; hsh*=1099511628211ull; } printf("output fnv1a64 = %016llx\n", hsh); }
    return 0;
}
