// tools/ubench2.cu -- what do the field primitives and their building blocks cost per SMSP on B200?
// Every kernel runs as ONE wave (148 x BPS CTAs of 256 threads), clock64() brackets the loop in CTA 0,
// and we report  cycles * 4 SMSP * 148 / (warp-level operations)  = cycles per operation per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../curve25519_b200/csrc/fe25519.cuh"
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* sink, const uint32_t* src, int iters, unsigned long long* cyc)
{
    fe x, y, z, w;
    const int t = blockIdx.x*256+threadIdx.x;
    for(int i=0;i<8;i++){ x.v[i]=src[(t*8+i)&1023]; y.v[i]=src[(t*8+i+3)&1023]|1; z.v[i]=src[(t*5+i)&1023]; w.v[i]=src[(t*7+i)&1023]; }
    __syncthreads();
    unsigned long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<iters;it++){
        if(MODE==0){ fe_mul(z,z,y); }
        else if(MODE==1){ fe_sqr(z,z); }
        else if(MODE==2){ fe_mul(z,z,y); fe_mul(w,w,x); }            // two independent chains
        else if(MODE==3){ fe_sqr(z,z); fe_sqr(w,w); }
        else if(MODE==4){ fe_add_nn(z,z,y); fe_add_nn(w,w,x); z.v[0]^=w.v[1]; }
        else if(MODE==5){ fe_sub(z,z,y); fe_sub(w,w,x); z.v[0]^=w.v[1]; }
        else if(MODE==6){ // plain IMAD.WIDE 64-bit accumulate, 8 independent, distinct multipliers
            unsigned long long a0=((unsigned long long)z.v[1]<<32)|z.v[0],a1=((unsigned long long)z.v[3]<<32)|z.v[2],a2=((unsigned long long)z.v[5]<<32)|z.v[4],a3=((unsigned long long)z.v[7]<<32)|z.v[6];
            unsigned long long b0=((unsigned long long)w.v[1]<<32)|w.v[0],b1=((unsigned long long)w.v[3]<<32)|w.v[2],b2=((unsigned long long)w.v[5]<<32)|w.v[4],b3=((unsigned long long)w.v[7]<<32)|w.v[6];
            #pragma unroll
            for(int r=0;r<8;r++){
              asm volatile("mad.wide.u32 %0,%8,%16,%0; mad.wide.u32 %1,%9,%16,%1; mad.wide.u32 %2,%10,%16,%2; mad.wide.u32 %3,%11,%16,%3;\n\t"
                           "mad.wide.u32 %4,%12,%16,%4; mad.wide.u32 %5,%13,%16,%5; mad.wide.u32 %6,%14,%16,%6; mad.wide.u32 %7,%15,%16,%7;"
                :"+l"(a0),"+l"(a1),"+l"(a2),"+l"(a3),"+l"(b0),"+l"(b1),"+l"(b2),"+l"(b3)
                :"r"(x.v[0]),"r"(x.v[1]),"r"(x.v[2]),"r"(x.v[3]),"r"(x.v[4]),"r"(x.v[5]),"r"(x.v[6]),"r"(x.v[7]),"r"(y.v[r]));
            }
            z.v[0]=(u32)a0; z.v[1]=(u32)(a0>>32); z.v[2]=(u32)a1; z.v[3]=(u32)(a1>>32); z.v[4]=(u32)a2; z.v[5]=(u32)(a2>>32); z.v[6]=(u32)a3; z.v[7]=(u32)(a3>>32);
            w.v[0]=(u32)b0; w.v[1]=(u32)(b0>>32); w.v[2]=(u32)b1; w.v[3]=(u32)(b1>>32); w.v[4]=(u32)b2; w.v[5]=(u32)(b2>>32); w.v[6]=(u32)b3; w.v[7]=(u32)(b3>>32);
        }
        else if(MODE==7){ // carry-chained rows exactly like fe_mul's product phase: 8 rows x (4+4) = 64 IMAD.WIDE[.X] + 7 IADD3.X, no merge/reduce
            u32 A[16],B[16];
            #pragma unroll
            for(int i=0;i<8;i++){A[i]=z.v[i];A[8+i]=w.v[i];B[i]=w.v[i];B[8+i]=z.v[i];}
            const u32*a=x.v;
            mad_row(B+0,A+2,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[1]);
            mad_row(A+2,B+2,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[2]);
            mad_row(B+2,A+4,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[3]);
            mad_row(A+4,B+4,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[4]);
            mad_row(B+4,A+6,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[5]);
            mad_row(A+6,B+6,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[6]);
            mad_row(B+6,A+8,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[7]);
            mad_row(A+0,B+0,a[0],a[2],a[4],a[6],a[1],a[3],a[5],a[7],y.v[0]);
            #pragma unroll
            for(int i=0;i<8;i++){z.v[i]=A[i]^A[8+i];w.v[i]=B[i]^B[8+i];}
        }


    }
    unsigned long long t1=clock64();
    uint32_t r=0; for(int i=0;i<8;i++) r^=z.v[i]^w.v[i];
    if(r==0x12345u) sink[0]=r;
    if(threadIdx.x==0 && blockIdx.x==0) cyc[0]=t1-t0;
}

template<int MODE> int run(const char* name, double ops_per_iter, int bps, uint32_t* sink, uint32_t* src, unsigned long long* cyc)
{
    int iters=400;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<MODE>);
    int maxb=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb,k<MODE>,256,0);
    if(bps>maxb) bps=maxb;
    int grid=148*bps;
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid,256>>>(sink,src,20,cyc); CHK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k<MODE><<<grid,256>>>(sink,src,iters,cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); unsigned long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    double warps_per_smsp = bps*8/4.0;
    double cyc_per_op_smsp = (double)c/(iters*ops_per_iter*warps_per_smsp);
    printf("%-34s regs=%3d warps/SMSP=%4.1f  cycles/op/SMSP=%8.1f   (wall %.3f ms, clk~%.0f MHz)\n", name, fa.numRegs, warps_per_smsp, cyc_per_op_smsp, ms, c/(ms*1e3));
    return 0;
}
int main(){
    uint32_t* sink; uint32_t* src; unsigned long long* cyc; CHK(cudaMalloc(&sink,64)); CHK(cudaMalloc(&cyc,64)); CHK(cudaMalloc(&src,4096));
    uint32_t h[1024]; for(int i=0;i<1024;i++) h[i]=0x9e3779b9u*(i+1)^(0x85ebca6bu*(i*i+7)); cudaMemcpy(src,h,4096,cudaMemcpyHostToDevice);
    int bpss[]={1,2,4,8};
    for(int b: bpss){
        run<6>("IMAD.WIDE plain x64 (per 64)",1,b,sink,src,cyc);
        run<7>("mad_row chains x64 (+8 IADD)",1,b,sink,src,cyc);
        run<0>("fe_mul (dependent)",1,b,sink,src,cyc);
        run<2>("fe_mul x2 independent (per mul)",2,b,sink,src,cyc);
        run<1>("fe_sqr (dependent)",1,b,sink,src,cyc);
        run<3>("fe_sqr x2 independent (per sqr)",2,b,sink,src,cyc);
        run<4>("fe_add_nn x2 (per add)",2,b,sink,src,cyc);
        run<5>("fe_sub x2 (per sub)",2,b,sink,src,cyc);
    }
    return 0;
}
