// tools/ubench.cu -- integer-pipe micro-benchmarks for sm_100a (B200).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Prints, for several instruction mixes, warp-instructions per clock per SM (max 4 = one per SMSP per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* sink, int iters, unsigned long long* cyc)
{
    uint32_t a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
    uint32_t b0=a0*3,b1=a1*3,b2=a2*3,b3=a3*3,b4=a4*3,b5=a5*3,b6=a6*3,b7=a7*3;
    uint32_t x=0x9e3779b9u^threadIdx.x, y=0x85ebca6bu+blockIdx.x;
    unsigned long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<iters;it++){
        #pragma unroll
        for(int r=0;r<4;r++){
        if(MODE==0){ // IMAD.WIDE.U32 64-bit accumulate, 4 independent chains of (lo,hi)
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.u32 %1,%8,%9,%1;\n\t"
                         "mad.lo.cc.u32 %2,%8,%9,%2; madc.hi.u32 %3,%8,%9,%3;\n\t"
                         "mad.lo.cc.u32 %4,%8,%9,%4; madc.hi.u32 %5,%8,%9,%5;\n\t"
                         "mad.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.u32 %1,%8,%9,%1;\n\t"
                         "mad.lo.cc.u32 %2,%8,%9,%2; madc.hi.u32 %3,%8,%9,%3;\n\t"
                         "mad.lo.cc.u32 %4,%8,%9,%4; madc.hi.u32 %5,%8,%9,%5;\n\t"
                         "mad.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(b0),"+r"(b1),"+r"(b2),"+r"(b3),"+r"(b4),"+r"(b5),"+r"(b6),"+r"(b7):"r"(x),"r"(y));
        } else if(MODE==1){ // carry-chained IMAD.WIDE.U32.X (two chains of 4)
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1;\n\t"
                         "madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;\n\t"
                         "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5;\n\t"
                         "madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1;\n\t"
                         "madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;\n\t"
                         "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5;\n\t"
                         "madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(b0),"+r"(b1),"+r"(b2),"+r"(b3),"+r"(b4),"+r"(b5),"+r"(b6),"+r"(b7):"r"(x),"r"(y));
        } else if(MODE==2){ // IMAD lo 32-bit, 8 independent
            asm volatile("mad.lo.u32 %0,%8,%9,%0; mad.lo.u32 %1,%8,%9,%1; mad.lo.u32 %2,%8,%9,%2; mad.lo.u32 %3,%8,%9,%3;\n\t"
                         "mad.lo.u32 %4,%8,%9,%4; mad.lo.u32 %5,%8,%9,%5; mad.lo.u32 %6,%8,%9,%6; mad.lo.u32 %7,%8,%9,%7;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
        } else if(MODE==3){ // IADD3 with carry chains: 8 adds
            asm volatile("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%8; addc.u32 %3,%3,%9;\n\t"
                         "add.cc.u32 %4,%4,%8; addc.cc.u32 %5,%5,%9; addc.cc.u32 %6,%6,%8; addc.u32 %7,%7,%9;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
        } else if(MODE==4){ // mix: 8 IMAD.WIDE (a-chains) + 8 IADD3.X (b-chains), independent
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.u32 %1,%8,%9,%1;\n\t"
                         "mad.lo.cc.u32 %2,%8,%9,%2; madc.hi.u32 %3,%8,%9,%3;\n\t"
                         "mad.lo.cc.u32 %4,%8,%9,%4; madc.hi.u32 %5,%8,%9,%5;\n\t"
                         "mad.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
            asm volatile("add.cc.u32 %0,%0,%4; addc.cc.u32 %1,%1,%5; addc.cc.u32 %2,%2,%4; addc.u32 %3,%3,%5;"
                         :"+r"(b0),"+r"(b1),"+r"(b2),"+r"(b3):"r"(x),"r"(y));
        } else if(MODE==5){ // mix 1:1 : 4 IMAD.WIDE + 4 IADD3
            asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.u32 %1,%8,%9,%1;\n\t"
                         "mad.lo.cc.u32 %2,%8,%9,%2; madc.hi.u32 %3,%8,%9,%3;\n\t"
                         "mad.lo.cc.u32 %4,%8,%9,%4; madc.hi.u32 %5,%8,%9,%5;\n\t"
                         "mad.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
            asm volatile("add.cc.u32 %0,%0,%4; addc.cc.u32 %1,%1,%5; addc.cc.u32 %2,%2,%4; addc.u32 %3,%3,%5;"
                         :"+r"(b0),"+r"(b1),"+r"(b2),"+r"(b3):"r"(x),"r"(y));
        } else if(MODE==6){ // SEL / LOP3 (alu pipe)
            asm volatile("lop3.b32 %0,%0,%8,%9,0x96; lop3.b32 %1,%1,%8,%9,0x96; lop3.b32 %2,%2,%8,%9,0x96; lop3.b32 %3,%3,%8,%9,0x96;\n\t"
                         "lop3.b32 %4,%4,%8,%9,0x96; lop3.b32 %5,%5,%8,%9,0x96; lop3.b32 %6,%6,%8,%9,0x96; lop3.b32 %7,%7,%8,%9,0x96;"
                         :"+r"(a0),"+r"(a1),"+r"(a2),"+r"(a3),"+r"(a4),"+r"(a5),"+r"(a6),"+r"(a7):"r"(x),"r"(y));
        }
        }
    }
    unsigned long long t1=clock64();
    uint32_t r=a0^a1^a2^a3^a4^a5^a6^a7^b0^b1^b2^b3^b4^b5^b6^b7;
    if(r==0x12345u) sink[0]=r;
    if(threadIdx.x==0 && blockIdx.x==0) cyc[0]=t1-t0;
}

template<int MODE> int run(const char* name, int instr_per_trip, int macs_per_trip, uint32_t* sink, unsigned long long* cyc, int warps_per_sm)
{
    int iters=2048; int threads=256; int blocks_per_sm=warps_per_sm*32/threads; if(blocks_per_sm<1){blocks_per_sm=1; threads=warps_per_sm*32;}
    int grid=148*blocks_per_sm;
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid,threads>>>(sink,64,cyc); CHK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k<MODE><<<grid,threads>>>(sink,iters,cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); unsigned long long c; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost);
    double winstr=(double)iters*4*instr_per_trip*warps_per_sm; // per SM
    printf("%-28s warps/SM=%2d  cycles=%llu  warp-instr/clk/SM=%.3f  (clk=%.0f MHz)  ", name, warps_per_sm, c, winstr/c, c/(ms*1e3));
    if(macs_per_trip) printf("MAC32/s=%.3e", (double)iters*4*macs_per_trip*grid*threads/(ms*1e-3));
    printf("\n"); return 0;
}
int main(){
    uint32_t* sink; unsigned long long* cyc; CHK(cudaMalloc(&sink,64)); CHK(cudaMalloc(&cyc,64));
    int ws[]={4,8,16,32};
    for(int w: ws){
        run<0>("IMAD.WIDE.U32 (acc64)",8,8,sink,cyc,w);
        run<1>("IMAD.WIDE.U32.X chained",8,8,sink,cyc,w);
        run<2>("IMAD lo",8,0,sink,cyc,w);
        run<3>("IADD3(.X)",8,0,sink,cyc,w);
        run<4>("4 IMAD.WIDE + 4 IADD3.X",8,4,sink,cyc,w);
        run<6>("LOP3",8,0,sink,cyc,w);
    }
    return 0;
}
