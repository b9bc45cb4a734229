// tools/ubench3.cu -- cost of the carry forms of IMAD.WIDE and of the ALU-pipe instructions, per SMSP, at full occupancy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
typedef uint32_t u32;

template<int MODE>
__global__ void __launch_bounds__(256) k(u32* sink, const u32* src, int iters, unsigned long long* cyc)
{
    u32 a[16], x[8], y;
    const int t=blockIdx.x*256+threadIdx.x;
    #pragma unroll
    for(int i=0;i<16;i++) a[i]=src[(t*16+i)&1023];
    #pragma unroll
    for(int i=0;i<8;i++) x[i]=src[(t*3+i+7)&1023]|1;
    y=src[(t+5)&1023]|1;
    u32 s0=0,s1=0,s2=0,s3=0,s4=0,s5=0,s6=0,s7=0;
    __syncthreads();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g0));
    unsigned long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int r=0;r<4;r++){
        if(MODE==0){ // 8 plain IMAD.WIDE (lo.cc + hi consumes carry: fused, no predicate I/O)
          asm volatile(
            "mad.lo.cc.u32 %0,%16,%24,%0; madc.hi.u32 %1,%16,%24,%1;\n\t" "mad.lo.cc.u32 %2,%17,%24,%2; madc.hi.u32 %3,%17,%24,%3;\n\t"
            "mad.lo.cc.u32 %4,%18,%24,%4; madc.hi.u32 %5,%18,%24,%5;\n\t" "mad.lo.cc.u32 %6,%19,%24,%6; madc.hi.u32 %7,%19,%24,%7;\n\t"
            "mad.lo.cc.u32 %8,%20,%24,%8; madc.hi.u32 %9,%20,%24,%9;\n\t" "mad.lo.cc.u32 %10,%21,%24,%10; madc.hi.u32 %11,%21,%24,%11;\n\t"
            "mad.lo.cc.u32 %12,%22,%24,%12; madc.hi.u32 %13,%22,%24,%13;\n\t" "mad.lo.cc.u32 %14,%23,%24,%14; madc.hi.u32 %15,%23,%24,%15;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==1){ // 8 IMAD.WIDE each with carry-OUT only, carry consumed by an IADD3.X into s_k  (8 IMAD + 8 IADD3.X)
          asm volatile(
            "mad.lo.cc.u32 %0,%24,%32,%0; madc.hi.cc.u32 %1,%24,%32,%1; addc.u32 %16,%16,0;\n\t"
            "mad.lo.cc.u32 %2,%25,%32,%2; madc.hi.cc.u32 %3,%25,%32,%3; addc.u32 %17,%17,0;\n\t"
            "mad.lo.cc.u32 %4,%26,%32,%4; madc.hi.cc.u32 %5,%26,%32,%5; addc.u32 %18,%18,0;\n\t"
            "mad.lo.cc.u32 %6,%27,%32,%6; madc.hi.cc.u32 %7,%27,%32,%7; addc.u32 %19,%19,0;\n\t"
            "mad.lo.cc.u32 %8,%28,%32,%8; madc.hi.cc.u32 %9,%28,%32,%9; addc.u32 %20,%20,0;\n\t"
            "mad.lo.cc.u32 %10,%29,%32,%10; madc.hi.cc.u32 %11,%29,%32,%11; addc.u32 %21,%21,0;\n\t"
            "mad.lo.cc.u32 %12,%30,%32,%12; madc.hi.cc.u32 %13,%30,%32,%13; addc.u32 %22,%22,0;\n\t"
            "mad.lo.cc.u32 %14,%31,%32,%14; madc.hi.cc.u32 %15,%31,%32,%15; addc.u32 %23,%23,0;\n\t"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15]),"+r"(s0),"+r"(s1),"+r"(s2),"+r"(s3),"+r"(s4),"+r"(s5),"+r"(s6),"+r"(s7)
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==2){ // 4 chains of 2: first carry-out only, second carry-in only
          asm volatile(
            "mad.lo.cc.u32 %0,%16,%24,%0; madc.hi.cc.u32 %1,%16,%24,%1; madc.lo.cc.u32 %2,%17,%24,%2; madc.hi.u32 %3,%17,%24,%3;\n\t"
            "mad.lo.cc.u32 %4,%18,%24,%4; madc.hi.cc.u32 %5,%18,%24,%5; madc.lo.cc.u32 %6,%19,%24,%6; madc.hi.u32 %7,%19,%24,%7;\n\t"
            "mad.lo.cc.u32 %8,%20,%24,%8; madc.hi.cc.u32 %9,%20,%24,%9; madc.lo.cc.u32 %10,%21,%24,%10; madc.hi.u32 %11,%21,%24,%11;\n\t"
            "mad.lo.cc.u32 %12,%22,%24,%12; madc.hi.cc.u32 %13,%22,%24,%13; madc.lo.cc.u32 %14,%23,%24,%14; madc.hi.u32 %15,%23,%24,%15;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==3){ // 2 chains of 4 (like mad_row): 2 out-only, 4 in+out, 2 in-only
          asm volatile(
            "mad.lo.cc.u32 %0,%16,%24,%0; madc.hi.cc.u32 %1,%16,%24,%1; madc.lo.cc.u32 %2,%17,%24,%2; madc.hi.cc.u32 %3,%17,%24,%3;\n\t"
            "madc.lo.cc.u32 %4,%18,%24,%4; madc.hi.cc.u32 %5,%18,%24,%5; madc.lo.cc.u32 %6,%19,%24,%6; madc.hi.u32 %7,%19,%24,%7;\n\t"
            "mad.lo.cc.u32 %8,%20,%24,%8; madc.hi.cc.u32 %9,%20,%24,%9; madc.lo.cc.u32 %10,%21,%24,%10; madc.hi.cc.u32 %11,%21,%24,%11;\n\t"
            "madc.lo.cc.u32 %12,%22,%24,%12; madc.hi.cc.u32 %13,%22,%24,%13; madc.lo.cc.u32 %14,%23,%24,%14; madc.hi.u32 %15,%23,%24,%15;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==4){ // 1 chain of 8: 1 out-only, 6 in+out, 1 in-only
          asm volatile(
            "mad.lo.cc.u32 %0,%16,%24,%0; madc.hi.cc.u32 %1,%16,%24,%1; madc.lo.cc.u32 %2,%17,%24,%2; madc.hi.cc.u32 %3,%17,%24,%3;\n\t"
            "madc.lo.cc.u32 %4,%18,%24,%4; madc.hi.cc.u32 %5,%18,%24,%5; madc.lo.cc.u32 %6,%19,%24,%6; madc.hi.cc.u32 %7,%19,%24,%7;\n\t"
            "madc.lo.cc.u32 %8,%20,%24,%8; madc.hi.cc.u32 %9,%20,%24,%9; madc.lo.cc.u32 %10,%21,%24,%10; madc.hi.cc.u32 %11,%21,%24,%11;\n\t"
            "madc.lo.cc.u32 %12,%22,%24,%12; madc.hi.cc.u32 %13,%22,%24,%13; madc.lo.cc.u32 %14,%23,%24,%14; madc.hi.u32 %15,%23,%24,%15;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==5){ // 8 x (IMAD lo + IMAD.HI) separate 32-bit accumulators: 16 fma-pipe instr
          asm volatile(
            "mad.lo.u32 %0,%16,%24,%0; mad.hi.u32 %1,%16,%24,%1; mad.lo.u32 %2,%17,%24,%2; mad.hi.u32 %3,%17,%24,%3;\n\t"
            "mad.lo.u32 %4,%18,%24,%4; mad.hi.u32 %5,%18,%24,%5; mad.lo.u32 %6,%19,%24,%6; mad.hi.u32 %7,%19,%24,%7;\n\t"
            "mad.lo.u32 %8,%20,%24,%8; mad.hi.u32 %9,%20,%24,%9; mad.lo.u32 %10,%21,%24,%10; mad.hi.u32 %11,%21,%24,%11;\n\t"
            "mad.lo.u32 %12,%22,%24,%12; mad.hi.u32 %13,%22,%24,%13; mad.lo.u32 %14,%23,%24,%14; mad.hi.u32 %15,%23,%24,%15;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==6){ // 16 IADD3 independent
          asm volatile(
            "add.u32 %0,%0,%16; add.u32 %1,%1,%17; add.u32 %2,%2,%18; add.u32 %3,%3,%19; add.u32 %4,%4,%20; add.u32 %5,%5,%21; add.u32 %6,%6,%22; add.u32 %7,%7,%23;\n\t"
            "add.u32 %8,%8,%16; add.u32 %9,%9,%17; add.u32 %10,%10,%18; add.u32 %11,%11,%19; add.u32 %12,%12,%20; add.u32 %13,%13,%21; add.u32 %14,%14,%22; add.u32 %15,%15,%23;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==7){ // 2 carry chains of 8 IADD3.X
          asm volatile(
            "add.cc.u32 %0,%0,%16; addc.cc.u32 %1,%1,%17; addc.cc.u32 %2,%2,%18; addc.cc.u32 %3,%3,%19; addc.cc.u32 %4,%4,%20; addc.cc.u32 %5,%5,%21; addc.cc.u32 %6,%6,%22; addc.u32 %7,%7,%23;\n\t"
            "add.cc.u32 %8,%8,%16; addc.cc.u32 %9,%9,%17; addc.cc.u32 %10,%10,%18; addc.cc.u32 %11,%11,%19; addc.cc.u32 %12,%12,%20; addc.cc.u32 %13,%13,%21; addc.cc.u32 %14,%14,%22; addc.u32 %15,%15,%23;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==8){ // 16 funnel shifts
          asm volatile(
            "shf.l.wrap.b32 %0,%0,%16,3; shf.l.wrap.b32 %1,%1,%17,3; shf.l.wrap.b32 %2,%2,%18,3; shf.l.wrap.b32 %3,%3,%19,3; shf.l.wrap.b32 %4,%4,%20,3; shf.l.wrap.b32 %5,%5,%21,3; shf.l.wrap.b32 %6,%6,%22,3; shf.l.wrap.b32 %7,%7,%23,3;\n\t"
            "shf.l.wrap.b32 %8,%8,%16,3; shf.l.wrap.b32 %9,%9,%17,3; shf.l.wrap.b32 %10,%10,%18,3; shf.l.wrap.b32 %11,%11,%19,3; shf.l.wrap.b32 %12,%12,%20,3; shf.l.wrap.b32 %13,%13,%21,3; shf.l.wrap.b32 %14,%14,%22,3; shf.l.wrap.b32 %15,%15,%23,3;"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15])
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        } else if(MODE==9){ // mix: 8 plain IMAD.WIDE + 8 independent IADD3
          asm volatile(
            "mad.lo.cc.u32 %0,%24,%32,%0; madc.hi.u32 %1,%24,%32,%1; add.u32 %16,%16,%24;\n\t"
            "mad.lo.cc.u32 %2,%25,%32,%2; madc.hi.u32 %3,%25,%32,%3; add.u32 %17,%17,%25;\n\t"
            "mad.lo.cc.u32 %4,%26,%32,%4; madc.hi.u32 %5,%26,%32,%5; add.u32 %18,%18,%26;\n\t"
            "mad.lo.cc.u32 %6,%27,%32,%6; madc.hi.u32 %7,%27,%32,%7; add.u32 %19,%19,%27;\n\t"
            "mad.lo.cc.u32 %8,%28,%32,%8; madc.hi.u32 %9,%28,%32,%9; add.u32 %20,%20,%28;\n\t"
            "mad.lo.cc.u32 %10,%29,%32,%10; madc.hi.u32 %11,%29,%32,%11; add.u32 %21,%21,%29;\n\t"
            "mad.lo.cc.u32 %12,%30,%32,%12; madc.hi.u32 %13,%30,%32,%13; add.u32 %22,%22,%30;\n\t"
            "mad.lo.cc.u32 %14,%31,%32,%14; madc.hi.u32 %15,%31,%32,%15; add.u32 %23,%23,%31;\n\t"
            :"+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]),"+r"(a[8]),"+r"(a[9]),"+r"(a[10]),"+r"(a[11]),"+r"(a[12]),"+r"(a[13]),"+r"(a[14]),"+r"(a[15]),"+r"(s0),"+r"(s1),"+r"(s2),"+r"(s3),"+r"(s4),"+r"(s5),"+r"(s6),"+r"(s7)
            :"r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(x[6]),"r"(x[7]),"r"(y));
        }
      }
    }
    unsigned long long t1=clock64();
    u32 r=s0^s1^s2^s3^s4^s5^s6^s7; for(int i=0;i<16;i++) r^=a[i];
    if(r==0x12345u) sink[0]=r;
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g1));
    if(threadIdx.x==0&&blockIdx.x==0){ cyc[0]=t1-t0; cyc[1]=g1-g0; }
}
template<int MODE> int run(const char* name, int instr_per_block, int bps, u32* sink, u32* src, unsigned long long* cyc)
{
    int iters=6000; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<MODE>);
    int maxb=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb,k<MODE>,256,0); if(bps>maxb)bps=maxb;
    int grid=148*bps;
    k<MODE><<<grid,256>>>(sink,src,2000,cyc); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid,256>>>(sink,src,iters,cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); unsigned long long c[2]; cudaMemcpy(c,cyc,16,cudaMemcpyDeviceToHost);
    double fghz=(double)c[0]/(double)c[1];                 // SM clock in GHz from clock64 vs globaltimer in CTA 0
    double wps=bps*8/4.0; double per=(double)ms*1e6*fghz/((double)iters*4*instr_per_block*wps);
    printf("%-44s regs=%3d warps/SMSP=%4.1f  cycles/instr/SMSP=%6.2f  (SM clk %.0f MHz, wall %.2f ms)\n",name,fa.numRegs,wps,per,fghz*1e3,ms);
    return 0;
}
int main(){
    u32*sink,*src; unsigned long long*cyc; CHK(cudaMalloc(&sink,64)); CHK(cudaMalloc(&cyc,64)); CHK(cudaMalloc(&src,4096));
    u32 h[1024]; for(int i=0;i<1024;i++) h[i]=0x9e3779b9u*(i+1)^(0x85ebca6bu*(i*i+7)); cudaMemcpy(src,h,4096,cudaMemcpyHostToDevice);
    for(int b: {2,4,8}){
        run<0>("IMAD.WIDE plain (8)",8,b,sink,src,cyc);
        run<1>("IMAD.WIDE carry-out + IADD3.X (8+8)",16,b,sink,src,cyc);
        run<2>("IMAD.WIDE chains of 2 (out | in)",8,b,sink,src,cyc);
        run<3>("IMAD.WIDE chains of 4",8,b,sink,src,cyc);
        run<4>("IMAD.WIDE chain of 8",8,b,sink,src,cyc);
        run<5>("IMAD lo + IMAD.HI (16)",16,b,sink,src,cyc);
        run<6>("IADD3 (16)",16,b,sink,src,cyc);
        run<7>("IADD3.X chains of 8 (16)",16,b,sink,src,cyc);
        run<8>("SHF funnel (16)",16,b,sink,src,cyc);
        run<9>("8 IMAD.WIDE plain + 8 IADD3 (16)",16,b,sink,src,cyc);
    }
    return 0;
}
