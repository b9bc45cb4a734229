// SUPERSEDED by tools/ubench6.cu (round 2): several loops here are not loop-variant in all chains, so ptxas hoists most of
// their products; see the note at the top of profiles/r1_ubench4.txt.  Kept for the record only.
// tools/ubench4.cu -- decisive per-instruction costs (cycles per warp-instruction per SMSP at saturation).
// Bodies are NOT unrolled (one asm block per loop trip) so ptxas cannot regroup chains; SASS checked by hand.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
typedef uint32_t u32; typedef unsigned long long u64;

template<int MODE>
__global__ void __launch_bounds__(256) k(u32* sink, const u32* src, int iters, u64* cyc)
{
    const int t=blockIdx.x*256+threadIdx.x;
    u64 a0=src[(t+0)&1023],a1=src[(t+1)&1023],a2=src[(t+2)&1023],a3=src[(t+3)&1023],a4=src[(t+4)&1023],a5=src[(t+5)&1023],a6=src[(t+6)&1023],a7=src[(t+7)&1023];
    u32 x0=src[(t+8)&1023]|1,x1=src[(t+9)&1023]|1,x2=src[(t+10)&1023]|1,x3=src[(t+11)&1023]|1,x4=src[(t+12)&1023]|1,x5=src[(t+13)&1023]|1,x6=src[(t+14)&1023]|1,x7=src[(t+15)&1023]|1;
    u32 y0=src[(t+16)&1023]|1,y1=src[(t+17)&1023]|1,y2=src[(t+18)&1023]|1,y3=src[(t+19)&1023]|1,y4=src[(t+20)&1023]|1,y5=src[(t+21)&1023]|1,y6=src[(t+22)&1023]|1,y7=src[(t+23)&1023]|1;
    u32 r0=x0,r1=x1,r2=x2,r3=x3,r4=x4,r5=x5,r6=x6,r7=x7,r8=y0,r9=y1,r10=y2,r11=y3,r12=y4,r13=y5,r14=y6,r15=y7;
    u64 g0; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g0));
    u64 t0=clock64();
    #pragma unroll 1
    for(int it=0;it<iters;it++){
        if(MODE==0) asm volatile("mad.wide.u32 %0,%8,%9,%0; mad.wide.u32 %1,%8,%9,%1; mad.wide.u32 %2,%8,%9,%2; mad.wide.u32 %3,%8,%9,%3; mad.wide.u32 %4,%8,%9,%4; mad.wide.u32 %5,%8,%9,%5; mad.wide.u32 %6,%8,%9,%6; mad.wide.u32 %7,%8,%9,%7;"
            :"+l"(a0),"+l"(a1),"+l"(a2),"+l"(a3),"+l"(a4),"+l"(a5),"+l"(a6),"+l"(a7):"r"(x0),"r"(y0));
        else if(MODE==1) asm volatile("mad.wide.u32 %0,%8,%16,%0; mad.wide.u32 %1,%9,%16,%1; mad.wide.u32 %2,%10,%16,%2; mad.wide.u32 %3,%11,%16,%3; mad.wide.u32 %4,%12,%16,%4; mad.wide.u32 %5,%13,%16,%5; mad.wide.u32 %6,%14,%16,%6; mad.wide.u32 %7,%15,%16,%7;"
            :"+l"(a0),"+l"(a1),"+l"(a2),"+l"(a3),"+l"(a4),"+l"(a5),"+l"(a6),"+l"(a7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0));
        else if(MODE==2) asm volatile("mad.wide.u32 %0,%8,%16,%0; mad.wide.u32 %1,%9,%17,%1; mad.wide.u32 %2,%10,%18,%2; mad.wide.u32 %3,%11,%19,%3; mad.wide.u32 %4,%12,%20,%4; mad.wide.u32 %5,%13,%21,%5; mad.wide.u32 %6,%14,%22,%6; mad.wide.u32 %7,%15,%23,%7;"
            :"+l"(a0),"+l"(a1),"+l"(a2),"+l"(a3),"+l"(a4),"+l"(a5),"+l"(a6),"+l"(a7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0),"r"(y1),"r"(y2),"r"(y3),"r"(y4),"r"(y5),"r"(y6),"r"(y7));
        else if(MODE==3){ // fresh products (Rc = RZ), consumed by 8 LOP3 (xor-3) on the ALU pipe
            u64 p0,p1,p2,p3,p4,p5,p6,p7;
            asm volatile("mul.wide.u32 %0,%8,%16; mul.wide.u32 %1,%9,%17; mul.wide.u32 %2,%10,%18; mul.wide.u32 %3,%11,%19; mul.wide.u32 %4,%12,%20; mul.wide.u32 %5,%13,%21; mul.wide.u32 %6,%14,%22; mul.wide.u32 %7,%15,%23;"
                :"=l"(p0),"=l"(p1),"=l"(p2),"=l"(p3),"=l"(p4),"=l"(p5),"=l"(p6),"=l"(p7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0),"r"(y1),"r"(y2),"r"(y3),"r"(y4),"r"(y5),"r"(y6),"r"(y7));
            r0^=(u32)p0^(u32)(p0>>32); r1^=(u32)p1^(u32)(p1>>32); r2^=(u32)p2^(u32)(p2>>32); r3^=(u32)p3^(u32)(p3>>32);
            r4^=(u32)p4^(u32)(p4>>32); r5^=(u32)p5^(u32)(p5>>32); r6^=(u32)p6^(u32)(p6>>32); r7^=(u32)p7^(u32)(p7>>32);
            x0+=1; // keep the products loop-variant (1 extra IADD)
        }
        else if(MODE==4){ // fresh products accumulated on the ALU pipe with 64-bit adds (IADD3 + IADD3.X per product)
            u64 p0,p1,p2,p3,p4,p5,p6,p7;
            asm volatile("mul.wide.u32 %0,%8,%16; mul.wide.u32 %1,%9,%17; mul.wide.u32 %2,%10,%18; mul.wide.u32 %3,%11,%19; mul.wide.u32 %4,%12,%20; mul.wide.u32 %5,%13,%21; mul.wide.u32 %6,%14,%22; mul.wide.u32 %7,%15,%23;"
                :"=l"(p0),"=l"(p1),"=l"(p2),"=l"(p3),"=l"(p4),"=l"(p5),"=l"(p6),"=l"(p7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0),"r"(y1),"r"(y2),"r"(y3),"r"(y4),"r"(y5),"r"(y6),"r"(y7));
            a0+=p0; a1+=p1; a2+=p2; a3+=p3; a4+=p4; a5+=p5; a6+=p6; a7+=p7; x0+=1;
        }
        else if(MODE==5){ // as 4 but two products per 3-input 64-bit add (IADD3 dual carry + IADD3.X)
            u64 p0,p1,p2,p3,p4,p5,p6,p7;
            asm volatile("mul.wide.u32 %0,%8,%16; mul.wide.u32 %1,%9,%17; mul.wide.u32 %2,%10,%18; mul.wide.u32 %3,%11,%19; mul.wide.u32 %4,%12,%20; mul.wide.u32 %5,%13,%21; mul.wide.u32 %6,%14,%22; mul.wide.u32 %7,%15,%23;"
                :"=l"(p0),"=l"(p1),"=l"(p2),"=l"(p3),"=l"(p4),"=l"(p5),"=l"(p6),"=l"(p7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0),"r"(y1),"r"(y2),"r"(y3),"r"(y4),"r"(y5),"r"(y6),"r"(y7));
            a0+=p0+p1; a1+=p2+p3; a2+=p4+p5; a3+=p6+p7; x0+=1;
        }
        else if(MODE==6) asm volatile("add.u32 %0,%0,%16; add.u32 %1,%1,%17; add.u32 %2,%2,%18; add.u32 %3,%3,%19; add.u32 %4,%4,%20; add.u32 %5,%5,%21; add.u32 %6,%6,%22; add.u32 %7,%7,%23;\n\t"
                                      "add.u32 %8,%8,%16; add.u32 %9,%9,%17; add.u32 %10,%10,%18; add.u32 %11,%11,%19; add.u32 %12,%12,%20; add.u32 %13,%13,%21; add.u32 %14,%14,%22; add.u32 %15,%15,%23;"
            :"+r"(r0),"+r"(r1),"+r"(r2),"+r"(r3),"+r"(r4),"+r"(r5),"+r"(r6),"+r"(r7),"+r"(r8),"+r"(r9),"+r"(r10),"+r"(r11),"+r"(r12),"+r"(r13),"+r"(r14),"+r"(r15)
            :"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7));
        else if(MODE==7) asm volatile( // 8 independent 64-bit adds: IADD3 (carry-out) + IADD3.X (carry-in)
            "add.cc.u32 %0,%0,%16; addc.u32 %1,%1,%17; add.cc.u32 %2,%2,%18; addc.u32 %3,%3,%19; add.cc.u32 %4,%4,%20; addc.u32 %5,%5,%21; add.cc.u32 %6,%6,%22; addc.u32 %7,%7,%23;\n\t"
            "add.cc.u32 %8,%8,%16; addc.u32 %9,%9,%17; add.cc.u32 %10,%10,%18; addc.u32 %11,%11,%19; add.cc.u32 %12,%12,%20; addc.u32 %13,%13,%21; add.cc.u32 %14,%14,%22; addc.u32 %15,%15,%23;"
            :"+r"(r0),"+r"(r1),"+r"(r2),"+r"(r3),"+r"(r4),"+r"(r5),"+r"(r6),"+r"(r7),"+r"(r8),"+r"(r9),"+r"(r10),"+r"(r11),"+r"(r12),"+r"(r13),"+r"(r14),"+r"(r15)
            :"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7));
        else if(MODE==8) asm volatile( // 4 independent 4-word chains: out, in+out, in+out, in
            "add.cc.u32 %0,%0,%16; addc.cc.u32 %1,%1,%17; addc.cc.u32 %2,%2,%18; addc.u32 %3,%3,%19; add.cc.u32 %4,%4,%20; addc.cc.u32 %5,%5,%21; addc.cc.u32 %6,%6,%22; addc.u32 %7,%7,%23;\n\t"
            "add.cc.u32 %8,%8,%16; addc.cc.u32 %9,%9,%17; addc.cc.u32 %10,%10,%18; addc.u32 %11,%11,%19; add.cc.u32 %12,%12,%20; addc.cc.u32 %13,%13,%21; addc.cc.u32 %14,%14,%22; addc.u32 %15,%15,%23;"
            :"+r"(r0),"+r"(r1),"+r"(r2),"+r"(r3),"+r"(r4),"+r"(r5),"+r"(r6),"+r"(r7),"+r"(r8),"+r"(r9),"+r"(r10),"+r"(r11),"+r"(r12),"+r"(r13),"+r"(r14),"+r"(r15)
            :"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7));
        else if(MODE==9){ // A2-style IMAD (Rb shared) + 8 independent plain IADD3
            asm volatile("mad.wide.u32 %0,%8,%16,%0; mad.wide.u32 %1,%9,%16,%1; mad.wide.u32 %2,%10,%16,%2; mad.wide.u32 %3,%11,%16,%3; mad.wide.u32 %4,%12,%16,%4; mad.wide.u32 %5,%13,%16,%5; mad.wide.u32 %6,%14,%16,%6; mad.wide.u32 %7,%15,%16,%7;"
            :"+l"(a0),"+l"(a1),"+l"(a2),"+l"(a3),"+l"(a4),"+l"(a5),"+l"(a6),"+l"(a7):"r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(y0));
            asm volatile("add.u32 %0,%0,%8; add.u32 %1,%1,%9; add.u32 %2,%2,%10; add.u32 %3,%3,%11; add.u32 %4,%4,%12; add.u32 %5,%5,%13; add.u32 %6,%6,%14; add.u32 %7,%7,%15;"
            :"+r"(r0),"+r"(r1),"+r"(r2),"+r"(r3),"+r"(r4),"+r"(r5),"+r"(r6),"+r"(r7):"r"(y0),"r"(y1),"r"(y2),"r"(y3),"r"(y4),"r"(y5),"r"(y6),"r"(y7));
        }
    }
    u64 t1=clock64(); u64 g1; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g1));
    u64 r=a0^a1^a2^a3^a4^a5^a6^a7; u32 q=r0^r1^r2^r3^r4^r5^r6^r7^r8^r9^r10^r11^r12^r13^r14^r15^x0;
    if((u32)r==0x12345u && q==77) sink[0]=(u32)(r>>32);
    if(threadIdx.x==0&&blockIdx.x==0){cyc[0]=t1-t0;cyc[1]=g1-g0;}
}
template<int MODE> int run(const char* name,int instr,int bps,u32*sink,u32*src,u64*cyc)
{
    int iters=24000; int maxb=0; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<MODE>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb,k<MODE>,256,0); if(bps>maxb)bps=maxb; int grid=148*bps;
    k<MODE><<<grid,256>>>(sink,src,8000,cyc); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid,256>>>(sink,src,iters,cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); u64 c[2]; cudaMemcpy(c,cyc,16,cudaMemcpyDeviceToHost);
    double f=(double)c[0]/(double)c[1]; double wps=bps*2.0; double per_trip=(double)ms*1e6*f/((double)iters*wps);
    printf("%-52s regs=%3d w/SMSP=%4.1f cycles/trip/SMSP=%7.2f = %5.2f per counted instr (%d)  clk=%.0f MHz\n",name,fa.numRegs,wps,per_trip,per_trip/instr,instr,f*1e3);
    return 0;
}
int main(){
    u32*sink,*src; u64*cyc; CHK(cudaMalloc(&sink,64)); CHK(cudaMalloc(&cyc,64)); CHK(cudaMalloc(&src,4096));
    u32 h[1024]; for(int i=0;i<1024;i++) h[i]=0x9e3779b9u*(i+1)^(0x85ebca6bu*(i*i+7)); cudaMemcpy(src,h,4096,cudaMemcpyHostToDevice);
    for(int b: {4,8}){
        run<0>("A1 8 IMAD.WIDE acc, same x,y",8,b,sink,src,cyc);
        run<1>("A2 8 IMAD.WIDE acc, distinct x, shared y",8,b,sink,src,cyc);
        run<2>("A3 8 IMAD.WIDE acc, distinct x, distinct y",8,b,sink,src,cyc);
        run<3>("A4 8 IMAD.WIDE fresh (Rc=RZ) + 8 LOP3",8,b,sink,src,cyc);
        run<4>("A5 8 IMAD.WIDE fresh + 8x(IADD3+IADD3.X)",8,b,sink,src,cyc);
        run<5>("A6 8 IMAD.WIDE fresh + 4x(IADD3 dual + IADD3.X)",8,b,sink,src,cyc);
        run<6>("B1 16 IADD3 plain",16,b,sink,src,cyc);
        run<7>("B2 8x(IADD3 cout + IADD3.X cin)",16,b,sink,src,cyc);
        run<8>("B3 4x 4-word carry chains",16,b,sink,src,cyc);
        run<9>("C1 8 IMAD.WIDE (A2) + 8 IADD3",16,b,sink,src,cyc);
    }
    return 0;
}
