// tools/ubench6.cu -- round-2 instruction-rate probe for the integer-multiply roofline (replaces ubench4).
//
// Every loop is LOOP-VARIANT in all eight chains (the round-1 probes let ptxas hoist seven of eight products):
// each chain feeds its own result back into its own operands, so nothing can be hoisted or CSE'd.  The SASS of
// every loop body is counted by tools/sass_mix.py (run here, no GPU needed) and the counts are printed next to the
// measured numbers in profiles/r2_ubench6.txt.  Output unit: cycles per loop trip per SM sub-partition (SMSP) at
// saturation = elapsed SM cycles / (trips x resident warps per SMSP); divide by the instruction count per trip for
// the issue interval of one warp-instruction.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I curve25519_b200/csrc -o tools/ubench6 tools/ubench6.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fe25519.cuh"
#include "x25519.cuh"
using namespace c25519;
#define CHK(x) do{cudaError_t e=(x); if(e){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
typedef unsigned long long ull;

#define REP8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

template<int MODE>
__global__ void __launch_bounds__(256) k(u32* sink, const u32* src, int iters, ull* cyc)
{
    const int t = blockIdx.x * 256 + threadIdx.x;
    u32 x[8], y[8], z[8]; ull a[8]; double d[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = src[(t + i) & 1023] | 1; y[i] = src[(t + 8 + i) & 1023] | 1; z[i] = src[(t + 16 + i) & 1023];
        a[i] = ((ull)src[(t + 24 + i) & 1023] << 32) | x[i];
        d[i] = 1.0 + x[i] * 1e-10; e[i] = 1.0 - y[i] * 1e-11;
    }
    fe fx, fy, fz, fw;
#pragma unroll
    for (int i = 0; i < 8; i++) { fx.v[i] = x[i]; fy.v[i] = y[i]; fz.v[i] = z[i]; fw.v[i] = x[i] ^ z[i]; }
    ull g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    ull t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // fresh product, both halves fed back in place:  IMAD.WIDE.U32 Rd, Ra, Rb, RZ
#define B(i) { ull p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(y[i])); x[i] = (u32)p; y[i] = (u32)(p >> 32); }
            REP8(B)
#undef B
        } else if (MODE == 1) {     // fresh product consumed by one LOP3 (ALU pipe) per product
#define B(i) { ull p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(y[i])); x[i] ^= (u32)p ^ (u32)(p >> 32); }
            REP8(B)
#undef B
        } else if (MODE == 2) {     // PTX mad.wide.u32 with the multiplicand fed back: ptxas SPLITS it into a fresh
                                    // IMAD.WIDE (RZ) + IADD3 (carry out) + IMAD.X -- the "split" form
#define B(i) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((u32)a[i]), "r"(y[i])); }
            REP8(B)
#undef B
        } else if (MODE == 4) {     // 32-bit IMAD (low half only)
#define B(i) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i])); }
            REP8(B)
#undef B
        } else if (MODE == 5) {     // IMAD.HI
#define B(i) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); }
            REP8(B)
#undef B
        } else if (MODE == 6) {     // the real carry-chained row of fe_mul: 16 IMAD.WIDE(.X) + 1 IADD3.X
            u32 lo[8], hi[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { lo[i] = x[i]; hi[i] = y[i]; }
            mad_row(lo, hi, z[0], z[1], z[2], z[3], z[4], z[5], z[6], z[7], x[0] ^ y[7]);
#pragma unroll
            for (int i = 0; i < 8; i++) { x[i] = lo[i]; y[i] = hi[i]; }
        } else if (MODE == 7) {     // 16 independent IADD3
#define B(i) { x[i] = x[i] + y[i] + z[i]; y[i] = y[i] + z[i] + x[i]; }
            REP8(B)
#undef B
        } else if (MODE == 20) {    // 16 additions ptxas places on the FMA pipe (IMAD.IADD)
#define B(i) { asm volatile("add.u32 %0, %0, %2; add.u32 %1, %1, %2;" : "+r"(x[i]), "+r"(y[i]) : "r"(z[i])); }
            REP8(B)
#undef B
        } else if (MODE == 8) {     // two 8-long carry chains (IADD3 Pout / IADD3.X)
            asm volatile("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11; addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,%15;"
                : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]));
            asm volatile("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11; addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,%15;"
                : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7])
                : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]));
        } else if (MODE == 9) {     // 16 LOP3
#define B(i) { asm volatile("lop3.b32 %0, %0, %2, %3, 0x96; lop3.b32 %1, %1, %2, %3, 0x96;" : "+r"(x[i]), "+r"(y[i]) : "r"(z[i]), "r"(z[(i + 1) & 7])); }
            REP8(B)
#undef B
        } else if (MODE == 10) {    // 8 DFMA chains
#define B(i) { asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) & 7])); }
            REP8(B)
#undef B
        } else if (MODE == 11) {    // 8 DADD chains
#define B(i) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i])); }
            REP8(B)
#undef B
        } else if (MODE == 12) {    // 8 fresh IMAD.WIDE + 16 IADD3
#define B(i) { ull p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(y[i])); x[i] ^= (u32)p ^ (u32)(p >> 32); fz.v[i] = fz.v[i] + fw.v[i] + z[i]; fw.v[i] = fw.v[i] + z[i] + fz.v[i]; }
            REP8(B)
#undef B
        } else if (MODE == 14) {    // 8 DFMA + 16 IADD3
#define B(i) { asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) & 7])); fz.v[i] = fz.v[i] + fw.v[i] + z[i]; fw.v[i] = fw.v[i] + z[i] + fz.v[i]; }
            REP8(B)
#undef B
        } else if (MODE == 16) {    // the shipped field multiplication, dependent chain
            fe_mul(fx, fx, fy);
        } else if (MODE == 17) {    // the shipped field squaring
            fe_sqr(fx, fx);
        } else if (MODE == 18) {    // one ladder step (5M + 4S + 1W + adds)
            mont_step(fx, fy, fz, fw, fx);
        }
    }
    ull t1 = clock64(); ull g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    u32 r = 0; double ds = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { r ^= x[i] ^ y[i] ^ z[i] ^ (u32)a[i] ^ (u32)(a[i] >> 32) ^ fx.v[i] ^ fy.v[i] ^ fz.v[i] ^ fw.v[i]; ds += d[i] + e[i]; }
    if (r == 0x12345u && ds == 77.0) sink[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = g1 - g0; }
}

// Accumulate form with scalar 64-bit accumulators (array-of-u64 state makes ptxas shuffle register pairs with IMAD.MOV):
//   EXTRA 0: 8 x IMAD.WIDE.U32 Rd, Ra, Rb, Rd       EXTRA 1: + 16 IADD3        EXTRA 2: + 8 DFMA
template<int EXTRA>
__global__ void __launch_bounds__(256) kacc(u32* sink, const u32* src, int iters, ull* cyc)
{
    const int t = blockIdx.x * 256 + threadIdx.x;
#define L(j) src[(t + (j)) & 1023]
    ull a0 = L(0), a1 = L(1), a2 = L(2), a3 = L(3), a4 = L(4), a5 = L(5), a6 = L(6), a7 = L(7);
    u32 x0 = L(8) | 1, x1 = L(9) | 1, x2 = L(10) | 1, x3 = L(11) | 1, x4 = L(12) | 1, x5 = L(13) | 1, x6 = L(14) | 1, x7 = L(15) | 1;
    u32 y0 = L(16) | 1, y1 = L(17) | 1, y2 = L(18) | 1, y3 = L(19) | 1, y4 = L(20) | 1, y5 = L(21) | 1, y6 = L(22) | 1, y7 = L(23) | 1;
    u32 p[8], q[8], r[8]; double d[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { p[i] = L(24 + i); q[i] = L(32 + i); r[i] = L(40 + i); d[i] = 1.0 + p[i] * 1e-10; e[i] = 1.0 - q[i] * 1e-11; }
#undef L
    ull g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    ull t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        ull p0, p1, p2, p3, p4, p5, p6, p7;
        asm volatile("mul.wide.u32 %0,%8,%16; mul.wide.u32 %1,%9,%17; mul.wide.u32 %2,%10,%18; mul.wide.u32 %3,%11,%19; "
                     "mul.wide.u32 %4,%12,%20; mul.wide.u32 %5,%13,%21; mul.wide.u32 %6,%14,%22; mul.wide.u32 %7,%15,%23;"
                     : "=l"(p0), "=l"(p1), "=l"(p2), "=l"(p3), "=l"(p4), "=l"(p5), "=l"(p6), "=l"(p7)
                     : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(x6), "r"(x7),
                       "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(y4), "r"(y5), "r"(y6), "r"(y7));
        a0 += p0; a1 += p1; a2 += p2; a3 += p3; a4 += p4; a5 += p5; a6 += p6; a7 += p7;
        x0 += 1;
        if (EXTRA == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) { p[i] = p[i] + q[i] + r[i]; q[i] = q[i] + r[i] + p[i]; }
        }
        if (EXTRA == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) & 7]));
        }
    }
    ull t1 = clock64(); ull g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    ull rr = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7; u32 w = x0; double ds = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { w ^= p[i] ^ q[i] ^ r[i]; ds += d[i]; }
    if ((u32)rr == 0x12345u && w == 77u && ds == 3.0) sink[0] = (u32)(rr >> 32);
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = g1 - g0; }
}

template<typename K> int run_k(K kern, const char* tag, const char* name, int per_trip, int bps, int iters, u32* sink, u32* src, ull* cyc)
{
    int maxb = 0; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kern, 256, 0); if (bps > maxb) bps = maxb; int grid = 148 * bps;
    kern<<<grid, 256>>>(sink, src, iters / 8 + 1, cyc); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); kern<<<grid, 256>>>(sink, src, iters, cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ull c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
    double f = (double)c[0] / (double)c[1]; double wps = bps * 2.0;
    double trip = (double)ms * 1e6 * f / ((double)iters * wps);
    printf("%s %-58s regs=%3d w/SMSP=%4.1f cyc/trip/SMSP=%8.2f  per-instr(%3d)=%6.3f  clk=%.0f MHz\n",
           tag, name, fa.numRegs, wps, trip, per_trip, trip / per_trip, f * 1e3);
    return 0;
}

template<int MODE> int run(const char* name, int per_trip, int bps, int iters, u32* sink, u32* src, ull* cyc)
{
    int maxb = 0; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<MODE>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k<MODE>, 256, 0); if (bps > maxb) bps = maxb; int grid = 148 * bps;
    k<MODE><<<grid, 256>>>(sink, src, iters / 8 + 1, cyc); CHK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(sink, src, iters, cyc); cudaEventRecord(e1); CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ull c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
    double f = (double)c[0] / (double)c[1];          // SM cycles per ns while the kernel ran
    double wps = bps * 2.0;                           // 256 threads = 8 warps per CTA over 4 SMSPs
    double trip = (double)ms * 1e6 * f / ((double)iters * wps);
    printf("m%02d %-58s regs=%3d w/SMSP=%4.1f cyc/trip/SMSP=%8.2f  per-instr(%3d)=%6.3f  clk=%.0f MHz\n",
           MODE, name, fa.numRegs, wps, trip, per_trip, trip / per_trip, f * 1e3);
    return 0;
}

int main()
{
    u32 *sink, *src; ull* cyc; CHK(cudaMalloc(&sink, 64)); CHK(cudaMalloc(&cyc, 64)); CHK(cudaMalloc(&src, 4096));
    u32 h[1024]; for (int i = 0; i < 1024; i++) h[i] = 0x9e3779b9u * (i + 1) ^ (0x85ebca6bu * (i * i + 7)); cudaMemcpy(src, h, 4096, cudaMemcpyHostToDevice);
    for (int b : {2, 4}) {
        const int N = 20000;
        run<0>("8 IMAD.WIDE fresh (in-place feedback)", 8, b, N, sink, src, cyc);
        run<1>("8 IMAD.WIDE fresh + 8 LOP3", 8, b, N, sink, src, cyc);
        run<2>("split: 8 x (IMAD.WIDE fresh + IADD3 + IMAD.X)", 8, b, N, sink, src, cyc);
        run_k(kacc<0>, "acc", "8 IMAD.WIDE accumulate (Rd,Ra,Rb,Rd)", 8, b, N, sink, src, cyc);
        run<4>("8 IMAD (32-bit)", 8, b, N, sink, src, cyc);
        run<5>("8 IMAD.HI", 8, b, N, sink, src, cyc);
        run<6>("fe_mul row: 8 IMAD.WIDE(.X) accumulate, carry-chained", 8, b, N, sink, src, cyc);
        run<7>("16 IADD3", 16, b, N, sink, src, cyc);
        run<20>("16 IMAD.IADD (adds on the FMA pipe)", 16, b, N, sink, src, cyc);
        run<8>("2 x 8-long IADD3/IADD3.X carry chains", 16, b, N, sink, src, cyc);
        run<9>("16 LOP3", 16, b, N, sink, src, cyc);
        run<10>("8 DFMA", 8, b, N, sink, src, cyc);
        run<11>("8 DADD", 8, b, N, sink, src, cyc);
        run<12>("8 IMAD.WIDE fresh + 8 LOP3 + 16 IADD3 (per-IMAD)", 8, b, N, sink, src, cyc);
        run_k(kacc<1>, "acc", "8 IMAD.WIDE accumulate + 16 IADD3 (per-IMAD)", 8, b, N, sink, src, cyc);
        run<14>("8 DFMA + 16 IADD3 (per-DFMA)", 8, b, N, sink, src, cyc);
        run_k(kacc<2>, "acc", "8 IMAD.WIDE accumulate + 8 DFMA (per-pair)", 8, b, N, sink, src, cyc);
        run<16>("fe_mul (shipped; 73 IMAD)", 73, b, N / 4, sink, src, cyc);
        run<17>("fe_sqr (shipped; 45 IMAD)", 45, b, N / 4, sink, src, cyc);
        run<18>("mont_step (shipped; 5M+4S+1W = 554 IMAD)", 554, b, N / 16, sink, src, cyc);
    }
    return 0;
}
