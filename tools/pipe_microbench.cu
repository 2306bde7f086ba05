// Dev tool: issue-port / FP64-pipe microbenchmark for sm_100a (B200).  Answers the questions DESIGN.md's k_step roofline needs:
// how many issue cycles a DFMA costs, whether integer / FP32 instructions issue in its shadow, and the dependent-issue latency
// of DFMA / MUFU.RSQ64H / F2I.F64 / I2F.F64.  One CTA per SM, W warps, in-kernel clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void bench(double *out, long long *cyc, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float f0 = (float)seed, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    const double b = 1.0000001, c = 1e-9;
    unsigned long long p0 = 0x3f8000013f800001ull + threadIdx.x, p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3, p4 = p0 + 4, p5 = p0 + 5, p6 = p0 + 6, p7 = p0 + 7;
    const unsigned long long pk = 0x3f8000013f800001ull;
    double q0 = a0 * 1.5, q1 = a1 * 1.5, q2 = a2 * 1.5, q3 = a3 * 1.5, q4 = a4 * 1.5, q5 = a5 * 1.5, q6 = a6 * 1.5, q7 = a7 * 1.5;
    double r0 = a0 * 2.5, r1 = a1 * 2.5, r2 = a2 * 2.5, r3 = a3 * 2.5, r4 = a4 * 2.5, r5 = a5 * 2.5, r6 = a6 * 2.5, r7 = a7 * 2.5;
    asm volatile("" : "+d"(q0), "+d"(q1), "+d"(q2), "+d"(q3), "+d"(q4), "+d"(q5), "+d"(q6), "+d"(q7));
    asm volatile("" : "+d"(r0), "+d"(r1), "+d"(r2), "+d"(r3), "+d"(r4), "+d"(r5), "+d"(r6), "+d"(r7));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // 8 independent DFMA
            asm volatile("fma.rn.f64 %0,%0,%8,%9; fma.rn.f64 %1,%1,%8,%9; fma.rn.f64 %2,%2,%8,%9; fma.rn.f64 %3,%3,%8,%9;"
                         "fma.rn.f64 %4,%4,%8,%9; fma.rn.f64 %5,%5,%8,%9; fma.rn.f64 %6,%6,%8,%9; fma.rn.f64 %7,%7,%8,%9;"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(b), "d"(c));
        } else if (MODE == 1) {  // 8 DFMA + 8 FFMA interleaved
            asm volatile("fma.rn.f64 %0,%0,%12,%13; fma.rn.f32 %8,%8,%14,%14; fma.rn.f64 %1,%1,%12,%13; fma.rn.f32 %9,%9,%14,%14;"
                         "fma.rn.f64 %2,%2,%12,%13; fma.rn.f32 %10,%10,%14,%14; fma.rn.f64 %3,%3,%12,%13; fma.rn.f32 %11,%11,%14,%14;"
                         "fma.rn.f64 %4,%4,%12,%13; fma.rn.f32 %8,%8,%14,%14; fma.rn.f64 %5,%5,%12,%13; fma.rn.f32 %9,%9,%14,%14;"
                         "fma.rn.f64 %6,%6,%12,%13; fma.rn.f32 %10,%10,%14,%14; fma.rn.f64 %7,%7,%12,%13; fma.rn.f32 %11,%11,%14,%14;"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7), "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3)
                         : "d"(b), "d"(c), "f"(1.0001f));
        } else if (MODE == 2) {  // 8 DFMA + 16 integer (mad.lo)
            asm volatile("fma.rn.f64 %0,%0,%12,%13; mad.lo.s32 %8,%8,%14,%14; mad.lo.s32 %9,%9,%14,%14; fma.rn.f64 %1,%1,%12,%13; mad.lo.s32 %10,%10,%14,%14; mad.lo.s32 %11,%11,%14,%14;"
                         "fma.rn.f64 %2,%2,%12,%13; mad.lo.s32 %8,%8,%14,%14; mad.lo.s32 %9,%9,%14,%14; fma.rn.f64 %3,%3,%12,%13; mad.lo.s32 %10,%10,%14,%14; mad.lo.s32 %11,%11,%14,%14;"
                         "fma.rn.f64 %4,%4,%12,%13; mad.lo.s32 %8,%8,%14,%14; mad.lo.s32 %9,%9,%14,%14; fma.rn.f64 %5,%5,%12,%13; mad.lo.s32 %10,%10,%14,%14; mad.lo.s32 %11,%11,%14,%14;"
                         "fma.rn.f64 %6,%6,%12,%13; mad.lo.s32 %8,%8,%14,%14; mad.lo.s32 %9,%9,%14,%14; fma.rn.f64 %7,%7,%12,%13; mad.lo.s32 %10,%10,%14,%14; mad.lo.s32 %11,%11,%14,%14;"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                         : "d"(b), "d"(c), "r"(3));
        } else if (MODE == 3) {  // 8 dependent DFMA (latency chain)
            asm volatile("fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2;"
                         "fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2;"
                         : "+d"(a0) : "d"(b), "d"(c));
        } else if (MODE == 4) {  // 8 independent DADD / DMUL alternating
            asm volatile("add.rn.f64 %0,%0,%9; mul.rn.f64 %1,%1,%8; add.rn.f64 %2,%2,%9; mul.rn.f64 %3,%3,%8;"
                         "add.rn.f64 %4,%4,%9; mul.rn.f64 %5,%5,%8; add.rn.f64 %6,%6,%9; mul.rn.f64 %7,%7,%8;"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(b), "d"(c));
        } else if (MODE == 5) {  // dependent chain: rsqrt.approx.f64 -> fma (latency of MUFU.RSQ64H + DFMA), x4
            asm volatile("rsqrt.approx.ftz.f64 %0,%0; fma.rn.f64 %0,%0,%1,%2; rsqrt.approx.ftz.f64 %0,%0; fma.rn.f64 %0,%0,%1,%2;"
                         "rsqrt.approx.ftz.f64 %0,%0; fma.rn.f64 %0,%0,%1,%2; rsqrt.approx.ftz.f64 %0,%0; fma.rn.f64 %0,%0,%1,%2;"
                         : "+d"(a0) : "d"(b), "d"(1.5));
        } else if (MODE == 6) {  // dependent chain: cvt.rni.s32.f64 -> cvt.rn.f64.s32 (F2I.F64 + I2F.F64), x4
            int t;
            asm volatile("cvt.rni.s32.f64 %1,%0; cvt.rn.f64.s32 %0,%1; cvt.rni.s32.f64 %1,%0; cvt.rn.f64.s32 %0,%1;"
                         "cvt.rni.s32.f64 %1,%0; cvt.rn.f64.s32 %0,%1; cvt.rni.s32.f64 %1,%0; cvt.rn.f64.s32 %0,%1;"
                         : "+d"(a0), "=r"(t));
        } else if (MODE == 7) {  // 8 independent FFMA-only (reference for the issue port)
            asm volatile("fma.rn.f32 %0,%0,%4,%4; fma.rn.f32 %1,%1,%4,%4; fma.rn.f32 %2,%2,%4,%4; fma.rn.f32 %3,%3,%4,%4;"
                         "fma.rn.f32 %0,%0,%4,%4; fma.rn.f32 %1,%1,%4,%4; fma.rn.f32 %2,%2,%4,%4; fma.rn.f32 %3,%3,%4,%4;"
                         : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3) : "f"(1.0001f));
        } else if (MODE == 9) {  // 8 independent packed FFMA2 (fma.rn.f32x2: two FP32 FMAs per instruction, 64-bit register pairs)
            asm volatile("fma.rn.f32x2 %0,%0,%8,%8; fma.rn.f32x2 %1,%1,%8,%8; fma.rn.f32x2 %2,%2,%8,%8; fma.rn.f32x2 %3,%3,%8,%8;"
                         "fma.rn.f32x2 %4,%4,%8,%8; fma.rn.f32x2 %5,%5,%8,%8; fma.rn.f32x2 %6,%6,%8,%8; fma.rn.f32x2 %7,%7,%8,%8;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(pk));
        } else if (MODE == 10) {  // 8 FFMA2 + 8 IMAD interleaved
            asm volatile("fma.rn.f32x2 %0,%0,%12,%12; mad.lo.s32 %8,%8,%13,%13; fma.rn.f32x2 %1,%1,%12,%12; mad.lo.s32 %9,%9,%13,%13;"
                         "fma.rn.f32x2 %2,%2,%12,%12; mad.lo.s32 %10,%10,%13,%13; fma.rn.f32x2 %3,%3,%12,%12; mad.lo.s32 %11,%11,%13,%13;"
                         "fma.rn.f32x2 %4,%4,%12,%12; mad.lo.s32 %8,%8,%13,%13; fma.rn.f32x2 %5,%5,%12,%12; mad.lo.s32 %9,%9,%13,%13;"
                         "fma.rn.f32x2 %6,%6,%12,%12; mad.lo.s32 %10,%10,%13,%13; fma.rn.f32x2 %7,%7,%12,%12; mad.lo.s32 %11,%11,%13,%13;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                         : "l"(pk), "r"(3));
        } else if (MODE == 11) {  // 8 FFMA + 8 IMAD interleaved (reference for MODE 10)
            asm volatile("fma.rn.f32 %0,%0,%8,%8; mad.lo.s32 %4,%4,%9,%9; fma.rn.f32 %1,%1,%8,%8; mad.lo.s32 %5,%5,%9,%9;"
                         "fma.rn.f32 %2,%2,%8,%8; mad.lo.s32 %6,%6,%9,%9; fma.rn.f32 %3,%3,%8,%8; mad.lo.s32 %7,%7,%9,%9;"
                         "fma.rn.f32 %0,%0,%8,%8; mad.lo.s32 %4,%4,%9,%9; fma.rn.f32 %1,%1,%8,%8; mad.lo.s32 %5,%5,%9,%9;"
                         "fma.rn.f32 %2,%2,%8,%8; mad.lo.s32 %6,%6,%9,%9; fma.rn.f32 %3,%3,%8,%8; mad.lo.s32 %7,%7,%9,%9;"
                         : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "f"(1.0001f), "r"(3));
        } else if (MODE >= 12 && MODE <= 21) {
            // Round 2: what DOES issue in the shadow of a DFMA?  8 DFMA + 16 companions from the ALU pipe (LOP3, SEL, IADD3 / VIADD,
            // VIMNMX), or 8 shared-memory loads / F2I+I2F / MUFU; odd modes run the companions alone.
            constexpr bool WITH = (MODE % 2) == 0;
#define DF(k) if (WITH) asm volatile("fma.rn.f64 %0,%0,%1,%2;" : "+d"(a##k) : "d"(b), "d"(c));
            if (MODE == 12 || MODE == 13) {
#define CO(r) asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(r) : "r"(i3 + 5), "r"(7));
                DF(0) CO(i0) CO(i1) DF(1) CO(i2) CO(i0) DF(2) CO(i1) CO(i2) DF(3) CO(i0) CO(i1) DF(4) CO(i2) CO(i0) DF(5) CO(i1) CO(i2) DF(6) CO(i0) CO(i1) DF(7) CO(i2) CO(i0)
#undef CO
            } else if (MODE == 14 || MODE == 15) {
#define CO(r) asm volatile("{.reg .pred p; setp.gt.s32 p,%0,%1; selp.b32 %0,%1,%0,p;}" : "+r"(r) : "r"(i3));
                DF(0) CO(i0) DF(1) CO(i1) DF(2) CO(i2) DF(3) CO(i0) DF(4) CO(i1) DF(5) CO(i2) DF(6) CO(i0) DF(7) CO(i1)
#undef CO
            } else if (MODE == 16 || MODE == 17) {
#define CO(r) asm volatile("max.s32 %0,%0,%1; add.s32 %0,%0,%2;" : "+r"(r) : "r"(i3), "r"(3));
                DF(0) CO(i0) DF(1) CO(i1) DF(2) CO(i2) DF(3) CO(i0) DF(4) CO(i1) DF(5) CO(i2) DF(6) CO(i0) DF(7) CO(i1)
#undef CO
            } else if (MODE == 18 || MODE == 19) {
                extern __shared__ double sm[];
#define CO(r) asm volatile("ld.shared.f64 %0,[%1];" : "=d"(r) : "r"((unsigned)(threadIdx.x * 8 + (it & 7) * 256)));
                double l0, l1, l2, l3, l4, l5, l6, l7;
                DF(0) CO(l0) DF(1) CO(l1) DF(2) CO(l2) DF(3) CO(l3) DF(4) CO(l4) DF(5) CO(l5) DF(6) CO(l6) DF(7) CO(l7)
                f0 += (float)(l0 + l1 + l2 + l3 + l4 + l5 + l6 + l7 > 1e300);
#undef CO
            } else {
#define CO(r) asm volatile("{.reg .f32 t; mov.b32 t,%0; ex2.approx.ftz.f32 t,t; mov.b32 %0,t;}" : "+r"(r));
                DF(0) CO(i0) DF(1) CO(i1) DF(2) CO(i2) DF(3) CO(i0) DF(4) CO(i1) DF(5) CO(i2) DF(6) CO(i0) DF(7) CO(i1)
#undef CO
            }
#undef DF
        } else if (MODE == 22 || MODE == 23 || MODE == 24) {
            // 8 independent FP64 instructions whose operands are all DISTINCT registers (no operand reuse, no constant-bank operand):
            // 22 DFMA (three 64-bit sources), 23 DMUL (two), 24 DFMA with the addend equal to the destination (accumulate form).
            double s0, s1, s2, s3, s4, s5, s6, s7;
            if (MODE == 22)
                asm volatile("fma.rn.f64 %0,%8,%16,%24; fma.rn.f64 %1,%9,%17,%25; fma.rn.f64 %2,%10,%18,%26; fma.rn.f64 %3,%11,%19,%27;"
                             "fma.rn.f64 %4,%12,%20,%28; fma.rn.f64 %5,%13,%21,%29; fma.rn.f64 %6,%14,%22,%30; fma.rn.f64 %7,%15,%23,%31;"
                             : "=d"(s0), "=d"(s1), "=d"(s2), "=d"(s3), "=d"(s4), "=d"(s5), "=d"(s6), "=d"(s7)
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(q0), "d"(q1), "d"(q2), "d"(q3), "d"(q4), "d"(q5),
                               "d"(q6), "d"(q7), "d"(r0), "d"(r1), "d"(r2), "d"(r3), "d"(r4), "d"(r5), "d"(r6), "d"(r7));
            else if (MODE == 23)
                asm volatile("mul.rn.f64 %0,%8,%16; mul.rn.f64 %1,%9,%17; mul.rn.f64 %2,%10,%18; mul.rn.f64 %3,%11,%19;"
                             "mul.rn.f64 %4,%12,%20; mul.rn.f64 %5,%13,%21; mul.rn.f64 %6,%14,%22; mul.rn.f64 %7,%15,%23;"
                             : "=d"(s0), "=d"(s1), "=d"(s2), "=d"(s3), "=d"(s4), "=d"(s5), "=d"(s6), "=d"(s7)
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(q0), "d"(q1), "d"(q2), "d"(q3), "d"(q4), "d"(q5),
                               "d"(q6), "d"(q7));
            else {
                s0 = r0; s1 = r1; s2 = r2; s3 = r3; s4 = r4; s5 = r5; s6 = r6; s7 = r7;
                asm volatile("fma.rn.f64 %0,%8,%16,%0; fma.rn.f64 %1,%9,%17,%1; fma.rn.f64 %2,%10,%18,%2; fma.rn.f64 %3,%11,%19,%3;"
                             "fma.rn.f64 %4,%12,%20,%4; fma.rn.f64 %5,%13,%21,%5; fma.rn.f64 %6,%14,%22,%6; fma.rn.f64 %7,%15,%23,%7;"
                             : "+d"(s0), "+d"(s1), "+d"(s2), "+d"(s3), "+d"(s4), "+d"(s5), "+d"(s6), "+d"(s7)
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(q0), "d"(q1), "d"(q2), "d"(q3), "d"(q4), "d"(q5),
                               "d"(q6), "d"(q7));
            }
            a0 = s0; a1 = s1; a2 = s2; a3 = s3; a4 = s4; a5 = s5; a6 = s6; a7 = s7;  // loop-carried through the first operand only (one chain per instruction)
        } else if (MODE == 8) {  // 8 independent shuffles of a 64-bit value (2 SHFL each) -> 16 SHFL
            a0 = __shfl_sync(0xffffffffu, a0, (threadIdx.x + 1) & 31); a1 = __shfl_sync(0xffffffffu, a1, (threadIdx.x + 2) & 31);
            a2 = __shfl_sync(0xffffffffu, a2, (threadIdx.x + 3) & 31); a3 = __shfl_sync(0xffffffffu, a3, (threadIdx.x + 4) & 31);
            a4 = __shfl_sync(0xffffffffu, a4, (threadIdx.x + 5) & 31); a5 = __shfl_sync(0xffffffffu, a5, (threadIdx.x + 6) & 31);
            a6 = __shfl_sync(0xffffffffu, a6, (threadIdx.x + 7) & 31); a7 = __shfl_sync(0xffffffffu, a7, (threadIdx.x + 8) & 31);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3 +
                                                 (double)(p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7);
}

template <int MODE> void run(const char *name, int per_iter, int warps) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    bench<MODE><<<sms, warps * 32, 16384>>>(out, cyc, 1.0);
    bench<MODE><<<sms, warps * 32, 16384>>>(out, cyc, 1.0);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    // cycles per instruction-group per SMSP: warps/4 warps share one scheduler
    const double per_smsp = avg / ITERS / (warps / 4.0 < 1 ? 1 : warps / 4.0);
    printf("%-44s warps/SM=%2d  cycles/iter/warp=%8.2f  cycles per instr per SMSP=%6.3f (%d instr/iter)\n", name, warps, avg / ITERS, per_smsp / per_iter, per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("8 independent DFMA", 8, w);
        run<4>("4 DADD + 4 DMUL independent", 8, w);
        run<1>("8 DFMA + 8 FFMA interleaved", 16, w);
        run<2>("8 DFMA + 16 IMAD interleaved", 24, w);
        run<7>("8 FFMA", 8, w);
        run<8>("16 SHFL (8 x 64-bit)", 16, w);
        run<9>("8 FFMA2 (packed fp32x2)", 8, w);
        run<10>("8 FFMA2 + 8 IMAD interleaved", 16, w);
        run<11>("8 FFMA + 8 IMAD interleaved", 16, w);
    }
    for (int w : {16}) {
        run<12>("8 DFMA + 16 LOP3", 24, w);   run<13>("16 LOP3 alone", 16, w);
        run<14>("8 DFMA + 8 (ISETP+SEL)", 24, w); run<15>("8 (ISETP+SEL) alone", 16, w);
        run<16>("8 DFMA + 8 (VIMNMX+IADD)", 24, w); run<17>("8 (VIMNMX+IADD) alone", 16, w);
        run<18>("8 DFMA + 8 LDS.64", 16, w);  run<19>("8 LDS.64 alone", 8, w);
        run<20>("8 DFMA + 8 MUFU.EX2", 16, w); run<21>("8 MUFU.EX2 alone", 8, w);
    }
    for (int w : {8, 16}) {
        run<22>("8 DFMA, 3 distinct register operands each", 8, w);
        run<23>("8 DMUL, 2 distinct register operands each", 8, w);
        run<24>("8 DFMA, accumulate form d = a*b + d", 8, w);
    }
    run<3>("dependent DFMA chain (latency)", 8, 4);
    run<5>("dependent MUFU.RSQ64H + DFMA (latency of pair)", 4, 4);
    run<6>("dependent F2I.F64 + I2F.F64 (latency of pair)", 4, 4);
    return 0;
}
