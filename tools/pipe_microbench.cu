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
    bench<MODE><<<sms, warps * 32>>>(out, cyc, 1.0);
    bench<MODE><<<sms, warps * 32>>>(out, cyc, 1.0);
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
    run<3>("dependent DFMA chain (latency)", 8, 4);
    run<5>("dependent MUFU.RSQ64H + DFMA (latency of pair)", 4, 4);
    run<6>("dependent F2I.F64 + I2F.F64 (latency of pair)", 4, 4);
    return 0;
}
