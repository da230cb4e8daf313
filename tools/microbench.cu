// microbench.cu -- per-SM issue rates of the instructions the modular butterflies are made of (B200, sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define ITERS 4096
#define UNROLL 8

// hand-scheduled 64x64 -> high 64: four 32x32+64 multiply-adds, no carry predicates
__device__ __forceinline__ u64 mulhi_mw(u64 x, u64 s) {
    u32 x0 = (u32) x, x1 = (u32) (x >> 32), s0 = (u32) s, s1 = (u32) (s >> 32);
    u64 A, B, C, D;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(A) : "r"(x0), "r"(s0));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(B) : "r"(x1), "r"(s0), "l"(A >> 32));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(C) : "r"(x0), "r"(s1), "l"(B & 0xffffffffull));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(D) : "r"(x1), "r"(s1), "l"(B >> 32));
    return D + (C >> 32);
}
// "sloppy" high word: drops x0*s0, result is floor or floor-1
__device__ __forceinline__ u64 mulhi_sloppy(u64 x, u64 s) {
    u32 x0 = (u32) x, x1 = (u32) (x >> 32), s0 = (u32) s, s1 = (u32) (s >> 32);
    u64 B, C, D;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(B) : "r"(x1), "r"(s0));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(C) : "r"(x0), "r"(s1), "l"(B & 0xffffffffull));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(D) : "r"(x1), "r"(s1), "l"(B >> 32));
    return D + (C >> 32);
}
// low 64 bits of a*b + c*d as one accumulate chain
__device__ __forceinline__ u64 lo_mac2(u64 a, u64 b, u64 c, u64 d) {
    u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
    u32 c0 = (u32) c, c1 = (u32) (c >> 32), d0 = (u32) d, d1 = (u32) (d >> 32);
    u64 acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(a0), "r"(b0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(c0), "r"(d0));
    u32 h = a0 * b1 + a1 * b0 + c0 * d1 + c1 * d0;
    return acc + ((u64) h << 32);
}

template<int MODE>
__global__ void __launch_bounds__(256) k(u64 *out, u64 seed, double dseed) {
    u64 a[UNROLL];
    double d[UNROLL];
    u32 lo[UNROLL], hi[UNROLL], x2[UNROLL];
    double e2[UNROLL];
    for (int i = 0; i < UNROLL; i++) {
        x2[i] = i;
        e2[i] = dseed * 7 + i;
        a[i] = seed + threadIdx.x * 977 + i * 13;
        d[i] = dseed + threadIdx.x * 0.5 + i;
        lo[i] = (u32) a[i];
        hi[i] = (u32) (a[i] >> 32) + 3;
    }
    const u64 w = seed * 3 + 1, ws = seed * 5 + 7, nq = 0 - (seed | 1);
    const double dw = dseed * 1.25, dq = dseed * 3.0 + 1.0, dinv = 1.0 / dq, magic = 6755399441055744.0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (MODE == 0) {   // IMAD.WIDE.U32 (32x32+64)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(lo[i]), "r"(hi[i]));
            } else if (MODE == 1) {   // IMAD (32x32+32 low)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(hi[i]), "r"((u32) it));
            } else if (MODE == 2) {   // mul.hi.u64
                a[i] = __umul64hi(a[i], ws) + it;
            } else if (MODE == 3) {   // Shoup lazy modmul (integer)
                a[i] = a[i] * w + __umul64hi(a[i], ws) * nq;
            } else if (MODE == 4) {   // DFMA
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dq));
            } else if (MODE == 5) {   // FP64 modmul: t = w*y mod q, error-free (6 fp64 ops)
                double y = d[i];
                double kq = __fma_rn(y, dinv * dw, magic) - magic;
                double p = y * dw;
                double e = __fma_rn(y, dw, -p);
                double r = __fma_rn(-kq, dq, p);
                d[i] = r + e;
            } else if (MODE == 6) {   // IADD3 64-bit add (2 instr)
                a[i] = a[i] + w + it;
            } else if (MODE == 7) {   // IMAD.HI.U32
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(hi[i]));
            } else if (MODE == 8) {   // 64-bit mul.lo
                a[i] = a[i] * w + it;
            } else if (MODE == 9) {
                a[i] = mulhi_mw(a[i], ws) + it;
            } else if (MODE == 10) {
                a[i] = lo_mac2(a[i], w, mulhi_mw(a[i], ws), nq);
            } else if (MODE == 11) {
                a[i] = lo_mac2(a[i], w, mulhi_sloppy(a[i], ws), nq);
            } else if (MODE == 12) {
                a[i] = a[i] * w + mulhi_mw(a[i], ws) * nq;
            } else if (MODE == 13) {   // FRND.F64
                asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(d[i]));
            } else if (MODE == 14) {   // F2I.S64.F64
                asm volatile("cvt.rni.s64.f64 %0, %1;" : "=l"(a[i]) : "d"(d[i]));
            } else if (MODE == 15) {   // I2F.F64.S64
                asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(d[i]) : "l"(a[i]));
            } else if (MODE == 16) {   // DFMA and IMAD.WIDE side by side (independent chains)
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dq));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(lo[i]), "r"(hi[i]));
            } else if (MODE == 17) {   // DFMA and FRND side by side
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dq));
                asm volatile("cvt.rni.f64.f64 %0, %1;" : "=d"(e2[i]) : "d"(d[i]));
            } else if (MODE == 18) {   // FP64 modmul with the quotient rounded by FRND (5 fp64 ops + FRND)
                double y = d[i];
                double kq;
                asm("cvt.rni.f64.f64 %0, %1;" : "=d"(kq) : "d"(y * (dinv * dw)));
                double p = y * dw;
                double e = __fma_rn(y, dw, -p);
                double r = __fma_rn(-kq, dq, p);
                d[i] = r + e;
            } else if (MODE == 19) {   // dual-pipe modmul: quotient on the FP64 pipe, remainder as low 64 bits on the integer pipe
                const u64 y = a[i] & 0xfffffffffffull;
                const double yd = __longlong_as_double((long long) (y | 0x4330000000000000ull)) - 4503599627370496.0;
                const u64 kb = (u64) __double_as_longlong(__fma_rn(yd, dinv * dw, magic));
                a[i] = lo_mac2(y, w, kb, nq);
            } else if (MODE == 20) {   // DFMA + 2 IMAD.WIDE + 2 IADD3-class per iteration
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dq));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(lo[i]), "r"(hi[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(hi[i]), "r"((u32) it));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"((u32) it));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x2[i]) : "r"((u32) it));
            } else if (MODE == 21) {   // two FP64 modmuls + one dual-pipe modmul, independent chains
                {
                    double y = d[i];
                    double kq = __fma_rn(y, dinv * dw, magic) - magic;
                    double p = y * dw;
                    double e = __fma_rn(y, dw, -p);
                    d[i] = __fma_rn(-kq, dq, p) + e;
                }
                {
                    double y = e2[i];
                    double kq = __fma_rn(y, dinv * dw, magic) - magic;
                    double p = y * dw;
                    double e = __fma_rn(y, dw, -p);
                    e2[i] = __fma_rn(-kq, dq, p) + e;
                }
                const u64 y = a[i] & 0xfffffffffffull;
                const double yd = __longlong_as_double((long long) (y | 0x4330000000000000ull)) - 4503599627370496.0;
                const u64 kb = (u64) __double_as_longlong(__fma_rn(yd, dinv * dw, magic));
                a[i] = lo_mac2(y, w, kb, nq);
            }
        }
    }
    u64 acc = 0;
    for (int i = 0; i < UNROLL; i++) acc += a[i] + (u64) d[i] + lo[i] + x2[i] + hi[i] + (u64) e2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int MODE>
void run(const char *name, double ops_per_iter) {
    u64 *out;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8;
    cudaMalloc(&out, (size_t) blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345, 3.0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 12345, 3.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double total = (double) blocks * 256 * ITERS * UNROLL * ops_per_iter;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_sm_clk = total / (ms * 1e-3) / sms / (clk * 1e3);
    printf("%-34s %8.3f ms  %9.2f Gop/s  %6.2f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, total / ms / 1e6,
           per_sm_clk, clk / 1000);
    cudaFree(out);
}

int main() {
    run<0>("mad.wide.u32 (IMAD.WIDE)", 1);
    run<1>("mad.lo.u32 (IMAD)", 1);
    run<7>("mul.hi.u32 (IMAD.HI)", 1);
    run<8>("mul.lo.u64 + add", 1);
    run<2>("mul.hi.u64 + add", 1);
    run<6>("64-bit add x2", 2);
    run<3>("Shoup lazy modmul (int)", 1);
    run<9>("mulhi via 4 mad.wide + add", 1);
    run<12>("Shoup lazy, mad.wide mulhi", 1);
    run<10>("Shoup lazy, mad.wide + lo_mac2", 1);
    run<11>("Shoup lazy sloppy-hi + lo_mac2", 1);
    run<4>("fma.rn.f64 (DFMA)", 1);
    run<5>("FP64 error-free modmul", 1);
    run<13>("cvt.rni.f64.f64 (FRND.F64)", 1);
    run<14>("cvt.rni.s64.f64 (F2I)", 1);
    run<15>("cvt.rn.f64.s64 (I2F)", 1);
    run<16>("DFMA + IMAD.WIDE pairs", 1);
    run<17>("DFMA + FRND pairs", 1);
    run<18>("FP64 modmul, FRND quotient", 1);
    run<19>("dual-pipe modmul (FP64 quotient, int lo64)", 1);
    run<20>("DFMA + IMAD.WIDE + IMAD + 2 ALU groups", 1);
    run<21>("2 FP64 modmuls + 1 dual-pipe modmul groups", 1);
    return 0;
}
