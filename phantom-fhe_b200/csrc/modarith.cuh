// modarith.cuh -- 64-bit modular arithmetic for the B200 RNS engine (device side).
//
// Semantics follow the reference's include/uintmodmath.cuh (file:line cited per function) but the code is
// organised around *lazy* value ranges tracked at compile time by the NTT kernels, so that conditional
// subtractions are only issued where a 64-bit overflow would otherwise occur.  Every value that leaves a
// kernel is a canonical residue in [0, q) -- the reference's store invariant.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pfhe {

// Programmatic dependent launch (sm_90+): every kernel of the engine lets its successor in the stream start
// launching right away (pdl_launch_dependents) and waits for its predecessor's results only after its own
// prologue (pdl_wait).  Kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch.hpp),
// so the ~4 us drain + launch gap between the 12 dependent kernels of a key switch overlaps with useful work.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

using u64 = unsigned long long;
using u32 = unsigned int;

// {q, floor(2^128/q) lo, hi} -- same content as the reference's DModulus (include/ntt.cuh:6-32)
struct __align__(8) Modulus {
    u64 q;
    u64 mu_lo;
    u64 mu_hi;
};

// (w, floor(w * 2^64 / q)) pair, 16 bytes so that one 128-bit load fetches both (the reference keeps two
// separate arrays, include/ntt.cuh:40-44)
using Tw = ulonglong2;

__device__ __forceinline__ u64 mulhi(u64 a, u64 b) { return __umul64hi(a, b); }

// x - q if x >= q  (csub_q, uintmodmath.cuh:18-21), branch-free min trick: valid for x < 2q, q < 2^63
__device__ __forceinline__ u64 csub(u64 x, u64 q) {
    u64 t = x - q;
    return (long long) t < 0 ? x : t;
}

// Shoup product, result in [0, 2q) for ANY 64-bit x (multiply_and_reduce_shoup_lazy, uintmodmath.cuh:226-231)
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 hi = mulhi(x, ws);
    return x * w - hi * q;
}

// same with the negated modulus nq = 2^64 - q: one multiply-accumulate chain, no subtraction
__device__ __forceinline__ u64 mul_shoup_lazy_neg(u64 x, u64 w, u64 ws, u64 nq) {
    return x * w + mulhi(x, ws) * nq;
}

// canonical Shoup product (multiply_and_reduce_shoup, uintmodmath.cuh:207-216)
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) { return csub(mul_shoup_lazy(x, w, ws, q), q); }
__device__ __forceinline__ u64 mul_shoup(u64 x, Tw w, u64 q) { return csub(mul_shoup_lazy(x, w.x, w.y, q), q); }

__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }          // :36-42
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return csub(a + q - b, q); }      // :47-53

// 128-bit value -> [0, q) (barrett_reduce_uint128_uint64, uintmodmath.cuh:96-136).  Valid for any
// (hi, lo) with q < 2^61: the quotient estimate is at most one too small.
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const Modulus &m) {
    // floor(((hi:lo) * (mu_hi:mu_lo)) / 2^128), low 64 bits only
    u64 t0 = mulhi(lo, m.mu_lo);
    u64 p1lo = lo * m.mu_hi, p1hi = mulhi(lo, m.mu_hi);
    u64 p2lo = hi * m.mu_lo, p2hi = mulhi(hi, m.mu_lo);
    u64 s = t0 + p1lo;
    u64 c1 = s < t0;
    u64 s2 = s + p2lo;
    u64 c2 = s2 < s;
    u64 quo = hi * m.mu_hi + p1hi + p2hi + c1 + c2;
    u64 r = lo - quo * m.q;
    return csub(r, m.q);
}

// a * b mod q, canonical (multiply_and_barrett_reduce_uint64, uintmodmath.cuh:160-198)
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const Modulus &m) {
    return barrett128(a * b, mulhi(a, b), m);
}

// x mod q for a single word (barrett_reduce_uint64_uint64, uintmodmath.cuh:144-151)
__device__ __forceinline__ u64 barrett64(u64 x, const Modulus &m) {
    u64 s = mulhi(m.mu_hi, x);
    return csub(x - s * m.q, m.q);
}

// full 64x64 -> 128 product from four 32x32+64 multiply-adds (the compiler's a*b and __umul64hi(a,b) do not
// share partial products: 5 wide + 2 narrow multiplies instead of 4 wide)
__device__ __forceinline__ void mul128(u64 a, u64 b, u64 &lo, u64 &hi) {
    const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
    u64 A, B, C, D;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(A) : "r"(a0), "r"(b0));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(B) : "r"(a1), "r"(b0), "l"(A >> 32));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(C) : "r"(a0), "r"(b1), "l"(B & 0xffffffffull));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(D) : "r"(a1), "r"(b1), "l"(B >> 32));
    lo = (A & 0xffffffffull) | (C << 32);
    hi = D + (C >> 32);
}

// Single-word Barrett for values known to be < 2^(2k+g) (k = bit length of q, g = growth bits of a sum of
// products): take the 64 bits of x starting at bit sh = max(0, 2k+g-64), one high multiply by
// mu = floor(2^(64+sh)/q), one low multiply, two conditional subtractions.  Valid when k + g <= 63
// (sh <= k-1, so the quotient estimate is at most 2 too small; derivation in DESIGN.md).  Same canonical
// result as barrett_reduce_uint128_uint64 (uintmodmath.cuh:96-136) at about a third of the multiplies.
struct BarG {
    u64 mu;
    u32 sh;      // 0..64, 0xff = not applicable for this modulus/growth: use the two-word Barrett
    u32 pad;
};
__device__ __forceinline__ u64 barrett_g(u64 lo, u64 hi, const BarG &b, const Modulus &m) {
    if (b.sh == 0xffu) return barrett128(lo, hi, m);   // CTA-uniform
    u64 x;
    if (b.sh == 0) x = lo;
    else if (b.sh == 64) x = hi;
    else x = (lo >> b.sh) | (hi << (64 - b.sh));
    const u64 q3 = mulhi(x, b.mu);
    u64 r = lo - q3 * m.q;
    r = csub(r, 2 * m.q);
    return csub(r, m.q);
}
__device__ __forceinline__ u64 mul_mod_g(u64 a, u64 b, const BarG &bg, const Modulus &m) {
    u64 lo, hi;
    mul128(a, b, lo, hi);
    return barrett_g(lo, hi, bg, m);
}

// 128-bit accumulator for inner products / base conversion (uintmath.cuh add_uint128_uint128).  The product and
// the carry chain are left to the compiler's native 128-bit arithmetic: 4 IMAD.WIDE + 4 carry-chained adds per
// multiply-accumulate, against 17 instructions for explicit partial products with compare-and-select carries.
struct Acc128 {
    u64 lo, hi;
    __device__ __forceinline__ void mac(u64 a, u64 b) {
        unsigned __int128 v = ((unsigned __int128) hi << 64) | lo;
        v += (unsigned __int128) a * b;
        lo = (u64) v, hi = (u64) (v >> 64);
    }
};

// ---------------------------------------------------------------------------------------------------
// FP64-pipe modular arithmetic for moduli below 2^46: values are exact integers held in doubles, products are
// error-free FMA sequences (see FpArith in ntt.cuh and tests/fp_modmul_check.c).  Used by every kernel for the
// small limbs of a chain; the results are converted back to the same canonical residues.
// ---------------------------------------------------------------------------------------------------
namespace fp {
constexpr double TWO52 = 4503599627370496.0;   // 2^52
constexpr double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
constexpr int MAX_BITS = 46;
constexpr int SPLIT_BITS = 30;                 // operands from a >= 2^46 modulus enter as two 30/31-bit halves

__device__ __forceinline__ double from_u64(u64 v) {   // exact for v < 2^52
    return __longlong_as_double((long long) (v | 0x4330000000000000ull)) - TWO52;
}
__device__ __forceinline__ u64 to_u64(double v) {     // exact for integers 0 <= v < 2^52
    return ((u64) __double_as_longlong(v + TWO52)) & 0x000fffffffffffffull;
}
// rint(y * c): quotient estimate of the FP64 modular product.  Two forms with the same accuracy for our purposes
// (off by at most one from the exact quotient): one FMA + one add against 1.5 * 2^52 on the FP64 pipe, or one
// multiply plus FRND.F64, which issues beside the FP64 pipe (tools/microbench.cu).
#ifndef PFHE_FRND
#define PFHE_FRND 0
#endif
__device__ __forceinline__ double rint_q(double y, double c) {
#if PFHE_FRND
    double k;
    asm("cvt.rni.f64.f64 %0, %1;" : "=d"(k) : "d"(y * c));
    return k;
#else
    return __fma_rn(y, c, MAGIC) - MAGIC;
#endif
}
__device__ __forceinline__ double reduce(double v, double q, double qinv) {   // -> [-q/2, q/2]
    const double k = rint_q(v, qinv);
    return __fma_rn(-k, q, v);
}
// y * w mod q for a constant w with precomputed winv = w/q: result in (-0.63q, 0.63q)
__device__ __forceinline__ double mulmod_c(double y, double w, double winv, double q) {
    const double k = rint_q(y, winv);
    const double p = y * w;
    const double e = __fma_rn(y, w, -p);
    return __fma_rn(-k, q, p) + e;
}
// a * b mod q, both variable, |a|,|b| < 2^47
__device__ __forceinline__ double mulmod_v(double a, double b, double q, double qinv) {
    const double p = a * b;
    const double e = __fma_rn(a, b, -p);
    const double k = rint_q(p, qinv);
    return __fma_rn(-k, q, p) + e;
}
__device__ __forceinline__ u64 canon(double v, double q) {   // v in (-q, q) -> [0, q)
    return to_u64(v < 0.0 ? v + q : v);
}
} // namespace fp

} // namespace pfhe
