// ntt.cuh -- negacyclic NTT kernels for sm_100a (forward: natural -> bit-reversed, inverse: the converse).
//
// Replaces the reference's nwt_2d_radix8_* family (include/ntt.cuh:157-226, src/ntt/*.cu).  Same transform
// (NTT(x)[k] = sum_j x_j psi^{j(2 bitrev(k)+1)}), different organisation:
//
//  * N = 2^P1 * 2^P2.  A "column pass" runs the P1 stages whose butterfly span is >= 2^P2 on tiles of
//    2^P1 rows x C adjacent columns, a "row pass" runs the remaining P2 stages on tiles of C contiguous rows.
//    Every CTA owns a 2048-element tile (256 threads x 8 elements); each thread keeps its 8 elements in
//    registers and runs up to 3 merged stages (radix-8) per round, rounds exchange through a skewed,
//    bank-conflict-free shared-memory tile, so a 2^16-point transform is 2 passes x 3 rounds.  (A 16-element
//    / radix-16 variant was measured first: fewer exchanges but half the warps, IPC 0.43 -- see DESIGN.md.)
//  * Lazy ranges are tracked at compile time: forward butterflies issue a conditional subtraction only on
//    every second stage (values < 8q < 2^64 for q < 2^61), the result is made canonical once at the end.
//  * Twiddles are (w, floor(w 2^64/q)) pairs fetched with one 128-bit load; the table is stored in a
//    kernel-native order so that the warp that owns a contiguous butterfly span reads contiguous twiddles
//    (the last row-pass round, where every butterfly has its own twiddle, is stored transposed).
//  * Prologue / epilogue functors fuse the neighbouring element-wise steps (n^-1 and digit scaling, the
//    mod-down epilogue, ...) into the passes.
#pragma once
#include <type_traits>
#include "modarith.cuh"
#include "tma.cuh"

namespace pfhe {

#ifndef PFHE_NTT_LOG_THREADS
#define PFHE_NTT_LOG_THREADS 8
#endif
#ifndef PFHE_NTT_LOG_EPT
#define PFHE_NTT_LOG_EPT 3
#endif
constexpr int NTT_LOG_THREADS = PFHE_NTT_LOG_THREADS;
constexpr int NTT_THREADS = 1 << NTT_LOG_THREADS;
constexpr int NTT_LOG_EPT = PFHE_NTT_LOG_EPT;         // elements per thread = one radix-8 group (A/B builds: 2 = radix-4, 512 threads)
constexpr int NTT_EPT = 1 << NTT_LOG_EPT;
constexpr int NTT_LOG_TILE = NTT_LOG_THREADS + NTT_LOG_EPT;
constexpr int NTT_TILE = 1 << NTT_LOG_TILE;
constexpr int NTT_SMEM_WORDS = NTT_TILE;
constexpr int NTT_MAX_LIMBS = 128;

// which limbs a launch touches: slot i works on data limb `data[i]` (units of N words from the base
// pointer) with the constant tables of key-level prime `row[i]`.  This one descriptor covers the
// reference's start_modulus_idx / include_special_mod / include_temp_mod / exclude_range variants.
struct LimbList {
    int count;
    short data[NTT_MAX_LIMBS];   // destination limb (and source limb when src_same)
    short row[NTT_MAX_LIMBS];    // key-level prime row
    short src[NTT_MAX_LIMBS];    // source limb of the first pass (out-of-place launches)
    u64 q[NTT_MAX_LIMBS];        // the modulus itself: kernels derive every per-limb constant from it without a
                                 // dependent global load at kernel start (q < 2^46 selects the FP64 butterflies)
};

// epilogue of the fused forward row pass:  out = (sub - NTT(x)) * mulc  (+ add)   mod q
// (nwt_2d_radix8_forward_inplace_fuse_moddown, reference src/ntt/ntt_moddown.cu:106-216,
//  + add_to_ct_kernel rns_bconv.cu:763-769; also divide_and_round_ntt_inv_scalar_kernel rns.cu:1141-1158)
struct EpiArgs {
    const u64 *sub_base;
    u64 *out_base;
    const u64 *add_base;
    const Tw *mulc;              // per slot
    short sub[NTT_MAX_LIMBS];
    short out[NTT_MAX_LIMBS];
    short add[NTT_MAX_LIMBS];    // -1: nothing to add
};

// round schedule of a pass with P stages: ceil(P/LOG_EPT) rounds, the earlier rounds take the remainder
template<int P>
struct Sched {
    static constexpr int NR = (P + NTT_LOG_EPT - 1) / NTT_LOG_EPT;
    static constexpr int r(int i) { return P / NR + (i < P % NR ? 1 : 0); }
    static constexpr int s0(int i) {
        int s = 0;
        for (int j = 0; j < i; j++) s += r(j);
        return s;
    }
};

__host__ __device__ constexpr int ntt_p1(int logn) { return logn / 2; }
__host__ __device__ constexpr int ntt_p2(int logn) { return logn - logn / 2; }

// position of the twiddle for global stage s, block B inside the per-limb table (host + device)
__host__ __device__ inline size_t tw_native_index(int logn, int s, size_t B) {
    const int p1 = ntt_p1(logn), p2 = ntt_p2(logn);
    const int nr = (p2 + NTT_LOG_EPT - 1) / NTT_LOG_EPT;
    const int rlast = p2 / nr;               // the last round never gets a remainder stage
    const int s0last = p2 - rlast;
    if (s < p1 + s0last) return ((size_t) 1 << s) + B;
    const int sigma = s - p1, u = sigma - s0last;
    const size_t row = B >> sigma, bl = B & (((size_t) 1 << sigma) - 1);
    const size_t hi = bl >> u, b = bl & (((size_t) 1 << u) - 1);
    return ((size_t) 1 << s) + (row << sigma) + (b << s0last) + hi;
}

// XOR swizzle of the exchange tile (64-bit words, 16 words per shared-memory wavefront).  The access
// patterns of the radix-8 rounds are: 16 contiguous words; 4 contiguous x 4 blocks 32 apart; 16 lanes 4
// apart; 8 contiguous x 2 blocks 32 apart.  Folding bits 4..6 into bits 0..3 makes all of them hit 16
// distinct bank pairs (derivation in DESIGN.md); it is a bijection inside every aligned 128-word block.
__host__ __device__ constexpr int skew_bits(int i) {   // GF(2)-linear in bits 4..6 of i, lands in bits 0..3
    return ((i >> 4) & 3) ^ (((i >> 5) & 1) << 3) ^ (((i >> 6) & 1) << 2);
}
__host__ __device__ constexpr int skew(int i) { return i ^ skew_bits(i); }

// Fused prologues / epilogues of the key-switch pipeline (engine.cu: Engine::keyswitch_fused) -----------
// Base conversion folded into the load of the forward column pass: slot s converts the `ni` coefficient-form
// limbs starting at limb in_limb[s] of in_base with matrix row mat_row[s] (bconv_matmul_*, reference
// src/rns_bconv.cu:109-210,455-485), so the converted polynomial never goes through memory before its NTT.
#ifndef PFHE_FUSE_MAX_IN
#define PFHE_FUSE_MAX_IN 6
#endif
constexpr int FUSE_MAX_IN = PFHE_FUSE_MAX_IN;   // special primes per digit the fused conversions take (benchmark/ckks_bench.cu goes up to 6)
struct BconvLoad {
    const u64 *in_base;
    const u64 *mat;          // [rows][ni] qhat_i mod p_j
    const double2 *matf;     // [rows][ni][2] FP64 form (see BconvJob::matf)
    const BarG *bar;         // [64][size_QP] single-word Barrett table
    int size_QP;
    int ni;                  // inputs per output (same for every slot of the launch), <= FUSE_MAX_IN
    int xbits;               // max input-prime bits + ceil(log2 ni)
    short in_limb[NTT_MAX_LIMBS];
    short mat_row[NTT_MAX_LIMBS];
    unsigned char in_big[NTT_MAX_LIMBS];   // per slot: bit i set = input i comes from a modulus >= 2^46
};

// HMult folded into the key switch: the tensor product d = (a0 b0, a0 b1 + a1 b0, a1 b1) is never stored;
// d2 is formed in the load of the first inverse pass, d0 / d1 in the epilogue of the last forward pass
// (tensor_prod_2x2_rns_poly, reference src/polymath.cu:463-498: same residues).
struct TensorSrc {
    const u64 *a;            // [2][l][n]
    const u64 *b;            // [2][l][n]
    int l;
};

// ----------------------------------------------------------------------------------------------------
// element <-> thread mapping of one round
// ----------------------------------------------------------------------------------------------------
template<int P, int RI, bool ROWS>
struct RoundMap {
    static constexpr int R = Sched<P>::r(RI);
    static constexpr int S0 = Sched<P>::s0(RI);
    static constexpr int LAM = P - S0 - R;           // bits of the low index
    static constexpr int GAM = NTT_LOG_TILE - P;     // bits of the batch index (columns resp. rows per tile)
    static constexpr int GB = NTT_LOG_EPT - R;
    static constexpr int G = 1 << GB;                // independent radix-2^R groups per thread
    static constexpr int T = 1 << P;
    static constexpr int T_LOG = P;
    static constexpr int C = 1 << GAM;
    static constexpr bool LAST = (RI == Sched<P>::NR - 1);
    // a thread's G groups sit NTT_THREADS apart in the packed (hi, lo, batch) index, i.e. they differ in its
    // top bits; they use the same twiddles only when those bits belong to the low index
    static constexpr bool SHARE = !ROWS && S0 == 0 && LAM >= GB;
    // row pass, one radix-2^R group per thread and 2^(P-R) = 32 threads per row: a warp owns a whole row of the tile in
    // this round.  An exchange between two such rounds stays inside the warp (__syncwarp instead of a CTA barrier)
    static constexpr bool WARP_ROW = ROWS && G == 1 && (P - R == 5) && (NTT_LOG_THREADS + NTT_LOG_EPT - P == 3);
    __device__ static __forceinline__ int mu(int tid, int g) { return (g << NTT_LOG_THREADS) | tid; }

    __device__ static __forceinline__ void decode(int mu, int &hi, int &lo, int &c) {
        if constexpr (!ROWS) {
            c = mu & (C - 1);
            lo = (mu >> GAM) & ((1 << LAM) - 1);
            hi = mu >> (GAM + LAM);
        } else {
            lo = mu & ((1 << LAM) - 1);
            hi = (mu >> LAM) & ((1 << S0) - 1);
            c = mu >> (LAM + S0);
        }
    }
    __device__ static __forceinline__ int elem(int hi, int k, int lo) { return (hi << (P - S0)) | (k << LAM) | lo; }
    // exchange-tile position of element k of a group: the k field occupies its own bits of the linear index, the
    // swizzle is XOR-linear, hence position = sidx0 ^ KC(k) with a compile-time KC -- one LOP per access
    static constexpr int KSHIFT = ROWS ? LAM : LAM + GAM;
    __device__ static __forceinline__ int sidx0(int hi, int lo, int c) {
        const int e = (hi << (P - S0)) | lo;
        return skew(ROWS ? (c << P) | e : (e << GAM) | c);
    }
    __host__ __device__ static constexpr int kc(int k) { return skew(k << KSHIFT); }
};

// synchronisation of an exchange between round RA (writer side) and round RB (reader side) of a pass
template<int P, int RA, int RB, bool ROWS>
__device__ __forceinline__ void exchange_sync() {
    if constexpr (RoundMap<P, RA, ROWS>::WARP_ROW && RoundMap<P, RB, ROWS>::WARP_ROW) __syncwarp();
    else __syncthreads();
}

// ----------------------------------------------------------------------------------------------------
// arithmetic policies.  Two exact implementations of the same modular butterflies:
//
//  IntArith  -- 64-bit Shoup multiplication on the integer pipe, any q < 2^61 (reference butterfly.cuh:10-37
//               semantics with a relaxed reduction schedule).
//  FpArith   -- for q < 2^46: values are kept as exact integers in FP64 registers and the modular product is an
//               error-free FMA sequence (p = y*w rounded, e = fma(y,w,-p) its exact error, k = rint(y*w/q),
//               r = fma(-k,q,p) + e exactly).  On B200 DFMA issues at the full 64 lanes/clk/SM and on its own
//               pipe, while a 64x64->128 integer product costs ~10 IMAD slots; measured 9.7 vs 3.9 modmuls/clk/SM
//               (tools/microbench.cu, profiles/).  Results are the same canonical residues bit for bit.
// ----------------------------------------------------------------------------------------------------
// forward stage s takes inputs < 4q (s = 0), < 6q (odd s) or < 8q (even s >= 2) and reduces on even s >= 2
__host__ __device__ constexpr bool fwd_csub(int s) { return s >= 2 && (s % 2) == 0; }

struct IntArith {
    using T = u64;
    struct Consts {
        u64 q, q2, q4, nq;
    };
    __device__ static __forceinline__ Consts consts(u64 q) { return Consts{q, 2 * q, 4 * q, 0 - q}; }
    __device__ static __forceinline__ T load(u64 canonical, const Consts &) { return canonical; }
    __device__ static __forceinline__ u64 raw(T v) { return v; }
    __device__ static __forceinline__ T from_raw(u64 bits) { return bits; }
    // forward (Cooley-Tukey, Harvey lazy)
    template<bool CSUB>
    __device__ static __forceinline__ void fwd(T &x, T &y, const Tw w, const Consts &c) {
        u64 X = x;
        if constexpr (CSUB) X = csub(X, c.q4);
        const u64 t = mul_shoup_lazy_neg(y, w.x, w.y, c.nq);
        x = X + t;
        y = X + c.q2 - t;
    }
    __device__ static __forceinline__ u64 canon_fwd(T v, const Consts &c) {   // < 8q -> [0, q)
        return csub(csub(csub(v, c.q4), c.q2), c.q);
    }
    // inverse (Gentleman-Sande): inputs and outputs in [0, 2q)
    __device__ static __forceinline__ void inv(T &x, T &y, const Tw w, const Consts &c) {
        const u64 s = csub(x + y, c.q2);
        const u64 d = x + c.q2 - y;
        x = s;
        y = mul_shoup_lazy_neg(d, w.x, w.y, c.nq);
    }
    // last inverse stage: folds n^-1 (and an optional per-limb scalar) into both outputs, canonical results
    // (replaces intt_2d.cu:195-203 "lower half times n^-1, upper half through itw[1]")
    __device__ static __forceinline__ void inv_last(T &x, T &y, const Tw cx, const Tw cy, const Consts &c) {
        const u64 s = x + y;
        const u64 d = x + c.q2 - y;
        x = mul_shoup(s, cx, c.q);
        y = mul_shoup(d, cy, c.q);
    }
    __device__ static __forceinline__ void inv_round_end(T &, const Consts &) {}
    __device__ static __forceinline__ u64 canon_inv(T v, const Consts &) { return v; }
};

struct FpArith {
    using T = double;
    struct Consts {
        double q, qinv;
    };
    static constexpr double TWO52 = 4503599627370496.0;          // 2^52
    static constexpr double MAGIC = 6755399441055744.0;          // 1.5 * 2^52: round-to-nearest-integer trick
    __device__ static __forceinline__ Consts consts(u64 q) {   // same values as the host table: IEEE division
        const double qd = (double) q;
        return Consts{qd, 1.0 / qd};
    }
    __device__ static __forceinline__ T load(u64 canonical, const Consts &) {   // exact for values < 2^52
        return __longlong_as_double((long long) (canonical | 0x4330000000000000ull)) - TWO52;
    }
    __device__ static __forceinline__ u64 raw(T v) { return (u64) __double_as_longlong(v); }
    __device__ static __forceinline__ T from_raw(u64 bits) { return __longlong_as_double((long long) bits); }
    __device__ static __forceinline__ u64 to_u64(T nonneg_int) {                // exact for 0 <= v < 2^52
        return ((u64) __double_as_longlong(nonneg_int + TWO52)) & 0x000fffffffffffffull;
    }
    // y * w mod q, result in (-0.62q, 0.62q); exact for |y| < 2^51, w < q < 2^47 (tests/test_fp_modmul.py)
    __device__ static __forceinline__ T mulmod(T y, const Tw w, const Consts &c) {
        const double wv = __longlong_as_double((long long) w.x), wi = __longlong_as_double((long long) w.y);
        const double k = fp::rint_q(y, wi);
        const double p = y * wv;
        const double e = __fma_rn(y, wv, -p);
        const double r = __fma_rn(-k, c.q, p);
        return r + e;
    }
    __device__ static __forceinline__ T reduce(T v, const Consts &c) {           // -> [-q/2, q/2]
        const double k = fp::rint_q(v, c.qinv);
        return __fma_rn(-k, c.q, v);
    }
    template<bool CSUB>
    __device__ static __forceinline__ void fwd(T &x, T &y, const Tw w, const Consts &c) {
        const double t = mulmod(y, w, c);   // |values| grow by < 0.62q per stage: < 12q < 2^50 after 17 stages
        const double X = x;
        x = X + t;
        y = X - t;
    }
    __device__ static __forceinline__ u64 canon_fwd(T v, const Consts &c) {
        double r = reduce(v, c);
        if (r < 0.0) r += c.q;
        return to_u64(r);
    }
    __device__ static __forceinline__ void inv(T &x, T &y, const Tw w, const Consts &c) {
        const double s = x + y, d = x - y;
        x = s;
        y = mulmod(d, w, c);
    }
    __device__ static __forceinline__ void inv_last(T &x, T &y, const Tw cx, const Tw cy, const Consts &c) {
        const double s = x + y, d = x - y;
        double a = mulmod(s, cx, c), b = mulmod(d, cy, c);
        if (a < 0.0) a += c.q;
        if (b < 0.0) b += c.q;
        x = a, y = b;
    }
    // only the all-sums element of a radix-2^R group can grow by 2^R per round: fold it back
    __device__ static __forceinline__ void inv_round_end(T &x0, const Consts &c) { x0 = reduce(x0, c); }
    __device__ static __forceinline__ u64 canon_inv(T v, const Consts &) { return to_u64(v); }
};

// Twiddles of a tile are staged in shared memory by TMA bulk copies (stage_twiddles below); position of the
// twiddle of (stage S0+u, hi, b) in that staging area.  Column pass: the 2^P1 - 1 twiddles of the first P1 stages
// in table order.  Row pass: per stage sigma the C rows of the tile are contiguous in the table
// (C * 2^sigma entries starting at 2^(P1+sigma) + row0 * 2^sigma), stored back to back: offset C * (2^sigma - 1).
template<class M, bool ROWS, int LOGN>
__device__ __forceinline__ int tw_index(int u, int hi, int b, int c, int tile) {
    if constexpr (!ROWS) {
        return (1 << (M::S0 + u)) + ((hi << u) | b);          // staged copy of the table head: same positions
    } else {
        // row pass: every twiddle is used by exactly one thread of the grid, so it is read straight from the table
        // (kernel-native order: contiguous across the lanes of a warp in the last round)
        constexpr int P1 = LOGN - M::T_LOG;
        const int sigma = M::S0 + u;
        const int row = (tile << M::GAM) + c;
        const int base = (1 << (P1 + sigma)) + (row << sigma);
        if constexpr (M::LAST) return base + (b << M::S0) + hi;
        else return base + ((hi << u) | b);
    }
}

constexpr int NTT_STW_ENTRIES = 256;   // column pass: the 2^P1 <= 256 leading table entries are staged per tile

// issue the bulk copy of the column-pass twiddles (one thread); completion is signalled on `bar`
template<int P>
__device__ __forceinline__ void stage_twiddles(Tw *stw, const Tw *tw_limb, uint64_t *bar) {
    constexpr uint32_t bytes = (1u << P) * sizeof(Tw);
    static_assert((1 << P) <= NTT_STW_ENTRIES, "staging area too small");
    mbar_expect_tx(bar, bytes);
    tma_load_1d(stw, tw_limb, bytes, bar);
}

template<class A>
struct PassCtx {
    const Tw *tw;               // column pass: twiddles staged in shared memory; row pass: this limb's table
    uint64_t *bar;              // mbarrier the staging copy completes on (column pass)
    typename A::Consts c;
    int tile;                   // tile index inside the limb (column block resp. row block)
    Tw fin_x, fin_y;            // constants of the last inverse stage
    uint32_t parity = 0;        // phase of `bar` this pass waits for (a persistent CTA reuses the barrier tile after tile)
};

// one forward round on the registers of a thread: x[g * 2^R + k]
template<class A, class M, bool ROWS, int LOGN, int SBASE>
__device__ __forceinline__ void fwd_round(typename A::T (&x)[NTT_EPT], const Tw *tw,
                                          const int (&hi)[M::G], const int (&cc)[M::G], int tile,
                                          const typename A::Consts &c) {
    constexpr int R = M::R;
#pragma unroll
    for (int u = 0; u < R; u++) {
        const int half = 1 << (R - 1 - u);
#pragma unroll
        for (int b = 0; b < (1 << u); b++) {
            Tw w[M::G];
#pragma unroll
            for (int g = 0; g < M::G; g++) {
                if (g == 0 || !M::SHARE) w[g] = ROWS ? __ldg(&tw[tw_index<M, ROWS, LOGN>(u, hi[g], b, cc[g], tile)]) : tw[tw_index<M, ROWS, LOGN>(u, hi[g], b, cc[g], tile)];
                else w[g] = w[0];
            }
#pragma unroll
            for (int g = 0; g < M::G; g++) {
#pragma unroll
                for (int t = 0; t < half; t++) {
                    const int i0 = (g << R) + b * 2 * half + t;
                    // compile-time reduction schedule (all loop variables are unrolled constants)
                    if (fwd_csub(SBASE + M::S0 + u)) A::template fwd<true>(x[i0], x[i0 + half], w[g], c);
                    else A::template fwd<false>(x[i0], x[i0 + half], w[g], c);
                }
            }
        }
    }
}

// one inverse round (stages S0+R-1 down to S0).  FINAL marks the round containing global stage 0.
template<class A, class M, bool ROWS, int LOGN, bool FINAL>
__device__ __forceinline__ void inv_round(typename A::T (&x)[NTT_EPT], const Tw *tw,
                                          const int (&hi)[M::G], const int (&cc)[M::G], int tile,
                                          const typename A::Consts &c, Tw fin_x, Tw fin_y) {
    constexpr int R = M::R;
#pragma unroll
    for (int u = R - 1; u >= 0; u--) {
        const int half = 1 << (R - 1 - u);
#pragma unroll
        for (int b = 0; b < (1 << u); b++) {
            Tw w[M::G];
            if (!(FINAL && u == 0)) {
#pragma unroll
                for (int g = 0; g < M::G; g++) {
                    if (g == 0 || !M::SHARE) w[g] = ROWS ? __ldg(&tw[tw_index<M, ROWS, LOGN>(u, hi[g], b, cc[g], tile)]) : tw[tw_index<M, ROWS, LOGN>(u, hi[g], b, cc[g], tile)];
                    else w[g] = w[0];
                }
            }
#pragma unroll
            for (int g = 0; g < M::G; g++) {
#pragma unroll
                for (int t = 0; t < half; t++) {
                    const int i0 = (g << R) + b * 2 * half + t;
                    if (FINAL && u == 0) A::inv_last(x[i0], x[i0 + half], fin_x, fin_y, c);
                    else A::inv(x[i0], x[i0 + half], w[g], c);
                }
            }
        }
    }
    if constexpr (!FINAL) {
#pragma unroll
        for (int g = 0; g < M::G; g++) A::inv_round_end(x[g << R], c);
    }
}

template<int P, bool ROWS, int LOGN>
__device__ __forceinline__ size_t gl_index(int e, int c, int tile) {
    constexpr int GAM = NTT_LOG_TILE - P;
    if constexpr (!ROWS) {
        constexpr int LOGN2 = LOGN - P;
        return ((size_t) e << LOGN2) + (tile << GAM) + c;
    } else {
        return ((size_t) ((tile << GAM) + c) << P) + e;
    }
}

// ----------------------------------------------------------------------------------------------------
// pass drivers.  Values enter through load.gather(idx[8], x[8]) and leave through store.scatter(idx[8], x[8]):
// the functor sees all eight coefficient indices of a thread at once so that it can issue every global load
// before the first use (memory-level parallelism; a per-element callback with data-dependent work in between
// serialised the loads -- profiles/r1_hmult_full.md, long_scoreboard).  per_elem() adapts a simple lambda.
// ----------------------------------------------------------------------------------------------------
// RUN = length of the runs of consecutive coefficient indices inside idx[] known at compile time: 1 in the column
// passes (strided), 2^R >= 4 at the outer end of a row pass, where a thread owns whole contiguous groups and the
// functors use 16-byte accesses (idx[k] is then even for even k and idx[k + 1] = idx[k] + 1).
template<class T, class F>
struct PerElemLoad {
    F f;
    template<int RUN>
    __device__ __forceinline__ void gather(const size_t (&idx)[NTT_EPT], T (&x)[NTT_EPT]) const {
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) x[k] = f(idx[k]);
    }
};
template<class T, class F>
struct PerElemStore {
    F f;
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const T (&x)[NTT_EPT]) const {
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) f(idx[k], x[k]);
    }
};
__device__ __forceinline__ ulonglong2 ldv2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
__device__ __forceinline__ void stv2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one instruction per run of four coefficients, and -- what
// matters for the stores -- every 32-byte sector is written whole by one instruction (two 16-byte stores per sector
// from different instructions measured 3.5 % slower than staging the tile through shared memory)
struct u64x4 {
    u64 v[4];
};
__device__ __forceinline__ u64x4 ldv4(const u64 *p) {
    u64x4 r;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stv4(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
// run of `RUN` consecutive words starting at p into / out of x[0..RUN)
template<int RUN>
__device__ __forceinline__ void ld_run(const u64 *p, u64 (&x)[RUN]) {
    if constexpr (RUN % 4 == 0) {
#pragma unroll
        for (int k = 0; k < RUN; k += 4) {
            const u64x4 v = ldv4(p + k);
            x[k] = v.v[0], x[k + 1] = v.v[1], x[k + 2] = v.v[2], x[k + 3] = v.v[3];
        }
    } else {
#pragma unroll
        for (int k = 0; k < RUN; k += 2) {
            const ulonglong2 v = ldv2(p + k);
            x[k] = v.x, x[k + 1] = v.y;
        }
    }
}
template<int RUN>
__device__ __forceinline__ void st_run(u64 *p, const u64 (&x)[RUN]) {
    if constexpr (RUN % 4 == 0) {
#pragma unroll
        for (int k = 0; k < RUN; k += 4) stv4(p + k, x[k], x[k + 1], x[k + 2], x[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < RUN; k += 2) stv2(p + k, x[k], x[k + 1]);
    }
}
// all NTT_EPT words of a thread: NTT_EPT / RUN runs starting at idx[0], idx[RUN], ...
template<int RUN>
__device__ __forceinline__ void ld_runs(const u64 *base, const size_t (&idx)[NTT_EPT], u64 (&w)[NTT_EPT]) {
#pragma unroll
    for (int g = 0; g < NTT_EPT; g += RUN) {
        u64 t[RUN];
        ld_run<RUN>(base + idx[g], t);
#pragma unroll
        for (int k = 0; k < RUN; k++) w[g + k] = t[k];
    }
}
template<int RUN>
__device__ __forceinline__ void st_runs(u64 *base, const size_t (&idx)[NTT_EPT], const u64 (&w)[NTT_EPT]) {
#pragma unroll
    for (int g = 0; g < NTT_EPT; g += RUN) {
        u64 t[RUN];
#pragma unroll
        for (int k = 0; k < RUN; k++) t[k] = w[g + k];
        st_run<RUN>(base + idx[g], t);
    }
}
// plain word source / sink with a per-element transform, vectorised over runs
template<class T, class F>
struct VecLoad {
    const u64 *src;
    F f;   // u64 -> T
    template<int RUN>
    __device__ __forceinline__ void gather(const size_t (&idx)[NTT_EPT], T (&x)[NTT_EPT]) const {
        if constexpr (RUN >= 2) {
            u64 w[NTT_EPT];
            ld_runs<RUN>(src, idx, w);
#pragma unroll
            for (int k = 0; k < NTT_EPT; k++) x[k] = f(w[k]);
        } else {
#pragma unroll
            for (int k = 0; k < NTT_EPT; k++) x[k] = f(src[idx[k]]);
        }
    }
};
template<class T, class F>
struct VecStore {
    u64 *dst;
    F f;   // T -> u64
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const T (&x)[NTT_EPT]) const {
        if constexpr (RUN >= 2) {
            u64 w[NTT_EPT];
#pragma unroll
            for (int k = 0; k < NTT_EPT; k++) w[k] = f(x[k]);
            st_runs<RUN>(dst, idx, w);
        } else {
#pragma unroll
            for (int k = 0; k < NTT_EPT; k++) dst[idx[k]] = f(x[k]);
        }
    }
};
template<class T, class F>
__device__ __forceinline__ VecLoad<T, F> vec_load(const u64 *src, F f) { return VecLoad<T, F>{src, f}; }
template<class T, class F>
__device__ __forceinline__ VecStore<T, F> vec_store(u64 *dst, F f) { return VecStore<T, F>{dst, f}; }
template<class T, class F>
__device__ __forceinline__ PerElemLoad<T, F> per_elem_load(F f) { return PerElemLoad<T, F>{f}; }
template<class T, class F>
__device__ __forceinline__ PerElemStore<T, F> per_elem_store(F f) { return PerElemStore<T, F>{f}; }

#ifdef PFHE_TIMELINE
// debug build (make timeline): thread 0 of every CTA of a forward pass leaves cycle stamps of its phases
static __device__ long long g_tl[(size_t) 2 * 65536 * 16];   // per translation unit: only ntt_kernels.cu uses it
#define TL_MARK(slot)                                                                                     \
    if (threadIdx.x == 0) g_tl[tl_base + (slot)] = clock64();
#else
#define TL_MARK(slot)
#endif

template<class A, int P, bool ROWS, int LOGN, int SBASE, class Load, class Store>
__device__ __forceinline__ void forward_pass(u64 *smem, const PassCtx<A> &cx, Load load, Store store) {
    constexpr int NR = Sched<P>::NR;
    typename A::T x[NTT_EPT];
    const int tid = threadIdx.x;
#ifdef PFHE_TIMELINE
    const size_t tl_base = (((size_t) (ROWS ? 1 : 0) << 16) + (size_t) blockIdx.y * gridDim.x + blockIdx.x) * 16;
    if (threadIdx.x == 0) {
        long long gt;
        unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_tl[tl_base + 0] = gt;
        g_tl[tl_base + 1] = sm;
    }
    TL_MARK(2)
#endif

    auto run_round = [&](auto ri_tag) {
        constexpr int RI = decltype(ri_tag)::value;
        using M = RoundMap<P, RI, ROWS>;
        int hi[M::G], lo[M::G], c[M::G], s0[M::G];
#pragma unroll
        for (int g = 0; g < M::G; g++) {
            M::decode(M::mu(tid, g), hi[g], lo[g], c[g]);
            s0[g] = M::sidx0(hi[g], lo[g], c[g]);
        }
        // gather
        if constexpr (RI == 0) {
            size_t gi[NTT_EPT];
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++)
                    gi[(g << M::R) + k] = gl_index<P, ROWS, LOGN>(M::elem(hi[g], k, lo[g]), c[g], cx.tile);
            load.template gather<1>(gi, x);
#ifdef PFHE_TIMELINE
#pragma unroll
            for (int k = 0; k < NTT_EPT; k++) asm volatile("" ::"l"(A::raw(x[k])));   // wait for the loads here
            TL_MARK(3)
#endif
        } else {
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++) x[(g << M::R) + k] = A::from_raw(smem[s0[g] ^ M::kc(k)]);
        }
        if constexpr (RI == 0 && !ROWS) mbar_wait(cx.bar, cx.parity);   // staged twiddles landed (overlapped the gather)
        fwd_round<A, M, ROWS, LOGN, SBASE>(x, cx.tw, hi, c, cx.tile, cx.c);
#ifdef PFHE_TIMELINE
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) asm volatile("" ::"l"(A::raw(x[k])));
        TL_MARK(4 + 2 * RI)
#endif
        // scatter.  The last round leaves the pass straight from the registers: in a column pass the stores are the
        // strided 64-byte segments of the tile, in a row pass a thread owns whole runs of 2^R consecutive coefficients
        // (LAM = 0) and a warp covers one contiguous stretch of a row, so no staging through shared memory is needed
        constexpr bool DIRECT_OUT = (RI == NR - 1);
        // (no barrier before the scatter: a tile position depends on the element alone, not on the round, so a thread
        // overwrites exactly the positions it gathered from)
        if constexpr (DIRECT_OUT) {
            size_t gi[NTT_EPT];
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++)
                    gi[(g << M::R) + k] = gl_index<P, ROWS, LOGN>(M::elem(hi[g], k, lo[g]), c[g], cx.tile);
            store.template scatter<(ROWS ? (1 << M::R) : 1)>(gi, x);
        } else {
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++) smem[s0[g] ^ M::kc(k)] = A::raw(x[(g << M::R) + k]);
            exchange_sync<P, RI, RI + 1, ROWS>();
        }
        TL_MARK(5 + 2 * RI)
    };

    run_round(std::integral_constant<int, 0>{});
    if constexpr (NR > 1) run_round(std::integral_constant<int, 1>{});
    if constexpr (NR > 2) run_round(std::integral_constant<int, 2>{});
    if constexpr (NR > 3) run_round(std::integral_constant<int, 3>{});
    if constexpr (NR > 4) run_round(std::integral_constant<int, 4>{});
    static_assert(NR <= 5, "round schedule too long");

    TL_MARK(12)
}

// Inverse pass over one tile (rounds in reverse order).
template<class A, int P, bool ROWS, int LOGN, bool FINAL, class Load, class Store>
__device__ __forceinline__ void inverse_pass(u64 *smem, const PassCtx<A> &cx, Load load, Store store) {
    constexpr int NR = Sched<P>::NR;
    typename A::T x[NTT_EPT];
    const int tid = threadIdx.x;

    auto run_round = [&](auto ri_tag) {
        constexpr int RI = decltype(ri_tag)::value;
        using M = RoundMap<P, RI, ROWS>;
        int hi[M::G], lo[M::G], c[M::G], s0[M::G];
#pragma unroll
        for (int g = 0; g < M::G; g++) {
            M::decode(M::mu(tid, g), hi[g], lo[g], c[g]);
            s0[g] = M::sidx0(hi[g], lo[g], c[g]);
        }
        constexpr bool DIRECT_IN = (RI == NR - 1);   // rows: runs of 2^R consecutive coefficients per thread
        if constexpr (DIRECT_IN) {
            size_t gi[NTT_EPT];
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++)
                    gi[(g << M::R) + k] = gl_index<P, ROWS, LOGN>(M::elem(hi[g], k, lo[g]), c[g], cx.tile);
            load.template gather<(ROWS ? (1 << M::R) : 1)>(gi, x);
        } else {
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++) x[(g << M::R) + k] = A::from_raw(smem[s0[g] ^ M::kc(k)]);
        }
        if constexpr (RI == NR - 1 && !ROWS) mbar_wait(cx.bar, cx.parity);
        inv_round<A, M, ROWS, LOGN, FINAL && RI == 0>(x, cx.tw, hi, c, cx.tile, cx.c, cx.fin_x, cx.fin_y);
        if constexpr (RI == 0) {
            // first round in index order = last in time: values leave the pass
            size_t gi[NTT_EPT];
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++)
                    gi[(g << M::R) + k] = gl_index<P, ROWS, LOGN>(M::elem(hi[g], k, lo[g]), c[g], cx.tile);
            store.template scatter<1>(gi, x);
        } else {
#pragma unroll
            for (int g = 0; g < M::G; g++)
#pragma unroll
                for (int k = 0; k < (1 << M::R); k++) {
                    const int e = M::elem(hi[g], k, lo[g]);
                    smem[s0[g] ^ M::kc(k)] = A::raw(x[(g << M::R) + k]);
                }
            exchange_sync<P, RI, RI - 1, ROWS>();
        }
    };

    static_assert(NR <= 5, "round schedule too long");
    if constexpr (NR > 4) run_round(std::integral_constant<int, 4>{});
    if constexpr (NR > 3) run_round(std::integral_constant<int, 3>{});
    if constexpr (NR > 2) run_round(std::integral_constant<int, 2>{});
    if constexpr (NR > 1) run_round(std::integral_constant<int, 1>{});
    run_round(std::integral_constant<int, 0>{});
}

} // namespace pfhe
