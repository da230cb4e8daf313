// poly_kernels.cuh -- dyadic, base-conversion and key-switch inner-product kernels.
//
// Layout everywhere: uint64 words, row-major [poly][limb][coeff] (reference include/ciphertext.h:15-25).
// All kernels are HBM/L2-bound integer streams: 128-bit vectorised accesses, one pass over the data,
// constants (moduli, conversion matrices) in registers / shared memory.
#pragma once
#include "engine.hpp"
#include "modarith.cuh"

namespace pfhe {

constexpr int EW_THREADS = 256;

// per-row arithmetic selector shared with the NTT kernels: rows below 2^46 use the FP64 pipe
struct RowArith {
    const unsigned char *is_fp;   // [size_QP]
    const double2 *fpc;           // [size_QP] {q, 1/q}
    u64 mask0, mask1;             // the same flags for rows 0..127 as a by-value bit mask: no dependent load
    int use_mask;
    __device__ __forceinline__ bool fp(int row) const {
        if (use_mask) return ((row < 64 ? mask0 >> row : mask1 >> (row - 64)) & 1) != 0;
        return is_fp[row] != 0;
    }
};

__device__ __forceinline__ ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
__device__ __forceinline__ ulonglong2 ld2_nc(const u64 *p) { return __ldg(reinterpret_cast<const ulonglong2 *>(p)); }
// streaming load: read once, do not keep in L1, first candidate for eviction in L2 (the 80 MiB switching key
// must not push the mod-up digits out of L2 between the NTT that wrote them and the inner product)
__device__ __forceinline__ u64 l2_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ ulonglong2 ld2_stream(const u64 *p, u64 pol) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
                 : "=l"(v.x), "=l"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }

// ---------------------------------------------------------------------------------------------------
// (c0,c1) x (c0',c1') -> (d0,d1,d2): d0 = c0c0', d2 = c1c1', d1 = (c0+c1)(c0'+c1') - d0 - d2
// (tensor_prod_2x2_rns_poly, reference src/polymath.cu:463-498).  out may alias a.
// grid.y = limb, grid.x * EW_THREADS * 2 = n
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) k_tensor_2x2(const u64 *a, const u64 *b, u64 *out, const Modulus *mod,
                                                            const BarG *bar, RowArith ra, size_t n, int l) {
    pdl_launch_dependents();
    pdl_wait();
    const int limb = blockIdx.y;
    const Modulus m = mod[limb];
    const BarG bg = bar[limb];   // growth class 2: (c0+c1)(c0'+c1') < 4 q^2
    const size_t poly = (size_t) l * n;
    const size_t i = (size_t) limb * n + ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 a0 = ld2(a + i), a1 = ld2(a + i + poly), b0 = ld2(b + i), b1 = ld2(b + i + poly);
    u64 d0[2], d1[2], d2[2];
    const u64 A0[2] = {a0.x, a0.y}, A1[2] = {a1.x, a1.y}, B0[2] = {b0.x, b0.y}, B1[2] = {b1.x, b1.y};
    if (ra.fp(limb)) {   // CTA-uniform: small limb, FP64 pipe
        const double q = ra.fpc[limb].x, qi = ra.fpc[limb].y;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double x0 = fp::from_u64(A0[k]), x1 = fp::from_u64(A1[k]);
            const double y0 = fp::from_u64(B0[k]), y1 = fp::from_u64(B1[k]);
            const double e0 = fp::mulmod_v(x0, y0, q, qi), e2 = fp::mulmod_v(x1, y1, q, qi);
            const double t = fp::mulmod_v(x0 + x1, y0 + y1, q, qi);
            d0[k] = fp::canon(e0, q);
            d2[k] = fp::canon(e2, q);
            d1[k] = fp::canon(fp::reduce(t - e0 - e2, q, qi), q);
        }
        st2(out + i, d0[0], d0[1]);
        st2(out + i + poly, d1[0], d1[1]);
        st2(out + i + 2 * poly, d2[0], d2[1]);
        return;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        d0[k] = mul_mod_g(A0[k], B0[k], bg, m);
        d2[k] = mul_mod_g(A1[k], B1[k], bg, m);
        const u64 t = mul_mod_g(A0[k] + A1[k], B0[k] + B1[k], bg, m);   // sums < 2q < 2^62: product < 2^124
        d1[k] = sub_mod(sub_mod(t, d0[k], m.q), d2[k], m.q);
    }
    st2(out + i, d0[0], d0[1]);
    st2(out + i + poly, d1[0], d1[1]);
    st2(out + i + 2 * poly, d2[0], d2[1]);
}

// tensor_prod_mxn_rns_poly (reference src/polymath.cu:546-594): ciphertexts of sizes sa x sb (not both 2),
// out[j] = sum_{i1 + i2 = j} a[i1] * b[i2], each sum accumulated in 128 bits and reduced once.  out may alias a:
// a thread reads every operand word of its coefficient before it writes.  One coefficient per thread (the
// operand sets live in registers; the reference allocates them with device-side new), grid.y = limb.
constexpr int MXN_MAX = 8;
__global__ void __launch_bounds__(EW_THREADS) k_tensor_mxn(const u64 *a, int sa, const u64 *b, int sb, u64 *out,
                                                            const Modulus *mod, size_t n, int l) {
    pdl_launch_dependents();
    pdl_wait();
    const int limb = blockIdx.y;
    const Modulus m = mod[limb];
    const size_t poly = (size_t) l * n;
    const size_t k = (size_t) limb * n + (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    u64 c1[MXN_MAX], c2[MXN_MAX], r[2 * MXN_MAX - 1];
#pragma unroll
    for (int u = 0; u < MXN_MAX; u++) {
        c1[u] = u < sa ? a[k + (size_t) u * poly] : 0;
        c2[u] = u < sb ? b[k + (size_t) u * poly] : 0;
    }
    // zero padding makes every anti-diagonal a fixed-shape sum: the extra terms are 0
#pragma unroll
    for (int j = 0; j < 2 * MXN_MAX - 1; j++) {
        Acc128 acc{0, 0};
#pragma unroll
        for (int u = 0; u < MXN_MAX; u++) {
            const int v = j - u;
            if (v >= 0 && v < MXN_MAX) acc.mac(c1[u], c2[v]);
        }
        r[j] = barrett128(acc.lo, acc.hi, m);
    }
    const int so = sa + sb - 1;
#pragma unroll
    for (int j = 0; j < 2 * MXN_MAX - 1; j++)
        if (j < so) out[k + (size_t) j * poly] = r[j];
}

// tensor_square_2x2_rns_poly (src/polymath.cu:500-532)
__global__ void __launch_bounds__(EW_THREADS) k_tensor_square(const u64 *a, u64 *out, const Modulus *mod,
                                                               const BarG *bar, size_t n, int l) {
    pdl_launch_dependents();
    pdl_wait();
    const int limb = blockIdx.y;
    const Modulus m = mod[limb];
    const BarG bg = bar[limb];
    const size_t poly = (size_t) l * n;
    const size_t i = (size_t) limb * n + ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 a0 = ld2(a + i), a1 = ld2(a + i + poly);
    const u64 A0[2] = {a0.x, a0.y}, A1[2] = {a1.x, a1.y};
    u64 d0[2], d1[2], d2[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        d0[k] = mul_mod_g(A0[k], A0[k], bg, m);
        const u64 t = mul_mod_g(A0[k], A1[k], bg, m);
        d1[k] = add_mod(t, t, m.q);
        d2[k] = mul_mod_g(A1[k], A1[k], bg, m);
    }
    st2(out + i, d0[0], d0[1]);
    st2(out + i + poly, d1[0], d1[1]);
    st2(out + i + 2 * poly, d2[0], d2[1]);
}

// generic two-operand limb-wise op (add_rns_poly / sub_rns_poly / multiply_rns_poly, polymath.cu:41-173)
template<int OP>
__global__ void __launch_bounds__(EW_THREADS) k_elementwise(const u64 *a, const u64 *b, u64 *out, const Modulus *mod,
                                                             const BarG *bar, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int limb = blockIdx.y;
    const Modulus m = mod[limb];
    const BarG bg = bar[limb];
    const size_t i = (size_t) limb * n + ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 x = ld2(a + i);
    ulonglong2 y = make_ulonglong2(0, 0);
    if (OP != EW_NEG) y = ld2(b + i);
    u64 r0, r1;
    if (OP == EW_ADD) r0 = add_mod(x.x, y.x, m.q), r1 = add_mod(x.y, y.y, m.q);
    else if (OP == EW_SUB) r0 = sub_mod(x.x, y.x, m.q), r1 = sub_mod(x.y, y.y, m.q);
    else if (OP == EW_MUL) r0 = mul_mod_g(x.x, y.x, bg, m), r1 = mul_mod_g(x.y, y.y, bg, m);
    else r0 = x.x ? m.q - x.x : 0, r1 = x.y ? m.q - x.y : 0;
    st2(out + i, r0, r1);
}

// ---------------------------------------------------------------------------------------------------
// Fast base conversion, matrix phase (bconv_matmul_*_kernel, reference src/rns_bconv.cu:109-210,455-485):
//   out[j][x] = sum_i y[i][x] * M[j][i]  mod p_j,   y already scaled by qhat_i^-1 (folded into the iNTT).
// One thread owns two adjacent coefficients, loads the NI inputs ONCE and produces every output limb
// (the reference re-reads all inputs per output limb).  Matrix + output moduli staged in shared memory.
// out_limb[j] gives the destination limb (units of n) so the mod-up "leap over own digit" layout and the
// plain P->Ql layout share this kernel.  grid.y = problem instance (digit or polynomial).
// ---------------------------------------------------------------------------------------------------
struct BconvJob {
    const u64 *in;        // [ni][n], limb stride n
    u64 *out;             // base pointer of the output limbs
    const u64 *mat;       // [no][ni]
    const short *omod;    // [no] key-level prime row of each output limb
    const short *olimb;   // [no] output limb index (units of n from `out`)
    int ni, no;
    int xbits;            // bit length bound of the accumulated sum minus the output modulus' own bits:
                          // max input-prime bits + ceil(log2 ni); selects the Barrett class per output limb
    const double2 *matf;  // [no][ni][2] FP64 form for outputs below 2^46: {M, M/p} and {M 2^30 mod p, that/p}
    unsigned in_big;      // bit i set: input limb i comes from a modulus >= 2^46 and enters as two 30-bit halves
};
constexpr int BCONV_MAX_IN = 8;   // alpha <= 8 per digit on this path (larger digits use the generic loop)
constexpr int BCONV_MAX_JOBS = 16;
struct BconvBatch {
    BconvJob job[BCONV_MAX_JOBS];
};

template<int NI>
__global__ void __launch_bounds__(EW_THREADS) k_bconv(BconvBatch batch, const Modulus *mod, const BarG *bar,
                                                       RowArith ra, int size_QP, size_t n) {
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char s_raw[];
    const BconvJob jb = batch.job[blockIdx.y];
    const int ni = NI > 0 ? NI : jb.ni;
    // shared (16-byte types first): [matf: no*ni*2 double2][fpc: no double2][bar: no][mat: no*ni u64][mod: no][fp flag: no]
    double2 *s_matf = reinterpret_cast<double2 *>(s_raw);
    double2 *s_fpc = s_matf + jb.no * ni * 2;
    BarG *s_bar = reinterpret_cast<BarG *>(s_fpc + jb.no);
    u64 *s_mat = reinterpret_cast<u64 *>(s_bar + jb.no);
    Modulus *s_mod = reinterpret_cast<Modulus *>(s_mat + jb.no * ni);
    int *s_isfp = reinterpret_cast<int *>(s_mod + jb.no);
    for (int i = threadIdx.x; i < jb.no * ni; i += blockDim.x) {
        s_mat[i] = jb.mat[i];
        s_matf[2 * i] = jb.matf[2 * i];
        s_matf[2 * i + 1] = jb.matf[2 * i + 1];
    }
    for (int i = threadIdx.x; i < jb.no; i += blockDim.x) {
        const int row = jb.omod[i];
        const Modulus mo = mod[row];
        s_mod[i] = mo;
        // sum < 2^(xbits + k_out): class = xbits - k_out above the same-modulus case 2 k_out
        const int cls = max(0, jb.xbits - (64 - __clzll((long long) mo.q)));
        s_bar[i] = bar[(size_t) min(cls, 63) * size_QP + row];
        s_fpc[i] = ra.fpc[row];
        s_isfp[i] = ra.fp(row);
    }
    pdl_wait();
    __syncthreads();
    const size_t x = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    if (NI > 0) {
        constexpr int NN = NI > 0 ? NI : 1;
        u64 y0[NN], y1[NN];
        double l0[NN], h0[NN], l1[NN], h1[NN];
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const ulonglong2 v = ld2(jb.in + (size_t) i * n + x);
            y0[i] = v.x, y1[i] = v.y;
            if ((jb.in_big >> i) & 1) {
                const u64 msk = (1ull << fp::SPLIT_BITS) - 1;
                l0[i] = fp::from_u64(v.x & msk), h0[i] = fp::from_u64(v.x >> fp::SPLIT_BITS);
                l1[i] = fp::from_u64(v.y & msk), h1[i] = fp::from_u64(v.y >> fp::SPLIT_BITS);
            } else {
                l0[i] = fp::from_u64(v.x), l1[i] = fp::from_u64(v.y);
                h0[i] = h1[i] = 0.0;
            }
        }
        for (int j = 0; j < jb.no; j++) {
            u64 r0, r1;
            if (s_isfp[j]) {   // CTA-uniform per output limb
                const double q = s_fpc[j].x, qi = s_fpc[j].y;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int i = 0; i < NI; i++) {
                    const double2 m0 = s_matf[2 * (j * NI + i)];
                    a0 += fp::mulmod_c(l0[i], m0.x, m0.y, q);
                    a1 += fp::mulmod_c(l1[i], m0.x, m0.y, q);
                    if ((jb.in_big >> i) & 1) {
                        const double2 m1 = s_matf[2 * (j * NI + i) + 1];
                        a0 += fp::mulmod_c(h0[i], m1.x, m1.y, q);
                        a1 += fp::mulmod_c(h1[i], m1.x, m1.y, q);
                    }
                }
                r0 = fp::canon(fp::reduce(a0, q, qi), q);
                r1 = fp::canon(fp::reduce(a1, q, qi), q);
            } else {
                Acc128 a0{0, 0}, a1{0, 0};
#pragma unroll
                for (int i = 0; i < NI; i++) {
                    const u64 mji = s_mat[j * NI + i];
                    a0.mac(y0[i], mji);
                    a1.mac(y1[i], mji);
                }
                const Modulus m = s_mod[j];
                const BarG bg = s_bar[j];
                r0 = barrett_g(a0.lo, a0.hi, bg, m);
                r1 = barrett_g(a1.lo, a1.hi, bg, m);
            }
            st2(jb.out + (size_t) jb.olimb[j] * n + x, r0, r1);
        }
    } else {
        for (int j = 0; j < jb.no; j++) {
            Acc128 a0{0, 0}, a1{0, 0};
            for (int i = 0; i < ni; i++) {
                const ulonglong2 v = ld2(jb.in + (size_t) i * n + x);
                const u64 mji = s_mat[j * ni + i];
                a0.mac(v.x, mji);
                a1.mac(v.y, mji);
            }
            const Modulus m = s_mod[j];
            const BarG bg = s_bar[j];
            st2(jb.out + (size_t) jb.olimb[j] * n + x, barrett_g(a0.lo, a0.hi, bg, m), barrett_g(a1.lo, a1.hi, bg, m));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// key-switch inner product (key_switch_inner_prod_c2_and_evk, reference src/eval_key_switch.cu:14-69):
//   cx[k][j][x] = sum_{d<beta} t[d][j][x] * evk[d][k][row(j)][x] mod p_row(j),  k = 0,1
// t = [beta][m][n]; evk = beta device pointers to [2][size_QP][n]; row(j) = j < l ? j : size_Q + (j - l).
// Streams the key exactly once; 128-bit accumulation, one Barrett reduction per output.
// grid.y = j, 2 coefficients per thread.
// ---------------------------------------------------------------------------------------------------
constexpr int KS_MAX_BETA = 64;
// where the digit's own limbs come from when mod-up no longer copies them into t_mod_up (fused pipeline):
// c2 in NTT form, or -- HMult fused -- the product a1 * b1 formed on the fly.  alpha = 0: everything is in t.
struct OwnSrc {
    const u64 *c2;
    const u64 *a1, *b1;
    int alpha;
};
struct InnerProdArgs {
    u64 *cx;
    const u64 *t;
    const u64 *const *evk;
    const Modulus *mod;
    const BarG *bar, *bar0;
    RowArith ra;
    OwnSrc os;
    const uint32_t *perm;
    int accumulate;
    int t_lazy;   // FP64 limbs of t hold lazy FP64 words (ntt_forward_bconv phase 3) instead of canonical residues
    size_t n;
    int l, m, size_Q, size_QP, beta;
    int j0, j_count;   // limbs [j0, j0 + j_count) of cx
};

// One tile: limb j of cx (both polynomials), IP_PAIRS * 2 * EW_THREADS consecutive coefficients; a thread owns
// IP_PAIRS coefficient pairs EW_THREADS pairs apart (every warp access is one contiguous 512-byte run).
// BETA > 0 fixes the digit count at compile time (full unroll, key pointers in registers); PLAIN = no Galois
// permutation and no accumulation into cx (the key-switch proper; hoisting takes the general form).  Everything that
// depends only on (j, launch) is set up once per tile, outside the pair loop: with two coefficients per thread and a
// run-time digit loop the set-up and addressing were 3/4 of the instructions (profiles/r1b_inner_prod_mix.md).
#ifndef PFHE_T_EVICT_FIRST
#define PFHE_T_EVICT_FIRST 1
#endif
constexpr int IP_PAIRS = 2;
constexpr int IP_TILE = IP_PAIRS * 2 * EW_THREADS;

template<int BETA, bool PLAIN, int PAIRS>
__device__ __forceinline__ void inner_prod_tile(const InnerProdArgs &A, const int j, const unsigned bx) {
    const int beta = BETA > 0 ? BETA : A.beta;
    const size_t n = A.n;
    const int l = A.l, m = A.m;
    const int row = j < l ? j : A.size_Q + (j - l);
    const size_t m_n = (size_t) m * n, qp_n = (size_t) A.size_QP * n;
    const OwnSrc &os = A.os;
    const int own_d = (os.alpha > 0 && j < l) ? j / os.alpha : -1;   // digit whose own limb this is
    const bool own_mul = own_d >= 0 && !os.c2;
    const uint32_t *perm = PLAIN ? nullptr : A.perm;
    const bool accumulate = PLAIN ? false : A.accumulate != 0;
    const u64 *tj = A.t + (size_t) j * n;
    const u64 *ownp = own_d < 0 ? nullptr : (os.c2 ? os.c2 : os.a1) + (size_t) j * n;
    const u64 *ownq = own_mul ? os.b1 + (size_t) j * n : nullptr;
    u64 *c0 = A.cx + (size_t) j * n, *c1 = c0 + m_n;
    const u64 pol = l2_evict_first_policy();
    // key rows of this limb, one pointer per digit (BETA > 0: registers)
    const u64 *kp[BETA > 0 ? BETA : 1];
    if constexpr (BETA > 0) {
#pragma unroll
        for (int d = 0; d < BETA; d++) kp[d] = A.evk[d] + (size_t) row * n;
    }
    const size_t x_base = ((size_t) bx * PAIRS * EW_THREADS + threadIdx.x) * 2;
    const bool is_fp = A.ra.fp(row);   // CTA-uniform
    double q = 0, qi = 0;
    Modulus md{};
    BarG bg{}, b0{};
    if (is_fp) q = A.ra.fpc[row].x, qi = A.ra.fpc[row].y;
    else md = A.mod[row], bg = A.bar[row], b0 = A.bar0[row];

#pragma unroll 1
    for (int ip = 0; ip < PAIRS; ip++) {
        const size_t x = x_base + (size_t) ip * 2 * EW_THREADS;
        size_t px0 = x, px1 = x + 1;
        if (perm) {   // hoisting (reference src/evaluate.cu:1775-1835): digits are read through the Galois permutation
            const uint2 pp = *reinterpret_cast<const uint2 *>(perm + x);
            px0 = pp.x, px1 = pp.y;
        }
        auto load_t = [&](int d) -> ulonglong2 {
            const u64 *base = tj + (size_t) d * m_n;
            if (perm) return make_ulonglong2(base[px0], base[px1]);
#if PFHE_T_EVICT_FIRST
            return ld2_stream(base + x, pol);   // last use of the mod-up digits: do not keep them in L2
#else
            return ld2(base + x);
#endif
        };
        auto key = [&](int d) -> const u64 * {
            if constexpr (BETA > 0) return kp[d] + x;
            else return A.evk[d] + (size_t) row * n + x;
        };
        ulonglong2 own = make_ulonglong2(0, 0), ownb = make_ulonglong2(0, 0);
        if (own_d >= 0) own = ld2(ownp + x);
        if (own_mul) ownb = ld2(ownq + x);
        ulonglong2 old0 = make_ulonglong2(0, 0), old1 = make_ulonglong2(0, 0);
        if (accumulate) old0 = ld2(c0 + x), old1 = ld2(c1 + x);
        constexpr int CH = 2;   // digits per batch: 6 independent 16-byte loads in flight per thread
        if (is_fp) {   // every term is reduced on the FP64 pipe, the small residues are summed
            double ox = fp::from_u64(own.x), oy = fp::from_u64(own.y);
            if (own_mul) {
                ox = fp::mulmod_v(ox, fp::from_u64(ownb.x), q, qi);
                oy = fp::mulmod_v(oy, fp::from_u64(ownb.y), q, qi);
            }
            double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
            for (int d0 = 0; d0 < beta; d0 += CH) {
                ulonglong2 v[CH], e0[CH], e1[CH];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int d = min(d0 + c, beta - 1);   // clamp: a tail lane re-reads the last digit, masked below
                    const u64 *k0 = key(d);
                    v[c] = d != own_d ? load_t(d) : own;
                    e0[c] = ld2_stream(k0, pol);
                    e1[c] = ld2_stream(k0 + qp_n, pol);
                }
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int d = d0 + c;
                    if (d < beta) {
                        const double vx = d == own_d ? ox : (A.t_lazy ? __longlong_as_double((long long) v[c].x) : fp::from_u64(v[c].x));
                        const double vy = d == own_d ? oy : (A.t_lazy ? __longlong_as_double((long long) v[c].y) : fp::from_u64(v[c].y));
                        s00 += fp::mulmod_v(vx, fp::from_u64(e0[c].x), q, qi);
                        s01 += fp::mulmod_v(vy, fp::from_u64(e0[c].y), q, qi);
                        s10 += fp::mulmod_v(vx, fp::from_u64(e1[c].x), q, qi);
                        s11 += fp::mulmod_v(vy, fp::from_u64(e1[c].y), q, qi);
                    }
                }
            }
            if (accumulate) {
                s00 += fp::from_u64(old0.x), s01 += fp::from_u64(old0.y);
                s10 += fp::from_u64(old1.x), s11 += fp::from_u64(old1.y);
            }
            st2(c0 + x, fp::canon(fp::reduce(s00, q, qi), q), fp::canon(fp::reduce(s01, q, qi), q));
            st2(c1 + x, fp::canon(fp::reduce(s10, q, qi), q), fp::canon(fp::reduce(s11, q, qi), q));
        } else {
            if (own_mul) own = make_ulonglong2(mul_mod_g(own.x, ownb.x, b0, md), mul_mod_g(own.y, ownb.y, b0, md));
            Acc128 a00{0, 0}, a01{0, 0}, a10{0, 0}, a11{0, 0};
#pragma unroll
            for (int d0 = 0; d0 < beta; d0 += CH) {
                ulonglong2 v[CH], e0[CH], e1[CH];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int d = min(d0 + c, beta - 1);
                    const u64 *k0 = key(d);
                    v[c] = d != own_d ? load_t(d) : own;
                    e0[c] = ld2_stream(k0, pol);
                    e1[c] = ld2_stream(k0 + qp_n, pol);
                }
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int d = d0 + c;
                    if (d < beta) {
                        const ulonglong2 w = d == own_d ? own : v[c];
                        a00.mac(w.x, e0[c].x);
                        a01.mac(w.y, e0[c].y);
                        a10.mac(w.x, e1[c].x);
                        a11.mac(w.y, e1[c].y);
                    }
                }
            }
            u64 r00 = barrett_g(a00.lo, a00.hi, bg, md), r01 = barrett_g(a01.lo, a01.hi, bg, md);
            u64 r10 = barrett_g(a10.lo, a10.hi, bg, md), r11 = barrett_g(a11.lo, a11.hi, bg, md);
            if (accumulate) {
                r00 = add_mod(r00, old0.x, md.q), r01 = add_mod(r01, old0.y, md.q);
                r10 = add_mod(r10, old1.x, md.q), r11 = add_mod(r11, old1.y, md.q);
            }
            st2(c0 + x, r00, r01);
            st2(c1 + x, r10, r11);
        }
    }
}

// grid-stride over the tiles (limb-major); a grid of all tiles runs the loop once, a smaller grid ("persist": leaves
// room on every SM for the higher-priority mod-down chain of the fused key switch) walks them
// PAIRS: coefficient pairs per thread.  2 for the streaming launches; 1 for the few-limb launch on the critical path of the
// key switch (the P limbs: twice the CTAs, half the serial work per thread)
template<int BETA, bool PLAIN, int PAIRS = IP_PAIRS>
__global__ void __launch_bounds__(EW_THREADS, 4) k_inner_prod(const InnerProdArgs A) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned nbx = (unsigned) (A.n / (PAIRS * 2 * EW_THREADS));   // power of two
    const unsigned sh = 31u - (unsigned) __clz(nbx);
    const unsigned total = nbx * (unsigned) A.j_count;
    for (unsigned tile = blockIdx.x; tile < total; tile += gridDim.x)
        inner_prod_tile<BETA, PLAIN, PAIRS>(A, A.j0 + (int) (tile >> sh), tile & (nbx - 1));
}

// ---------------------------------------------------------------------------------------------------
// coefficient-domain pieces of the BFV / BGV key switch and modulus switch (integer Shoup constants)
// ---------------------------------------------------------------------------------------------------
// bconv_mult_kernel (reference src/rns_bconv.cu:22-31): dst[i][x] = src[i][x] * c[i] mod q_i; grid.y = limb
__global__ void __launch_bounds__(EW_THREADS) k_scale_limbs(u64 *dst, const u64 *src, const Tw *c, const Modulus *mod,
                                                             size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.y;
    const u64 q = mod[i].q;
    const Tw k = c[i];
    const size_t x = (size_t) i * n + ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 v = ld2(src + x);
    st2(dst + x, mul_shoup(v.x, k, q), mul_shoup(v.y, k, q));
}

// moddown_kernel (src/rns_bconv.cu:680-689) + add_to_ct_kernel (:763-769): out = (cx - delta) * P^-1 (+ add)
__global__ void __launch_bounds__(EW_THREADS) k_moddown_coeff(u64 *out, const u64 *cx, const u64 *delta, const Tw *pinv,
                                                               const u64 *add, const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const u64 q = mod[j].q;
    const Tw k = pinv[j];
    const size_t x = (size_t) j * n + ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 c = ld2(cx + x), d = ld2(delta + x);
    u64 r0 = mul_shoup(sub_mod(c.x, d.x, q), k, q), r1 = mul_shoup(sub_mod(c.y, d.y, q), k, q);
    if (add) {
        const ulonglong2 a = ld2(add + x);
        r0 = add_mod(r0, a.x, q), r1 = add_mod(r1, a.y, q);
    }
    st2(out + x, r0, r1);
}

// bgv_moddown_kernel (src/rns_bconv.cu:636-652): dst = ((cx - delta) + [cp_t * P^-1]_t * P) * P^-1 mod q_j
__global__ void __launch_bounds__(EW_THREADS) k_bgv_moddown(u64 *dst, const u64 *cx, const u64 *delta, const u64 *cp_t,
                                                             const Tw *P_mod_q, const Tw *pinv, Tw pinv_t, u64 t,
                                                             const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const u64 q = mod[j].q;
    const size_t xo = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const size_t x = (size_t) j * n + xo;
    const ulonglong2 c = ld2(cx + x), d = ld2(delta + x), ct = ld2(cp_t + xo);
    const Tw pq = P_mod_q[j], pi = pinv[j];
    u64 r[2];
    const u64 cc[2] = {c.x, c.y}, dd[2] = {d.x, d.y}, tt[2] = {ct.x, ct.y};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const u64 tmp = mul_shoup(tt[k], pinv_t, t);
        const u64 corr = mul_shoup(tmp, pq, q);
        r[k] = mul_shoup(add_mod(sub_mod(cc[k], dd[k], q), corr, q), pi, q);
    }
    st2(dst + x, r[0], r[1]);
}

// BFV mod-down in one pass over the output limbs (bConv_BEHZ matmul of the P limbs, rns_bconv.cu:143-168 / :691-707, +
// moddown_kernel :680-689 + add_to_ct :763-769): out[k][j] = (cx[k][j] - sum_i cx[k][l + i] * (phat_i mod q_j)) * P^-1 (+ add).
// cx = [npoly][m][n] in coefficient form, its P limbs already scaled by phat_i^-1.  grid = (n / (2 EW_THREADS), l, npoly)
__global__ void __launch_bounds__(EW_THREADS) k_moddown_coeff_conv(u64 *out, const u64 *cx, const u64 *mat, int alpha, int m,
                                                                    const Tw *pinv, const u64 *add, unsigned add_mask,
                                                                    const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y, k = blockIdx.z, l = gridDim.y;
    const Modulus mj = mod[j];
    const size_t xo = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const u64 *c = cx + (size_t) k * m * n;
    Acc128 a0{0, 0}, a1{0, 0};
    for (int i = 0; i < alpha; i++) {
        const ulonglong2 y = ld2(c + (size_t) (l + i) * n + xo);
        const u64 M = mat[(size_t) j * alpha + i];
        a0.mac(y.x, M), a1.mac(y.y, M);
    }
    const u64 d0 = barrett128(a0.lo, a0.hi, mj), d1 = barrett128(a1.lo, a1.hi, mj);
    const ulonglong2 v = ld2(c + (size_t) j * n + xo);
    const Tw pi = pinv[j];
    u64 r0 = mul_shoup(sub_mod(v.x, d0, mj.q), pi, mj.q), r1 = mul_shoup(sub_mod(v.y, d1, mj.q), pi, mj.q);
    const size_t o = ((size_t) k * l + j) * n + xo;
    if (add && ((add_mask >> k) & 1)) {
        const ulonglong2 a = ld2(add + o);
        r0 = add_mod(r0, a.x, mj.q), r1 = add_mod(r1, a.y, mj.q);
    }
    st2(out + o, r0, r1);
}

// plain-modulus correction of the BGV mod-down (bgv_moddown_kernel, src/rns_bconv.cu:636-652, with base_P_to_t_conv):
// buf = [npoly][alpha + 1][n]; limb alpha <- ((sum_k buf[k] * (phat_k mod t)) mod t) * P^-1 mod t.  The inputs already
// carry the phat_k^-1 scaling (folded into the inverse transform).  grid = (n / EW_THREADS, npoly)
__global__ void __launch_bounds__(EW_THREADS) k_bgv_corr(u64 *buf, const u64 *mat_t, int alpha, Tw pinv_t, Modulus mt, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t x = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    u64 *b = buf + (size_t) blockIdx.y * (alpha + 1) * n + x;
    Acc128 acc{0, 0};
    for (int k = 0; k < alpha; k++) acc.mac(b[(size_t) k * n], mat_t[k]);
    const u64 dt = barrett128(acc.lo, acc.hi, mt);
    b[(size_t) alpha * n] = mul_shoup(dt, pinv_t, mt.q);
}

// divide_and_round_q_last_kernel (src/rns.cu:1082-1108), BFV modulus switch in the coefficient domain:
// dst[j] = (src[j] - (src[last] mod q_j)) * q_last^-1 mod q_j
__global__ void __launch_bounds__(EW_THREADS) k_divide_round_last(u64 *dst, const u64 *src, const Tw *qlast_inv,
                                                                   const Modulus *mod, size_t n, int nl) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const Modulus m = mod[j];
    const size_t xo = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 last = ld2(src + (size_t) nl * n + xo), c = ld2(src + (size_t) j * n + xo);
    const Tw k = qlast_inv[j];
    st2(dst + (size_t) j * n + xo, mul_shoup(sub_mod(c.x, barrett64(last.x, m), m.q), k, m.q),
        mul_shoup(sub_mod(c.y, barrett64(last.y, m), m.q), k, m.q));
}

// bgv_mod_t_divide_q_kernel (src/rns.cu:1186-1207)
__global__ void __launch_bounds__(EW_THREADS) k_bgv_mod_t_divide(u64 *dst, const u64 *cx, const u64 *ci_last,
                                                                  const Tw *qlast_mod_q, const Tw *qlast_inv,
                                                                  Tw inv_qlast_t, Modulus tm, const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const Modulus m = mod[j];
    const size_t xo = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const size_t x = (size_t) j * n + xo;
    const ulonglong2 last = ld2(ci_last + xo), c = ld2(cx + x);
    const Tw qm = qlast_mod_q[j], qi = qlast_inv[j];
    const u64 ll[2] = {last.x, last.y}, cc[2] = {c.x, c.y};
    u64 r[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const u64 delta = barrett64(ll[k], m);
        const u64 tmp = mul_shoup(barrett64(ll[k], tm), inv_qlast_t, tm.q);
        const u64 corr = mul_shoup(tmp, qm, m.q);
        r[k] = mul_shoup(add_mod(sub_mod(cc[k], delta, m.q), corr, m.q), qi, m.q);
    }
    st2(dst + x, r[0], r[1]);
}

// ---------------------------------------------------------------------------------------------------
// CKKS encoder, encoding direction (PhantomCKKSEncoder::encode_internal, reference src/ckks.cu:66-135): bit-reversed
// placement, special inverse FFT over the slots, rounding and RNS decomposition.  The floating-point operations are the
// ones the reference's compiled kernels execute (Gentleman-Sande butterfly (x0 + x1, (x0 - x1) w) with
// re = fma(d.x, w.x, -(d.y w.y)), im = fma(d.x, w.y, d.y w.x); then * scalar; then round()), so the residues are the same
// words; only the schedule differs: the stages whose butterflies stay inside a block of CKKS_FFT_BLOCK slots run in one
// shared-memory kernel, the rest one launch per stage.
// ---------------------------------------------------------------------------------------------------
constexpr int CKKS_FFT_LOG_BLOCK = 11;
constexpr int CKKS_FFT_BLOCK = 1 << CKKS_FFT_LOG_BLOCK;

__global__ void __launch_bounds__(EW_THREADS) k_ckks_place(double2 *x, const double2 *values, size_t count, int logs) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t r = blockIdx.x * EW_THREADS + threadIdx.x;   // output position
    const uint32_t i = logs ? (__brev(r) >> (32 - logs)) : 0;    // bit_reverse_kernel, ckks.cu:9-15
    x[r] = i < count ? values[i] : make_double2(0.0, 0.0);
}

// one inverse butterfly on the pair (a, a + pairs).  nvcc contracted the imaginary part of the complex product
// d.x w.y + d.y w.x differently in the reference's two kernels (SASS of oracle/_ref): the shared-memory kernel (stages
// with pairs <= 1024) computes fma(d.y, w.x, d.x * w.y), the one-stage kernel fma(d.x, w.y, d.y * w.x); the two round
// differently in about a quarter of the butterflies, which flips about one encoded coefficient in 10^4, so the form is
// part of the result.  BLOCK_FORM selects the first.
template<bool BLOCK_FORM>
__device__ __forceinline__ void ckks_gs(double2 &x0, double2 &x1, const double2 w) {
    const double sx = __dadd_rn(x0.x, x1.x), sy = __dadd_rn(x0.y, x1.y);
    const double dx = __dadd_rn(x0.x, -x1.x), dy = __dadd_rn(x0.y, -x1.y);
    const double t1 = __dmul_rn(dy, w.y);
    const double re = __fma_rn(dx, w.x, -t1);
    const double im = BLOCK_FORM ? __fma_rn(dy, w.x, __dmul_rn(dx, w.y)) : __fma_rn(dx, w.y, __dmul_rn(dy, w.x));
    x0 = make_double2(sx, sy);
    x1 = make_double2(re, im);
}
__device__ __forceinline__ double2 ckks_tw(const double2 *tw, const uint32_t *group, uint32_t k, int logPairs, int logs,
                                           uint32_t M) {
    // psiIdx = group[brev(k << logPairs) >> (33 - logs)] << logPairs mod M; inverse transform: twiddles[M - psiIdx]
    uint32_t psi = group[__brev(k << logPairs) >> (33 - logs)];
    psi = (psi << logPairs) & (M - 1);
    return tw[M - psi];
}

// stages iter = logs-1 down to iter_end (pairs 1 .. BLOCK/2) inside blocks of CKKS_FFT_BLOCK (or all slots) elements
__global__ void __launch_bounds__(CKKS_FFT_BLOCK / 2) k_ckks_ifft_block(double2 *x, const double2 *tw, const uint32_t *group,
                                                                        int logs, int iter_end, uint32_t M, double scalar) {
    extern __shared__ __align__(16) unsigned char ckks_smem[];
    double2 *buf = reinterpret_cast<double2 *>(ckks_smem);
    pdl_launch_dependents();
    pdl_wait();
    const int nb = blockDim.x * 2;                         // slots per block
    const uint32_t base = blockIdx.x * nb, t = threadIdx.x;
    buf[t] = x[base + t], buf[t + blockDim.x] = x[base + t + blockDim.x];
    __syncthreads();
    for (int iter = logs - 1; iter >= iter_end; iter--) {
        const int logPairs = logs - iter - 1;
        const uint32_t pairs = 1u << logPairs;
        const uint32_t gt = blockIdx.x * blockDim.x + t;   // global butterfly index of this stage
        const uint32_t k = gt >> logPairs, j = gt & (pairs - 1);
        const uint32_t a = 2 * k * pairs + j - base;
        double2 x0 = buf[a], x1 = buf[a + pairs];
        ckks_gs<true>(x0, x1, ckks_tw(tw, group, k, logPairs, logs, M));
        if (iter == 0 && scalar != 0.0) {
            x0 = make_double2(__dmul_rn(x0.x, scalar), __dmul_rn(x0.y, scalar));
            x1 = make_double2(__dmul_rn(x1.x, scalar), __dmul_rn(x1.y, scalar));
        }
        __syncthreads();
        buf[a] = x0, buf[a + pairs] = x1;
        __syncthreads();
    }
    x[base + t] = buf[t], x[base + t + blockDim.x] = buf[t + blockDim.x];
}

// one stage in global memory (pairs >= CKKS_FFT_BLOCK / 2)
__global__ void __launch_bounds__(EW_THREADS) k_ckks_ifft_stage(double2 *x, const double2 *tw, const uint32_t *group, int logs,
                                                               int iter, uint32_t M, double scalar) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t gt = blockIdx.x * EW_THREADS + threadIdx.x;
    const int logPairs = logs - iter - 1;
    const uint32_t pairs = 1u << logPairs;
    const uint32_t k = gt >> logPairs, j = gt & (pairs - 1), a = 2 * k * pairs + j;
    double2 x0 = x[a], x1 = x[a + pairs];
    ckks_gs<false>(x0, x1, ckks_tw(tw, group, k, logPairs, logs, M));
    if (iter == 0 && scalar != 0.0) {
        x0 = make_double2(__dmul_rn(x0.x, scalar), __dmul_rn(x0.y, scalar));
        x1 = make_double2(__dmul_rn(x1.x, scalar), __dmul_rn(x1.y, scalar));
    }
    x[a] = x0, x[a + pairs] = x1;
}

// ---- decoding direction (PhantomCKKSEncoder::decode_internal, src/ckks.cu:137-190) ---------------------------------
// forward butterfly (x0 + x1 w, x0 - x1 w); imaginary part of the product in the reference's two contraction forms (see
// ckks_gs): shared-memory kernel fma(x1.y, w.x, x1.x * w.y), one-stage kernel fma(x1.x, w.y, x1.y * w.x)
template<bool BLOCK_FORM>
__device__ __forceinline__ void ckks_ct(double2 &x0, double2 &x1, const double2 w) {
    const double t1 = __dmul_rn(x1.y, w.y);
    const double re = __fma_rn(x1.x, w.x, -t1);
    const double im = BLOCK_FORM ? __fma_rn(x1.y, w.x, __dmul_rn(x1.x, w.y)) : __fma_rn(x1.x, w.y, __dmul_rn(x1.y, w.x));
    const double2 a = x0;
    x0 = make_double2(__dadd_rn(a.x, re), __dadd_rn(a.y, im));
    x1 = make_double2(__dadd_rn(a.x, -re), __dadd_rn(a.y, -im));
}
__device__ __forceinline__ double2 ckks_tw_fwd(const double2 *tw, const uint32_t *group, uint32_t k, int logPairs, int logs,
                                               uint32_t M) {
    uint32_t psi = group[__brev(k << logPairs) >> (33 - logs)];
    return tw[(psi << logPairs) & (M - 1)];
}
// stages iter_begin .. logs-1 (pairs <= CKKS_FFT_BLOCK / 2) inside blocks
__global__ void __launch_bounds__(CKKS_FFT_BLOCK / 2) k_ckks_fft_block(double2 *x, const double2 *tw, const uint32_t *group,
                                                                       int logs, int iter_begin, uint32_t M) {
    extern __shared__ __align__(16) unsigned char ckks_smem[];
    double2 *buf = reinterpret_cast<double2 *>(ckks_smem);
    pdl_launch_dependents();
    pdl_wait();
    const int nb = blockDim.x * 2;
    const uint32_t base = blockIdx.x * nb, t = threadIdx.x;
    buf[t] = x[base + t], buf[t + blockDim.x] = x[base + t + blockDim.x];
    __syncthreads();
    for (int iter = iter_begin; iter < logs; iter++) {
        const int logPairs = logs - iter - 1;
        const uint32_t pairs = 1u << logPairs;
        const uint32_t gt = blockIdx.x * blockDim.x + t;
        const uint32_t k = gt >> logPairs, j = gt & (pairs - 1);
        const uint32_t a = 2 * k * pairs + j - base;
        double2 x0 = buf[a], x1 = buf[a + pairs];
        ckks_ct<true>(x0, x1, ckks_tw_fwd(tw, group, k, logPairs, logs, M));
        buf[a] = x0, buf[a + pairs] = x1;
        __syncthreads();
    }
    x[base + t] = buf[t], x[base + t + blockDim.x] = buf[t + blockDim.x];
}
__global__ void __launch_bounds__(EW_THREADS) k_ckks_fft_stage(double2 *x, const double2 *tw, const uint32_t *group, int logs,
                                                              int iter, uint32_t M) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t gt = blockIdx.x * EW_THREADS + threadIdx.x;
    const int logPairs = logs - iter - 1;
    const uint32_t pairs = 1u << logPairs;
    const uint32_t k = gt >> logPairs, j = gt & (pairs - 1), a = 2 * k * pairs + j;
    double2 x0 = x[a], x1 = x[a + pairs];
    ckks_ct<false>(x0, x1, ckks_tw_fwd(tw, group, k, logPairs, logs, M));
    x[a] = x0, x[a + pairs] = x1;
}
__global__ void __launch_bounds__(EW_THREADS) k_ckks_unplace(double2 *out, const double2 *x, int logs) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t i = blockIdx.x * EW_THREADS + threadIdx.x;
    out[logs ? (__brev(i) >> (32 - logs)) : 0] = x[i];   // bit_reverse_kernel, ckks.cu:9-15
}

// compose_array_kernel (reference src/rns_base.cu:174-244): CRT composition of one coefficient to a multi-word integer
// mod Q, centring against (Q + 1) / 2, conversion to double word by word (multiply, then add: not fused in the
// reference), times 1 / scale.  Coefficient x < n/2 becomes the real part of slot x, the others the imaginary parts.
constexpr int CKKS_MAX_WORDS = 32;
struct CkksComposeArgs {
    double2 *x;
    const u64 *w;        // [l][n] coefficient form
    const u64 *hat;      // [l][l] punctured products, little-endian words
    const u64 *Qw, *thr; // [l]
    const Tw *hinv;      // [l]
    const Modulus *mod;
    double inv_scale;
    int l;
    size_t n;
};
__global__ void __launch_bounds__(EW_THREADS) k_ckks_compose(const CkksComposeArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const int l = a.l;
    u64 acc[CKKS_MAX_WORDS];
    for (int k = 0; k < l; k++) acc[k] = 0;
    if (l > 1) {
        for (int i = 0; i < l; i++) {
            const u64 prod = mul_shoup(a.w[(size_t) i * a.n + c], a.hinv[i], a.mod[i].q);
            const u64 *h = a.hat + (size_t) i * l;
            u64 carry = 0, cy = 0;
            for (int k = 0; k < l; k++) {   // acc += hat_i * prod (low l words: the product is below Q)
                const unsigned __int128 t = (unsigned __int128) h[k] * prod + carry;
                carry = (u64) (t >> 64);
                const unsigned __int128 s = (unsigned __int128) acc[k] + (u64) t + cy;
                acc[k] = (u64) s, cy = (u64) (s >> 64);
            }
            bool ge = cy != 0;   // acc >= Q ?
            if (!ge) {
                ge = true;
                for (int k = l - 1; k >= 0; k--)
                    if (acc[k] != a.Qw[k]) {
                        ge = acc[k] > a.Qw[k];
                        break;
                    }
            }
            if (ge) {
                u64 bw = 0;
                for (int k = 0; k < l; k++) {
                    const unsigned __int128 t = (unsigned __int128) acc[k] - a.Qw[k] - bw;
                    acc[k] = (u64) t, bw = (u64) (t >> 64) & 1;
                }
            }
        }
    } else {
        acc[0] = a.w[c];
    }
    bool upper = true;   // acc >= (Q + 1) / 2 ?
    for (int k = l - 1; k >= 0; k--)
        if (acc[k] != a.thr[k]) {
            upper = acc[k] > a.thr[k];
            break;
        }
    double res = 0.0, s2 = a.inv_scale;
    for (int k = 0; k < l; k++, s2 = __dmul_rn(s2, 18446744073709551616.0)) {
        if (upper) {
            if (acc[k] > a.Qw[k]) {
                const u64 d = acc[k] - a.Qw[k];
                res = __dadd_rn(res, d ? __dmul_rn((double) d, s2) : 0.0);
            } else {
                const u64 d = a.Qw[k] - acc[k];
                res = __dadd_rn(res, -(d ? __dmul_rn((double) d, s2) : 0.0));
            }
        } else {
            const u64 d = acc[k];
            res = __dadd_rn(res, d ? __dmul_rn((double) d, s2) : 0.0);
        }
    }
    const size_t slots = a.n >> 1;
    if (c < slots) a.x[c].x = res;
    else a.x[c - slots].y = res;
}

// max |component| over the slots as the bit pattern of a non-negative double (orders like the value)
__global__ void __launch_bounds__(EW_THREADS) k_ckks_absmax(const double2 *x, unsigned long long *out) {
    pdl_launch_dependents();
    pdl_wait();
    const double2 v = x[blockIdx.x * EW_THREADS + threadIdx.x];
    const double m = fmax(fabs(v.x), fabs(v.y));
    unsigned long long b = (unsigned long long) __double_as_longlong(m);
    for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, b);
}

// decompose_array (reference src/rns_base.cu:49-103): coefficient x of limb i from round(re) (x < n/2) or round(im);
// wide = 0: values below 2^64, wide = 1: below 2^128.  grid.y = limb
__global__ void __launch_bounds__(EW_THREADS) k_ckks_decompose(u64 *out, const double2 *x, const Modulus *mod, size_t n, int wide) {
    pdl_launch_dependents();
    pdl_wait();
    const Modulus m = mod[blockIdx.y];
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x, slots = n >> 1;
    const double cd = round(c < slots ? x[c].x : x[c - slots].y);
    const bool negative = signbit(cd);
    const double ad = fabs(cd);
    u64 r;
    if (!wide) {
        r = barrett64((u64) ad, m);
    } else {
        const u64 lo = (u64) fmod(ad, 18446744073709551616.0), hi = (u64) (ad / 18446744073709551616.0);
        r = barrett128(lo, hi, m);
    }
    out[(size_t) blockIdx.y * n + c] = negative ? m.q - r : r;   // a negative zero yields q, like the reference
}

// PhantomBatchEncoder (reference src/batchencoder.cu:51-60,91-95): slot i <-> position map[i] of the NTT-form vector
__global__ void __launch_bounds__(EW_THREADS) k_batch_encode(u64 *out, const u64 *in, size_t count, const uint32_t *map, u64 t) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    u64 v = 0;
    if (i < count) {
        v = in[i];
        v += (v >> 63) * t;   // negative values arrive as two's complement
    }
    out[map[i]] = v;
}
__global__ void __launch_bounds__(EW_THREADS) k_batch_decode(u64 *out, const u64 *in, const uint32_t *map) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    out[i] = in[map[i]];
}

// dst[limb i] = src[perm...] helpers -------------------------------------------------------------------

// apply_galois_ntt_permutation (reference src/galois.cu:11-18): dst[l][i] = src[l][perm[i]]
__global__ void __launch_bounds__(EW_THREADS) k_galois_ntt(u64 *dst, const u64 *src, const uint32_t *perm, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t limb = blockIdx.y;
    const size_t i = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const uint2 p = *reinterpret_cast<const uint2 *>(perm + i);
    st2(dst + limb * n + i, src[limb * n + p.x], src[limb * n + p.y]);
}
// dst[l][i] += src[l][perm[i]]  (hoisting: accumulated automorphisms of c0, evaluate.cu:1797-1810)
__global__ void __launch_bounds__(EW_THREADS) k_galois_ntt_acc(u64 *dst, const u64 *src, const uint32_t *perm,
                                                                const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t limb = blockIdx.y;
    const u64 q = mod[limb].q;
    const size_t i = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const uint2 p = *reinterpret_cast<const uint2 *>(perm + i);
    const ulonglong2 a = ld2(dst + limb * n + i);
    st2(dst + limb * n + i, add_mod(a.x, src[limb * n + p.x], q), add_mod(a.y, src[limb * n + p.y], q));
}

// apply_galois_permutation (reference src/galois.cu:20-39), coefficient domain (BFV): x^i -> x^(i * elt mod 2n),
// sign flip when the exponent lands in [n, 2n).  grid.y = poly * l + limb
__global__ void __launch_bounds__(EW_THREADS) k_galois_coeff(u64 *dst, const u64 *src, const Modulus *mod, uint32_t elt,
                                                              size_t n, int l) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t pl = blockIdx.y;
    const u64 q = mod[pl % l].q;
    const size_t i = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const size_t idx = (i * elt) & (2 * n - 1);
    u64 v = src[pl * n + i];
    if (idx >= n) v = v ? q - v : 0;
    dst[pl * n + (idx & (n - 1))] = v;
}

// CKKS rescale pieces (divide_and_round_q_last_ntt, reference src/rns.cu:1128-1184)
// r[j][x] = last[x] mod q_j
__global__ void __launch_bounds__(EW_THREADS) k_reduce_last(u64 *dst, const u64 *last, const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const Modulus m = mod[j];
    const size_t x = ((size_t) blockIdx.x * EW_THREADS + threadIdx.x) * 2;
    const ulonglong2 v = ld2(last + x);
    st2(dst + (size_t) j * n + x, barrett64(v.x, m), barrett64(v.y, m));
}


// ---------------------------------------------------------------------------------------------------
// per-call constants of the kernel-level launchers that take device arrays from the caller (nwt_2d_radix8_*_scale,
// ..._fuse_moddown): one thread per limb slot
// ---------------------------------------------------------------------------------------------------
struct FinList {
    int count;
    short row[NTT_MAX_LIMBS];   // table row of the slot
    short idx[NTT_MAX_LIMBS];   // index into the caller's array
};

__device__ __forceinline__ Tw tw_of_row(u64 w, u64 q, bool fp) {
    if (fp) {
        const double dw = (double) w, wi = dw / (double) q;
        return make_ulonglong2((u64) __double_as_longlong(dw), (u64) __double_as_longlong(wi));
    }
    return make_ulonglong2(w, (u64) ((((unsigned __int128) w) << 64) / q));
}

// fin[2 s] = tw(n^-1 scale), fin[2 s + 1] = tw(itw[1] n^-1 scale) in the representation of the slot's row
__global__ void k_make_fin(Tw *fin, const u64 *scale, FinList fl, const u64 *fin_int, const Modulus *mod,
                           const unsigned char *is_fp, int fp_enabled) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = threadIdx.x;
    if (s >= fl.count) return;
    const int row = fl.row[s];
    const u64 q = mod[row].q;
    const u64 c = scale[fl.idx[s]] % q;
    const bool fp = fp_enabled && is_fp[row];
    fin[2 * s] = tw_of_row((u64) ((unsigned __int128) fin_int[2 * row] * c % q), q, fp);
    fin[2 * s + 1] = tw_of_row((u64) ((unsigned __int128) fin_int[2 * row + 1] * c % q), q, fp);
}

// epilogue constants {c, floor(c 2^64 / q)} from the caller's two arrays
__global__ void k_zip_tw(Tw *out, const u64 *c, const u64 *c_shoup, FinList fl) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = threadIdx.x;
    if (s < fl.count) out[s] = make_ulonglong2(c[fl.idx[s]], c_shoup[fl.idx[s]]);
}

} // namespace pfhe
