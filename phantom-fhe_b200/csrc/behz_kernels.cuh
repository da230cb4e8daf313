// behz_kernels.cuh -- the two per-coefficient kernels of the BEHZ BFV multiplication.
//
// Every step of BEHZ between the NTTs is coefficient-local: all residues of one coefficient are combined with
// small constant matrices.  Two kernels cover them:
//   k_behz_lift      fastbconv_m_tilde + sm_mrq  (reference rns.cu:1249-1339): q -> Bsk with the m_tilde correction
//   k_behz_floor_sk  fast_floor + fastbconv_sk   (reference rns.cu:1343-1517, polymath.cu:596-634): Bsk -> q
// One thread owns one coefficient; its residues live in a thread-local array, the matrices are warp-uniform
// read-only loads.  The auxiliary base is 61-bit, so this is integer-pipe work (128-bit accumulate, one Barrett
// reduction per output residue) -- the same arithmetic as the reference's base-conversion kernels.
#pragma once
#include "modarith.cuh"

namespace pfhe {

constexpr int BEHZ_MAX_LIMBS = 64;   // q limbs and Bsk limbs per coefficient (thread-local array size)
constexpr int BEHZ_THREADS = 128;

struct BehzLiftArgs {
    const u64 *ct1, *ct2;   // [2][l][n] each, coefficient form
    u64 *out;               // [4][nbsk][n]: Bsk residues of ct1.c0, ct1.c1, ct2.c0, ct2.c1
    const Tw *k;            // [l]  m_tilde * qhat_i^-1 mod q_i
    const u64 *mat;         // [nbsk][l]  qhat_i mod Bsk_j
    const u32 *mat_mt;      // [l]  qhat_i mod m_tilde
    const u64 *q_mod_bsk;   // [nbsk]
    const Tw *inv_mt;       // [nbsk]
    const Modulus *mod_q, *mod_bsk;
    u32 neg_inv_q;
    int l, nbsk;
    size_t n;
};

__global__ void __launch_bounds__(BEHZ_THREADS) k_behz_lift(const BehzLiftArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int poly = blockIdx.y;
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *src = (poly < 2 ? a.ct1 : a.ct2) + (size_t) (poly & 1) * a.l * a.n + x;
    u64 y[BEHZ_MAX_LIMBS];
    u32 mt = 0;
    for (int i = 0; i < a.l; i++) {
        y[i] = mul_shoup(src[(size_t) i * a.n], a.k[i], a.mod_q[i].q);
        mt += (u32) y[i] * a.mat_mt[i];   // m_tilde = 2^32: the reduction is the wrap-around
    }
    // sm_mrq: r = mt * (-Q^-1) mod m_tilde, centred
    const u32 r = mt * a.neg_inv_q;
    u64 *dst = a.out + (size_t) poly * a.nbsk * a.n + x;
    for (int j = 0; j < a.nbsk; j++) {
        const Modulus m = a.mod_bsk[j];
        Acc128 acc{0, 0};
        const u64 *row = a.mat + (size_t) j * a.l;
        for (int i = 0; i < a.l; i++) acc.mac(y[i], row[i]);
        const u64 rr = (r >> 31) ? (u64) r + m.q - ((u64) 1 << 32) : (u64) r;
        acc.mac(rr, a.q_mod_bsk[j]);
        const u64 v = barrett128(acc.lo, acc.hi, m);
        dst[(size_t) j * a.n] = mul_shoup(v, a.inv_mt[j], m.q);
    }
}

struct BehzFloorArgs {
    const u64 *xq;          // [3][l][n]     t * tensor result, base q, coefficient form
    const u64 *xb;          // [3][nbsk][n]  the same in Bsk = [m_sk, B...]
    u64 *out;               // [3][l][n]
    const Tw *q_hinv;       // [l]
    const u64 *q_to_bsk;    // [nbsk][l]
    const Tw *inv_q_bsk;    // [nbsk]
    const Tw *b_hinv;       // [nB]
    const u64 *b_to_q;      // [l][nB]
    const u64 *b_to_msk;    // [nB]
    const u64 *B_mod_q;     // [l]
    Tw inv_B_msk;
    const Modulus *mod_q, *mod_bsk;
    int l, nbsk;
    size_t n;
};

__global__ void __launch_bounds__(BEHZ_THREADS) k_behz_floor_sk(const BehzFloorArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int poly = blockIdx.y;
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *sq = a.xq + (size_t) poly * a.l * a.n + x;
    const u64 *sb = a.xb + (size_t) poly * a.nbsk * a.n + x;
    u64 y[BEHZ_MAX_LIMBS], fl[BEHZ_MAX_LIMBS];
    // fast_floor: fl_j = (xb_j - FastBconv(xq)_j) * Q^-1 mod Bsk_j
    for (int i = 0; i < a.l; i++) y[i] = mul_shoup(sq[(size_t) i * a.n], a.q_hinv[i], a.mod_q[i].q);
    for (int j = 0; j < a.nbsk; j++) {
        const Modulus m = a.mod_bsk[j];
        Acc128 acc{0, 0};
        const u64 *row = a.q_to_bsk + (size_t) j * a.l;
        for (int i = 0; i < a.l; i++) acc.mac(y[i], row[i]);
        const u64 v = barrett128(acc.lo, acc.hi, m);
        fl[j] = mul_shoup(sub_mod(sb[(size_t) j * a.n], v, m.q), a.inv_q_bsk[j], m.q);
    }
    // fastbconv_sk: Shenoy-Kumaresan conversion B -> q with alpha_sk from the m_sk residue
    const int nB = a.nbsk - 1;
    const Modulus msk = a.mod_bsk[0];
    Acc128 am{0, 0};
    for (int i = 0; i < nB; i++) {
        y[i] = mul_shoup(fl[i + 1], a.b_hinv[i], a.mod_bsk[i + 1].q);
        am.mac(y[i], a.b_to_msk[i]);
    }
    const u64 alpha = mul_shoup(sub_mod(barrett128(am.lo, am.hi, msk), fl[0], msk.q), a.inv_B_msk, msk.q);
    const bool neg = alpha > (msk.q >> 1);
    u64 *dst = a.out + (size_t) poly * a.l * a.n + x;
    for (int k = 0; k < a.l; k++) {
        const Modulus m = a.mod_q[k];
        Acc128 acc{0, 0};
        const u64 *row = a.b_to_q + (size_t) k * nB;
        for (int i = 0; i < nB; i++) acc.mac(y[i], row[i]);
        const u64 v = barrett128(acc.lo, acc.hi, m);
        const u64 Bq = a.B_mod_q[k];
        const u64 corr = neg ? mul_mod(msk.q - alpha, Bq, m) : mul_mod(alpha, m.q - Bq, m);
        dst[(size_t) k * a.n] = add_mod(v, corr, m.q);
    }
}

// ---------------------------------------------------------------------------------------------------
// HPS (mul_tech_type::hps).  The floating-point parts reproduce the reference's compiled arithmetic: one
// I2F.F64.U64 and one DFMA per term, accumulated in index order (nvcc contracts `acc += double(x) * c`), llround for
// the overflow estimate of bConv_HPS, truncation for the rounding term of scaleAndRound.
// ---------------------------------------------------------------------------------------------------
struct HpsLiftArgs {
    const u64 *ct1, *ct2;   // [2][l][n] each, coefficient form
    u64 *out;               // [4][nR][n]
    const Tw *q_hinv;       // [l]
    const double *q_inv;    // [l]
    const u64 *mat;         // [nR][l]
    const u64 *Q_mod_r;     // [nR]
    const Modulus *mod_q, *mod_r;
    int l, nR;
    size_t n;
};

// bConv_HPS Q -> R (rns_bconv.cu:278-304,354-372)
__global__ void __launch_bounds__(BEHZ_THREADS) k_hps_lift(const HpsLiftArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int poly = blockIdx.y;
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *src = (poly < 2 ? a.ct1 : a.ct2) + (size_t) (poly & 1) * a.l * a.n + x;
    u64 y[BEHZ_MAX_LIMBS];
    double frac = 0.0;
    for (int i = 0; i < a.l; i++) {
        y[i] = mul_shoup(src[(size_t) i * a.n], a.q_hinv[i], a.mod_q[i].q);
        frac = __fma_rn((double) y[i], a.q_inv[i], frac);
    }
    const u64 v = (u64) llround(frac);
    u64 *dst = a.out + (size_t) poly * a.nR * a.n + x;
    for (int j = 0; j < a.nR; j++) {
        const Modulus m = a.mod_r[j];
        Acc128 acc{0, 0};
        const u64 *row = a.mat + (size_t) j * a.l;
        for (int i = 0; i < a.l; i++) acc.mac(y[i], row[i]);
        const u64 r = barrett128(acc.lo, acc.hi, m);
        dst[(size_t) j * a.n] = sub_mod(r, mul_mod(v, a.Q_mod_r[j], m), m.q);
    }
}

struct HpsScaleArgs {
    const u64 *xq;          // [3][l][n]   tensor result over Q, coefficient form
    const u64 *xr;          // [3][nR][n]  and over R
    u64 *out;               // [3][l][n]
    const double *sr_frac;  // [l]
    const u64 *sr_tab;      // [nR][l+1]
    const Tw *r_hinv;       // [nR]
    const double *r_inv;    // [nR]
    const u64 *r_to_q;      // [l][nR]
    const u64 *R_mod_q;     // [l]
    const Modulus *mod_q, *mod_r;
    int l, nR;
    size_t n;
};

// scaleAndRound_HPS_QR_R (rns.cu:1699-1733) followed by bConv_HPS R -> Q
__global__ void __launch_bounds__(BEHZ_THREADS) k_hps_scale_round(const HpsScaleArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int poly = blockIdx.y;
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *sq = a.xq + (size_t) poly * a.l * a.n + x;
    const u64 *sr = a.xr + (size_t) poly * a.nR * a.n + x;
    u64 xi[BEHZ_MAX_LIMBS], y[BEHZ_MAX_LIMBS];
    double nu = 0.5;
    for (int i = 0; i < a.l; i++) {
        xi[i] = sq[(size_t) i * a.n];
        nu = __fma_rn((double) xi[i], a.sr_frac[i], nu);
    }
    u64 alpha = (u64) nu;   // cvt.rzi.u64.f64 (saturating), as static_cast<uint64_t> compiles to
    double frac = 0.0;
    for (int j = 0; j < a.nR; j++) {
        const Modulus m = a.mod_r[j];
        const u64 *row = a.sr_tab + (size_t) j * (a.l + 1);
        Acc128 acc{0, 0};
        for (int i = 0; i < a.l; i++) acc.mac(xi[i], row[i]);
        acc.mac(sr[(size_t) j * a.n], row[a.l]);
        const u64 cur = barrett128(acc.lo, acc.hi, m);
        alpha = barrett64(alpha, m);   // the reference re-reduces the same variable limb after limb
        const u64 v = add_mod(cur, alpha, m.q);
        y[j] = mul_shoup(v, a.r_hinv[j], m.q);
        frac = __fma_rn((double) y[j], a.r_inv[j], frac);
    }
    const u64 v = (u64) llround(frac);
    u64 *dst = a.out + (size_t) poly * a.l * a.n + x;
    for (int i = 0; i < a.l; i++) {
        const Modulus m = a.mod_q[i];
        Acc128 acc{0, 0};
        const u64 *row = a.r_to_q + (size_t) i * a.nR;
        for (int j = 0; j < a.nR; j++) acc.mac(y[j], row[j]);
        const u64 r = barrett128(acc.lo, acc.hi, m);
        dst[(size_t) i * a.n] = sub_mod(r, mul_mod(v, a.R_mod_q[i], m), m.q);
    }
}

// ---------------------------------------------------------------------------------------------------
// HPS over Q (mul_tech_type::hps_overq, and the arithmetic of hps_overq_leveled; evaluate.cu:647-801).  Three generic
// thread-per-coefficient kernels on strided [limb][n] views; grid.y = polynomial.
// ---------------------------------------------------------------------------------------------------
struct PolyView {        // limb i of polynomial p at base + p * poly_stride + i * n
    u64 *base;
    size_t poly_stride;
};

// bConv_HPS (rns_bconv.cu:278-304,354-372): exact conversion with the floating-point overflow estimate
struct BconvHpsArgs {
    PolyView in, out;
    const Tw *hinv;         // [ni]      ihat_i^-1 mod i_i
    const double *inv;      // [ni]      1.0 / i_i
    const u64 *mat;         // [no][ni]  ihat_i mod o_j
    const u64 *Imod;        // [no]      I mod o_j
    const Modulus *mod_in, *mod_out;
    int ni, no;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_bconv_hps(const BconvHpsArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *src = a.in.base + (size_t) blockIdx.y * a.in.poly_stride + x;
    u64 y[BEHZ_MAX_LIMBS];
    double frac = 0.0;
    for (int i = 0; i < a.ni; i++) {
        y[i] = mul_shoup(src[(size_t) i * a.n], a.hinv[i], a.mod_in[i].q);
        frac = __fma_rn((double) y[i], a.inv[i], frac);
    }
    const u64 v = (u64) llround(frac);
    u64 *dst = a.out.base + (size_t) blockIdx.y * a.out.poly_stride + x;
    for (int j = 0; j < a.no; j++) {
        const Modulus m = a.mod_out[j];
        Acc128 acc{0, 0};
        const u64 *row = a.mat + (size_t) j * a.ni;
        for (int i = 0; i < a.ni; i++) acc.mac(y[i], row[i]);
        const u64 r = barrett128(acc.lo, acc.hi, m);
        dst[(size_t) j * a.n] = sub_mod(r, mul_mod(v, a.Imod[j], m), m.q);
    }
}

// bConv_BEHZ_var1 (rns_bconv.cu:231-246): y_i = x_i * (-O * ihat_i^-1) mod i_i, out_j = sum_i y_i * (i_i^-1 mod o_j)
struct BconvVar1Args {
    PolyView in, out;
    const Tw *c1;           // [ni]
    const u64 *mat;         // [no][ni]
    const Modulus *mod_in, *mod_out;
    int ni, no;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_bconv_var1(const BconvVar1Args a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *src = a.in.base + (size_t) blockIdx.y * a.in.poly_stride + x;
    u64 y[BEHZ_MAX_LIMBS];
    for (int i = 0; i < a.ni; i++) {
        // the constant may equal the modulus itself (host/rns.cu:481-482): reduce through the general product
        y[i] = mul_mod(src[(size_t) i * a.n], a.c1[i].x, a.mod_in[i]);
    }
    u64 *dst = a.out.base + (size_t) blockIdx.y * a.out.poly_stride + x;
    for (int j = 0; j < a.no; j++) {
        const Modulus m = a.mod_out[j];
        Acc128 acc{0, 0};
        const u64 *row = a.mat + (size_t) j * a.ni;
        for (int i = 0; i < a.ni; i++) acc.mac(y[i], row[i]);
        dst[(size_t) j * a.n] = barrett128(acc.lo, acc.hi, m);
    }
}

// scaleAndRound_HPS_QlRl_Ql_kernel (rns.cu:1748-1784), also scaleAndRound_HPS_Q_Ql (:1797-1805):
//   out_i = sum_j xb_j tab[i][j] + xa_i tab[i][nb] + alpha mod a_i,  alpha = trunc(0.5 + sum_j double(xb_j) frac_j),
// alpha re-reduced limb after limb as the reference does.  Optionally followed by ExpandCRTBasis_Ql_Q (:1810-1834):
// times expand[i], and `zero` further limbs of the output cleared.
struct ScaleRoundArgs {
    PolyView xa, xb, out;
    const u64 *tab;         // [na][nb + 1]
    const double *frac;     // [nb]
    const Tw *expand;       // [na] or null
    const Modulus *mod_a;
    int na, nb, zero;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_scale_round(const ScaleRoundArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t x = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 *sa = a.xa.base + (size_t) blockIdx.y * a.xa.poly_stride + x;
    const u64 *sb = a.xb.base + (size_t) blockIdx.y * a.xb.poly_stride + x;
    u64 xb[BEHZ_MAX_LIMBS], xi[BEHZ_MAX_LIMBS];
    double nu = 0.5;
    for (int j = 0; j < a.nb; j++) {
        xb[j] = sb[(size_t) j * a.n];
        nu = __fma_rn((double) xb[j], a.frac[j], nu);
    }
    for (int i = 0; i < a.na; i++) xi[i] = sa[(size_t) i * a.n];   // out may alias xa
    u64 alpha = (u64) nu;   // cvt.rzi.u64.f64, as static_cast<uint64_t> compiles to
    u64 *dst = a.out.base + (size_t) blockIdx.y * a.out.poly_stride + x;
    for (int i = 0; i < a.na; i++) {
        const Modulus m = a.mod_a[i];
        const u64 *row = a.tab + (size_t) i * (a.nb + 1);
        Acc128 acc{0, 0};
        for (int j = 0; j < a.nb; j++) acc.mac(xb[j], row[j]);
        acc.mac(xi[i], row[a.nb]);
        const u64 cur = barrett128(acc.lo, acc.hi, m);
        alpha = barrett64(alpha, m);
        u64 v = add_mod(cur, alpha, m.q);
        if (a.expand) v = mul_shoup(v, a.expand[i], m.q);
        dst[(size_t) i * a.n] = v;
    }
    for (int i = 0; i < a.zero; i++) dst[(size_t) (a.na + i) * a.n] = 0;
}

// ExpandCRTBasis_Ql_Q_add_to_ct (rns.cu:1836-1856): ct[p][i] += in[p][i] * expand_i mod q_i for the ll kept limbs;
// grid.y = polynomial * ll + limb, two coefficients per thread
__global__ void __launch_bounds__(BEHZ_THREADS) k_expand_add(PolyView ct, PolyView in, const Tw *expand, const Modulus *mod,
                                                            int ll, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int p = blockIdx.y / ll, i = blockIdx.y % ll;
    const size_t x = ((size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x) * 2;
    const u64 q = mod[i].q;
    const Tw f = expand[i];
    u64 *d = ct.base + (size_t) p * ct.poly_stride + (size_t) i * n + x;
    const u64 *s = in.base + (size_t) p * in.poly_stride + (size_t) i * n + x;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(d), b = *reinterpret_cast<const ulonglong2 *>(s);
    *reinterpret_cast<ulonglong2 *>(d) =
            make_ulonglong2(add_mod(a.x, mul_shoup(b.x, f, q), q), add_mod(a.y, mul_shoup(b.y, f, q), q));
}

// ---------------------------------------------------------------------------------------------------
// decryption tails (PhantomSecretKey::bfv_decrypt / bgv_decrypt, reference src/secretkey.cu:571-691): one thread per
// coefficient turns the RNS residues of c_0 + c_1 s + ... into the plaintext residue mod t.
// ---------------------------------------------------------------------------------------------------
// c_0 + sum_k c_k s^k, limb-wise (multiply_and_add_rns_poly chain, secretkey.cu:553-562): ct = [size][l][n] (polys
// 1.. already in NTT form; BFV passes its transformed copy and adds c_0 after the inverse transform), sk = powers of the
// secret key, stride sk_stride words.  grid.y = limb, 2 coefficients per thread.
__global__ void __launch_bounds__(BEHZ_THREADS) k_decrypt_inner(u64 *acc, const u64 *c0, const u64 *cts, size_t ct_stride,
                                                               const u64 *sk, size_t sk_stride, int terms,
                                                               const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.y;
    const Modulus m = mod[i];
    const size_t x = (size_t) i * n + ((size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x) * 2;
    Acc128 a0{0, 0}, a1{0, 0};
    if (c0) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(c0 + x);
        a0.lo = v.x, a1.lo = v.y;
    }
    for (int k = 0; k < terms; k++) {
        const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(cts + (size_t) k * ct_stride + x);
        const ulonglong2 s = *reinterpret_cast<const ulonglong2 *>(sk + (size_t) k * sk_stride + x);
        a0.mac(c.x, s.x), a1.mac(c.y, s.y);
    }
    *reinterpret_cast<ulonglong2 *>(acc + x) = make_ulonglong2(barrett128(a0.lo, a0.hi, m), barrett128(a1.lo, a1.hi, m));
}

// hps_decrypt_scale_and_round (rns.cu:1519-1692): round(t/Q x) mod t.  `add` = c_0 (coefficient form) is added to the
// inner product first (secretkey.cu:613-615).  large / lazy select the reference's four kernel variants; the FP sums
// are one fma per term in index order, like the reference's compiled code.
struct HpsDecryptArgs {
    const u64 *x, *add;     // [l][n]
    u64 *out;               // [n]
    const u64 *mt, *mtB;    // [l]
    const double *fr, *frB; // [l]
    const Modulus *mod_q;
    Modulus tm;
    int l, large, lazy, hf;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_hps_decrypt(const HpsDecryptArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t k = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    const u64 t = a.tm.q;
    double fs = 0.0;
    u64 is = 0;
    for (int i = 0; i < a.l; i++) {
        const u64 v = add_mod(a.add[(size_t) i * a.n + k], a.x[(size_t) i * a.n + k], a.mod_q[i].q);
        if (!a.large) {
            fs = __fma_rn((double) v, a.fr[i], fs);
            is += a.lazy ? v * a.mt[i] : mul_mod(v, a.mt[i], a.tm);
        } else {
            const u64 hi = v >> a.hf, lo = v & (((u64) 1 << a.hf) - 1);
            fs = __fma_rn((double) lo, a.fr[i], fs);
            fs = __fma_rn((double) hi, a.frB[i], fs);
            is += a.lazy ? lo * a.mt[i] : mul_mod(lo, a.mt[i], a.tm);
            is += a.lazy ? hi * a.mtB[i] : mul_mod(hi, a.mtB[i], a.tm);
        }
    }
    fs = __dadd_rn(fs, (double) is);
    const u64 quot = (u64) __dmul_rn(fs, 1. / (double) t);
    fs = __dadd_rn(fs, -(double) (t * quot));
    a.out[k] = (u64) llround(fs);
}

// behz_decrypt_scale_and_round (rns.cu:1008-1080): |gamma t|_q scaling, fast conversion to {t, gamma}, times
// -Q^-1, centred correction through gamma, times gamma^-1 mod t
struct BehzDecryptArgs {
    const u64 *x, *add;
    u64 *out;
    const Tw *tg_mod_q;     // [l]   t * gamma mod q_i
    const Tw *q_hinv;       // [l]   qhat_i^-1 mod q_i
    const u64 *mat;         // [2][l] qhat_i mod t, mod gamma
    const Modulus *mod_q;
    Modulus tm, gm;
    u64 ninv_t, ninv_g, inv_gamma_t;
    int l;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_behz_decrypt(const BehzDecryptArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t k = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    Acc128 at{0, 0}, ag{0, 0};
    for (int i = 0; i < a.l; i++) {
        const u64 q = a.mod_q[i].q;
        const u64 v = add_mod(a.add[(size_t) i * a.n + k], a.x[(size_t) i * a.n + k], q);
        const u64 y = mul_shoup(mul_shoup(v, a.tg_mod_q[i], q), a.q_hinv[i], q);
        at.mac(y, a.mat[i]);
        ag.mac(y, a.mat[a.l + i]);
    }
    const u64 t = a.tm.q, g = a.gm.q;
    const u64 rt = mul_mod(barrett128(at.lo, at.hi, a.tm), a.ninv_t, a.tm);
    const u64 rg = mul_mod(barrett128(ag.lo, ag.hi, a.gm), a.ninv_g, a.gm);
    u64 tmp;
    if (rg > (g >> 1)) tmp = add_mod(rt, barrett128(g - rg, 0, a.tm), t);
    else tmp = sub_mod(rt, barrett128(rg, 0, a.tm), t);
    a.out[k] = mul_mod(tmp, a.inv_gamma_t, a.tm);
}

// exact_convert_array Q_l -> t (rns_bconv.cu:374-430; DRNSTool::decrypt_mod_t) and the BGV correction factor
struct ExactConvertArgs {
    const u64 *x;
    u64 *out;
    const Tw *q_hinv;       // [l]
    const u64 *mat;         // [l] qhat_i mod t
    const Modulus *mod_q;
    Modulus tm;
    u64 q_mod_t, fix;       // fix = correction_factor^-1 mod t (1: none)
    int l;
    size_t n;
};
__global__ void __launch_bounds__(BEHZ_THREADS) k_exact_convert_t(const ExactConvertArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t k = (size_t) blockIdx.x * BEHZ_THREADS + threadIdx.x;
    double v = 0.0;
    Acc128 ip{0, 0};
    for (int i = 0; i < a.l; i++) {
        const u64 q = a.mod_q[i].q;
        const u64 y = mul_shoup(a.x[(size_t) i * a.n + k], a.q_hinv[i], q);
        ip.mac(y, a.mat[i]);
        v = __dadd_rn(v, __ddiv_rn((double) y, (double) q));   // IEEE quotient then sum, as the reference compiles it
    }
    const u64 rv = (u64) round(v);
    u64 r = sub_mod(barrett128(ip.lo, ip.hi, a.tm), mul_mod(barrett128(rv, 0, a.tm), a.q_mod_t, a.tm), a.tm.q);
    if (a.fix != 1) r = mul_mod(r, a.fix, a.tm);
    a.out[k] = r;
}

} // namespace pfhe
