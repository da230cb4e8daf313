// Samplers and the element-wise steps of key generation / encryption.
//
// The reference draws every random polynomial from a counter-mode generator keyed by a 64-byte seed: one 64-byte block
// per nonce from the Salsa20 core (20 rounds) over a state of its own layout -- no "expand" constants: key words 0..7,
// the 64-bit nonce, key words 8..13 (salsa20_gpu, src/prng.cu:17-133; only 56 of the 64 seed bytes are read).  The three
// samplers (src/prng.cu:142-244) are reproduced bit for bit given the seed:
//   ternary  byte 0 of block(nonce = coefficient) mod 3 - 1, the same value in every limb
//   error    centred binomial, 21 bits against 21 bits out of bytes 0..5 of block(nonce = coefficient), same in every limb
//   uniform  the eight 64-bit words of block(nonce = limb * N/8 + group) reduced mod q_limb, with rejection above the
//            largest multiple of q: a rejected word re-draws the whole block with nonce + tries * N * limbs and carries on
//            at the same word
// The reference spends one generator block per (limb, coefficient) for the first two; here a thread computes the block
// of its coefficient once and writes every limb.
#pragma once
#include "engine.hpp"
#include "poly_kernels.cuh"

namespace pfhe {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

#define PFHE_SALSA_QR(a, b, c, d)        \
    x[b] ^= rotl32(x[a] + x[d], 7);      \
    x[c] ^= rotl32(x[b] + x[a], 9);      \
    x[d] ^= rotl32(x[c] + x[b], 13);     \
    x[a] ^= rotl32(x[d] + x[c], 18);

__device__ __forceinline__ void salsa_block(uint32_t (&x)[16], const Seed &key, u64 nonce) {
    uint32_t in[16];
#pragma unroll
    for (int i = 0; i < 8; i++) in[i] = key.w[i];
    in[8] = (uint32_t) nonce, in[9] = (uint32_t) (nonce >> 32);
#pragma unroll
    for (int i = 0; i < 6; i++) in[10 + i] = key.w[8 + i];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        PFHE_SALSA_QR(0, 4, 8, 12) PFHE_SALSA_QR(5, 9, 13, 1) PFHE_SALSA_QR(10, 14, 2, 6) PFHE_SALSA_QR(15, 3, 7, 11)
        PFHE_SALSA_QR(0, 1, 2, 3) PFHE_SALSA_QR(5, 6, 7, 4) PFHE_SALSA_QR(10, 11, 8, 9) PFHE_SALSA_QR(15, 12, 13, 14)
    }
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] += in[i];
}
#undef PFHE_SALSA_QR

enum { SAMPLE_TERNARY = 0, SAMPLE_ERROR = 1, SAMPLE_UNIFORM = 2 };

// sample_ternary_poly / sample_error_poly (src/prng.cu:142-164, 222-244): out = [limbs][n], limb i over modulus row i
template<int KIND>
__global__ void __launch_bounds__(EW_THREADS) k_sample_small(u64 *out, const Seed seed, const Modulus *mod, size_t n,
                                                             int limbs) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    uint32_t x[16];
    salsa_block(x, seed, c);
    int v;
    if (KIND == SAMPLE_TERNARY) {
        v = (int) ((x[0] & 0xFF) % 3) - 1;
    } else {
        v = __popc(x[0] & 0x1FFFFF) - __popc((x[0] >> 24) | ((x[1] & 0x1FFF) << 8));   // bytes 0,1,2&1F - bytes 3,4,5&1F
    }
    for (int i = 0; i < limbs; i++) out[(size_t) i * n + c] = v < 0 ? mod[i].q + (long long) v : (u64) v;
}

// sample_uniform_poly (src/prng.cu:174-204): one thread per eight consecutive coefficients of one limb
__global__ void __launch_bounds__(EW_THREADS) k_sample_uniform(u64 *out, const Seed seed, const Modulus *mod, size_t n,
                                                               int limbs) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t groups = n >> 3;
    const size_t g = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    if (g >= groups) return;
    const int limb = blockIdx.y;
    const Modulus m = mod[limb];
    const u64 tid = (u64) limb * groups + g;
    const u64 max_multiple = ~0ull - (~0ull % m.q) - 1;
    uint32_t x[16];
    salsa_block(x, seed, tid);
    u64 tries = 1;
    u64 *dst = out + (size_t) limb * n + g * 8;
    for (int k = 0; k < 8; k++) {
        u64 r = (u64) x[2 * k] | ((u64) x[2 * k + 1] << 32);
        while (r > max_multiple) {
            salsa_block(x, seed, tid + tries * n * (u64) limbs);
            tries++;
            r = (u64) x[2 * k] | ((u64) x[2 * k + 1] << 32);
        }
        dst[k] = r % m.q;
    }
}

// element-wise steps of encrypt_zero_* (polymath.cu:349-411): out = +-(a * b + e) [* then e scaled by t first, BGV]
//   NEG = 1: multiply_and_add_negate_rns_poly (symmetric: -(a s + e));  NEG = 0: multiply_and_add_rns_poly (u pk + e)
// (the keys live at the key level, [size_QP][n]; a data level uses their first l limbs, so one limb stride serves all)
template<bool NEG>
__global__ void __launch_bounds__(EW_THREADS) k_enc_fma(u64 *out, const u64 *a, const u64 *b, const u64 *e, const Modulus *mod,
                                                        size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const int i = blockIdx.y;
    const Modulus m = mod[i];
    const size_t o = (size_t) i * n + c;
    const u64 v = add_mod(mul_mod(a[o], b[o], m), e[o], m.q);
    out[o] = NEG ? (v ? m.q - v : 0) : v;
}

// multiply_scalar_rns_poly by the plain modulus (BGV noise t * e, secretkey.cu:60-66, 271-276)
__global__ void __launch_bounds__(EW_THREADS) k_scale_by(u64 *inout, u64 scalar, const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const Modulus m = mod[blockIdx.y];
    const size_t o = (size_t) blockIdx.y * n + c;
    inout[o] = mul_mod(inout[o], barrett64(scalar, m), m);
}

// multiply_temp_mod_and_add_rns_poly (polymath.cu:318-338): digit d of a key-switching key gets P * new_key added to the
// limbs of its own digit: key[d][0][j] += (P mod q_j) * new_key[j] for j in [d alpha, (d + 1) alpha)
__global__ void __launch_bounds__(EW_THREADS) k_kswitch_target(u64 *const *digits, const u64 *new_key, const u64 *p_mod_q,
                                                               const Modulus *mod, size_t n, int alpha) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const int j = blockIdx.y;
    const Modulus m = mod[j];
    const size_t o = (size_t) j * n + c;
    u64 *key = digits[j / alpha];
    key[o] = add_mod(key[o], mul_mod(new_key[o], p_mod_q[j], m), m.q);
}

// ---- plaintext operands (add_plain / sub_plain / multiply_plain, evaluate.cu:1106-1340, and the last step of encryption) --

// bfv_add_timesQ_overt_kernel / bfv_sub_timesQ_overt_kernel (polymath.cu:413-461): c0 +-= [m * (-Q_l mod t)]_t * t^-1 mod q_i
// (t may be any modulus here, not only a table row: plain 128-bit remainder)
template<bool SUB>
__global__ void __launch_bounds__(EW_THREADS) k_bfv_add_plain(u64 *ct0, const u64 *plain, u64 neg_q_mod_t, u64 t,
                                                              const u64 *tinv_mod_q, const Modulus *mod, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const int i = blockIdx.y;
    const Modulus m = mod[i];
    const u64 scaled = (u64) (((unsigned __int128) plain[c] * neg_q_mod_t) % t);
    const size_t o = (size_t) i * n + c;
    const u64 v = mul_mod(scaled, tinv_mod_q[i], m);
    ct0[o] = SUB ? sub_mod(ct0[o], v, m.q) : add_mod(ct0[o], v, m.q);
}

// a plaintext [n] (residues mod t) under every limb.  ABS = false: reduced mod q_i (nwt_2d_radix8_forward_modup_fuse's
// load; encrypt_symmetric copies it unreduced, the same thing whenever t < q_i).  ABS = true: abs_plain_rns_poly
// (polymath.cu:645-664), values from (t + 1) / 2 up stand for negative numbers and move to q_i - (t - value)
template<bool ABS>
__global__ void __launch_bounds__(EW_THREADS) k_lift_plain(u64 *out, const u64 *plain, const Modulus *mod, size_t n, u64 t) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const Modulus m = mod[blockIdx.y];
    u64 v = plain[c];
    if (ABS) v = v >= ((t + 1) >> 1) ? v + (m.q - t) : v;
    else v = barrett64(v, m);
    out[(size_t) blockIdx.y * n + c] = v;
}

// multiply_scalar_and_add_rns_poly / multiply_scalar_and_sub_rns_poly / multiply_scalar_rns_poly (polymath.cu:210-285):
// out = a +- scalar * b, or scalar * b alone when a is null
template<bool SUB>
__global__ void __launch_bounds__(EW_THREADS) k_axpy(u64 *out, const u64 *a, const u64 *b, u64 scalar, const Modulus *mod,
                                                     size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t c = (size_t) blockIdx.x * EW_THREADS + threadIdx.x;
    const Modulus m = mod[blockIdx.y];
    const size_t o = (size_t) blockIdx.y * n + c;
    const u64 v = mul_mod(b[o], barrett64(scalar, m), m);
    out[o] = !a ? v : SUB ? sub_mod(a[o], v, m.q) : add_mod(a[o], v, m.q);
}

}   // namespace pfhe
