/*
 * pfhe_b200.h -- C-ABI of the B200-native RNS polynomial-arithmetic engine (libpfhe_b200.so).
 *
 * Drop-in boundary for the hot path of encryptorion-lab/phantom-fhe: every entry
 * point replaces one reference function and keeps its argument meaning; the reference-side binding a
 * maintainer would add is shown in INTEGRATION.md.  Conventions (same as the reference's kernel-level
 * interface, include/ntt.cuh:157-226, include/rns.cuh:156-205, include/evaluate.cuh:18-32):
 *
 *   - all data pointers are DEVICE pointers to uint64 words laid out [poly][limb][coeff]
 *     (include/ciphertext.h:15-25); `stream` is a cudaStream_t passed as void*;
 *   - calls enqueue work on `stream` and return without synchronising;
 *   - `chain_index` follows PhantomCiphertext::chain_index(): 1 = top data level, level c has
 *     size_Q - (c - 1) limbs (src/context.cu:145-159, src/eval_key_switch.cu:125,130);
 *   - switching keys are passed exactly as PhantomRelinKey::public_keys_ptr() (include/secretkey.h:102-127):
 *     a DEVICE array of dnum device pointers, each to a [2][size_QP][N] buffer in NTT form;
 *   - errors: the reference throws std::invalid_argument / std::logic_error / std::runtime_error("CUDA
 *     Runtime Error"); here every call returns a status code and pfhe_last_error() holds the message so the
 *     C++ shim can re-throw the same exception type.
 *   - the *_host entry points take HOST buffers and perform the host<->device copies on `stream`
 *     (used for end-to-end measurement; pass pinned memory for asynchronous copies).
 *
 * There is no CPU fallback: without a CUDA device every call fails with PFHE_ERR_CUDA.
 */
#ifndef PFHE_B200_H
#define PFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfhe_engine pfhe_engine;

enum {
    PFHE_OK = 0,
    PFHE_ERR_INVALID_ARGUMENT = 1, /* reference: std::invalid_argument */
    PFHE_ERR_LOGIC = 2,            /* reference: std::logic_error      */
    PFHE_ERR_CUDA = 3,             /* reference: std::runtime_error("CUDA Runtime Error") */
    PFHE_ERR_UNSUPPORTED = 4
};

enum { PFHE_SCHEME_BGV = 1, PFHE_SCHEME_BFV = 2, PFHE_SCHEME_CKKS = 3 }; /* host/encryptionparams.h:14-22 */

/* thread-local message of the last failing call */
const char *pfhe_last_error(void);

/* ---- context (replaces PhantomContext::PhantomContext, src/context.cu:121-232, for this path) -------- */
/* CoeffModulus::Create (src/host/modulus.cu:79-110): primes for the given bit sizes, host output */
int pfhe_create_primes(uint64_t n, const int *bit_sizes, int count, uint64_t *primes_out);
/* primes = key-level chain: size_QP - size_P data primes followed by size_P special primes
 * (EncryptionParameters::set_coeff_modulus + set_special_modulus_size, host/encryptionparams.h:92-135);
 * galois_elts as EncryptionParameters::set_galois_elts (may be NULL/0) */
int pfhe_engine_create(pfhe_engine **out, int scheme, uint64_t n, const uint64_t *primes, int size_QP, int size_P,
                       uint64_t plain_modulus, const uint32_t *galois_elts, int n_galois);
void pfhe_engine_destroy(pfhe_engine *e);
/* EncryptionParameters::set_mul_tech (include/host/encryptionparams.h:25-35,57-69): 1 = behz, 2 = hps (the default
 * of a BFV engine, like the reference), 3 = hps_overq, 4 = hps_overq_leveled (the level-dropping forms take the number of
 * levels from the caller: pfhe_find_levels_to_drop, pfhe_multiply_leveled, ... below).  BFV engines only. */
int pfhe_engine_set_mul_tech(pfhe_engine *e, int mul_tech);
uint64_t pfhe_poly_degree(const pfhe_engine *e);
int pfhe_size_QP(const pfhe_engine *e);
int pfhe_size_P(const pfhe_engine *e);
int pfhe_dnum(const pfhe_engine *e, size_t chain_index); /* beta at that level, src/rns.cu:152 */
/* The engine's Galois elements in key order (PhantomGaloisTool::galois_elts(), include/galois.cuh:140): the list given to
 * pfhe_engine_create, or -- when that list was empty -- the reference's default get_elts_all() (src/galois.cu:41-65:
 * 2N-1, then 5^(2^i) and 5^-(2^i) for i < log2(N)-1).  Writes at most `capacity` entries, returns the count (-1: no engine). */
int pfhe_galois_elts(const pfhe_engine *e, uint32_t *elts_out, int capacity);
/* get_elt_from_step (include/galois.cuh:16-49) */
int pfhe_galois_elt_from_step(int step, uint64_t n, uint32_t *elt_out);

/* ---- NTT (replaces include/ntt.cuh:172-226) --------------------------------------------------------- */
/* Addressing of every launcher in this section is the reference's: `inout` / `in` / `out` point at limb 0 of a buffer
 * [..][N]; the call works on limbs [start_modulus_idx, start_modulus_idx + coeff_modulus_size) of it, limb i with the
 * constants of table entry i (fntt_2d.cu:35-40: data_ptr = inout + twr_idx * n, twr_idx = i + start_mod_idx).
 * nwt_2d_radix8_forward_inplace (src/ntt/fntt_2d.cu:620-653) */
int pfhe_ntt_forward_inplace(pfhe_engine *e, uint64_t *inout, size_t coeff_modulus_size, size_t start_modulus_idx,
                             void *stream);
/* nwt_2d_radix8_backward_inplace (src/ntt/intt_2d.cu:724-757) */
int pfhe_ntt_backward_inplace(pfhe_engine *e, uint64_t *inout, size_t coeff_modulus_size, size_t start_modulus_idx,
                              void *stream);
/* same transforms for n_poly polynomials laid out [n_poly][coeff_modulus_size][N] in ONE launch pair (the
 * reference loops over polynomials on the host, e.g. src/rns.cu:1165-1183) */
int pfhe_ntt_forward_inplace_batch(pfhe_engine *e, uint64_t *inout, size_t n_poly, size_t coeff_modulus_size,
                                   size_t start_modulus_idx, void *stream);
int pfhe_ntt_backward_inplace_batch(pfhe_engine *e, uint64_t *inout, size_t n_poly, size_t coeff_modulus_size,
                                    size_t start_modulus_idx, void *stream);
/* nwt_2d_radix8_backward (out of place, src/ntt/ntt_modup.cu:320-354) */
int pfhe_ntt_backward(pfhe_engine *e, uint64_t *out, const uint64_t *in, size_t coeff_modulus_size,
                      size_t start_modulus_idx, void *stream);
/* nwt_2d_radix8_forward_inplace_include_special_mod (src/ntt/fntt_2d.cu:694-736): limbs
 * [start, start+count) of a packed Ql u P buffer, P limbs use table rows size_QP - size_P + ... */
int pfhe_ntt_forward_inplace_include_special_mod(pfhe_engine *e, uint64_t *inout, size_t coeff_modulus_size,
                                                 size_t start_modulus_idx, size_t size_QP, size_t size_P,
                                                 void *stream);
int pfhe_ntt_backward_inplace_include_special_mod(pfhe_engine *e, uint64_t *inout, size_t coeff_modulus_size,
                                                  size_t start_modulus_idx, size_t size_QP, size_t size_P,
                                                  void *stream);


/* ---- the remaining launchers of include/ntt.cuh:172-226, same addressing ----------------------------------------
 * `table` names the DNTTTable the reference passes: PFHE_TABLE_RNS = context.gpu_rns_tables() (key-level primes),
 * PFHE_TABLE_BSK = rns_tool.gpu_Bsk_tables() (B_0.., m_sk; BFV), PFHE_TABLE_QLRL = gpu_QlRl_tables() (Q then R; BFV HPS),
 * PFHE_TABLE_PLAIN = gpu_plain_tables() (the batching plain modulus).  The Bsk and QlRl tables belong to a level's DRNSTool:
 * pass family | chain_index << 8 for a level other than the first data level. */
enum { PFHE_TABLE_RNS = 0, PFHE_TABLE_BSK = 1, PFHE_TABLE_QLRL = 2, PFHE_TABLE_PLAIN = 3 };
int pfhe_table_size(const pfhe_engine *e, int table);
/* modulus of entry idx of that table (0 if out of range) */
uint64_t pfhe_table_modulus(const pfhe_engine *e, int table, size_t idx);
/* nwt_2d_radix8_forward_inplace / _backward_inplace / _backward on any table (fntt_2d.cu:620-653, intt_2d.cu:724-757,
 * ntt_modup.cu:320-354) */
int pfhe_nwt_2d_radix8_forward_inplace(pfhe_engine *e, int table, uint64_t *inout, size_t coeff_modulus_size,
                                       size_t start_modulus_idx, void *stream);
int pfhe_nwt_2d_radix8_backward_inplace(pfhe_engine *e, int table, uint64_t *inout, size_t coeff_modulus_size,
                                        size_t start_modulus_idx, void *stream);
int pfhe_nwt_2d_radix8_backward(pfhe_engine *e, int table, uint64_t *out, const uint64_t *in, size_t coeff_modulus_size,
                                size_t start_modulus_idx, void *stream);
/* nwt_2d_radix8_forward_inplace_fuse_moddown (ntt_moddown.cu:222-261): ct[i] = (cx[i] - NTT(delta[i])) * bigPInv_mod_q[i];
 * delta is clobbered */
int pfhe_nwt_2d_radix8_forward_inplace_fuse_moddown(pfhe_engine *e, uint64_t *ct, const uint64_t *cx,
                                                    const uint64_t *bigPInv_mod_q, const uint64_t *bigPInv_mod_q_shoup,
                                                    uint64_t *delta, size_t coeff_modulus_size, size_t start_modulus_idx,
                                                    void *stream);
/* nwt_2d_radix8_forward_inplace_include_temp_mod (fntt_2d.cu:655-692): limb coeff_modulus_size - 1 uses table entry
 * total_modulus_size - 1 */
int pfhe_nwt_2d_radix8_forward_inplace_include_temp_mod(pfhe_engine *e, int table, uint64_t *inout,
                                                        size_t coeff_modulus_size, size_t start_modulus_idx,
                                                        size_t total_modulus_size, void *stream);
/* nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range (ntt_modup.cu:610-657) */
int pfhe_nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(pfhe_engine *e, uint64_t *inout,
                                                                         size_t coeff_modulus_size, size_t start_modulus_idx,
                                                                         size_t size_QP, size_t size_P,
                                                                         size_t excluded_range_start,
                                                                         size_t excluded_range_end, void *stream);
/* nwt_2d_radix8_forward_modup_fuse (ntt_keyswitch_old.cu:225-265): out[i] = NTT of in[i] with the constants of table entry
 * modulus_index for every limb of the window (inputs below that modulus, e.g. a plaintext modulo t lifted under q_i) */
int pfhe_nwt_2d_radix8_forward_modup_fuse(pfhe_engine *e, uint64_t *out, const uint64_t *in, size_t modulus_index,
                                          size_t coeff_modulus_size, size_t start_modulus_idx, void *stream);
/* nwt_2d_radix8_backward_scale / _backward_inplace_scale (ntt_modup.cu:356-393, intt_2d.cu:759-794): inverse transform
 * times scale[i]; scale / scale_shoup are device arrays indexed by limb */
int pfhe_nwt_2d_radix8_backward_scale(pfhe_engine *e, int table, uint64_t *out, const uint64_t *in,
                                      size_t coeff_modulus_size, size_t start_modulus_idx, const uint64_t *scale,
                                      const uint64_t *scale_shoup, void *stream);
int pfhe_nwt_2d_radix8_backward_inplace_scale(pfhe_engine *e, int table, uint64_t *inout, size_t coeff_modulus_size,
                                              size_t start_modulus_idx, const uint64_t *scale, const uint64_t *scale_shoup,
                                              void *stream);
/* nwt_2d_radix8_backward_inplace_include_temp_mod_scale (intt_2d.cu:835-873): scale indexed by table entry */
int pfhe_nwt_2d_radix8_backward_inplace_include_temp_mod_scale(pfhe_engine *e, int table, uint64_t *inout,
                                                               size_t coeff_modulus_size, size_t start_modulus_idx,
                                                               size_t total_modulus_size, const uint64_t *scale,
                                                               const uint64_t *scale_shoup, void *stream);

/* ---- DBaseConverter::bConv_BEHZ / bConv_BEHZ_var1 / bConv_HPS (include/rns_bconv.cuh:62-68, src/rns_bconv.cu:212-246,
 * 354-372).  The converter is named by its bases: ibase / obase are (table, entry) pairs flattened as table * 65536 + entry.
 * src = [ni][N], dst = [no][N] (coefficient form).  BEHZ: y_i = x_i qhat_i^-1, dst_j = sum y_i (qhat_i mod p_j);
 * var1: y_i = x_i (-P qhat_i^-1), dst_j = sum y_i (q_i^-1 mod p_j); HPS: BEHZ minus round(sum y_i / q_i) * Q mod p_j with
 * the reference's FP64 accumulation order. */
enum { PFHE_BCONV_BEHZ = 0, PFHE_BCONV_BEHZ_VAR1 = 1, PFHE_BCONV_HPS = 2 };
int pfhe_bconv(pfhe_engine *e, int mode, const uint32_t *ibase, int ni, const uint32_t *obase, int no, uint64_t *dst,
               const uint64_t *src, void *stream);
/* DRNSTool::moddown (src/rns_bconv.cu:712-761): ct_i[l][N] = moddown of cx_i[l + size_P][N]; CKKS / BGV input in NTT form,
 * BFV input in coefficient form.  cx_i is clobbered. */
int pfhe_moddown(pfhe_engine *e, size_t chain_index, uint64_t *ct_i, uint64_t *cx_i, void *stream);
/* DRNSTool::divide_and_round_q_last (src/rns.cu:1082-1126, coefficient form), divide_and_round_q_last_ntt (:1160-1184,
 * NTT form) and mod_t_and_divide_q_last_ntt (:1186-1235, NTT form with the plain-modulus correction): src = [size][l][N] at
 * chain_index, dst = [size][l-1][N].  Unlike the reference, src is left intact. */
int pfhe_divide_and_round_q_last(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t cipher_size,
                                 uint64_t *dst, void *stream);
int pfhe_divide_and_round_q_last_ntt(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t cipher_size,
                                     uint64_t *dst, void *stream);
int pfhe_mod_t_and_divide_q_last_ntt(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t cipher_size,
                                     uint64_t *dst, void *stream);
/* add_to_ct_kernel (src/rns_bconv.cu:763-769): ct[l][N] += cx[l][N] */
int pfhe_add_to_ct(pfhe_engine *e, uint64_t *ct, const uint64_t *cx, size_t size_Ql, void *stream);

/* ---- dyadic kernels (replace the __global__ symbols evaluate.cu launches, include/polymath.cuh) ------ */
/* tensor_prod_2x2_rns_poly (src/polymath.cu:463-498); result = [3][l][n], may alias operand1 */
int pfhe_tensor_prod_2x2(pfhe_engine *e, const uint64_t *operand1, const uint64_t *operand2, uint64_t *result,
                         size_t coeff_mod_size, void *stream);
/* tensor_square_2x2_rns_poly (:500-532) */
int pfhe_tensor_square_2x2(pfhe_engine *e, const uint64_t *operand, uint64_t *result, size_t coeff_mod_size,
                           void *stream);
/* add_rns_poly / sub_rns_poly / multiply_rns_poly / negate_rns_poly (:17-173) on `coeff_mod_size` limbs */
int pfhe_add_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *result, size_t coeff_mod_size,
                      void *stream);
int pfhe_sub_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *result, size_t coeff_mod_size,
                      void *stream);
int pfhe_multiply_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *result,
                           size_t coeff_mod_size, void *stream);
int pfhe_negate_rns_poly(pfhe_engine *e, const uint64_t *a, uint64_t *result, size_t coeff_mod_size, void *stream);

/* ---- key switching (replaces DRNSTool::modup / moddown_from_NTT, key_switch_inner_prod, keyswitch_inplace) */
/* DRNSTool::modup (src/rns_bconv.cu:530-628): dst = [beta][l+alpha][n] */
int pfhe_modup(pfhe_engine *e, size_t chain_index, uint64_t *dst, const uint64_t *cks, void *stream);
/* key_switch_inner_prod (src/eval_key_switch.cu:71-92): p_cx = [2][l+alpha][n] */
int pfhe_key_switch_inner_prod(pfhe_engine *e, size_t chain_index, uint64_t *p_cx, const uint64_t *p_t_mod_up,
                               const uint64_t *const *rlk, void *stream);
/* DRNSTool::moddown_from_NTT (src/rns_bconv.cu:776-828): ct_i = [l][n] result, cx_i = [l+alpha][n] (clobbered) */
int pfhe_moddown_from_ntt(pfhe_engine *e, size_t chain_index, uint64_t *ct_i, uint64_t *cx_i, void *stream);
/* keyswitch_inplace (src/eval_key_switch.cu:95-182): encrypted[2][l][n] += key-switch of c2[l][n] */
int pfhe_keyswitch_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted, const uint64_t *c2,
                           const uint64_t *const *relin_keys, void *stream);

/* ---- scheme level (replace the namespace phantom free functions of include/evaluate.cuh:37-245) --------- */
/* multiply_inplace + relinearize_inplace (src/evaluate.cu:1029-1104,1342-1374):
 * encrypted1 = [2][l][n] in/out, encrypted2 = [2][l][n].  CKKS/BGV: NTT form (bgv_ckks_multiply :345-397);
 * BFV: coefficient form, BEHZ multiplication (bfv_multiply_behz :451-548, mul_tech_type::behz) */
int pfhe_multiply_and_relin_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted1,
                                    const uint64_t *encrypted2, const uint64_t *const *relin_keys, void *stream);
/* same, result written to a separate [2][l][n] buffer (no operand copy; what the in-place form's resize does in
 * the reference, include/ciphertext.h:44-72, is left to the caller) */
int pfhe_multiply_and_relin(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted1,
                            const uint64_t *encrypted2, uint64_t *destination, const uint64_t *const *relin_keys,
                            void *stream);
/* PhantomCKKSEncoder::encode_internal (src/ckks.cu:66-135; special inverse FFT src/fft.cu; decompose_array
 * src/rns_base.cu:49-173): `values` = count <= N/2 complex numbers on the device (re, im interleaved, 16-byte aligned),
 * `plain` = [l][N] residues in NTT form at chain_index with the given scale.  Same floating-point operations in the same
 * order as the reference's kernels: the same words.  Synchronises `stream` once, like the reference (the magnitude of the
 * encoded coefficients selects the decomposition and is validated against the modulus); coefficients above 128 bits
 * (the reference's slow multi-word path) are refused. */
int pfhe_ckks_encode(pfhe_engine *e, size_t chain_index, const double *values, size_t count, double scale,
                     uint64_t *plain, void *stream);
/* ---- samplers, key generation, encryption ------------------------------------------------
 * The reference draws a fresh 64-byte seed from std::random_device for every random polynomial (random_bytes,
 * include/prng.cuh:10-32) and expands it on the device with its Salsa20-core generator (src/prng.cu:17-133).  Here the
 * caller supplies the seeds (`seed*` = 64 bytes on the host each); given the seed, every polynomial below is the one the
 * reference's kernels produce.  All buffers are device memory; nothing synchronises unless stated.
 *
 * sample_ternary_poly / sample_error_poly / sample_uniform_poly (src/prng.cu:142-244): kind 0 / 1 / 2, out = [limbs][N]
 * over the first `limbs` key primes. */
int pfhe_sample_poly(pfhe_engine *e, int kind, size_t limbs, const uint8_t *seed, uint64_t *out, void *stream);
/* PhantomSecretKey::gen_secretkey (src/secretkey.cu:345-378): secret_key = [size_QP][N], NTT form (secret_key_array()) */
int pfhe_gen_secretkey(pfhe_engine *e, const uint8_t *seed, uint64_t *secret_key, void *stream);
/* PhantomSecretKey::encrypt_zero_symmetric (src/secretkey.cu:232-295): ct = [2][l][N] = (-(a s + e), a) (e times t for BGV),
 * NTT form (BFV at a data level: coefficient form).  chain_index 0 = the key level (l = size_QP): that is gen_publickey
 * (src/secretkey.cu:380-392) and each digit of a key-switching key. */
int pfhe_encrypt_zero_symmetric(pfhe_engine *e, size_t chain_index, const uint64_t *secret_key, const uint8_t *seed_a,
                                const uint8_t *seed_e, uint64_t *ct, void *stream);
/* PhantomPublicKey::encrypt_zero_asymmetric_internal (src/secretkey.cu:10-128): public_key = [2][size_QP][N] (key level, NTT
 * form), ct = [2][size_Q][N]: (u pk + e) at the key level divided by P.  chain_index must be 1 (the reference's own
 * mod-down step addresses the first data level only). */
int pfhe_encrypt_zero_asymmetric(pfhe_engine *e, size_t chain_index, const uint64_t *public_key, const uint8_t *seed_u,
                                 const uint8_t *seed_e, uint64_t *ct, void *stream);
/* PhantomSecretKey::generate_one_kswitch_key (src/secretkey.cu:297-343): digits = HOST array of dnum = size_Q / size_P device
 * buffers [2][size_QP][N]; new_key = [size_QP][N] NTT form (s^2 for gen_relinkey, the rotated key for a Galois key);
 * seeds = dnum pairs (a, e) of 64 bytes.  Synchronises the stream once. */
int pfhe_gen_kswitch_key(pfhe_engine *e, const uint64_t *new_key, const uint64_t *secret_key, const uint8_t *seeds,
                         uint64_t *const *digits, void *stream);
/* the secret key under a Galois automorphism (create_galois_keys, src/secretkey.cu:443-451): rotated = [size_QP][N] */
int pfhe_galois_secret_key(pfhe_engine *e, const uint64_t *secret_key, uint32_t galois_elt, uint64_t *rotated, void *stream);
/* PhantomGaloisTool::apply_galois_ntt (src/galois.cu:86-102, NTT form: an index permutation, the element must be one of the
 * context's) and apply_galois (src/galois.cu:20-39,104-120, coefficient form: x^i -> +-x^(i elt mod 2N), any odd element):
 * result[limb] = automorphism of operand[limb] over the first coeff_mod_size key primes; not in place. */
int pfhe_apply_galois_ntt(pfhe_engine *e, const uint64_t *operand, size_t coeff_mod_size, uint32_t galois_elt, uint64_t *result,
                          void *stream);
int pfhe_apply_galois(pfhe_engine *e, const uint64_t *operand, size_t coeff_mod_size, uint32_t galois_elt, uint64_t *result,
                      void *stream);
/* the last step of encrypt_symmetric / encrypt_asymmetric (src/secretkey.cu:130-190, 463-530): ct[0] += plaintext.
 * BFV: multiply_add_plain_with_scaling_variant (src/scalingvariant.cu:10-34), plain = [N] mod t; CKKS: plain = [l][N] NTT
 * form; BGV: plain = [N] mod t, lifted to every limb and transformed. */
int pfhe_encrypt_add_plain(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, void *stream);
/* add_plain_inplace / sub_plain_inplace (src/evaluate.cu:1106-1224): ct[0] +-= plaintext, plaintext shapes as above; BGV
 * multiplies the lifted plaintext by the ciphertext's correction factor (multiply_scalar_and_add / _sub_rns_poly). */
int pfhe_add_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t correction_factor,
                           void *stream);
int pfhe_sub_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t correction_factor,
                           void *stream);
/* multiply_plain_inplace (src/evaluate.cu:1226-1340): every polynomial of ct = [size][l][N] times the plaintext (BFV:
 * multiply_plain_normal through NTT form and back; CKKS: multiply_plain_ntt; BGV: lifted plaintext). */
int pfhe_multiply_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, size_t size, const uint64_t *plain,
                                void *stream);
/* multiply_scalar_rns_poly (src/polymath.cu:210-228) on `size` polynomials of coeff_mod_size limbs: the correction-factor
 * balancing of BGV add / sub (src/evaluate.cu:148-165). */
int pfhe_multiply_scalar_rns_poly(pfhe_engine *e, uint64_t *inout, size_t size, uint64_t scalar, size_t coeff_mod_size,
                                  void *stream);
/* PhantomCKKSEncoder::decode_internal (src/ckks.cu:137-190; compose_array src/rns_base.cu:174-258; special forward FFT
 * src/fft.cu:90-218,352-384): `plain` = [l][N] residues in NTT form with the given scale, `values` = N/2 complex numbers
 * on the device.  Same operations in the same order as the reference's kernels; up to 32 limbs.  Does not synchronise. */
int pfhe_ckks_decode(pfhe_engine *e, size_t chain_index, const uint64_t *plain, double scale, double *values,
                     void *stream);

/* PhantomBatchEncoder::encode / decode for BFV / BGV (src/batchencoder.cu:62-118): `values` = count <= N slot values on
 * the device (the reference copies its std::vector there first), `plain` = the [N] plaintext polynomial mod t,
 * coefficient form.  Needs a batching plain modulus (prime, 1 mod 2N): the engine then keeps NTT tables for it
 * (gpu_plain_tables).  Decode writes all N slots. */
int pfhe_batch_encode(pfhe_engine *e, const uint64_t *values, size_t count, uint64_t *plain, void *stream);
int pfhe_batch_decode(pfhe_engine *e, const uint64_t *plain, uint64_t *values, void *stream);

/* PhantomSecretKey::decrypt (src/secretkey.cu:533-723: ckks_decrypt / bgv_decrypt / bfv_decrypt) -- the step after the
 * hot path.  encrypted = [size][l][n] at chain_index (CKKS / BGV: NTT form; BFV: coefficient form);
 * secret_key_array = PhantomSecretKey::secret_key_array(): the powers s, s^2, .. s^(size-1) of the secret key in NTT form,
 * [size-1][size_QP][n] at the key level (compute_secret_key_array, secretkey.cu:247-295).  destination: CKKS [l][n] in
 * NTT form (the plaintext the decoder takes); BGV [n] = exact_convert_array to t (times correction_factor^-1); BFV [n] =
 * behz_ / hps_decrypt_scale_and_round by the engine's mul_tech.  correction_factor: BGV only, 1 otherwise. */
int pfhe_decrypt(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted, size_t size,
                 const uint64_t *secret_key_array, uint64_t correction_factor, uint64_t *destination, void *stream);

/* ---- BFV with mul_tech_type::hps_overq_leveled (pfhe_engine_set_mul_tech(e, 4)) ---------------------------------
 * The reference decides per operation how many levels to drop from the ciphertexts' noise-scale degree
 * (FindLevelsToDrop, evaluate.cu:550-643; the degree lives on PhantomCiphertext, ciphertext.h:21,97-131): the caller
 * passes multiplicative_depth = max(noiseScaleDeg) - 1 exactly like bfv_multiply_hps (:680-690) / keyswitch_inplace
 * (eval_key_switch.cu:113-123) do, and hands the result to the three leveled entry points.  All buffers are at the
 * first data level ([.][size_Q][n], coefficient form).  pfhe_multiply / pfhe_multiply_and_relin with this mul_tech
 * run with 0 levels dropped (= hps_overq). */
int pfhe_find_levels_to_drop(pfhe_engine *e, size_t multiplicative_depth, int is_key_switch, int is_asymmetric,
                             int *levels);
/* bfv_multiply_hps (evaluate.cu:647-801) with levels_dropped levels: destination = [3][size_Q][n] */
int pfhe_multiply_leveled(pfhe_engine *e, const uint64_t *encrypted1, const uint64_t *encrypted2, uint64_t *destination,
                          int levels_dropped, void *stream);
/* bfv_mul_relin_hps (evaluate.cu:819-1026): destination = [2][size_Q][n] */
int pfhe_multiply_and_relin_leveled(pfhe_engine *e, const uint64_t *encrypted1, const uint64_t *encrypted2,
                                    uint64_t *destination, const uint64_t *const *relin_keys, int levels_dropped,
                                    void *stream);
/* keyswitch_inplace (eval_key_switch.cu:95-182) with levels dropped: encrypted = [2][size_Q][n] in/out, c2 = [size_Q][n] */
int pfhe_keyswitch_leveled_inplace(pfhe_engine *e, uint64_t *encrypted, const uint64_t *c2, const uint64_t *const *keys,
                                   int levels_dropped, void *stream);

/* fnwt_1d[_opt] / inwt_1d[_opt] (include/ntt.cuh:157-170, src/ntt/ntt_1d.cu:146-292): single-block negacyclic
 * transforms for dim <= 2048 on CALLER-SUPPLIED device tables in the reference's order (twiddles[bitrev(i)] = psi^i with
 * separate Shoup arrays, itwiddles[1] already times n^-1), modulus = DModulus array {value, const_ratio[0], const_ratio[1]}
 * (include/ntt.cuh:6-32).  Limb i of the call is absolute index start_modulus_idx + i in inout, the tables, modulus and
 * scalar.  Forward: natural -> bit-reversed, output in [0, q).  Inverse: lower half times scalar[idx] (pass n^-1 for
 * the plain inverse), upper half unscaled, output in [0, q).  No engine handle: nothing but the tables is needed. */
int pfhe_fnwt_1d(uint64_t *inout, const uint64_t *twiddles, const uint64_t *twiddles_shoup, const uint64_t *modulus,
                 size_t dim, size_t coeff_modulus_size, size_t start_modulus_idx, void *stream);
int pfhe_inwt_1d(uint64_t *inout, const uint64_t *itwiddles, const uint64_t *itwiddles_shoup, const uint64_t *modulus,
                 const uint64_t *scalar, const uint64_t *scalar_shoup, size_t dim, size_t coeff_modulus_size,
                 size_t start_modulus_idx, void *stream);
/* multiply_inplace for ciphertext sizes other than 2 x 2 (bgv_ckks_multiply's tensor_prod_mxn_rns_poly branch,
 * evaluate.cu:382-386, polymath.cu:546-594): encrypted1 = [size1][l][n], encrypted2 = [size2][l][n], destination =
 * [size1 + size2 - 1][l][n] (may alias encrypted1, as the reference's in-place form does); sizes up to 8.  CKKS / BGV */
int pfhe_multiply_sizes(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted1, size_t size1,
                        const uint64_t *encrypted2, size_t size2, uint64_t *destination, void *stream);
/* `count` independent multiply_and_relin ops on device buffers (arrays of `count` device pointers held in HOST
 * memory).  The reference runs independent ciphertexts on independent host threads / cudaStreamPerThread
 * (src/CMakeLists.txt:39, evaluate.cu:1079); here the ops go round-robin over pfhe_engine_lanes() internal streams,
 * each with its own workspace, forked from and joined back into `stream`.  destination[i] must not alias an operand.
 * All schemes (BFV: by the engine's mul_tech, 0 levels dropped for hps_overq_leveled). */
int pfhe_multiply_and_relin_batch(pfhe_engine *e, size_t chain_index, const uint64_t *const *encrypted1,
                                  const uint64_t *const *encrypted2, uint64_t *const *destination, size_t count,
                                  const uint64_t *const *relin_keys, void *stream);
/* number of lanes of the batched entry points (1..4, default 2; 1 = strictly one op at a time) */
int pfhe_engine_set_lanes(pfhe_engine *e, int lanes);
int pfhe_engine_lanes(const pfhe_engine *e);
/* multiply_inplace alone: destination = [3][l][n] (BFV: must not alias an operand) */
int pfhe_multiply(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted1, const uint64_t *encrypted2,
                  uint64_t *destination, void *stream);
/* relinearize_inplace: encrypted = [3][l][n], the first two polynomials are updated */
int pfhe_relinearize_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted,
                             const uint64_t *const *relin_keys, void *stream);
/* apply_galois_inplace (src/evaluate.cu:1567-1630): galois_key = relin-key pointer array of that element */
int pfhe_apply_galois_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted, uint32_t galois_elt,
                              const uint64_t *const *galois_key, void *stream);
/* rotate_inplace (src/evaluate.cu:1633-1668) for a step whose element is in the engine's Galois set */
/* `count` independent rotate_inplace calls (e.g. the steps of a baby-step / giant-step loop, each on its own copy):
 * encrypted[i] <- rotate(encrypted[i], steps[i]) with galois_keys[i] = get_relin_keys(idx_i).public_keys_ptr(),
 * interleaved over the engine's lanes like pfhe_multiply_and_relin_batch.  The ciphertexts must be distinct buffers. */
int pfhe_rotate_batch(pfhe_engine *e, size_t chain_index, uint64_t *const *encrypted, const int *steps,
                      const uint64_t *const *const *galois_keys, size_t count, void *stream);
int pfhe_rotate_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted, int step,
                        const uint64_t *const *galois_key, void *stream);
/* hoisting_inplace (src/evaluate.cu:1670-1865): encrypted <- sum_i rotate(encrypted, steps[i]) with one shared mod-up
 * and one mod-down; galois_keys[i] = PhantomGaloisKey::get_relin_keys(index of steps[i]).public_keys_ptr() (host
 * array of n_steps device pointer arrays).  All three schemes; BFV ciphertexts are in coefficient form (apply_galois on c0,
 * mod-up from and mod-down to coefficient form, evaluate.cu:1745-1747) and, like the reference's, at the first data level. */
int pfhe_hoisting_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted, const int *steps, size_t n_steps,
                          const uint64_t *const *const *galois_keys, void *stream);
/* the same under mul_tech hps_overq_leveled with `levels_dropped` levels dropped (pfhe_find_levels_to_drop with
 * is_key_switch = 1; evaluate.cu:1690-1701): encrypted = [2][size_Q][n] is scaled to Ql, rotated and summed there, expanded back */
int pfhe_hoisting_leveled_inplace(pfhe_engine *e, uint64_t *encrypted, const int *steps, size_t n_steps,
                                  const uint64_t *const *const *galois_keys, int levels_dropped, void *stream);
/* rescale_to_next (src/evaluate.cu:1545-1565): destination = [size][l-1][n] */
int pfhe_rescale_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted, size_t size,
                         uint64_t *destination, void *stream);
/* mod_switch_to_next (src/evaluate.cu:1505-1543): CKKS drops the last limb (:1429-1472); BFV divide_and_round_q_last
 * (src/rns.cu:1082-1126); BGV mod_t_and_divide_q_last_ntt (src/rns.cu:1186-1235) */
int pfhe_mod_switch_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *encrypted, size_t size,
                            uint64_t *destination, void *stream);

/* ---- host-buffer variants: H2D of the operands, the op, D2H of the result, all on `stream` --------------- */
int pfhe_multiply_and_relin_host(pfhe_engine *e, size_t chain_index, const uint64_t *h_encrypted1,
                                 const uint64_t *h_encrypted2, uint64_t *h_destination,
                                 const uint64_t *const *relin_keys, void *stream);
/* `count` independent ops; copies are double-buffered on internal streams so they overlap the kernels */
int pfhe_multiply_and_relin_host_batch(pfhe_engine *e, size_t chain_index, const uint64_t *const *h_encrypted1,
                                       const uint64_t *const *h_encrypted2, uint64_t *const *h_destination,
                                       size_t count, const uint64_t *const *relin_keys, void *stream);
int pfhe_rotate_host(pfhe_engine *e, size_t chain_index, const uint64_t *h_encrypted, int step,
                     uint64_t *h_destination, const uint64_t *const *galois_key, void *stream);
int pfhe_rescale_host(pfhe_engine *e, size_t chain_index, const uint64_t *h_encrypted, size_t size,
                      uint64_t *h_destination, void *stream);
int pfhe_ntt_forward_host(pfhe_engine *e, const uint64_t *h_in, uint64_t *h_out, size_t coeff_modulus_size,
                          size_t start_modulus_idx, void *stream);

/* Multi-GPU batches (the reference is single-device): lets the kernels of the calling thread's current
 * device dereference memory of `peer_device` over NVLink (cudaDeviceEnablePeerAccess; already enabled is not an error),
 * so that operand / destination pointers of the entry points above may point into another GPU's HBM of the same node
 * (CUDA IPC mappings of another rank's buffers included).  PFHE_ERR_UNSUPPORTED if the devices are not peers. */
int pfhe_enable_peer_access(int peer_device);
/* Zero-copy batches across the ranks of one node (one process per GPU): the owner of a device buffer exports it
 * (cudaIpcGetMemHandle of the allocation that contains `device_ptr`; `handle_out` = 64 opaque bytes, `offset_out` =
 * byte offset of device_ptr inside that allocation), another process maps it with its OWN device current
 * (cudaIpcOpenMemHandle with lazy peer access: the mapping is addressable by that device's kernels over NVLink) and
 * receives the address that corresponds to the exported pointer.  pfhe_ipc_close takes the address pfhe_ipc_open gave. */
int pfhe_ipc_export(const void *device_ptr, unsigned char handle_out[64], uint64_t *offset_out);
int pfhe_ipc_open(const unsigned char handle[64], uint64_t offset, void **mapped_out);
int pfhe_ipc_close(void *mapped, uint64_t offset);

/* number of kernel launches issued by this engine since creation (bench.py's gpu_launches) */
uint64_t pfhe_launch_count(const pfhe_engine *e);

#ifdef __cplusplus
}
#endif
#endif
