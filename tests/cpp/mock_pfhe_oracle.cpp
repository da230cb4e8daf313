// TEST INFRASTRUCTURE ONLY -- never part of the product.
//
// A stand-in for libpfhe_b200.so + the CUDA runtime that lets applications written against include/phantom_b200.hpp run on
// a machine without a GPU: "device" memory is host memory, and every C-ABI entry point the C++ mirror calls is answered by
// the CPU oracle (oracle/liboracle.so).  It exists so that the host logic of the mirror -- call sequences, buffer sizes,
// bookkeeping of levels, scales, correction factors and noise degrees -- is exercised by the CPU test suite
// (tests/test_host_logic.py::test_cpp_mirror_application_on_the_oracle).  The real library is what the GPU tests link.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pfhe_b200.h"
#include "../../oracle/fhe_oracle.h"

// ---- CUDA runtime stand-ins: host memory -------------------------------------------------------------------------------
extern "C" {
cudaError_t cudaMalloc(void **p, size_t bytes) {
    *p = std::malloc(bytes ? bytes : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) {
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "mock"; }
}

// ---- the engine handle -------------------------------------------------------------------------------------------------
struct pfhe_engine {
    orc_ctx *c = nullptr, *flat = nullptr;   // flat: every key prime a data limb (limb-wise ops at the key level)
    int scheme = 0, size_QP = 0, size_P = 0, size_Q = 0, mul_tech = 2;
    uint64_t n = 0, t = 0;
    std::vector<uint32_t> elts;
};
static thread_local std::string g_error;
static int fail(int code, const char *msg) {
    g_error = msg;
    return code;
}
static int limbs_at(const pfhe_engine *e, size_t chain_index) { return e->size_Q - (int) (chain_index - 1); }
static int dnum_of(const pfhe_engine *e) { return e->size_Q / e->size_P; }
// the mirror hands keys over as an array of per-digit pointers; the oracle reads [dnum][2][size_QP][n] in one piece
static std::vector<uint64_t> gather_key(const pfhe_engine *e, const uint64_t *const *digits) {
    const size_t words = (size_t) 2 * e->size_QP * e->n;
    std::vector<uint64_t> key(words * dnum_of(e));
    for (int d = 0; d < dnum_of(e); d++) std::memcpy(key.data() + d * words, digits[d], words * 8);
    return key;
}

extern "C" {
const char *pfhe_last_error(void) { return g_error.c_str(); }
int pfhe_create_primes(uint64_t n, const int *bits, int count, uint64_t *out) {
    return orc_create_primes(n, bits, count, out) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "failed to find enough qualifying primes");
}
int pfhe_galois_elt_from_step(int step, uint64_t n, uint32_t *elt) {
    *elt = orc_galois_elt_from_step(step, n);
    return PFHE_OK;
}
int pfhe_engine_create(pfhe_engine **out, int scheme, uint64_t n, const uint64_t *primes, int size_QP, int size_P, uint64_t t,
                       const uint32_t *elts, int n_elts) {
    auto *e = new pfhe_engine;
    e->c = orc_create(scheme, n, primes, size_QP, size_P, t);
    e->flat = orc_create(scheme, n, primes, size_QP, 0, t);
    if (!e->c || !e->flat) {
        delete e;
        return fail(PFHE_ERR_INVALID_ARGUMENT, "invalid parameters");
    }
    e->scheme = scheme, e->n = n, e->t = t, e->size_QP = size_QP, e->size_P = size_P, e->size_Q = size_QP - size_P;
    if (elts && n_elts > 0) e->elts.assign(elts, elts + n_elts);
    if (e->elts.empty()) {   // get_elts_all (reference src/galois.cu:41-65)
        const uint32_t m = (uint32_t) (2 * n);
        int logn = 0;
        while (((uint64_t) 1 << logn) < n) logn++;
        e->elts.push_back(m - 1);
        uint64_t pos = 5, neg = 1;
        while ((neg * 5) % m != 1) neg += 2;
        for (int i = 0; i < logn - 1; i++) {
            e->elts.push_back((uint32_t) pos), pos = (pos * pos) & (m - 1);
            e->elts.push_back((uint32_t) neg), neg = (neg * neg) & (m - 1);
        }
    }
    *out = e;
    return PFHE_OK;
}
int pfhe_galois_elts(const pfhe_engine *e, uint32_t *out, int cap) {
    if (out)
        for (int i = 0; i < cap && i < (int) e->elts.size(); i++) out[i] = e->elts[i];
    return (int) e->elts.size();
}
void pfhe_engine_destroy(pfhe_engine *e) {
    if (!e) return;
    orc_destroy(e->c), orc_destroy(e->flat);
    delete e;
}
int pfhe_engine_set_mul_tech(pfhe_engine *e, int m) {
    e->mul_tech = m;
    return PFHE_OK;
}
int pfhe_find_levels_to_drop(pfhe_engine *, size_t, int, int, int *levels) {
    // the rule itself is the engine's (and is tested there); here: no level, or PFHE_MOCK_DROP levels so that the
    // level-dropping branches of the mirror run too
    const char *v = std::getenv("PFHE_MOCK_DROP");
    *levels = v ? std::atoi(v) : 0;
    return PFHE_OK;
}
cudaError_t cudaMemsetAsync(void *p, int value, size_t bytes, cudaStream_t) {
    std::memset(p, value, bytes);
    return cudaSuccess;
}
int pfhe_apply_galois(pfhe_engine *e, const uint64_t *operand, size_t l, uint32_t elt, uint64_t *result, void *) {
    orc_apply_galois_coeff(e->flat, operand, result, (int) l, elt);
    return PFHE_OK;
}

// ---- limb-wise ---------------------------------------------------------------------------------------------------------
int pfhe_add_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t l, void *) {
    orc_poly_add(e->flat, a, b, r, (int) l);
    return PFHE_OK;
}
int pfhe_sub_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t l, void *) {
    orc_poly_sub(e->flat, a, b, r, (int) l);
    return PFHE_OK;
}
int pfhe_multiply_rns_poly(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t l, void *) {
    orc_poly_mul(e->flat, a, b, r, (int) l);
    return PFHE_OK;
}
int pfhe_negate_rns_poly(pfhe_engine *e, const uint64_t *a, uint64_t *r, size_t l, void *) {
    orc_poly_negate(e->flat, a, r, (int) l);
    return PFHE_OK;
}
int pfhe_multiply_scalar_rns_poly(pfhe_engine *e, uint64_t *inout, size_t size, uint64_t scalar, size_t l, void *) {
    for (size_t k = 0; k < size * l; k++) {
        const uint64_t q = orc_prime(e->c, (int) (k % l));
        for (uint64_t x = 0; x < e->n; x++) inout[k * e->n + x] = orc_mulmod(inout[k * e->n + x], scalar % q, q);
    }
    return PFHE_OK;
}

int pfhe_sample_poly(pfhe_engine *e, int kind, size_t limbs, const uint8_t *seed, uint64_t *out, void *) {
    return orc_sample_poly(e->c, kind, (int) limbs, seed, out) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "unknown sampler");
}
int pfhe_ntt_backward_inplace(pfhe_engine *e, uint64_t *inout, size_t limbs, size_t start, void *) {
    std::vector<int> rows(limbs);
    for (size_t i = 0; i < limbs; i++) rows[i] = (int) (start + i);
    orc_ntt_inverse(e->c, inout, (int) limbs, rows.data());
    return PFHE_OK;
}

// ---- keys, encryption, decryption ----------------------------------------------------------------------------------------
int pfhe_gen_secretkey(pfhe_engine *e, const uint8_t *seed, uint64_t *sk, void *) {
    orc_gen_secretkey(e->c, seed, sk);
    return PFHE_OK;
}
int pfhe_encrypt_zero_symmetric(pfhe_engine *e, size_t chain_index, const uint64_t *sk, const uint8_t *sa, const uint8_t *se, uint64_t *ct,
                                void *) {
    return orc_encrypt_zero_symmetric(e->c, (int) chain_index, sk, sa, se, ct) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "index is invalid!");
}
int pfhe_encrypt_zero_asymmetric(pfhe_engine *e, size_t chain_index, const uint64_t *pk, const uint8_t *su, const uint8_t *se, uint64_t *ct,
                                 void *) {
    if (chain_index != 1) return fail(PFHE_ERR_INVALID_ARGUMENT, "asymmetric encryption is built for the first data level");
    return orc_encrypt_zero_asymmetric(e->c, pk, su, se, ct) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "no special modulus");
}
int pfhe_encrypt_add_plain(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, void *) {
    return orc_encrypt_add_plain(e->c, limbs_at(e, chain_index), ct, plain) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "index is invalid!");
}
int pfhe_gen_kswitch_key(pfhe_engine *e, const uint64_t *new_key, const uint64_t *sk, const uint8_t *seeds, uint64_t *const *digits, void *) {
    if (e->size_P < 1 || e->size_Q % e->size_P) return fail(PFHE_ERR_INVALID_ARGUMENT, "size_Q must be a multiple of size_P");
    const size_t words = (size_t) 2 * e->size_QP * e->n;
    std::vector<uint64_t> key(words * dnum_of(e));
    if (orc_gen_kswitch_key(e->c, new_key, sk, seeds, key.data())) return fail(PFHE_ERR_INVALID_ARGUMENT, "key generation failed");
    for (int d = 0; d < dnum_of(e); d++) std::memcpy(digits[d], key.data() + d * words, words * 8);
    return PFHE_OK;
}
int pfhe_galois_secret_key(pfhe_engine *e, const uint64_t *sk, uint32_t elt, uint64_t *rotated, void *) {
    std::vector<uint32_t> table(e->n);
    orc_galois_table(e->n, elt, table.data());
    orc_apply_galois_ntt(e->flat, sk, rotated, e->size_QP, table.data());
    return PFHE_OK;
}
int pfhe_decrypt(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, const uint64_t *sk_pow, uint64_t cf, uint64_t *out, void *) {
    return orc_decrypt(e->c, limbs_at(e, chain_index), ct, (int) size, sk_pow, e->mul_tech, cf, out) == 0
               ? PFHE_OK
               : fail(PFHE_ERR_INVALID_ARGUMENT, "decrypt failed");
}

// ---- encoders ------------------------------------------------------------------------------------------------------------
int pfhe_batch_encode(pfhe_engine *e, const uint64_t *values, size_t count, uint64_t *plain, void *) {
    return orc_batch_encode(e->n, e->t, values, count, plain) == 0 ? PFHE_OK : fail(PFHE_ERR_LOGIC, "values_matrix size is too large");
}
int pfhe_batch_decode(pfhe_engine *e, const uint64_t *plain, uint64_t *values, void *) {
    return orc_batch_decode(e->n, e->t, plain, values) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "decode failed");
}
int pfhe_ckks_encode(pfhe_engine *e, size_t chain_index, const double *values, size_t count, double scale, uint64_t *plain, void *) {
    return orc_ckks_encode(e->c, limbs_at(e, chain_index), values, count, scale, plain) == 0 ? PFHE_OK
                                                                                           : fail(PFHE_ERR_INVALID_ARGUMENT, "scale out of bounds");
}
int pfhe_ckks_decode(pfhe_engine *e, size_t chain_index, const uint64_t *plain, double scale, double *values, void *) {
    return orc_ckks_decode(e->c, limbs_at(e, chain_index), plain, scale, values) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "scale out of bounds");
}

// ---- evaluator -----------------------------------------------------------------------------------------------------------
static int bfv_multiply(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *out3, int drop) {
    switch (e->mul_tech) {
        case 1: return orc_bfv_multiply_behz(e->c, a, b, out3);
        case 2: return orc_bfv_multiply_hps(e->c, a, b, out3);
        default: return orc_bfv_multiply_hps_overq(e->c, a, b, out3, drop);
    }
}
int pfhe_multiply(pfhe_engine *e, size_t chain_index, const uint64_t *a, const uint64_t *b, uint64_t *dst, void *) {
    if (e->scheme == PFHE_SCHEME_BFV) {
        if (chain_index != 1) return fail(PFHE_ERR_INVALID_ARGUMENT, "the oracle multiplies BFV at the first level");
        return bfv_multiply(e, a, b, dst, 0) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "multiply failed");
    }
    orc_tensor_2x2(e->c, a, b, dst, limbs_at(e, chain_index));
    return PFHE_OK;
}
int pfhe_multiply_leveled(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *dst, int drop, void *) {
    return bfv_multiply(e, a, b, dst, drop) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "multiply failed");
}
int pfhe_multiply_sizes(pfhe_engine *e, size_t chain_index, const uint64_t *a, size_t sa, const uint64_t *b, size_t sb, uint64_t *dst, void *) {
    orc_tensor_mxn(e->c, a, (int) sa, b, (int) sb, dst, limbs_at(e, chain_index));
    return PFHE_OK;
}
int pfhe_relinearize_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *const *keys, void *) {
    const int l = limbs_at(e, chain_index);
    const auto key = gather_key(e, keys);
    orc_keyswitch(e->c, l, ct, ct + (size_t) 2 * l * e->n, key.data());
    return PFHE_OK;
}
int pfhe_keyswitch_leveled_inplace(pfhe_engine *e, uint64_t *ct, const uint64_t *c2, const uint64_t *const *keys, int drop, void *) {
    const auto key = gather_key(e, keys);
    return orc_bfv_keyswitch_leveled(e->c, ct, c2, key.data(), drop, 0) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "key switch failed");
}
int pfhe_multiply_and_relin(pfhe_engine *e, size_t chain_index, const uint64_t *a, const uint64_t *b, uint64_t *dst, const uint64_t *const *keys,
                            void *) {
    const auto key = gather_key(e, keys);
    if (e->scheme == PFHE_SCHEME_BFV) {
        int rc;
        if (e->mul_tech == 1) rc = orc_bfv_multiply_relin_behz(e->c, a, b, key.data(), dst);
        else if (e->mul_tech == 2) rc = orc_bfv_multiply_relin_hps(e->c, a, b, key.data(), dst);
        else rc = orc_bfv_multiply_relin_hps_overq(e->c, a, b, key.data(), dst, 0);
        return rc == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "multiply failed");
    }
    orc_multiply_relin(e->c, limbs_at(e, chain_index), a, b, key.data(), dst);
    return PFHE_OK;
}
int pfhe_multiply_and_relin_leveled(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *dst, const uint64_t *const *keys, int drop,
                                    void *) {
    const auto key = gather_key(e, keys);
    return orc_bfv_multiply_relin_hps_overq(e->c, a, b, key.data(), dst, drop) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "multiply failed");
}
int pfhe_apply_galois_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, uint32_t elt, const uint64_t *const *keys, void *) {
    const auto key = gather_key(e, keys);
    orc_apply_galois(e->c, limbs_at(e, chain_index), ct, elt, key.data());
    return PFHE_OK;
}
int pfhe_hoisting_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const int *steps, size_t n_steps,
                          const uint64_t *const *const *keys, void *) {
    std::vector<std::vector<uint64_t>> gathered;
    std::vector<const uint64_t *> key_ptrs;
    std::vector<uint32_t> elts;
    for (size_t i = 0; i < n_steps; i++) {
        gathered.push_back(gather_key(e, keys[i]));
        elts.push_back(orc_galois_elt_from_step(steps[i], e->n));
    }
    for (auto &k : gathered) key_ptrs.push_back(k.data());
    orc_hoisting(e->c, limbs_at(e, chain_index), ct, elts.data(), (int) n_steps, key_ptrs.data());
    return PFHE_OK;
}
int pfhe_hoisting_leveled_inplace(pfhe_engine *e, uint64_t *ct, const int *steps, size_t n_steps,
                                  const uint64_t *const *const *keys, int drop, void *) {
    std::vector<std::vector<uint64_t>> gathered;
    std::vector<const uint64_t *> key_ptrs;
    std::vector<uint32_t> elts;
    for (size_t i = 0; i < n_steps; i++) {
        gathered.push_back(gather_key(e, keys[i]));
        elts.push_back(orc_galois_elt_from_step(steps[i], e->n));
    }
    for (auto &g : gathered) key_ptrs.push_back(g.data());
    return orc_bfv_hoisting_leveled(e->c, ct, elts.data(), (int) n_steps, key_ptrs.data(), drop) == 0
                   ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "hoisting failed");
}
int pfhe_rescale_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *dst, void *) {
    const int l = limbs_at(e, chain_index);
    std::vector<uint64_t> copy(ct, ct + size * l * e->n);   // the oracle uses its input as scratch
    orc_rescale(e->c, l, copy.data(), (int) size, dst);
    return PFHE_OK;
}
int pfhe_mod_switch_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *dst, void *) {
    const int l = limbs_at(e, chain_index);
    std::vector<uint64_t> copy(ct, ct + size * l * e->n);
    if (e->scheme == PFHE_SCHEME_CKKS) orc_mod_switch_drop(e->c, l, copy.data(), (int) size, dst);
    else if (e->scheme == PFHE_SCHEME_BFV) orc_divide_round_q_last(e->c, l, copy.data(), (int) size, dst);
    else orc_bgv_mod_switch(e->c, l, copy.data(), (int) size, dst);
    return PFHE_OK;
}
int pfhe_add_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t cf, void *) {
    return orc_plain_add(e->c, limbs_at(e, chain_index), ct, plain, 0, cf) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "index is invalid!");
}
int pfhe_sub_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t cf, void *) {
    return orc_plain_add(e->c, limbs_at(e, chain_index), ct, plain, 1, cf) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "index is invalid!");
}
int pfhe_multiply_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, size_t size, const uint64_t *plain, void *) {
    return orc_plain_multiply(e->c, limbs_at(e, chain_index), ct, (int) size, plain) == 0 ? PFHE_OK : fail(PFHE_ERR_INVALID_ARGUMENT, "index is invalid!");
}
}   // extern "C"
