/*
 * ref_shim.cu -- thin C-ABI harness around the UNMODIFIED reference (encryptorion-lab/phantom-fhe).
 * TEST / BENCH INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile.ref together with the reference's own
 * sources (from /root/reference, never copied) into oracle/_ref/libphantom_ref.so.
 *
 * Purpose: (1) host-only table generators (primes, NTT tables, base-conversion matrices) to pin the CPU
 * oracle in this GPU-less container, (2) on the GPU box: run the reference's public API
 * (multiply_inplace, relinearize_inplace, rotate_inplace, rescale_to_next, mod_switch_to_next,
 * nwt_2d_radix8_*) on caller-supplied words so the B200 engine can be compared bit for bit, and time the
 * same calls with the reference's own method (cudaEvent pair around the op on a fresh copy,
 * benchmark/ckks_bench.cu:167-176).  Contains no arithmetic of its own.
 */
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <sstream>
#include <vector>

#include "phantom.h"

using namespace phantom;
using namespace phantom::arith;
using namespace phantom::util;

#define SHIM_TRY try {
#define SHIM_CATCH                                                                                       \
    }                                                                                                    \
    catch (const std::exception &e) {                                                                    \
        snprintf(g_err, sizeof(g_err), "%s", e.what());                                                  \
        return -1;                                                                                       \
    }

static char g_err[512] = {0};

struct RefCtx {
    std::unique_ptr<PhantomContext> ctx;
    std::unique_ptr<PhantomSecretKey> sk;
    std::unique_ptr<PhantomRelinKey> rlk;
    std::unique_ptr<PhantomGaloisKey> glk;
    scheme_type scheme;
    size_t n, size_QP, size_P;
    double scale;
};

static PhantomCiphertext make_ct(RefCtx *h, size_t chain_index, size_t size, const uint64_t *host, bool ntt) {
    const auto &s = cudaStreamPerThread;
    PhantomCiphertext ct;
    ct.resize(*h->ctx, chain_index, size, s);
    ct.set_ntt_form(ntt);
    ct.set_scale(h->scale);
    size_t words = size * ct.coeff_modulus_size() * ct.poly_modulus_degree();
    cudaMemcpyAsync(ct.data(), host, words * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    return ct;
}

static void fetch_ct(const PhantomCiphertext &ct, uint64_t *host) {
    size_t words = ct.size() * ct.coeff_modulus_size() * ct.poly_modulus_degree();
    cudaStreamSynchronize(cudaStreamPerThread);
    cudaMemcpy(host, ct.data(), words * sizeof(uint64_t), cudaMemcpyDeviceToHost);
}

extern "C" {

const char *ref_last_error() { return g_err; }

/* ---------------------------------------------------------------------------------------------------
 * host-only (no CUDA calls; usable without a GPU)
 * ------------------------------------------------------------------------------------------------- */
int ref_host_create_primes(size_t n, const int *bit_sizes, int count, uint64_t *out) {
    SHIM_TRY
    std::vector<int> bits(bit_sizes, bit_sizes + count);
    auto mods = CoeffModulus::Create(n, bits);
    for (int i = 0; i < count; i++) out[i] = mods[i].value();
    return 0;
    SHIM_CATCH
}

/* tw/tws/itw/itws: n words each; misc = {root, n_inv, n_inv_shoup, ratio0, ratio1, ratio2} */
int ref_host_ntt_table(int log_n, uint64_t q, uint64_t *tw, uint64_t *tws, uint64_t *itw, uint64_t *itws,
                       uint64_t *misc) {
    SHIM_TRY
    Modulus mod(q);
    NTT t(log_n, mod);
    size_t n = size_t(1) << log_n;
    memcpy(tw, t.get_from_root_powers().data(), n * 8);
    memcpy(tws, t.get_from_root_powers_shoup().data(), n * 8);
    memcpy(itw, t.get_from_inv_root_powers().data(), n * 8);
    memcpy(itws, t.get_from_inv_root_powers_shoup().data(), n * 8);
    misc[0] = t.get_root();
    misc[1] = t.inv_degree_modulo();
    misc[2] = t.inv_degree_modulo_shoup();
    misc[3] = mod.const_ratio()[0];
    misc[4] = mod.const_ratio()[1];
    misc[5] = mod.const_ratio()[2];
    return 0;
    SHIM_CATCH
}

/* BaseConverter(ibase, obase): qhat_mod_p = [no][ni], qhatinv_mod_q = [ni] */
int ref_host_bconv_tables(const uint64_t *ibase, int ni, const uint64_t *obase, int no, uint64_t *qhat_mod_p,
                          uint64_t *qhatinv_mod_q) {
    SHIM_TRY
    std::vector<Modulus> iv, ov;
    for (int i = 0; i < ni; i++) iv.emplace_back(ibase[i]);
    for (int j = 0; j < no; j++) ov.emplace_back(obase[j]);
    RNSBase ib(iv), ob(ov);
    BaseConverter conv(ib, ob);
    for (int j = 0; j < no; j++)
        for (int i = 0; i < ni; i++) qhat_mod_p[j * ni + i] = conv.QHatModp(j)[i];
    for (int i = 0; i < ni; i++) qhatinv_mod_q[i] = ib.QHatInvModq()[i];
    return 0;
    SHIM_CATCH
}

uint32_t ref_host_galois_elt(int step, size_t n) { return get_elt_from_step(step, n); }

/* ---------------------------------------------------------------------------------------------------
 * GPU side
 * ------------------------------------------------------------------------------------------------- */
void *ref_create(int scheme, size_t n, const uint64_t *primes, int size_QP, int size_P, uint64_t plain_mod,
                 int mul_tech, const int *galois_steps, int n_steps, double scale, int gen_keys) {
    try {
        auto h = new RefCtx();
        h->scheme = static_cast<scheme_type>(scheme);
        EncryptionParameters parms(h->scheme);
        parms.set_poly_modulus_degree(n);
        std::vector<Modulus> mods;
        for (int i = 0; i < size_QP; i++) mods.emplace_back(primes[i]);
        parms.set_coeff_modulus(mods);
        parms.set_special_modulus_size(size_P);
        if (h->scheme != scheme_type::ckks) parms.set_plain_modulus(Modulus(plain_mod));
        if (h->scheme == scheme_type::bfv) parms.set_mul_tech(static_cast<mul_tech_type>(mul_tech));
        if (n_steps > 0) {
            std::vector<int> steps(galois_steps, galois_steps + n_steps);
            parms.set_galois_elts(get_elts_from_steps(steps, n));
        }
        h->ctx = std::make_unique<PhantomContext>(parms);
        h->n = n;
        h->size_QP = size_QP;
        h->size_P = size_P;
        h->scale = scale;
        if (gen_keys) {
            h->sk = std::make_unique<PhantomSecretKey>(*h->ctx);
            h->rlk = std::make_unique<PhantomRelinKey>(h->sk->gen_relinkey(*h->ctx));
            if (n_steps > 0 || gen_keys == 2) h->glk = std::make_unique<PhantomGaloisKey>(h->sk->create_galois_keys(*h->ctx));   // gen_keys 2: keys of the default elements
        }
        cudaStreamSynchronize(cudaStreamPerThread);
        return h;
    } catch (const std::exception &e) {
        snprintf(g_err, sizeof(g_err), "%s", e.what());
        return nullptr;
    }
}

void ref_destroy(void *p) { delete static_cast<RefCtx *>(p); }

int ref_dnum(void *p) {
    auto h = static_cast<RefCtx *>(p);
    auto &tool = h->ctx->get_context_data(1).gpu_rns_tool();
    return (int) tool.v_base_part_Ql_to_compl_part_QlP_conv().size();
}

int ref_galois_count(void *p) {
    auto h = static_cast<RefCtx *>(p);
    return (int) h->ctx->key_galois_tool_->galois_elts().size();
}

uint32_t ref_galois_elt_at(void *p, int idx) {
    auto h = static_cast<RefCtx *>(p);
    return h->ctx->key_galois_tool_->galois_elts().at(idx);
}

/* key digit d = [2][size_QP][n] words.  which < 0: relin key; which >= 0: galois key index. dir 0 = get. */
static int key_xfer(RefCtx *h, int which, int d, uint64_t *host, int dir) {
    cudaGetLastError();   /* a launch the reference left unchecked must not fail the transfer */
    uint64_t *const *dev_ptrs =
            which < 0 ? h->rlk->public_keys_ptr() : h->glk->get_relin_keys(which).public_keys_ptr();
    uint64_t *ptr = nullptr;
    cudaMemcpy(&ptr, dev_ptrs + d, sizeof(uint64_t *), cudaMemcpyDeviceToHost);
    size_t bytes = 2 * h->size_QP * h->n * sizeof(uint64_t);
    if (dir == 0) cudaMemcpy(host, ptr, bytes, cudaMemcpyDeviceToHost);
    else cudaMemcpy(ptr, host, bytes, cudaMemcpyHostToDevice);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int ref_key_get(void *p, int which, int d, uint64_t *host) {
    SHIM_TRY return key_xfer(static_cast<RefCtx *>(p), which, d, host, 0);
    SHIM_CATCH
}

int ref_key_set(void *p, int which, int d, const uint64_t *host) {
    SHIM_TRY return key_xfer(static_cast<RefCtx *>(p), which, d, const_cast<uint64_t *>(host), 1);
    SHIM_CATCH
}

/* forward / inverse NTT of `limbs` limbs starting at table row start_idx (in place on host data) */
int ref_ntt(void *p, uint64_t *host, size_t limbs, size_t start_idx, int inverse) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    size_t words = limbs * h->n;
    auto buf = make_cuda_auto_ptr<uint64_t>(words, s);
    cudaMemcpyAsync(buf.get(), host, words * 8, cudaMemcpyHostToDevice, s);
    if (inverse) nwt_2d_radix8_backward_inplace(buf.get(), h->ctx->gpu_rns_tables(), limbs, start_idx, s);
    else nwt_2d_radix8_forward_inplace(buf.get(), h->ctx->gpu_rns_tables(), limbs, start_idx, s);
    cudaMemcpyAsync(host, buf.get(), words * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* multiply_inplace + relinearize_inplace (evaluate.cu:1029-1057,1342-1374).  ct = [2][l][n] */
int ref_multiply_relin(void *p, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    bool ntt = h->scheme != scheme_type::bfv;
    auto a = make_ct(h, chain_index, 2, ct1, ntt);
    auto b = make_ct(h, chain_index, 2, ct2, ntt);
    multiply_inplace(*h->ctx, a, b);
    relinearize_inplace(*h->ctx, a, *h->rlk);
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* multiply_inplace only: out = [3][l][n] */
int ref_multiply(void *p, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    bool ntt = h->scheme != scheme_type::bfv;
    auto a = make_ct(h, chain_index, 2, ct1, ntt);
    auto b = make_ct(h, chain_index, 2, ct2, ntt);
    multiply_inplace(*h->ctx, a, b);
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* fnwt_1d_opt / inwt_1d_opt on tables built the way test/ntt_test.cu:7-60 builds them: `count` primes of `bits` bits
 * for degree 2^log_dim, limbs [start, count) transformed in place on the host copy `data` = [count][dim] */
#define STEP(name)                                                                                       \
    do {                                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) {                                                                         \
            snprintf(g_err, sizeof(g_err), "ref_nwt_1d after %s: %s", name, cudaGetErrorString(e_));     \
            return -1;                                                                                   \
        }                                                                                                \
    } while (0)
int ref_nwt_1d(int log_dim, int count, int bits, int start, uint64_t *data, int inverse) {
    SHIM_TRY
    const auto &s = cudaStreamPerThread;
    cudaGetLastError(); /* start from a clean per-thread error state */
    size_t dim = size_t(1) << log_dim;
    const auto h_modulus = CoeffModulus::Create(dim, std::vector<int>(count, bits));
    auto modulus = make_cuda_auto_ptr<DModulus>(count, s);
    std::vector<DModulus> hm(count);
    for (int i = 0; i < count; i++)
        hm[i] = DModulus(h_modulus[i].value(), h_modulus[i].const_ratio()[0], h_modulus[i].const_ratio()[1]);
    cudaMemcpyAsync(modulus.get(), hm.data(), count * sizeof(DModulus), cudaMemcpyHostToDevice, s);
    STEP("modulus");
    auto tw = make_cuda_auto_ptr<uint64_t>(count * dim, s);
    auto tws = make_cuda_auto_ptr<uint64_t>(count * dim, s);
    auto itw = make_cuda_auto_ptr<uint64_t>(count * dim, s);
    auto itws = make_cuda_auto_ptr<uint64_t>(count * dim, s);
    auto ninv = make_cuda_auto_ptr<uint64_t>(count, s);
    auto ninvs = make_cuda_auto_ptr<uint64_t>(count, s);
    for (int i = 0; i < count; i++) {
        auto t = NTT(log_dim, h_modulus[i]);
        cudaMemcpyAsync(tw.get() + i * dim, t.get_from_root_powers().data(), dim * 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(tws.get() + i * dim, t.get_from_root_powers_shoup().data(), dim * 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(itw.get() + i * dim, t.get_from_inv_root_powers().data(), dim * 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(itws.get() + i * dim, t.get_from_inv_root_powers_shoup().data(), dim * 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ninv.get() + i, &t.inv_degree_modulo(), 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ninvs.get() + i, &t.inv_degree_modulo_shoup(), 8, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);   /* the host table dies at the end of the iteration */
    }
    STEP("tables");
    auto d = make_cuda_auto_ptr<uint64_t>(count * dim, s);
    cudaMemcpyAsync(d.get(), data, count * dim * 8, cudaMemcpyHostToDevice, s);
    STEP("input");
    /* fnwt_1d_opt ignores start_modulus_idx (ntt_1d.cu:76-86): the plain form is used when a start index is given */
    if (inverse) inwt_1d_opt(d.get(), itw.get(), itws.get(), modulus.get(), ninv.get(), ninvs.get(), dim, count - start, start, s);
    else if (start == 0) fnwt_1d_opt(d.get(), tw.get(), tws.get(), modulus.get(), dim, count, 0, s);
    else fnwt_1d(d.get(), tw.get(), tws.get(), modulus.get(), dim, count - start, start, s);
    STEP("launch");
    cudaMemcpyAsync(data, d.get(), count * dim * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "ref_nwt_1d: %s", cudaGetErrorString(ce));
        return -1;
    }
    return 0;
    SHIM_CATCH
}

/* multiply_inplace on ciphertexts of sizes size1 x size2 (tensor_prod_mxn_rns_poly branch): out = [size1+size2-1][l][n] */
int ref_multiply_sizes(void *p, size_t chain_index, const uint64_t *ct1, size_t size1, const uint64_t *ct2, size_t size2,
                       uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    bool ntt = h->scheme != scheme_type::bfv;
    auto a = make_ct(h, chain_index, size1, ct1, ntt);
    auto b = make_ct(h, chain_index, size2, ct2, ntt);
    multiply_inplace(*h->ctx, a, b);
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* multiply_inplace (relin = 0: out = [3][l][n]), multiply_inplace + relinearize_inplace (1) or multiply_and_relin_inplace
 * (2) (out = [2][l][n]) on ciphertexts carrying the given noise-scale degrees: what mul_tech hps_overq_leveled reads to
 * decide how many levels to drop (evaluate.cu:680-690) */
int ref_multiply_deg(void *p, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, size_t deg1, size_t deg2,
                     int relin, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    bool ntt = h->scheme != scheme_type::bfv;
    auto a = make_ct(h, chain_index, 2, ct1, ntt);
    auto b = make_ct(h, chain_index, 2, ct2, ntt);
    a.SetNoiseScaleDeg(deg1);
    b.SetNoiseScaleDeg(deg2);
    if (relin == 2) multiply_and_relin_inplace(*h->ctx, a, b, *h->rlk);
    else {
        multiply_inplace(*h->ctx, a, b);
        if (relin == 1) relinearize_inplace(*h->ctx, a, *h->rlk);
    }
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* the secret key (first power, NTT form, [size_QP][n]) through PhantomSecretKey::save (secretkey.h:346-364) */
int ref_secret_key(void *p, uint64_t *host) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    if (!h->sk) throw std::invalid_argument("context was created without keys");
    std::stringstream ss;
    h->sk->save(ss);
    const std::string blob = ss.str();
    size_t hdr[3];
    std::memcpy(hdr, blob.data(), sizeof(hdr));   /* sk_max_power, poly_modulus_degree, coeff_modulus_size */
    std::memcpy(host, blob.data() + sizeof(hdr), hdr[1] * hdr[2] * sizeof(uint64_t));
    return 0;
    SHIM_CATCH
}

/* PhantomSecretKey::decrypt (secretkey.cu:693-723) of caller-supplied ciphertext words: out = [l][n] (CKKS) or [n] */
int ref_decrypt(void *p, size_t chain_index, const uint64_t *ct, size_t size, uint64_t correction_factor, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    if (!h->sk) throw std::invalid_argument("context was created without keys");
    bool ntt = h->scheme != scheme_type::bfv;
    auto c = make_ct(h, chain_index, size, ct, ntt);
    if (h->scheme == scheme_type::bgv) c.set_correction_factor(correction_factor);
    PhantomPlaintext plain;
    h->sk->decrypt(*h->ctx, c, plain);
    cudaStreamSynchronize(cudaStreamPerThread);
    size_t words = (h->scheme == scheme_type::ckks ? c.coeff_modulus_size() : 1) * h->n;   /* secretkey.cu:705-712 */
    cudaMemcpy(out, plain.data(), words * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    return 0;
    SHIM_CATCH
}

/* PhantomCiphertext::save of caller-supplied words and metadata; returns the stream length (or -1), bytes in out */
long ref_save_ct(void *p, size_t chain_index, const uint64_t *ct, size_t size, double scale, size_t noise_deg,
                 unsigned char *out, size_t cap) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto c = make_ct(h, chain_index, size, ct, h->scheme != scheme_type::bfv);
    c.set_scale(scale);
    c.SetNoiseScaleDeg(noise_deg);
    std::stringstream ss;
    c.save(ss);
    const std::string blob = ss.str();
    if (blob.size() > cap) throw std::invalid_argument("buffer too small");
    std::memcpy(out, blob.data(), blob.size());
    return (long) blob.size();
    SHIM_CATCH
}
/* PhantomCiphertext::load of a caller-supplied stream: words to out, {chain_index, size, N, l, noise_deg, ntt} to meta */
int ref_load_ct(const unsigned char *bytes, size_t len, uint64_t *out, size_t *meta, double *scale) {
    SHIM_TRY
    std::stringstream ss(std::string(reinterpret_cast<const char *>(bytes), len));
    PhantomCiphertext c;
    c.load(ss);
    meta[0] = c.chain_index(), meta[1] = c.size(), meta[2] = c.poly_modulus_degree(), meta[3] = c.coeff_modulus_size();
    meta[4] = c.GetNoiseScaleDeg(), meta[5] = c.is_ntt_form();
    *scale = c.scale();
    fetch_ct(c, out);
    return 0;
    SHIM_CATCH
}
/* save() of the context's own keys: which = 0 relin key, 1 Galois key, 2 secret key; returns the length */
long ref_save_key(void *p, int which, unsigned char *out, size_t cap) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    std::stringstream ss;
    if (which == 0) h->rlk->save(ss);
    else if (which == 1) h->glk->save(ss);
    else h->sk->save(ss);
    const std::string blob = ss.str();
    if (out) {
        if (blob.size() > cap) throw std::invalid_argument("buffer too small");
        std::memcpy(out, blob.data(), blob.size());
    }
    return (long) blob.size();
    SHIM_CATCH
}

/* PhantomCKKSEncoder::encode (ckks.cu:66-135): count complex values (re, im interleaved) -> [l][n] residues, NTT form */
int ref_ckks_encode(void *p, const double *values, size_t count, size_t chain_index, double scale, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    PhantomCKKSEncoder enc(*h->ctx);
    std::vector<cuDoubleComplex> v(count);
    for (size_t i = 0; i < count; i++) v[i] = make_cuDoubleComplex(values[2 * i], values[2 * i + 1]);
    PhantomPlaintext pt;
    enc.encode(*h->ctx, v, scale, pt, chain_index);
    cudaStreamSynchronize(cudaStreamPerThread);
    size_t l = h->ctx->get_context_data(chain_index).parms().coeff_modulus().size();
    cudaMemcpy(out, pt.data(), l * h->n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    return 0;
    SHIM_CATCH
}

/* PhantomCKKSEncoder::decode (ckks.cu:137-190): [l][n] NTT-form residues at chain_index with the given scale -> n/2
 * complex values (re, im interleaved) */
int ref_ckks_decode(void *p, const uint64_t *plain, size_t chain_index, double scale, double *values) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    PhantomCKKSEncoder enc(*h->ctx);
    std::vector<cuDoubleComplex> zero(1, make_cuDoubleComplex(0.0, 0.0));
    PhantomPlaintext pt;
    enc.encode(*h->ctx, zero, scale, pt, chain_index);   /* a plaintext object of the right shape, level and scale */
    size_t l = h->ctx->get_context_data(chain_index).parms().coeff_modulus().size();
    cudaStreamSynchronize(cudaStreamPerThread);
    cudaMemcpy(pt.data(), plain, l * h->n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    std::vector<cuDoubleComplex> out;
    enc.decode(*h->ctx, pt, out);
    for (size_t i = 0; i < out.size(); i++) values[2 * i] = cuCreal(out[i]), values[2 * i + 1] = cuCimag(out[i]);
    return 0;
    SHIM_CATCH
}

/* encode then decode inside the reference, no host round trip of the plaintext */
int ref_ckks_roundtrip(void *p, const double *values, size_t count, size_t chain_index, double scale, double *out_values) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    PhantomCKKSEncoder enc(*h->ctx);
    std::vector<cuDoubleComplex> v(count);
    for (size_t i = 0; i < count; i++) v[i] = make_cuDoubleComplex(values[2 * i], values[2 * i + 1]);
    PhantomPlaintext pt;
    enc.encode(*h->ctx, v, scale, pt, chain_index);
    std::vector<cuDoubleComplex> out;
    enc.decode(*h->ctx, pt, out);
    for (size_t i = 0; i < out.size(); i++) out_values[2 * i] = cuCreal(out[i]), out_values[2 * i + 1] = cuCimag(out[i]);
    return 0;
    SHIM_CATCH
}

/* the reference's sampler kernels (prng.cu:142-244) with a caller-supplied 64-byte seed, launched the way secretkey.cu
 * launches them: kind 0 ternary, 1 error, 2 uniform; out = [limbs][n] */
int ref_sample_poly(void *p, int kind, const uint8_t *seed, size_t limbs, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto d_seed = phantom::util::make_cuda_auto_ptr<uint8_t>(phantom::util::global_variables::prng_seed_byte_count, s);
    auto d_out = phantom::util::make_cuda_auto_ptr<uint64_t>(limbs * h->n, s);
    cudaMemcpyAsync(d_seed.get(), seed, phantom::util::global_variables::prng_seed_byte_count, cudaMemcpyHostToDevice, s);
    auto base_rns = h->ctx->gpu_rns_tables().modulus();
    uint64_t grid = h->n * limbs / blockDimGlb.x;
    if (kind == 0) sample_ternary_poly<<<grid, blockDimGlb, 0, s>>>(d_out.get(), d_seed.get(), base_rns, h->n, limbs);
    else if (kind == 1) sample_error_poly<<<grid, blockDimGlb, 0, s>>>(d_out.get(), d_seed.get(), base_rns, h->n, limbs);
    else if (kind == 2) sample_uniform_poly<<<grid, blockDimGlb, 0, s>>>(d_out.get(), d_seed.get(), base_rns, h->n, limbs);
    else throw std::invalid_argument("unknown sampler");
    cudaStreamSynchronize(s);
    cudaMemcpy(out, d_out.get(), limbs * h->n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* PhantomSecretKey::encrypt_symmetric / PhantomPublicKey::encrypt_asymmetric (secretkey.cu:130-190, 463-530) of a
 * caller-supplied plaintext with the context's own keys: plain = [n] mod t (BFV, BGV) or [l][n] NTT form (CKKS) */
int ref_encrypt(void *p, int asymmetric, size_t chain_index, const uint64_t *plain, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    if (!h->sk) throw std::invalid_argument("context was created without keys");
    const auto &s = cudaStreamPerThread;
    PhantomPlaintext pt;
    size_t l = 1;
    if (h->scheme == scheme_type::ckks) {   /* a plaintext object of the right level and scale, then the caller's words */
        PhantomCKKSEncoder enc(*h->ctx);
        std::vector<cuDoubleComplex> zero(1, make_cuDoubleComplex(0.0, 0.0));
        enc.encode(*h->ctx, zero, h->scale, pt, chain_index);
        l = h->ctx->get_context_data(chain_index).parms().coeff_modulus().size();
    } else {
        pt.resize(1, h->n, s);
    }
    cudaStreamSynchronize(s);
    cudaMemcpy(pt.data(), plain, l * h->n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    PhantomCiphertext ct;
    if (asymmetric) {
        PhantomPublicKey pk = h->sk->gen_publickey(*h->ctx);
        pk.encrypt_asymmetric(*h->ctx, pt, ct);
    } else {
        h->sk->encrypt_symmetric(*h->ctx, pt, ct);
    }
    fetch_ct(ct, out);
    return 0;
    SHIM_CATCH
}

/* encrypt_symmetric with the context's key, then PhantomCiphertext::save_symmetric (ciphertext.h:216-245): the stream to
 * `bytes` (returns its length), the full ciphertext words to `words` */
long ref_encrypt_save_symmetric(void *p, size_t chain_index, const uint64_t *plain, unsigned char *bytes, size_t cap,
                                uint64_t *words) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    if (!h->sk) throw std::invalid_argument("context was created without keys");
    const auto &s = cudaStreamPerThread;
    PhantomPlaintext pt;
    size_t l = 1;
    if (h->scheme == scheme_type::ckks) {
        PhantomCKKSEncoder enc(*h->ctx);
        std::vector<cuDoubleComplex> zero(1, make_cuDoubleComplex(0.0, 0.0));
        enc.encode(*h->ctx, zero, h->scale, pt, chain_index);
        l = h->ctx->get_context_data(chain_index).parms().coeff_modulus().size();
    } else {
        pt.resize(1, h->n, s);
    }
    cudaStreamSynchronize(s);
    cudaMemcpy(pt.data(), plain, l * h->n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    PhantomCiphertext ct;
    h->sk->encrypt_symmetric(*h->ctx, pt, ct);
    fetch_ct(ct, words);
    std::stringstream ss;
    ct.save_symmetric(ss);
    const std::string blob = ss.str();
    if (blob.size() > cap) throw std::invalid_argument("buffer too small");
    std::memcpy(bytes, blob.data(), blob.size());
    return (long) blob.size();
    SHIM_CATCH
}
/* PhantomCiphertext::load_symmetric (ciphertext.h:247-307) of a caller-supplied stream: the rebuilt words to `words` */
int ref_load_symmetric(void *p, const unsigned char *bytes, size_t len, uint64_t *words) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    std::stringstream ss(std::string(reinterpret_cast<const char *>(bytes), len));
    PhantomCiphertext ct;
    ct.load_symmetric(*h->ctx, ss);
    fetch_ct(ct, words);
    return 0;
    SHIM_CATCH
}

/* a PhantomPlaintext holding caller-supplied words: [n] mod t (BFV, BGV) or [l][n] NTT form at chain_index with the
 * context's scale (CKKS; built by encoding zero, then overwritten: the class has no public setters for level and scale) */
static PhantomPlaintext make_plain(RefCtx *h, size_t chain_index, const uint64_t *plain) {
    const auto &s = cudaStreamPerThread;
    PhantomPlaintext pt;
    size_t l = 1;
    if (h->scheme == scheme_type::ckks) {
        PhantomCKKSEncoder enc(*h->ctx);
        std::vector<cuDoubleComplex> zero(1, make_cuDoubleComplex(0.0, 0.0));
        enc.encode(*h->ctx, zero, h->scale, pt, chain_index);
        l = h->ctx->get_context_data(chain_index).parms().coeff_modulus().size();
    } else {
        pt.resize(1, h->n, s);
    }
    cudaStreamSynchronize(s);
    cudaMemcpy(pt.data(), plain, l * h->n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    return pt;
}

/* add_plain_inplace / sub_plain_inplace / multiply_plain_inplace (evaluate.cu:1106-1340) on caller-supplied words:
 * op 0 / 1 / 2 */
int ref_plain_op(void *p, int op, size_t chain_index, const uint64_t *ct, size_t size, const uint64_t *plain,
                 uint64_t correction_factor, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto c = make_ct(h, chain_index, size, ct, h->scheme != scheme_type::bfv);
    if (h->scheme == scheme_type::bgv) c.set_correction_factor(correction_factor);
    PhantomPlaintext pt = make_plain(h, chain_index, plain);
    if (op == 0) add_plain_inplace(*h->ctx, c, pt);
    else if (op == 1) sub_plain_inplace(*h->ctx, c, pt);
    else if (op == 2) multiply_plain_inplace(*h->ctx, c, pt);
    else throw std::invalid_argument("unknown op");
    fetch_ct(c, out);
    return 0;
    SHIM_CATCH
}

/* add_inplace / sub_inplace / sub_inplace(negate) / negate_inplace (evaluate.cu:106-338): op 0 / 1 / 2 / 3; the BGV
 * correction factors go in, the balanced one comes back */
int ref_add_sub(void *p, int op, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, size_t size, uint64_t cf1,
                uint64_t cf2, uint64_t *out, uint64_t *cf_out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    bool ntt = h->scheme != scheme_type::bfv;
    auto a = make_ct(h, chain_index, size, ct1, ntt);
    if (h->scheme == scheme_type::bgv) a.set_correction_factor(cf1);
    if (op == 3) {
        negate_inplace(*h->ctx, a);
    } else {
        auto b = make_ct(h, chain_index, size, ct2, ntt);
        if (h->scheme == scheme_type::bgv) b.set_correction_factor(cf2);
        if (op == 0) add_inplace(*h->ctx, a, b);
        else sub_inplace(*h->ctx, a, b, op == 2);
    }
    fetch_ct(a, out);
    *cf_out = a.correction_factor();
    return 0;
    SHIM_CATCH
}

/* gen_publickey with the context's secret key, then PhantomPublicKey::save (secretkey.h:85-90): returns the stream length */
long ref_public_key_stream(void *p, unsigned char *out, size_t cap) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    if (!h->sk) throw std::invalid_argument("context was created without keys");
    PhantomPublicKey pk = h->sk->gen_publickey(*h->ctx);
    cudaStreamSynchronize(cudaStreamPerThread);
    std::stringstream ss;
    pk.save(ss);
    const std::string blob = ss.str();
    if (blob.size() > cap) throw std::invalid_argument("buffer too small");
    std::memcpy(out, blob.data(), blob.size());
    return (long) blob.size();
    SHIM_CATCH
}

/* PhantomBatchEncoder::encode / decode (batchencoder.cu:62-118): values[count] -> plain[n]; plain[n] -> values[n] */
int ref_batch_encode(void *p, const uint64_t *values, size_t count, uint64_t *plain) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    PhantomBatchEncoder enc(*h->ctx);
    std::vector<uint64_t> v(values, values + count);
    PhantomPlaintext pt = enc.encode(*h->ctx, v);
    cudaStreamSynchronize(cudaStreamPerThread);
    cudaMemcpy(plain, pt.data(), h->n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    return 0;
    SHIM_CATCH
}
int ref_batch_decode(void *p, const uint64_t *plain, uint64_t *values) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    PhantomBatchEncoder enc(*h->ctx);
    std::vector<uint64_t> zeros(h->n, 0);
    PhantomPlaintext pt = enc.encode(*h->ctx, zeros);   /* a plaintext object of the right shape */
    cudaMemcpy(pt.data(), plain, h->n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    std::vector<uint64_t> out = enc.decode(*h->ctx, pt);
    std::memcpy(values, out.data(), h->n * sizeof(uint64_t));
    return 0;
    SHIM_CATCH
}

/* stage-wise key-switch taps (eval_key_switch.cu:95-182) for differential debugging */
int ref_modup(void *p, size_t chain_index, const uint64_t *c2, uint64_t *t_mod_up) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    size_t l = tool.base_Ql().size(), m = l + h->size_P, beta = tool.v_base_part_Ql_to_compl_part_QlP_conv().size();
    auto in = make_cuda_auto_ptr<uint64_t>(l * h->n, s);
    auto out = make_cuda_auto_ptr<uint64_t>(beta * m * h->n, s);
    cudaMemcpyAsync(in.get(), c2, l * h->n * 8, cudaMemcpyHostToDevice, s);
    tool.modup(out.get(), in.get(), h->ctx->gpu_rns_tables(), h->scheme, s);
    cudaMemcpyAsync(t_mod_up, out.get(), beta * m * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return 0;
    SHIM_CATCH
}

int ref_inner_prod(void *p, size_t chain_index, int which_key, const uint64_t *t_mod_up, uint64_t *cx) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    size_t l = tool.base_Ql().size(), m = l + h->size_P, beta = tool.v_base_part_Ql_to_compl_part_QlP_conv().size();
    auto in = make_cuda_auto_ptr<uint64_t>(beta * m * h->n, s);
    auto out = make_cuda_auto_ptr<uint64_t>(2 * m * h->n, s);
    cudaMemcpyAsync(in.get(), t_mod_up, beta * m * h->n * 8, cudaMemcpyHostToDevice, s);
    auto keys = which_key < 0 ? h->rlk->public_keys_ptr() : h->glk->get_relin_keys(which_key).public_keys_ptr();
    key_switch_inner_prod(out.get(), in.get(), keys, tool, h->ctx->gpu_rns_tables().modulus(), 0, s);
    cudaMemcpyAsync(cx, out.get(), 2 * m * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return 0;
    SHIM_CATCH
}

/* moddown_from_NTT(cx_i, cx_i): returns the first l limbs */
int ref_moddown(void *p, size_t chain_index, const uint64_t *cx_i, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    size_t l = tool.base_Ql().size(), m = l + h->size_P;
    auto buf = make_cuda_auto_ptr<uint64_t>(m * h->n, s);
    cudaMemcpyAsync(buf.get(), cx_i, m * h->n * 8, cudaMemcpyHostToDevice, s);
    tool.moddown_from_NTT(buf.get(), buf.get(), h->ctx->gpu_rns_tables(), h->scheme, s);
    cudaMemcpyAsync(out, buf.get(), l * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return 0;
    SHIM_CATCH
}

/* rotate_inplace (evaluate.cu:1633-1668) */
int ref_rotate(void *p, size_t chain_index, const uint64_t *ct, int step, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto a = make_ct(h, chain_index, 2, ct, h->scheme != scheme_type::bfv);
    rotate_inplace(*h->ctx, a, step, *h->glk);
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* rescale_to_next (CKKS) / mod_switch_to_next; out = [size][l-1][n] */
int ref_rescale(void *p, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto a = make_ct(h, chain_index, size, ct, h->scheme != scheme_type::bfv);
    auto r = rescale_to_next(*h->ctx, a);
    fetch_ct(r, out);
    return 0;
    SHIM_CATCH
}

int ref_mod_switch(void *p, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto a = make_ct(h, chain_index, size, ct, h->scheme != scheme_type::bfv);
    auto r = mod_switch_to_next(*h->ctx, a);
    fetch_ct(r, out);
    return 0;
    SHIM_CATCH
}

/* ---------------------------------------------------------------------------------------------------
 * timing, the reference's own method: cudaEvent pair around the op on a fresh copy
 * (include/cuda_wrapper.cuh:191-283, benchmark/ckks_bench.cu:167-176).  times_us has `trials` entries.
 * op: 0 = multiply+relin, 1 = rotate(step), 2 = rescale (after mult+relin outside the timer as in the
 * bench), 3 = forward NTT of `aux` limbs, 4 = inverse NTT of `aux` limbs
 * mode 0: device-resident inputs (the reference bench's timed region)
 * mode 1: end-to-end from pinned host buffers: H2D of the inputs, op, D2H of the result, all inside the
 *         timed region (wall clock around stream sync).
 * ------------------------------------------------------------------------------------------------- */
int ref_time_op(void *p, int op, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, int aux, int mode,
                int trials, double *times_us) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    bool ntt = h->scheme != scheme_type::bfv;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (op == 3 || op == 4) {
        size_t limbs = aux, words = limbs * h->n;
        auto buf = make_cuda_auto_ptr<uint64_t>(words, s);
        cudaMemsetAsync(buf.get(), 0, words * 8, s);
        for (int t = 0; t < trials; t++) {
            cudaEventRecord(e0, s);
            if (op == 3) nwt_2d_radix8_forward_inplace(buf.get(), h->ctx->gpu_rns_tables(), limbs, 0, s);
            else nwt_2d_radix8_backward_inplace(buf.get(), h->ctx->gpu_rns_tables(), limbs, 0, s);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            times_us[t] = ms * 1000.0;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return 0;
    }
    auto a0 = make_ct(h, chain_index, 2, ct1, ntt);
    auto b0 = make_ct(h, chain_index, 2, ct2 ? ct2 : ct1, ntt);
    size_t words = 2 * a0.coeff_modulus_size() * h->n;
    uint64_t *pin_a = nullptr, *pin_b = nullptr, *pin_o = nullptr;
    if (mode == 1) {
        cudaMallocHost(&pin_a, words * 8);
        cudaMallocHost(&pin_b, words * 8);
        cudaMallocHost(&pin_o, words * 8);
        memcpy(pin_a, ct1, words * 8);
        memcpy(pin_b, ct2 ? ct2 : ct1, words * 8);
    }
    for (int t = 0; t < trials; t++) {
        if (mode == 0) {
            PhantomCiphertext tmp(a0);
            if (op == 2) {
                multiply_inplace(*h->ctx, tmp, b0);
                relinearize_inplace(*h->ctx, tmp, *h->rlk);
            }
            cudaEventRecord(e0, s);
            if (op == 0) {
                multiply_inplace(*h->ctx, tmp, b0);
                relinearize_inplace(*h->ctx, tmp, *h->rlk);
            } else if (op == 1) {
                rotate_inplace(*h->ctx, tmp, aux, *h->glk);
            } else if (op == 2) {
                auto r = rescale_to_next(*h->ctx, tmp);
            }
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            times_us[t] = ms * 1000.0;
        } else {
            cudaStreamSynchronize(s);
            auto w0 = std::chrono::steady_clock::now();
            PhantomCiphertext a, b;
            a.resize(*h->ctx, chain_index, 2, s);
            a.set_ntt_form(ntt);
            a.set_scale(h->scale);
            cudaMemcpyAsync(a.data(), pin_a, words * 8, cudaMemcpyHostToDevice, s);
            size_t out_words = words;
            if (op == 0) {
                b.resize(*h->ctx, chain_index, 2, s);
                b.set_ntt_form(ntt);
                b.set_scale(h->scale);
                cudaMemcpyAsync(b.data(), pin_b, words * 8, cudaMemcpyHostToDevice, s);
                multiply_inplace(*h->ctx, a, b);
                relinearize_inplace(*h->ctx, a, *h->rlk);
                cudaMemcpyAsync(pin_o, a.data(), out_words * 8, cudaMemcpyDeviceToHost, s);
            } else if (op == 1) {
                rotate_inplace(*h->ctx, a, aux, *h->glk);
                cudaMemcpyAsync(pin_o, a.data(), out_words * 8, cudaMemcpyDeviceToHost, s);
            } else {
                auto r = rescale_to_next(*h->ctx, a);
                out_words = 2 * r.coeff_modulus_size() * h->n;
                cudaMemcpyAsync(pin_o, r.data(), out_words * 8, cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
            }
            cudaStreamSynchronize(s);
            auto w1 = std::chrono::steady_clock::now();
            times_us[t] = std::chrono::duration<double, std::micro>(w1 - w0).count();
        }
    }
    if (pin_a) cudaFreeHost(pin_a);
    if (pin_b) cudaFreeHost(pin_b);
    if (pin_o) cudaFreeHost(pin_o);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* ---------------------------------------------------------------------------------------------------
 * kernel-level launchers of include/ntt.cuh:172-226 on caller-supplied words.  buf = [total_limbs][n] host words
 * (in place); aux = second host buffer where the launcher has one (input of the out-of-place forms, cx of
 * fuse_moddown).  variant: 0 forward_inplace, 1 backward_inplace, 2 backward (out of place: aux -> buf),
 * 3 forward_inplace_fuse_moddown (buf = delta on entry, ct on return; aux = cx; scale = bigPInv_mod_q),
 * 4 forward_inplace_include_temp_mod (prm[0] = total), 5 forward_inplace_include_special_mod (prm = size_QP, size_P),
 * 6 ..._include_special_mod_exclude_range (prm = size_QP, size_P, excl_start, excl_end),
 * 7 forward_modup_fuse (aux -> buf, prm[0] = modulus_index), 8 backward_scale (aux -> buf), 9 backward_inplace_scale,
 * 10 backward_inplace_include_special_mod (prm = size_QP, size_P), 11 backward_inplace_include_temp_mod_scale (prm[0] = total).
 * table: 0 gpu_rns_tables, 1 gpu_Bsk_tables, 2 gpu_QlRl_tables.  scale: host values, Shoup companions made here
 * with the table's own modulus (compute_shoup).
 * ------------------------------------------------------------------------------------------------- */
int ref_nwt(void *p, int variant, int table, uint64_t *buf, const uint64_t *aux, size_t total_limbs, size_t count,
            size_t start, const size_t *prm, const uint64_t *scale, size_t n_scale, const uint64_t *scale_mod) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(1).gpu_rns_tool();
    const DNTTTable &tab = table == 0 ? h->ctx->gpu_rns_tables() : table == 1 ? tool.gpu_Bsk_tables() : tool.gpu_QlRl_tables();
    size_t words = total_limbs * h->n;
    auto d = make_cuda_auto_ptr<uint64_t>(words, s);
    auto d2 = make_cuda_auto_ptr<uint64_t>(words, s);
    cudaMemcpyAsync(d.get(), buf, words * 8, cudaMemcpyHostToDevice, s);
    if (aux) cudaMemcpyAsync(d2.get(), aux, words * 8, cudaMemcpyHostToDevice, s);
    auto sc = make_cuda_auto_ptr<uint64_t>(n_scale + 1, s);
    auto scs = make_cuda_auto_ptr<uint64_t>(n_scale + 1, s);
    if (scale) {
        std::vector<uint64_t> sh(n_scale);
        for (size_t i = 0; i < n_scale; i++) sh[i] = compute_shoup(scale[i], scale_mod[i]);
        cudaMemcpyAsync(sc.get(), scale, n_scale * 8, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(scs.get(), sh.data(), n_scale * 8, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
    }
    switch (variant) {
        case 0: nwt_2d_radix8_forward_inplace(d.get(), tab, count, start, s); break;
        case 1: nwt_2d_radix8_backward_inplace(d.get(), tab, count, start, s); break;
        case 2: nwt_2d_radix8_backward(d.get(), d2.get(), tab, count, start, s); break;
        case 3: {
            auto ct = make_cuda_auto_ptr<uint64_t>(words, s);
            cudaMemsetAsync(ct.get(), 0, words * 8, s);
            nwt_2d_radix8_forward_inplace_fuse_moddown(ct.get(), d2.get(), sc.get(), scs.get(), d.get(), tab, count, start, s);
            cudaMemcpyAsync(d.get(), ct.get(), words * 8, cudaMemcpyDeviceToDevice, s);
            cudaStreamSynchronize(s);
            break;
        }
        case 4: nwt_2d_radix8_forward_inplace_include_temp_mod(d.get(), tab, count, start, prm[0], s); break;
        case 5: nwt_2d_radix8_forward_inplace_include_special_mod(d.get(), tab, count, start, prm[0], prm[1], s); break;
        case 6:
            nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(d.get(), tab, count, start, prm[0], prm[1], prm[2],
                                                                            prm[3], s);
            break;
        case 7: nwt_2d_radix8_forward_modup_fuse(d.get(), d2.get(), prm[0], tab, count, start, s); break;
        case 8: nwt_2d_radix8_backward_scale(d.get(), d2.get(), tab, count, start, sc.get(), scs.get(), s); break;
        case 9: nwt_2d_radix8_backward_inplace_scale(d.get(), tab, count, start, sc.get(), scs.get(), s); break;
        case 10: nwt_2d_radix8_backward_inplace_include_special_mod(d.get(), tab, count, start, prm[0], prm[1], s); break;
        case 11:
            nwt_2d_radix8_backward_inplace_include_temp_mod_scale(d.get(), tab, count, start, prm[0], sc.get(), scs.get(), s);
            break;
        default: throw std::invalid_argument("unknown launcher");
    }
    cudaMemcpyAsync(buf, d.get(), words * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* moduli of a table family (for building test inputs): 1 Bsk, 2 QlRl */
int ref_table_moduli(void *p, int table, uint64_t *out, int cap) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto &tool = h->ctx->get_context_data(1).gpu_rns_tool();
    const DNTTTable &tab = table == 0 ? h->ctx->gpu_rns_tables() : table == 1 ? tool.gpu_Bsk_tables() : tool.gpu_QlRl_tables();
    int cnt = (int) tab.size();
    std::vector<DModulus> m(cnt);
    cudaMemcpy(m.data(), tab.modulus(), cnt * sizeof(DModulus), cudaMemcpyDeviceToHost);
    for (int i = 0; i < cnt && i < cap; i++) out[i] = m[i].value();
    return cnt;
    SHIM_CATCH
}

/* DBaseConverter::bConv_* of the converters a level owns.  which: 0 base_P_to_Ql_conv (key switch), 1 digit d's
 * part_Ql -> compl_part_QlP converter (aux = d), 2 base_Ql_to_Rl_conv, 3 base_Rl_to_Ql_conv (BFV HPS).
 * mode: 0 bConv_BEHZ, 1 bConv_BEHZ_var1, 2 bConv_HPS.  src = [ni][n], dst = [no][n] host words; returns no. */
int ref_bconv(void *p, size_t chain_index, int which, int aux, int mode, const uint64_t *src, size_t ni, uint64_t *dst,
              size_t no) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    const DBaseConverter &cv = which == 0 ? tool.base_P_to_Ql_conv()
                             : which == 1 ? tool.base_part_Ql_to_compl_part_QlP_conv(aux)
                             : which == 2 ? tool.base_Ql_to_Rl_conv() : tool.base_Rl_to_Ql_conv();
    if (cv.ibase().size() != ni || cv.obase().size() != no) throw std::invalid_argument("base sizes differ");
    auto in = make_cuda_auto_ptr<uint64_t>(ni * h->n, s);
    auto out = make_cuda_auto_ptr<uint64_t>(no * h->n, s);
    cudaMemcpyAsync(in.get(), src, ni * h->n * 8, cudaMemcpyHostToDevice, s);
    if (mode == 0) cv.bConv_BEHZ(out.get(), in.get(), h->n, s);
    else if (mode == 1) cv.bConv_BEHZ_var1(out.get(), in.get(), h->n, s);
    else cv.bConv_HPS(out.get(), in.get(), h->n, s);
    cudaMemcpyAsync(dst, out.get(), no * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* moduli of a converter's bases: out = ibase then obase; returns ni * 65536 + no */
int ref_bconv_bases(void *p, size_t chain_index, int which, int aux, uint64_t *out, int cap) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    const DBaseConverter &cv = which == 0 ? tool.base_P_to_Ql_conv()
                             : which == 1 ? tool.base_part_Ql_to_compl_part_QlP_conv(aux)
                             : which == 2 ? tool.base_Ql_to_Rl_conv() : tool.base_Rl_to_Ql_conv();
    int ni = (int) cv.ibase().size(), no = (int) cv.obase().size();
    std::vector<DModulus> m(ni + no);
    cudaMemcpy(m.data(), cv.ibase().base(), ni * sizeof(DModulus), cudaMemcpyDeviceToHost);
    cudaMemcpy(m.data() + ni, cv.obase().base(), no * sizeof(DModulus), cudaMemcpyDeviceToHost);
    for (int i = 0; i < ni + no && i < cap; i++) out[i] = m[i].value();
    return ni * 65536 + no;
    SHIM_CATCH
}

/* DRNSTool::moddown (rns_bconv.cu:712-761): cx_i = [l + size_P][n] -> out [l][n] */
int ref_moddown_plain(void *p, size_t chain_index, const uint64_t *cx_i, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    size_t l = tool.base_Ql().size(), m = l + h->size_P;
    auto buf = make_cuda_auto_ptr<uint64_t>(m * h->n, s);
    auto ct = make_cuda_auto_ptr<uint64_t>(l * h->n, s);
    cudaMemcpyAsync(buf.get(), cx_i, m * h->n * 8, cudaMemcpyHostToDevice, s);
    tool.moddown(ct.get(), buf.get(), h->ctx->gpu_rns_tables(), h->scheme, s);
    cudaMemcpyAsync(out, ct.get(), l * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* DRNSTool::divide_and_round_q_last (variant 1), divide_and_round_q_last_ntt (0), mod_t_and_divide_q_last_ntt (2):
 * src = [size][l][n] -> dst = [size][l-1][n] */
int ref_divide_round(void *p, int variant, size_t chain_index, const uint64_t *src, size_t size, uint64_t *dst) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto &tool = h->ctx->get_context_data(chain_index).gpu_rns_tool();
    size_t l = tool.base_Ql().size();
    auto in = make_cuda_auto_ptr<uint64_t>(size * l * h->n, s);
    auto out = make_cuda_auto_ptr<uint64_t>(size * (l - 1) * h->n, s);
    cudaMemcpyAsync(in.get(), src, size * l * h->n * 8, cudaMemcpyHostToDevice, s);
    if (variant == 0) tool.divide_and_round_q_last_ntt(in.get(), size, h->ctx->gpu_rns_tables(), out.get(), s);
    else if (variant == 1) tool.divide_and_round_q_last(in.get(), size, out.get(), s);
    else tool.mod_t_and_divide_q_last_ntt(in.get(), size, h->ctx->gpu_rns_tables(), out.get(), s);
    cudaMemcpyAsync(dst, out.get(), size * (l - 1) * h->n * 8, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
    SHIM_CATCH
}

/* hoisting_inplace (evaluate.cu:1670-1865) over `n_steps` steps; noise_deg = noiseScaleDeg of the input (BFV leveled) */
int ref_hoisting(void *p, size_t chain_index, const uint64_t *ct, const int *steps, int n_steps, size_t noise_deg,
                 uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    auto a = make_ct(h, chain_index, 2, ct, h->scheme != scheme_type::bfv);
    a.SetNoiseScaleDeg(noise_deg);
    hoisting_inplace(*h->ctx, a, *h->glk, std::vector<int>(steps, steps + n_steps));
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

/* keyswitch_inplace (eval_key_switch.cu:95-182) with the relin key: ct = [2][l][n] += keyswitch(c2 [l][n]) */
int ref_keyswitch(void *p, size_t chain_index, const uint64_t *ct, const uint64_t *c2, uint64_t *out) {
    SHIM_TRY
    auto h = static_cast<RefCtx *>(p);
    const auto &s = cudaStreamPerThread;
    auto a = make_ct(h, chain_index, 2, ct, h->scheme != scheme_type::bfv);
    size_t words = a.coeff_modulus_size() * h->n;
    auto d = make_cuda_auto_ptr<uint64_t>(words, s);
    cudaMemcpyAsync(d.get(), c2, words * 8, cudaMemcpyHostToDevice, s);
    keyswitch_inplace(*h->ctx, a, d.get(), *h->rlk, true, s);
    fetch_ct(a, out);
    return 0;
    SHIM_CATCH
}

} // extern "C"
