// phantom_b200.hpp -- header-only C++17 mirror of the reference's class and function names for the hot path and the
// steps either side of it, on top of the C-ABI (pfhe_b200.h).  For applications written against
// encryptorion-lab/phantom-fhe's phantom.h: the same names (PhantomContext, PhantomCiphertext, PhantomSecretKey,
// multiply_and_relin_inplace, rotate_inplace, rescale_to_next, ...), the same argument order and the same exception
// types, with device memory owned through cudaMalloc.  Everything runs on the stream given to the constructor of the
// context (default: cudaStreamPerThread, like the reference).  Differences from the reference are noted where they occur;
// the Python mirror phantom-fhe_b200/api.py is the same layer in Python.
//
// Reference: include/context.cuh, include/ciphertext.h, include/plaintext.h, include/secretkey.h, include/evaluate.cuh,
// include/batchencoder.h, include/ckks.h, include/host/encryptionparams.h, include/host/modulus.h.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <string>
#include <cstdint>
#include <istream>
#include <ostream>
#include <random>
#include <stdexcept>
#include <utility>
#include <vector>

#include "pfhe_b200.h"

namespace phantom_b200 {

enum class scheme_type : int { none = 0, bgv = 1, bfv = 2, ckks = 3 };                                     // encryptionparams.h:14-22
enum class mul_tech_type : int { none = 0, behz = 1, hps = 2, hps_overq = 3, hps_overq_leveled = 4 };      // :25-35

// status codes back to the exceptions the reference throws for the same conditions
inline void rethrow(int rc) {
    switch (rc) {
        case PFHE_OK: return;
        case PFHE_ERR_INVALID_ARGUMENT: throw std::invalid_argument(pfhe_last_error());
        case PFHE_ERR_LOGIC:
        case PFHE_ERR_UNSUPPORTED: throw std::logic_error(pfhe_last_error());
        default: throw std::runtime_error(pfhe_last_error());
    }
}
inline void cuda_check(cudaError_t e) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA Runtime Error: ") + cudaGetErrorString(e));
}

struct CoeffModulus {   // host/modulus.h:225-263
    static std::vector<uint64_t> Create(size_t poly_modulus_degree, const std::vector<int> &bit_sizes) {
        std::vector<uint64_t> out(bit_sizes.size());
        rethrow(pfhe_create_primes(poly_modulus_degree, bit_sizes.data(), (int) bit_sizes.size(), out.data()));
        return out;
    }
};
struct PlainModulus {   // host/modulus.h:317-319
    static uint64_t Batching(size_t poly_modulus_degree, int bit_size) {
        return CoeffModulus::Create(poly_modulus_degree, {bit_size})[0];
    }
};
inline uint32_t get_elt_from_step(int step, size_t coeff_count) {   // galois.cuh:16-49
    uint32_t elt = 0;
    rethrow(pfhe_galois_elt_from_step(step, coeff_count, &elt));
    return elt;
}
inline std::vector<uint32_t> get_elts_from_steps(const std::vector<int> &steps, size_t coeff_count) {
    std::vector<uint32_t> out;
    for (int s : steps) out.push_back(get_elt_from_step(s, coeff_count));
    return out;
}

class EncryptionParameters {   // host/encryptionparams.h:57-150
public:
    explicit EncryptionParameters(scheme_type scheme) : scheme_(scheme) {
        if (scheme == scheme_type::bfv) mul_tech_ = mul_tech_type::hps;
    }
    void set_poly_modulus_degree(size_t n) { n_ = n; }
    void set_coeff_modulus(const std::vector<uint64_t> &primes) { coeff_modulus_ = primes; }
    void set_special_modulus_size(size_t size_P) { size_P_ = size_P; }
    void set_plain_modulus(uint64_t t) {
        if (scheme_ == scheme_type::ckks && t) throw std::logic_error("plain_modulus is not supported for this scheme");
        t_ = t;
    }
    void set_galois_elts(const std::vector<uint32_t> &elts) { galois_elts_ = elts; }
    void set_mul_tech(mul_tech_type m) { mul_tech_ = m; }
    scheme_type scheme() const { return scheme_; }
    size_t poly_modulus_degree() const { return n_; }
    const std::vector<uint64_t> &coeff_modulus() const { return coeff_modulus_; }
    size_t special_modulus_size() const { return size_P_; }
    uint64_t plain_modulus() const { return t_; }
    const std::vector<uint32_t> &galois_elts() const { return galois_elts_; }
    mul_tech_type mul_tech() const { return mul_tech_; }

private:
    scheme_type scheme_;
    mul_tech_type mul_tech_ = mul_tech_type::none;
    size_t n_ = 0, size_P_ = 1;
    uint64_t t_ = 0;
    std::vector<uint64_t> coeff_modulus_;
    std::vector<uint32_t> galois_elts_;
};

// device words, freed with the object (the reference's cuda_auto_ptr, include/cuda_wrapper.cuh)
class DeviceWords {
public:
    DeviceWords() = default;
    explicit DeviceWords(size_t words) { resize(words); }
    DeviceWords(const DeviceWords &o) { *this = o; }
    DeviceWords &operator=(const DeviceWords &o) {
        if (this != &o) {
            resize(o.words_);
            if (words_) cuda_check(cudaMemcpy(p_, o.p_, words_ * 8, cudaMemcpyDeviceToDevice));
        }
        return *this;
    }
    DeviceWords(DeviceWords &&o) noexcept : p_(o.p_), words_(o.words_) { o.p_ = nullptr, o.words_ = 0; }
    DeviceWords &operator=(DeviceWords &&o) noexcept {
        if (this != &o) {
            release();
            p_ = o.p_, words_ = o.words_;
            o.p_ = nullptr, o.words_ = 0;
        }
        return *this;
    }
    ~DeviceWords() { release(); }
    void resize(size_t words) {
        if (words == words_) return;
        release();
        if (words) cuda_check(cudaMalloc(&p_, words * 8));
        words_ = words;
    }
    uint64_t *get() const { return p_; }
    size_t size() const { return words_; }
    void upload(const uint64_t *host, size_t words) {
        resize(words);
        cuda_check(cudaMemcpy(p_, host, words * 8, cudaMemcpyHostToDevice));
    }
    std::vector<uint64_t> download() const {
        std::vector<uint64_t> out(words_);
        cuda_check(cudaDeviceSynchronize());
        if (words_) cuda_check(cudaMemcpy(out.data(), p_, words_ * 8, cudaMemcpyDeviceToHost));
        return out;
    }

private:
    void release() {
        if (p_) cudaFree(p_);
        p_ = nullptr, words_ = 0;
    }
    uint64_t *p_ = nullptr;
    size_t words_ = 0;
};

namespace detail {
// the reference's streams are raw struct members, little-endian (include/ciphertext.h:173-213, include/plaintext.h:69-97)
template<class T>
inline void put(std::ostream &s, const T &v) {
    s.write(reinterpret_cast<const char *>(&v), sizeof(T));
}
template<class T>
inline T get(std::istream &s) {
    T v{};
    s.read(reinterpret_cast<char *>(&v), sizeof(T));
    if (!s) throw std::invalid_argument("truncated stream");
    return v;
}
inline void put_words(std::ostream &s, const DeviceWords &d, size_t first, size_t words) {
    const std::vector<uint64_t> host = d.download();
    s.write(reinterpret_cast<const char *>(host.data() + first), (std::streamsize) (words * 8));
}
inline void get_words(std::istream &s, DeviceWords &d, size_t words) {
    std::vector<uint64_t> host(words);
    s.read(reinterpret_cast<char *>(host.data()), (std::streamsize) (words * 8));
    if (!s) throw std::invalid_argument("truncated stream");
    d.upload(host.data(), words);
}
struct CipherHeader {   // 58 bytes on the wire
    uint64_t chain_index, size, n, l;
    double scale;
    uint64_t correction_factor, noise_scale_deg;
    bool is_ntt_form, is_asymmetric;
    void write(std::ostream &s) const {
        put(s, chain_index), put(s, size), put(s, n), put(s, l), put(s, scale), put(s, correction_factor), put(s, noise_scale_deg);
        put(s, is_ntt_form), put(s, is_asymmetric);
    }
    static CipherHeader read(std::istream &s) {
        CipherHeader h;
        h.chain_index = get<uint64_t>(s), h.size = get<uint64_t>(s), h.n = get<uint64_t>(s), h.l = get<uint64_t>(s);
        h.scale = get<double>(s), h.correction_factor = get<uint64_t>(s), h.noise_scale_deg = get<uint64_t>(s);
        h.is_ntt_form = get<bool>(s), h.is_asymmetric = get<bool>(s);
        return h;
    }
};
}   // namespace detail

class PhantomContext {   // include/context.cuh:118-214 as far as this path reads it
public:
    explicit PhantomContext(const EncryptionParameters &parms, cudaStream_t stream = cudaStreamPerThread)
        : parms_(parms), stream_(stream) {
        const auto &primes = parms.coeff_modulus();
        const auto &elts = parms.galois_elts();
        rethrow(pfhe_engine_create(&engine_, (int) parms.scheme(), parms.poly_modulus_degree(), primes.data(), (int) primes.size(),
                                   (int) parms.special_modulus_size(), parms.plain_modulus(), elts.empty() ? nullptr : elts.data(),
                                   (int) elts.size()));
        if (parms.scheme() == scheme_type::bfv) rethrow(pfhe_engine_set_mul_tech(engine_, (int) parms.mul_tech()));
        // key order of the context: the given elements or the reference's default set (include/galois.cuh:84-89)
        std::vector<uint32_t> effective((size_t) std::max(0, pfhe_galois_elts(engine_, nullptr, 0)));
        pfhe_galois_elts(engine_, effective.data(), (int) effective.size());
        parms_.set_galois_elts(effective);
    }
    PhantomContext(const PhantomContext &) = delete;
    PhantomContext &operator=(const PhantomContext &) = delete;
    ~PhantomContext() {
        if (engine_) pfhe_engine_destroy(engine_);
    }
    pfhe_engine *engine() const { return engine_; }
    cudaStream_t stream() const { return stream_; }
    // a ciphertext handed to the engine must have the shape its chain_index implies: the engine derives limb counts
    // from the level, not from the buffer
    template<class Ct>
    void check(const Ct &ct) const {
        if (ct.poly_modulus_degree() != poly_degree() || ct.chain_index() < 1 || ct.chain_index() > size_Q() ||
            ct.coeff_modulus_size() != coeff_modulus_size(ct.chain_index()))
            throw std::invalid_argument("ciphertext does not belong to this context");
    }
    const EncryptionParameters &parms() const { return parms_; }
    size_t poly_degree() const { return parms_.poly_modulus_degree(); }
    size_t size_QP() const { return parms_.coeff_modulus().size(); }
    size_t size_P() const { return parms_.special_modulus_size(); }
    size_t size_Q() const { return size_QP() - size_P(); }
    size_t get_first_index() const { return 1; }
    size_t coeff_modulus_size(size_t chain_index) const {   // limbs at a data level; 0 = the key level
        if (chain_index == 0) return size_QP();
        if (chain_index > size_Q()) throw std::invalid_argument("index is invalid!");
        return size_Q() - (chain_index - 1);
    }
    bool leveled() const {
        return parms_.scheme() == scheme_type::bfv && parms_.mul_tech() == mul_tech_type::hps_overq_leveled;
    }

private:
    EncryptionParameters parms_;
    cudaStream_t stream_;
    pfhe_engine *engine_ = nullptr;
};

class PhantomPlaintext {   // include/plaintext.h: BFV / BGV [N] mod t (chain_index 0), CKKS [l][N] NTT form + scale
public:
    DeviceWords data_;
    size_t chain_index_ = 0, poly_modulus_degree_ = 0;
    double scale_ = 1.0;
    uint64_t *data() const { return data_.get(); }
    size_t chain_index() const { return chain_index_; }
    double scale() const { return scale_; }
    void save(std::ostream &stream) const {   // plaintext.h:69-81
        if (!poly_modulus_degree_) throw std::invalid_argument("empty plaintext");
        const uint64_t l = data_.size() / poly_modulus_degree_;
        detail::put<uint64_t>(stream, chain_index_), detail::put<uint64_t>(stream, poly_modulus_degree_), detail::put<uint64_t>(stream, l);
        detail::put(stream, scale_);
        detail::put_words(stream, data_, 0, data_.size());
    }
    void load(std::istream &stream) {   // plaintext.h:83-97
        chain_index_ = detail::get<uint64_t>(stream);
        const uint64_t n = detail::get<uint64_t>(stream), l = detail::get<uint64_t>(stream);
        scale_ = detail::get<double>(stream);
        poly_modulus_degree_ = n;
        detail::get_words(stream, data_, n * l);
    }
};

class PhantomCiphertext {   // include/ciphertext.h:10-170
public:
    void resize(const PhantomContext &context, size_t chain_index, size_t size) {
        chain_index_ = chain_index, size_ = size;
        coeff_modulus_size_ = context.coeff_modulus_size(chain_index);
        poly_modulus_degree_ = context.poly_degree();
        data_.resize(size_ * coeff_modulus_size_ * poly_modulus_degree_);
    }
    uint64_t *data() const { return data_.get(); }
    size_t size() const { return size_; }
    size_t chain_index() const { return chain_index_; }
    size_t coeff_modulus_size() const { return coeff_modulus_size_; }
    size_t poly_modulus_degree() const { return poly_modulus_degree_; }
    double scale() const { return scale_; }
    void set_scale(double s) { scale_ = s; }
    bool is_ntt_form() const { return is_ntt_form_; }
    void set_ntt_form(bool f) { is_ntt_form_ = f; }
    uint64_t correction_factor() const { return correction_factor_; }
    void set_correction_factor(uint64_t f) { correction_factor_ = f; }
    size_t GetNoiseScaleDeg() const { return noiseScaleDeg_; }
    void SetNoiseScaleDeg(size_t d) { noiseScaleDeg_ = d; }
    bool is_asymmetric() const { return is_asymmetric_; }
    void set_asymmetric(bool a) { is_asymmetric_ = a; }
    DeviceWords &words() { return data_; }
    std::vector<uint8_t> &seed_ptr() { return seed_; }   // seed of c1 after symmetric encryption (ciphertext.h:16,163-170)
    const std::vector<uint8_t> &seed_ptr() const { return seed_; }
    void save(std::ostream &stream) const {   // ciphertext.h:173-190
        header().write(stream);
        detail::put_words(stream, data_, 0, data_.size());
    }
    void load(std::istream &stream) {   // ciphertext.h:192-213
        const auto h = detail::CipherHeader::read(stream);
        // refuse headers that cannot describe a ciphertext before sizing a device buffer from them
        if (h.size > 64 || h.l > 16384 || h.n > (size_t(1) << 17) || (h.n & (h.n - 1)) != 0)
            throw std::invalid_argument("ciphertext stream is not valid");
        adopt(h);
        detail::get_words(stream, data_, size_ * coeff_modulus_size_ * poly_modulus_degree_);
    }
    // the same, checked against the context the ciphertext is going to be used with: degree and limb count of its level
    inline void load(const class PhantomContext &context, std::istream &stream);
    void save_symmetric(std::ostream &stream) const {   // ciphertext.h:216-245: c0 and the seed of c1
        if (is_asymmetric_ || seed_.size() != 64) throw std::runtime_error("Asymmetric ciphertext does not have seed.");
        if (size_ != 2) throw std::runtime_error("This method is only for 2-polynomial ciphertext.");
        header().write(stream);
        detail::put_words(stream, data_, 0, coeff_modulus_size_ * poly_modulus_degree_);
        stream.write(reinterpret_cast<const char *>(seed_.data()), 64);
    }
    inline void load_symmetric(const class PhantomContext &context, std::istream &stream);   // ciphertext.h:247-307, below
    // a ciphertext with this one's attributes and no words yet (what the reference's resize leaves of the old object)
    PhantomCiphertext attributes_only() const {
        PhantomCiphertext c;
        c.scale_ = scale_, c.correction_factor_ = correction_factor_, c.noiseScaleDeg_ = noiseScaleDeg_;
        c.is_ntt_form_ = is_ntt_form_, c.is_asymmetric_ = is_asymmetric_;
        return c;
    }

private:
    detail::CipherHeader header() const {
        return detail::CipherHeader{chain_index_, size_, poly_modulus_degree_, coeff_modulus_size_, scale_, correction_factor_, noiseScaleDeg_,
                                    is_ntt_form_, is_asymmetric_};
    }
    void adopt(const detail::CipherHeader &h) {
        chain_index_ = h.chain_index, size_ = h.size, poly_modulus_degree_ = h.n, coeff_modulus_size_ = h.l, scale_ = h.scale;
        correction_factor_ = h.correction_factor, noiseScaleDeg_ = h.noise_scale_deg, is_ntt_form_ = h.is_ntt_form, is_asymmetric_ = h.is_asymmetric;
    }
    DeviceWords data_;
    size_t size_ = 0, chain_index_ = 0, coeff_modulus_size_ = 0, poly_modulus_degree_ = 0;
    double scale_ = 1.0;
    uint64_t correction_factor_ = 1;
    size_t noiseScaleDeg_ = 1;
    bool is_ntt_form_ = true, is_asymmetric_ = false;
    std::vector<uint8_t> seed_;
};
inline void PhantomCiphertext::load(const PhantomContext &context, std::istream &stream) {
    load(stream);
    context.check(*this);
}
inline void PhantomCiphertext::load_symmetric(const PhantomContext &context, std::istream &stream) {
    const auto h = detail::CipherHeader::read(stream);
    if (h.is_asymmetric) throw std::runtime_error("Asymmetric ciphertext does not have seed.");
    if (h.size != 2) throw std::runtime_error("This method is only for 2-polynomial ciphertext.");
    if (h.l != context.coeff_modulus_size(context.get_first_index()) || h.chain_index != context.get_first_index())
        throw std::runtime_error("Only support ciphertext without modulus switching.");
    // the sampler below writes h.l limbs of the CONTEXT's degree: a stream of another degree must not size the buffer
    if (h.n != context.poly_degree()) throw std::invalid_argument("ciphertext stream does not belong to this context");
    adopt(h);
    const size_t words = h.l * h.n;
    std::vector<uint64_t> c0(words);
    stream.read(reinterpret_cast<char *>(c0.data()), (std::streamsize) (words * 8));
    seed_.resize(64);
    stream.read(reinterpret_cast<char *>(seed_.data()), 64);
    if (!stream) throw std::invalid_argument("truncated stream");
    data_.resize(2 * words);
    cuda_check(cudaMemcpy(data_.get(), c0.data(), words * 8, cudaMemcpyHostToDevice));
    rethrow(pfhe_sample_poly(context.engine(), 2, h.l, seed_.data(), data_.get() + words, context.stream()));   // c1 from its seed
    if (!h.is_ntt_form) rethrow(pfhe_ntt_backward_inplace(context.engine(), data_.get() + words, h.l, 0, context.stream()));
}

class PhantomRelinKey {   // include/secretkey.h:102-166: dnum buffers [2][size_QP][N] + a device array of their addresses
public:
    PhantomRelinKey() = default;
    void adopt(std::vector<DeviceWords> &&digits, size_t poly_modulus_degree, size_t size_QP) {
        n_ = poly_modulus_degree, size_QP_ = size_QP;
        digits_ = std::move(digits);
        std::vector<uint64_t> addr;
        for (auto &d : digits_) addr.push_back(reinterpret_cast<uint64_t>(d.get()));
        ptrs_.upload(addr.data(), addr.size());
    }
    const uint64_t *const *public_keys_ptr() const { return reinterpret_cast<const uint64_t *const *>(ptrs_.get()); }
    size_t dnum() const { return digits_.size(); }
    // secretkey.h:129-162: dnum, then every digit as the ciphertext stream of a public key (chain_index 0, NTT form)
    void save(std::ostream &stream) const {
        detail::put<uint64_t>(stream, digits_.size());
        for (const auto &d : digits_) {
            detail::CipherHeader{0, 2, n_, size_QP_, 1.0, 1, 1, true, false}.write(stream);
            detail::put_words(stream, d, 0, d.size());
        }
    }
    void load(std::istream &stream) {
        const uint64_t dnum = detail::get<uint64_t>(stream);
        std::vector<DeviceWords> digits(dnum);
        size_t n = 0, size_QP = 0;
        for (auto &d : digits) {
            const auto h = detail::CipherHeader::read(stream);
            n = h.n, size_QP = h.l;
            detail::get_words(stream, d, h.size * h.l * h.n);
        }
        adopt(std::move(digits), n, size_QP);
    }
    // checked against the context: the key inner product walks dnum digit pointers of [2][size_QP][N] words each
    template<class Ctx>
    void load(const Ctx &context, std::istream &stream) {
        load(stream);
        if (digits_.size() != (size_t) pfhe_dnum(context.engine(), 1) || n_ != context.poly_degree() || size_QP_ != context.size_QP())
            throw std::invalid_argument("relinearisation key stream does not belong to this context");
    }

private:
    std::vector<DeviceWords> digits_;
    DeviceWords ptrs_;
    size_t n_ = 0, size_QP_ = 0;
};

class PhantomGaloisKey {   // include/secretkey.h:168-224
public:
    std::vector<PhantomRelinKey> relin_keys_;
    const PhantomRelinKey &get_relin_keys(size_t index) const { return relin_keys_.at(index); }
    void save(std::ostream &stream) const {   // secretkey.h:194-205
        detail::put<uint64_t>(stream, relin_keys_.size());
        for (const auto &k : relin_keys_) k.save(stream);
    }
    void load(std::istream &stream) {   // secretkey.h:207-219
        relin_keys_.clear();
        relin_keys_.resize(detail::get<uint64_t>(stream));
        for (auto &k : relin_keys_) k.load(stream);
    }
};

namespace detail {
struct Seed {
    uint8_t bytes[64];
};
inline Seed random_seed() {   // random_bytes, include/prng.cuh:10-32
    std::random_device rd;
    Seed s;
    for (auto &b : s.bytes) b = (uint8_t) (rd() & 0xFF);
    return s;
}
inline void require_form(const PhantomContext &context, const PhantomCiphertext &ct) {
    const auto scheme = context.parms().scheme();
    if (scheme == scheme_type::ckks && !ct.is_ntt_form()) throw std::invalid_argument("CKKS encrypted must be in NTT form");
    if (scheme == scheme_type::bgv && !ct.is_ntt_form()) throw std::invalid_argument("BGV encrypted must be in NTT form");
    if (scheme == scheme_type::bfv && ct.is_ntt_form()) throw std::invalid_argument("BFV encrypted cannot be in NTT form");
}
// bookkeeping of bgv_ckks_multiply (src/evaluate.cu:388-396): CKKS scales multiply, BGV correction factors multiply mod t
inline void after_product(const PhantomContext &context, PhantomCiphertext &dst, const PhantomCiphertext &a, const PhantomCiphertext &b) {
    if (context.parms().scheme() == scheme_type::ckks) dst.set_scale(a.scale() * b.scale());
    if (context.parms().scheme() == scheme_type::bgv)
        dst.set_correction_factor((uint64_t) ((unsigned __int128) a.correction_factor() * b.correction_factor() % context.parms().plain_modulus()));
}
inline int levels_to_drop(const PhantomContext &context, size_t depth, bool is_key_switch, bool is_asymmetric) {
    int levels = 0;
    rethrow(pfhe_find_levels_to_drop(context.engine(), depth, is_key_switch, is_asymmetric, &levels));
    return levels;
}
}   // namespace detail

class PhantomPublicKey {   // include/secretkey.h:25-100
public:
    DeviceWords pk_;   // [2][size_QP][N], NTT form
    size_t n_ = 0, size_QP_ = 0;
    void save(std::ostream &stream) const {   // secretkey.h:85-90
        if (!pk_.size()) throw std::invalid_argument("PhantomPublicKey has not been generated");
        detail::CipherHeader{0, 2, n_, size_QP_, 1.0, 1, 1, true, false}.write(stream);
        detail::put_words(stream, pk_, 0, pk_.size());
    }
    void load(std::istream &stream) {   // secretkey.h:92-96
        const auto h = detail::CipherHeader::read(stream);
        n_ = h.n, size_QP_ = h.l;
        detail::get_words(stream, pk_, h.size * h.l * h.n);
    }
    // encrypt_asymmetric (src/secretkey.cu:130-190); first data level (see pfhe_encrypt_zero_asymmetric)
    void encrypt_asymmetric(const PhantomContext &context, const PhantomPlaintext &plain, PhantomCiphertext &cipher) const {
        const auto scheme = context.parms().scheme();
        const size_t chain_index = scheme == scheme_type::ckks ? plain.chain_index() : context.get_first_index();
        cipher.resize(context, chain_index, 2);
        const auto su = detail::random_seed(), se = detail::random_seed();
        rethrow(pfhe_encrypt_zero_asymmetric(context.engine(), chain_index, pk_.get(), su.bytes, se.bytes, cipher.data(), context.stream()));
        rethrow(pfhe_encrypt_add_plain(context.engine(), chain_index, cipher.data(), plain.data(), context.stream()));
        cipher.set_ntt_form(scheme != scheme_type::bfv);
        cipher.set_scale(scheme == scheme_type::ckks ? plain.scale() : 1.0);
        cipher.set_correction_factor(1), cipher.SetNoiseScaleDeg(1), cipher.set_asymmetric(true);
    }
};

class PhantomSecretKey {   // include/secretkey.h:226-338
public:
    explicit PhantomSecretKey(const PhantomContext &context) {   // gen_secretkey, src/secretkey.cu:345-378
        n_ = context.poly_degree(), size_QP_ = context.size_QP();
        pow_.resize(context.size_QP() * context.poly_degree());
        const auto seed = detail::random_seed();
        rethrow(pfhe_gen_secretkey(context.engine(), seed.bytes, pow_.get(), context.stream()));
        powers_ = 1;
    }
    PhantomSecretKey() = default;   // to be filled by load()
    const uint64_t *secret_key_array() const { return pow_.get(); }
    void save(std::ostream &stream) const {   // secretkey.h:346-364: every power computed so far
        detail::put<uint64_t>(stream, powers_), detail::put<uint64_t>(stream, n_), detail::put<uint64_t>(stream, size_QP_);
        detail::put_words(stream, pow_, 0, pow_.size());
    }
    void load(std::istream &stream) {   // secretkey.h:366-390
        powers_ = detail::get<uint64_t>(stream);
        n_ = detail::get<uint64_t>(stream), size_QP_ = detail::get<uint64_t>(stream);
        detail::get_words(stream, pow_, powers_ * n_ * size_QP_);
    }

    PhantomPublicKey gen_publickey(const PhantomContext &context) const {   // :380-392
        PhantomPublicKey pk;
        pk.pk_.resize(2 * context.size_QP() * context.poly_degree());
        pk.n_ = context.poly_degree(), pk.size_QP_ = context.size_QP();
        const auto sa = detail::random_seed(), se = detail::random_seed();
        rethrow(pfhe_encrypt_zero_symmetric(context.engine(), 0, pow_.get(), sa.bytes, se.bytes, pk.pk_.get(), context.stream()));
        return pk;
    }
    PhantomRelinKey gen_relinkey(const PhantomContext &context) {   // :394-418
        compute_secret_key_array(context, 2);
        return kswitch_key(context, pow_.get() + context.size_QP() * context.poly_degree());
    }
    PhantomGaloisKey create_galois_keys(const PhantomContext &context) const {   // :420-461
        PhantomGaloisKey keys;
        DeviceWords rotated(context.size_QP() * context.poly_degree());
        for (uint32_t elt : context.parms().galois_elts()) {
            rethrow(pfhe_galois_secret_key(context.engine(), pow_.get(), elt, rotated.get(), context.stream()));
            keys.relin_keys_.push_back(kswitch_key(context, rotated.get()));
        }
        return keys;
    }
    void encrypt_symmetric(const PhantomContext &context, const PhantomPlaintext &plain, PhantomCiphertext &cipher) const {   // :463-530
        const auto scheme = context.parms().scheme();
        const size_t chain_index = scheme == scheme_type::ckks ? plain.chain_index() : context.get_first_index();
        cipher.resize(context, chain_index, 2);
        const auto sa = detail::random_seed(), se = detail::random_seed();
        rethrow(pfhe_encrypt_zero_symmetric(context.engine(), chain_index, pow_.get(), sa.bytes, se.bytes, cipher.data(), context.stream()));
        rethrow(pfhe_encrypt_add_plain(context.engine(), chain_index, cipher.data(), plain.data(), context.stream()));
        cipher.set_ntt_form(scheme != scheme_type::bfv);
        cipher.set_scale(scheme == scheme_type::ckks ? plain.scale() : 1.0);
        cipher.set_correction_factor(1), cipher.SetNoiseScaleDeg(1), cipher.set_asymmetric(false);
        cipher.seed_ptr().assign(sa.bytes, sa.bytes + 64);   // kept for save_symmetric, like the reference's seed_ptr()
    }
    void decrypt(const PhantomContext &context, const PhantomCiphertext &cipher, PhantomPlaintext &plain) {   // :693-723
        detail::require_form(context, cipher);
        compute_secret_key_array(context, cipher.size() > 1 ? cipher.size() - 1 : 1);
        const bool ckks = context.parms().scheme() == scheme_type::ckks;
        plain.data_.resize((ckks ? cipher.coeff_modulus_size() : 1) * context.poly_degree());
        rethrow(pfhe_decrypt(context.engine(), cipher.chain_index(), cipher.data(), cipher.size(), pow_.get(),
                             context.parms().scheme() == scheme_type::bgv ? cipher.correction_factor() : 1, plain.data(), context.stream()));
        plain.chain_index_ = ckks ? cipher.chain_index() : 0;
        plain.scale_ = ckks ? cipher.scale() : 1.0;
        plain.poly_modulus_degree_ = context.poly_degree();
    }

private:
    void compute_secret_key_array(const PhantomContext &context, size_t max_power) {   // :196-230
        const size_t words = context.size_QP() * context.poly_degree();
        if (max_power <= powers_) return;
        DeviceWords grown(max_power * words);
        cuda_check(cudaMemcpyAsync(grown.get(), pow_.get(), powers_ * words * 8, cudaMemcpyDeviceToDevice, context.stream()));
        for (size_t k = powers_; k < max_power; k++)
            rethrow(pfhe_multiply_rns_poly(context.engine(), grown.get() + (k - 1) * words, grown.get(), grown.get() + k * words,
                                           context.size_QP(), context.stream()));
        cuda_check(cudaStreamSynchronize(context.stream()));   // the old buffer is freed on return
        pow_ = std::move(grown);
        powers_ = max_power;
    }
    PhantomRelinKey kswitch_key(const PhantomContext &context, const uint64_t *new_key) const {   // :297-343
        if (context.size_P() == 0 || context.size_Q() % context.size_P()) throw std::invalid_argument("size_Q must be a multiple of size_P");
        const size_t dnum = context.size_Q() / context.size_P();
        std::vector<DeviceWords> digits;
        std::vector<uint64_t *> ptrs;
        std::vector<uint8_t> seeds;
        for (size_t d = 0; d < dnum; d++) {
            digits.emplace_back(2 * context.size_QP() * context.poly_degree());
            ptrs.push_back(digits.back().get());
            for (int k = 0; k < 2; k++) {
                const auto s = detail::random_seed();
                seeds.insert(seeds.end(), s.bytes, s.bytes + 64);
            }
        }
        rethrow(pfhe_gen_kswitch_key(context.engine(), new_key, pow_.get(), seeds.data(), ptrs.data(), context.stream()));
        PhantomRelinKey key;
        key.adopt(std::move(digits), context.poly_degree(), context.size_QP());
        return key;
    }
    DeviceWords pow_;   // [sk_max_power][size_QP][N], NTT form
    size_t powers_ = 0, n_ = 0, size_QP_ = 0;
};

class PhantomBatchEncoder {   // include/batchencoder.h, src/batchencoder.cu
public:
    explicit PhantomBatchEncoder(const PhantomContext &context) : slots_(context.poly_degree()) {
        const auto s = context.parms().scheme();
        if (s != scheme_type::bfv && s != scheme_type::bgv) throw std::invalid_argument("PhantomBatchEncoder only supports BFV/BGV scheme");
    }
    size_t slot_count() const { return slots_; }
    void encode(const PhantomContext &context, const std::vector<uint64_t> &values, PhantomPlaintext &plain) const {
        if (values.size() > slots_) throw std::logic_error("values_matrix size is too large");
        DeviceWords in;
        in.upload(values.data(), values.size());
        plain.data_.resize(slots_);
        rethrow(pfhe_batch_encode(context.engine(), in.get(), values.size(), plain.data(), context.stream()));
        cuda_check(cudaStreamSynchronize(context.stream()));   // `in` is freed on return
        plain.chain_index_ = 0, plain.scale_ = 1.0, plain.poly_modulus_degree_ = slots_;
    }
    std::vector<uint64_t> decode(const PhantomContext &context, const PhantomPlaintext &plain) const {
        DeviceWords out(slots_);
        rethrow(pfhe_batch_decode(context.engine(), plain.data(), out.get(), context.stream()));
        return out.download();
    }

private:
    size_t slots_;
};

class PhantomCKKSEncoder {   // include/ckks.h, src/ckks.cu
public:
    explicit PhantomCKKSEncoder(const PhantomContext &context) : slots_(context.poly_degree() >> 1) {
        if (context.parms().scheme() != scheme_type::ckks) throw std::invalid_argument("unsupported scheme");
    }
    size_t slot_count() const { return slots_; }
    void encode(const PhantomContext &context, const std::vector<std::complex<double>> &values, double scale, PhantomPlaintext &plain,
                size_t chain_index = 1) const {
        if (values.empty()) throw std::invalid_argument("Input vector is empty");
        if (values.size() > slots_) throw std::invalid_argument("Input vector exceeds max slots");
        double *d_in = nullptr;
        cuda_check(cudaMalloc(&d_in, values.size() * 16));
        cuda_check(cudaMemcpy(d_in, values.data(), values.size() * 16, cudaMemcpyHostToDevice));
        plain.data_.resize(context.coeff_modulus_size(chain_index) * context.poly_degree());
        const int rc = pfhe_ckks_encode(context.engine(), chain_index, d_in, values.size(), scale, plain.data(), context.stream());
        cudaStreamSynchronize(context.stream());
        cudaFree(d_in);
        rethrow(rc);
        plain.chain_index_ = chain_index, plain.scale_ = scale, plain.poly_modulus_degree_ = context.poly_degree();
    }
    void encode(const PhantomContext &context, const std::vector<double> &values, double scale, PhantomPlaintext &plain,
                size_t chain_index = 1) const {
        std::vector<std::complex<double>> z(values.begin(), values.end());
        encode(context, z, scale, plain, chain_index);
    }
    void decode(const PhantomContext &context, const PhantomPlaintext &plain, std::vector<std::complex<double>> &destination) const {
        double *d_out = nullptr;
        cuda_check(cudaMalloc(&d_out, slots_ * 16));
        const int rc = pfhe_ckks_decode(context.engine(), plain.chain_index(), plain.data(), plain.scale(), d_out, context.stream());
        destination.resize(slots_);
        cudaStreamSynchronize(context.stream());
        if (rc == PFHE_OK) cudaMemcpy(destination.data(), d_out, slots_ * 16, cudaMemcpyDeviceToHost);
        cudaFree(d_out);
        rethrow(rc);
    }
    void decode(const PhantomContext &context, const PhantomPlaintext &plain, std::vector<double> &destination) const {
        std::vector<std::complex<double>> z;
        decode(context, plain, z);
        destination.resize(z.size());
        for (size_t i = 0; i < z.size(); i++) destination[i] = z[i].real();
    }

private:
    size_t slots_;
};

// ---- evaluator (include/evaluate.cuh:37-245) ---------------------------------------------------------------------------
inline void negate_inplace(const PhantomContext &context, PhantomCiphertext &encrypted) {   // evaluate.cu:83-108
    const size_t words = encrypted.coeff_modulus_size() * encrypted.poly_modulus_degree();
    for (size_t k = 0; k < encrypted.size(); k++)
        rethrow(pfhe_negate_rns_poly(context.engine(), encrypted.data() + k * words, encrypted.data() + k * words,
                                     encrypted.coeff_modulus_size(), context.stream()));
}
namespace detail {
inline void check_pair(const PhantomCiphertext &a, const PhantomCiphertext &b) {
    if (a.chain_index() != b.chain_index()) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (a.is_ntt_form() != b.is_ntt_form()) throw std::invalid_argument("NTT form mismatch");
    const double s1 = a.scale(), s2 = b.scale();
    if (!(std::fabs(s1 - s2) < std::numeric_limits<double>::epsilon() * std::max({std::fabs(s1), std::fabs(s2), 1.0})))
        throw std::invalid_argument("scale mismatch");
    if (a.size() != b.size()) throw std::invalid_argument("poly number mismatch");
}
}   // namespace detail
namespace detail {
// balance_correction_factors (evaluate.cu:14-72): (f, e1, e2) with e1 * factor1 = e2 * factor2 = f mod t, e1 and e2
// invertible, |e1| + |e2| minimal over the remainders of the extended Euclidean algorithm on (t, factor2 / factor1)
struct Balanced {
    uint64_t f, e1, e2;
};
inline Balanced balance_correction_factors(uint64_t factor1, uint64_t factor2, uint64_t t) {
    using i128 = __int128;
    auto gcd = [](uint64_t a, uint64_t b) {
        while (b) {
            const uint64_t r = a % b;
            a = b, b = r;
        }
        return a;
    };
    auto mag = [&](uint64_t x) { return (i128) (x > t / 2 ? t - x : x); };
    uint64_t inv1;
    {   // factor1^-1 mod t
        i128 r0 = t, r1 = factor1 % t, s0 = 0, s1 = 1;
        while (r1) {
            const i128 k = r0 / r1, r2 = r0 - k * r1, s2 = s0 - k * s1;
            r0 = r1, r1 = r2, s0 = s1, s1 = s2;
        }
        if (r0 != 1) throw std::logic_error("invalid correction factor1");
        inv1 = (uint64_t) (s0 < 0 ? s0 + (i128) t : s0);
    }
    const uint64_t ratio = (uint64_t) ((unsigned __int128) inv1 * (factor2 % t) % t);
    uint64_t e1 = ratio, e2 = 1;
    i128 best = mag(e1) + mag(e2);
    i128 prev_a = t, prev_b = 0, a = ratio, b = 1;
    while (a != 0) {
        const i128 q = prev_a / a, rem = prev_a % a;
        prev_a = a, a = rem;
        const i128 nb = prev_b - b * q;
        prev_b = b, b = nb;
        const uint64_t a_mod = (uint64_t) (((a % (i128) t) + (i128) t) % (i128) t), b_mod = (uint64_t) (((b % (i128) t) + (i128) t) % (i128) t);
        if (a_mod != 0 && gcd(a_mod, t) == 1) {
            const i128 cand = mag(a_mod) + mag(b_mod);
            if (cand < best) best = cand, e1 = a_mod, e2 = b_mod;
        }
    }
    return Balanced{(uint64_t) ((unsigned __int128) e1 * (factor1 % t) % t), e1, e2};
}
// add_inplace / sub_inplace (evaluate.cu:115-338): BGV operands with different correction factors are scaled to a common
// one first (multiply_scalar_rns_poly, :148-165)
inline void add_sub(const PhantomContext &context, PhantomCiphertext &encrypted1, const PhantomCiphertext &encrypted2, bool sub, bool negate) {
    check_pair(encrypted1, encrypted2);
    const size_t l = encrypted1.coeff_modulus_size(), words = l * encrypted1.poly_modulus_degree();
    const uint64_t *other = encrypted2.data();
    DeviceWords scaled;
    if (encrypted1.correction_factor() != encrypted2.correction_factor()) {
        const Balanced bal = balance_correction_factors(encrypted1.correction_factor(), encrypted2.correction_factor(),
                                                        context.parms().plain_modulus());
        scaled.resize(encrypted2.size() * words);
        cuda_check(cudaMemcpyAsync(scaled.get(), encrypted2.data(), encrypted2.size() * words * 8, cudaMemcpyDeviceToDevice, context.stream()));
        rethrow(pfhe_multiply_scalar_rns_poly(context.engine(), encrypted1.data(), encrypted1.size(), bal.e1, l, context.stream()));
        rethrow(pfhe_multiply_scalar_rns_poly(context.engine(), scaled.get(), encrypted2.size(), bal.e2, l, context.stream()));
        encrypted1.set_correction_factor(bal.f);
        other = scaled.get();
    }
    for (size_t k = 0; k < encrypted1.size(); k++) {
        const uint64_t *a = encrypted1.data() + k * words, *b = other + k * words;
        if (sub) rethrow(pfhe_sub_rns_poly(context.engine(), negate ? b : a, negate ? a : b, encrypted1.data() + k * words, l, context.stream()));
        else rethrow(pfhe_add_rns_poly(context.engine(), a, b, encrypted1.data() + k * words, l, context.stream()));
    }
    if (scaled.size()) cuda_check(cudaStreamSynchronize(context.stream()));   // the scaled copy is freed on return
}
}   // namespace detail
inline void add_inplace(const PhantomContext &context, PhantomCiphertext &encrypted1, const PhantomCiphertext &encrypted2) {
    detail::add_sub(context, encrypted1, encrypted2, false, false);
}
inline void sub_inplace(const PhantomContext &context, PhantomCiphertext &encrypted1, const PhantomCiphertext &encrypted2,
                        bool negate = false) {
    detail::add_sub(context, encrypted1, encrypted2, true, negate);
}
inline void add_plain_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, const PhantomPlaintext &plain) {   // :1106-1164
    detail::require_form(context, encrypted);
    rethrow(pfhe_add_plain_inplace(context.engine(), encrypted.chain_index(), encrypted.data(), plain.data(),
                                   encrypted.correction_factor(), context.stream()));
}
inline void sub_plain_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, const PhantomPlaintext &plain) {   // :1166-1224
    detail::require_form(context, encrypted);
    rethrow(pfhe_sub_plain_inplace(context.engine(), encrypted.chain_index(), encrypted.data(), plain.data(),
                                   encrypted.correction_factor(), context.stream()));
}
inline void multiply_plain_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, const PhantomPlaintext &plain) {   // :1226-1340
    detail::require_form(context, encrypted);
    rethrow(pfhe_multiply_plain_inplace(context.engine(), encrypted.chain_index(), encrypted.data(), encrypted.size(), plain.data(),
                                        context.stream()));
    encrypted.set_scale(encrypted.scale() * plain.scale());
}

// multiply_inplace (evaluate.cu:1029-1057): two-polynomial operands -> three polynomials
inline void multiply_inplace(const PhantomContext &context, PhantomCiphertext &encrypted1, const PhantomCiphertext &encrypted2) {
    context.check(encrypted1);
    context.check(encrypted2);
    detail::require_form(context, encrypted1);
    detail::require_form(context, encrypted2);
    if (encrypted1.chain_index() != encrypted2.chain_index()) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (encrypted1.size() != encrypted2.size()) throw std::invalid_argument("poly number mismatch");
    const size_t s1 = encrypted1.size(), s2 = encrypted2.size();
    PhantomCiphertext dst = encrypted1.attributes_only();
    dst.resize(context, encrypted1.chain_index(), s1 + s2 - 1);
    if (context.leveled()) {
        if (s1 != 2) throw std::logic_error("dest_size must be 3 when computing BFV multiplication using HPS");
        const size_t deg = std::max(encrypted1.GetNoiseScaleDeg(), encrypted2.GetNoiseScaleDeg());
        const int drop = detail::levels_to_drop(context, deg - 1, false, encrypted1.is_asymmetric());
        rethrow(pfhe_multiply_leveled(context.engine(), encrypted1.data(), encrypted2.data(), dst.data(), drop, context.stream()));
        dst.SetNoiseScaleDeg(deg + 1);
    } else if (s1 == 2) {
        rethrow(pfhe_multiply(context.engine(), encrypted1.chain_index(), encrypted1.data(), encrypted2.data(), dst.data(), context.stream()));
    } else {
        rethrow(pfhe_multiply_sizes(context.engine(), encrypted1.chain_index(), encrypted1.data(), s1, encrypted2.data(), s2, dst.data(),
                                    context.stream()));
    }
    detail::after_product(context, dst, encrypted1, encrypted2);
    cuda_check(cudaStreamSynchronize(context.stream()));   // encrypted1's old words are freed by the move below
    encrypted1 = std::move(dst);
}
// relinearize_inplace (evaluate.cu:1342-1374)
inline void relinearize_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, const PhantomRelinKey &relin_keys) {
    context.check(encrypted);
    if (encrypted.size() != 3) throw std::invalid_argument("destination_size must be 3");
    detail::require_form(context, encrypted);
    const size_t words = encrypted.coeff_modulus_size() * encrypted.poly_modulus_degree();
    if (context.leveled()) {
        const int drop = detail::levels_to_drop(context, encrypted.GetNoiseScaleDeg() - 1, false, encrypted.is_asymmetric());
        rethrow(pfhe_keyswitch_leveled_inplace(context.engine(), encrypted.data(), encrypted.data() + 2 * words, relin_keys.public_keys_ptr(),
                                               drop, context.stream()));
    } else {
        rethrow(pfhe_relinearize_inplace(context.engine(), encrypted.chain_index(), encrypted.data(), relin_keys.public_keys_ptr(),
                                         context.stream()));
    }
    PhantomCiphertext two = encrypted.attributes_only();   // keep the first two polynomials
    two.resize(context, encrypted.chain_index(), 2);
    cuda_check(cudaMemcpyAsync(two.data(), encrypted.data(), 2 * words * 8, cudaMemcpyDeviceToDevice, context.stream()));
    cuda_check(cudaStreamSynchronize(context.stream()));
    encrypted = std::move(two);
}
// multiply_and_relin_inplace (evaluate.cu:1061-1104): the fused tensor + key switch of the engine
inline void multiply_and_relin_inplace(const PhantomContext &context, PhantomCiphertext &encrypted1, const PhantomCiphertext &encrypted2,
                                       const PhantomRelinKey &relin_keys) {
    context.check(encrypted1);
    context.check(encrypted2);
    detail::require_form(context, encrypted1);
    detail::require_form(context, encrypted2);
    if (encrypted1.chain_index() != encrypted2.chain_index()) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    detail::check_pair(encrypted1, encrypted2);
    if (encrypted1.size() != 2) throw std::invalid_argument("poly number mismatch");
    PhantomCiphertext dst = encrypted1.attributes_only();   // a separate buffer: the fused form does not work in place
    dst.resize(context, encrypted1.chain_index(), 2);
    if (context.leveled()) {
        const size_t deg = std::max(encrypted1.GetNoiseScaleDeg(), encrypted2.GetNoiseScaleDeg());
        const int drop = detail::levels_to_drop(context, deg - 1, false, encrypted1.is_asymmetric());
        rethrow(pfhe_multiply_and_relin_leveled(context.engine(), encrypted1.data(), encrypted2.data(), dst.data(), relin_keys.public_keys_ptr(),
                                                drop, context.stream()));
        dst.SetNoiseScaleDeg(deg + 1);
    } else {
        rethrow(pfhe_multiply_and_relin(context.engine(), encrypted1.chain_index(), encrypted1.data(), encrypted2.data(), dst.data(),
                                        relin_keys.public_keys_ptr(), context.stream()));
    }
    detail::after_product(context, dst, encrypted1, encrypted2);
    cuda_check(cudaStreamSynchronize(context.stream()));
    encrypted1 = std::move(dst);
}
namespace detail {
// non-adjacent form of a step as signed powers of two, lowest first (naf, include/host/numth.h:17-34)
inline std::vector<int> naf(int step) {
    std::vector<int> out;
    const bool negative = step < 0;
    long long value = negative ? -(long long) step : (long long) step;
    for (int i = 0; value; i++) {
        int zi = 0;
        if (value & 1) zi = 2 - (int) (value & 3);
        value = (value - zi) >> 1;
        if (zi) out.push_back((negative ? -1 : 1) * zi * (1 << i));
    }
    return out;
}
}   // namespace detail
// apply_galois_inplace (evaluate.cu:1567-1630)
inline void apply_galois_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, uint32_t galois_elt,
                                 const PhantomGaloisKey &galois_keys) {
    context.check(encrypted);
    if (encrypted.size() > 2) throw std::invalid_argument("encrypted size must be 2");
    const auto &elts = context.parms().galois_elts();
    const auto it = std::find(elts.begin(), elts.end(), galois_elt);
    if (it == elts.end()) throw std::invalid_argument("Galois elt not present");
    const PhantomRelinKey &key = galois_keys.get_relin_keys((size_t) (it - elts.begin()));
    if (context.leveled() && encrypted.chain_index() == 1) {
        // keyswitch_inplace with is_relin = false under mul_tech hps_overq_leveled (eval_key_switch.cu:111-123, 141-146,
        // 168-174): the switched polynomial is scaled down by the levels FindLevelsToDrop allows, switched there, expanded
        const int drop = detail::levels_to_drop(context, encrypted.GetNoiseScaleDeg() - 1, true, encrypted.is_asymmetric());
        if (drop) {
            const size_t l = encrypted.coeff_modulus_size(), words = l * encrypted.poly_modulus_degree();
            DeviceWords moved(2 * words);
            for (size_t k = 0; k < 2; k++)
                rethrow(pfhe_apply_galois(context.engine(), encrypted.data() + k * words, l, galois_elt, moved.get() + k * words, context.stream()));
            cuda_check(cudaMemcpyAsync(encrypted.data(), moved.get(), words * 8, cudaMemcpyDeviceToDevice, context.stream()));
            cuda_check(cudaMemsetAsync(encrypted.data() + words, 0, words * 8, context.stream()));
            rethrow(pfhe_keyswitch_leveled_inplace(context.engine(), encrypted.data(), moved.get() + words, key.public_keys_ptr(), drop,
                                                   context.stream()));
            cuda_check(cudaStreamSynchronize(context.stream()));   // `moved` is freed on return
            return;
        }
    }
    rethrow(pfhe_apply_galois_inplace(context.engine(), encrypted.chain_index(), encrypted.data(), galois_elt, key.public_keys_ptr(),
                                      context.stream()));
}
// rotate_inplace / rotate_internal (evaluate.cu:1633-1668): a step whose element the context holds is one automorphism;
// any other step is composed from the powers of two of its non-adjacent form (include/host/numth.h:17-34)
inline void rotate_inplace(const PhantomContext &context, PhantomCiphertext &encrypted, int step, const PhantomGaloisKey &galois_key) {
    const size_t n = context.poly_degree();
    const uint32_t elt = get_elt_from_step(step, n);
    const auto &elts = context.parms().galois_elts();
    if (std::find(elts.begin(), elts.end(), elt) != elts.end()) {
        apply_galois_inplace(context, encrypted, elt, galois_key);
        return;
    }
    const std::vector<int> naf = detail::naf(step);
    if (naf.size() == 1) throw std::invalid_argument("Galois key not present");
    for (int s : naf)
        if ((size_t) std::abs(s) != (n >> 1)) rotate_inplace(context, encrypted, s, galois_key);
}
// hoisting_inplace (evaluate.cu:1670-1865): ct <- sum over the steps of rotate(ct, step), one shared mod-up and mod-down
inline void hoisting_inplace(const PhantomContext &context, PhantomCiphertext &ct, const PhantomGaloisKey &glk, const std::vector<int> &steps) {
    context.check(ct);
    if (ct.size() > 2) throw std::invalid_argument("ciphertext size must be 2");
    if (context.parms().scheme() == scheme_type::bfv && ct.chain_index() != 1)
        throw std::invalid_argument("BFV hoisting is built for the first data level");   // evaluate.cu:1688-1708
    const auto &elts = context.parms().galois_elts();
    std::vector<const uint64_t *const *> keys;
    for (int step : steps) {
        const uint32_t elt = get_elt_from_step(step, context.poly_degree());
        const auto it = std::find(elts.begin(), elts.end(), elt);
        if (it == elts.end()) throw std::logic_error("Galois key not present in hoisting");
        keys.push_back(glk.get_relin_keys((size_t) (it - elts.begin())).public_keys_ptr());
    }
    if (context.leveled()) {   // hps_overq_leveled: a key switch at the depth of the ciphertext (evaluate.cu:1690-1701)
        const int drop = detail::levels_to_drop(context, ct.GetNoiseScaleDeg() - 1, true, ct.is_asymmetric());
        rethrow(pfhe_hoisting_leveled_inplace(context.engine(), ct.data(), steps.data(), steps.size(), keys.data(), drop, context.stream()));
        return;
    }
    rethrow(pfhe_hoisting_inplace(context.engine(), ct.chain_index(), ct.data(), steps.data(), steps.size(), keys.data(), context.stream()));
}
// rescale_to_next (evaluate.cu:1545-1565)
inline PhantomCiphertext rescale_to_next(const PhantomContext &context, const PhantomCiphertext &encrypted) {
    context.check(encrypted);
    if (context.parms().scheme() != scheme_type::ckks) throw std::invalid_argument("unsupported scheme");
    if (encrypted.chain_index() == context.size_Q()) throw std::invalid_argument("end of modulus switching chain reached");
    PhantomCiphertext dst;   // a fresh object, like the reference's `destination`: only scale and form are set below
    dst.set_ntt_form(encrypted.is_ntt_form());
    dst.resize(context, encrypted.chain_index() + 1, encrypted.size());
    rethrow(pfhe_rescale_to_next(context.engine(), encrypted.chain_index(), encrypted.data(), encrypted.size(), dst.data(), context.stream()));
    dst.set_scale(encrypted.scale() / (double) context.parms().coeff_modulus()[encrypted.coeff_modulus_size() - 1]);
    return dst;
}
// mod_switch_to_next (evaluate.cu:1505-1543)
inline PhantomCiphertext mod_switch_to_next(const PhantomContext &context, const PhantomCiphertext &encrypted) {
    context.check(encrypted);
    if (encrypted.chain_index() == context.size_Q()) throw std::invalid_argument("end of modulus switching chain reached");
    detail::require_form(context, encrypted);
    PhantomCiphertext dst;   // fresh: noiseScaleDeg and is_asymmetric restart at their defaults (evaluate.cu:1527-1541)
    dst.set_ntt_form(encrypted.is_ntt_form());
    dst.set_scale(encrypted.scale());
    dst.resize(context, encrypted.chain_index() + 1, encrypted.size());
    rethrow(pfhe_mod_switch_to_next(context.engine(), encrypted.chain_index(), encrypted.data(), encrypted.size(), dst.data(), context.stream()));
    if (context.parms().scheme() == scheme_type::bgv) {   // correction factor times q_last^-1 mod t (evaluate.cu:1420-1425)
        const uint64_t t = context.parms().plain_modulus();
        const uint64_t q = context.parms().coeff_modulus()[encrypted.coeff_modulus_size() - 1] % t;
        long long r0 = (long long) t, r1 = (long long) q, s0 = 0, s1 = 1;   // extended Euclid: s1 * q = gcd mod t
        while (r1) {
            const long long k = r0 / r1, r2 = r0 - k * r1, s2 = s0 - k * s1;
            r0 = r1, r1 = r2, s0 = s1, s1 = s2;
        }
        if (r0 != 1) throw std::logic_error("q_last is not invertible modulo the plain modulus");
        const uint64_t inv = (uint64_t) (s0 < 0 ? s0 + (long long) t : s0);
        dst.set_correction_factor((uint64_t) ((unsigned __int128) encrypted.correction_factor() * inv % t));
    }
    return dst;
}

}   // namespace phantom_b200
