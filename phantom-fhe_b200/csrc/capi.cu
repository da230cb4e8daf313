// capi.cu -- extern "C" boundary (include/pfhe_b200.h).  Translates C++ exceptions of the engine into the
// status codes documented there; no arithmetic lives here.
#include "../../include/pfhe_b200.h"

#include <cstring>
#include <string>

#include "engine.hpp"
#include "hostmath.hpp"

using namespace pfhe;

struct pfhe_engine {
    Engine impl;
    template<class... A>
    explicit pfhe_engine(A &&...a) : impl(std::forward<A>(a)...) {}
};

static thread_local std::string g_error;

const char *pfhe_last_error(void) { return g_error.c_str(); }

#define API_BEGIN try {
#define API_END                                                                                           \
    return PFHE_OK;                                                                                       \
    }                                                                                                     \
    catch (const CudaError &ex) {                                                                         \
        g_error = ex.what();                                                                              \
        return PFHE_ERR_CUDA;                                                                             \
    }                                                                                                     \
    catch (const std::invalid_argument &ex) {                                                             \
        g_error = ex.what();                                                                              \
        return PFHE_ERR_INVALID_ARGUMENT;                                                                 \
    }                                                                                                     \
    catch (const std::logic_error &ex) {                                                                  \
        g_error = ex.what();                                                                              \
        return PFHE_ERR_LOGIC;                                                                            \
    }                                                                                                     \
    catch (const std::exception &ex) {                                                                    \
        g_error = ex.what();                                                                              \
        return PFHE_ERR_CUDA;                                                                             \
    }

static cudaStream_t S(void *s) { return static_cast<cudaStream_t>(s); }
static u64 *U(uint64_t *p) { return reinterpret_cast<u64 *>(p); }
static const u64 *U(const uint64_t *p) { return reinterpret_cast<const u64 *>(p); }
static const u64 *const *K(const uint64_t *const *p) { return reinterpret_cast<const u64 *const *>(p); }

static void require(bool ok, const char *msg) {
    if (!ok) throw std::invalid_argument(msg);
}

extern "C" {

int pfhe_create_primes(uint64_t n, const int *bit_sizes, int count, uint64_t *primes_out) {
    API_BEGIN
    require(bit_sizes && primes_out && count > 0, "bit_sizes is invalid");
    auto v = host::create_primes(n, std::vector<int>(bit_sizes, bit_sizes + count));
    for (int i = 0; i < count; i++) primes_out[i] = v[i];
    API_END
}

int pfhe_engine_create(pfhe_engine **out, int scheme, uint64_t n, const uint64_t *primes, int size_QP, int size_P,
                       uint64_t plain_modulus, const uint32_t *galois_elts, int n_galois) {
    API_BEGIN
    require(out && primes && size_QP > 0, "coeff_modulus is invalid");
    require(scheme >= 1 && scheme <= 3, "unsupported scheme");
    int dev_count = 0;
    cudaError_t ce = cudaGetDeviceCount(&dev_count);
    if (ce != cudaSuccess || dev_count == 0)
        throw CudaError(ce != cudaSuccess ? ce : cudaErrorNoDevice, "pfhe_engine_create: this engine has no CPU path");
    std::vector<u64> p(primes, primes + size_QP);
    std::vector<uint32_t> g;
    if (galois_elts && n_galois > 0) g.assign(galois_elts, galois_elts + n_galois);
    *out = new pfhe_engine(static_cast<Scheme>(scheme), (size_t) n, p, size_P, (u64) plain_modulus, g);
    API_END
}

void pfhe_engine_destroy(pfhe_engine *e) { delete e; }

int pfhe_engine_set_mul_tech(pfhe_engine *e, int mul_tech) {
    API_BEGIN
    require(e != nullptr, "engine is null");
    e->impl.set_mul_tech(mul_tech);
    API_END
}
uint64_t pfhe_poly_degree(const pfhe_engine *e) { return e->impl.n(); }
int pfhe_size_QP(const pfhe_engine *e) { return e->impl.size_QP(); }
int pfhe_size_P(const pfhe_engine *e) { return e->impl.size_P(); }
int pfhe_dnum(const pfhe_engine *e, size_t chain_index) {
    try {
        return e->impl.beta(e->impl.limbs_at(chain_index));
    } catch (...) { return -1; }
}
uint64_t pfhe_launch_count(const pfhe_engine *) { return g_launches.load(); }

int pfhe_galois_elts(const pfhe_engine *e, uint32_t *elts_out, int capacity) {
    if (!e) return -1;
    const auto &g = e->impl.galois_elts();
    if (elts_out)
        for (int i = 0; i < capacity && i < (int) g.size(); i++) elts_out[i] = g[i];
    return (int) g.size();
}

int pfhe_ipc_export(const void *device_ptr, unsigned char handle_out[64], uint64_t *offset_out) {
    API_BEGIN
    require(device_ptr && handle_out && offset_out, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    // base of the allocation: the handle names the whole allocation, the mapping starts at its base
    using RangeFn = int (*)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    PFHE_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) throw std::logic_error("cuMemGetAddressRange is unavailable");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<RangeFn>(fn)(&base, &size, (unsigned long long) (uintptr_t) device_ptr) != 0)
        throw std::invalid_argument("not a device allocation");
    cudaIpcMemHandle_t h;
    PFHE_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t) base)));
    std::memcpy(handle_out, &h, 64);
    *offset_out = (uint64_t) ((uintptr_t) device_ptr - (uintptr_t) base);
    API_END
}

int pfhe_ipc_open(const unsigned char handle[64], uint64_t offset, void **mapped_out) {
    API_BEGIN
    require(handle && mapped_out, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void *base = nullptr;
    PFHE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *mapped_out = static_cast<unsigned char *>(base) + offset;
    API_END
}

int pfhe_ipc_close(void *mapped, uint64_t offset) {
    API_BEGIN
    if (mapped) PFHE_CUDA(cudaIpcCloseMemHandle(static_cast<unsigned char *>(mapped) - offset));
    API_END
}

int pfhe_enable_peer_access(int peer_device) {
    API_BEGIN
    int dev = 0, can = 0;
    PFHE_CUDA(cudaGetDevice(&dev));
    if (peer_device == dev) return PFHE_OK;
    PFHE_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) {
        g_error = "devices are not peers";
        return PFHE_ERR_UNSUPPORTED;
    }
    const cudaError_t ce = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();   // clear the sticky-less error
    else if (ce != cudaSuccess) throw CudaError(ce, "cudaDeviceEnablePeerAccess");
    API_END
}

int pfhe_galois_elt_from_step(int step, uint64_t n, uint32_t *elt_out) {
    API_BEGIN
    // get_elt_from_step, include/galois.cuh:16-49
    const uint32_t m32 = (uint32_t) (2 * n);
    if (step == 0) {
        *elt_out = m32 - 1;
    } else {
        const bool neg = step < 0;
        uint32_t pos = (uint32_t) (neg ? -step : step);
        require(pos < (n >> 1), "step count too large");
        pos &= m32 - 1;
        int s = neg ? (int) (n >> 1) - (int) pos : (int) pos;
        uint64_t elt = 1;
        while (s--) elt = (elt * 5) & ((uint64_t) m32 - 1);
        *elt_out = (uint32_t) elt;
    }
    API_END
}

// ---- NTT --------------------------------------------------------------------------------------------
int pfhe_ntt_forward_inplace(pfhe_engine *e, uint64_t *inout, size_t count, size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(start + count <= (size_t) e->impl.size_QP(), "modulus index out of range");
    // reference addressing: limbs [start, start + count) of the buffer, limb i with table row i (fntt_2d.cu:35-40)
    e->impl.ntt_fwd_rows_range(U(inout) + start * e->impl.n(), (int) count, (int) start, S(stream));
    API_END
}

int pfhe_ntt_backward_inplace(pfhe_engine *e, uint64_t *inout, size_t count, size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(start + count <= (size_t) e->impl.size_QP(), "modulus index out of range");
    u64 *p = U(inout) + start * e->impl.n();
    e->impl.ntt_inv_rows_range(p, p, (int) count, (int) start, S(stream));
    API_END
}

int pfhe_ntt_forward_inplace_batch(pfhe_engine *e, uint64_t *inout, size_t n_poly, size_t count, size_t start,
                                   void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(start + count <= (size_t) e->impl.size_QP() && n_poly * count < 32768, "modulus index out of range");
    e->impl.ntt_batch(U(inout), (int) n_poly, (int) count, (int) start, false, S(stream));
    API_END
}

int pfhe_ntt_backward_inplace_batch(pfhe_engine *e, uint64_t *inout, size_t n_poly, size_t count, size_t start,
                                    void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(start + count <= (size_t) e->impl.size_QP() && n_poly * count < 32768, "modulus index out of range");
    e->impl.ntt_batch(U(inout), (int) n_poly, (int) count, (int) start, true, S(stream));
    API_END
}

int pfhe_ntt_backward(pfhe_engine *e, uint64_t *out, const uint64_t *in, size_t count, size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(start + count <= (size_t) e->impl.size_QP(), "modulus index out of range");
    e->impl.ntt_inv_rows_range(U(out) + start * e->impl.n(), U(in) + start * e->impl.n(), (int) count, (int) start, S(stream));
    API_END
}

// the reference's "include_special_mod" launchers address limb (start+i) of the buffer and map it to
// table row (start+i) if it is below size_Ql = coeff_modulus_size_total - size_P ... : with its calling
// convention (fntt_2d.cu:434-436) limb index t of the packed buffer uses row t for t < size_QlP - size_P and
// row size_QP - size_QlP + t otherwise, where size_QlP = start + count at every reference call site.
static void special_mod(pfhe_engine *e, uint64_t *inout, size_t count, size_t start, size_t size_QP, size_t size_P,
                        bool inverse, void *stream) {
    require(size_QP <= (size_t) e->impl.size_QP() && size_P <= count, "modulus index out of range");
    Engine::NttCall c;
    c.inverse = inverse, c.table = Engine::TABLE_RNS, c.count = count, c.start = start;
    c.remap = 1, c.a = size_QP, c.b = size_P;
    e->impl.ntt_call(U(inout), U(inout), c, nullptr, nullptr, S(stream));
}

int pfhe_ntt_forward_inplace_include_special_mod(pfhe_engine *e, uint64_t *inout, size_t count, size_t start,
                                                 size_t size_QP, size_t size_P, void *stream) {
    API_BEGIN
    SmallNttScope small;
    special_mod(e, inout, count, start, size_QP, size_P, false, stream);
    API_END
}

int pfhe_ntt_backward_inplace_include_special_mod(pfhe_engine *e, uint64_t *inout, size_t count, size_t start,
                                                  size_t size_QP, size_t size_P, void *stream) {
    API_BEGIN
    SmallNttScope small;
    special_mod(e, inout, count, start, size_QP, size_P, true, stream);
    API_END
}


// ---- the remaining launchers of include/ntt.cuh:172-226 and the DRNSTool / DBaseConverter members -------------------
int pfhe_table_size(const pfhe_engine *e, int table) {
    try {
        return e ? e->impl.table_size(table) : 0;
    } catch (...) { return 0; }
}
uint64_t pfhe_table_modulus(const pfhe_engine *e, int table, size_t idx) {
    try {
        return e ? e->impl.row_modulus(e->impl.table_row(table, idx)) : 0;
    } catch (...) { return 0; }
}
static Engine::NttCall ntt_call(bool inverse, int table, size_t count, size_t start) {
    Engine::NttCall c;
    c.inverse = inverse, c.table = table, c.count = count, c.start = start;
    return c;
}
int pfhe_nwt_2d_radix8_forward_inplace(pfhe_engine *e, int table, uint64_t *inout, size_t count, size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    e->impl.ntt_call(U(inout), U(inout), ntt_call(false, table, count, start), nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_backward_inplace(pfhe_engine *e, int table, uint64_t *inout, size_t count, size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    e->impl.ntt_call(U(inout), U(inout), ntt_call(true, table, count, start), nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_backward(pfhe_engine *e, int table, uint64_t *out, const uint64_t *in, size_t count, size_t start,
                                void *stream) {
    API_BEGIN
    SmallNttScope small;
    e->impl.ntt_call(U(out), U(in), ntt_call(true, table, count, start), nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_forward_inplace_fuse_moddown(pfhe_engine *e, uint64_t *ct, const uint64_t *cx, const uint64_t *pinv,
                                                    const uint64_t *pinv_shoup, uint64_t *delta, size_t count, size_t start,
                                                    void *stream) {
    API_BEGIN
    require(ct && cx && pinv && pinv_shoup && delta, "null argument");
    e->impl.ntt_fuse_moddown(U(ct), U(cx), U(pinv), U(pinv_shoup), U(delta), count, start, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_forward_inplace_include_temp_mod(pfhe_engine *e, int table, uint64_t *inout, size_t count, size_t start,
                                                        size_t total, void *stream) {
    API_BEGIN
    SmallNttScope small;
    auto c = ntt_call(false, table, count, start);
    c.remap = 2, c.a = total;
    e->impl.ntt_call(U(inout), U(inout), c, nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(pfhe_engine *e, uint64_t *inout, size_t count,
                                                                         size_t start, size_t size_QP, size_t size_P,
                                                                         size_t excl_lo, size_t excl_hi, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(size_QP <= (size_t) e->impl.size_QP() && size_P <= count, "modulus index out of range");
    auto c = ntt_call(false, Engine::TABLE_RNS, count, start);
    c.remap = 1, c.a = size_QP, c.b = size_P, c.excl_lo = excl_lo, c.excl_hi = excl_hi;
    e->impl.ntt_call(U(inout), U(inout), c, nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_forward_modup_fuse(pfhe_engine *e, uint64_t *out, const uint64_t *in, size_t modulus_index, size_t count,
                                          size_t start, void *stream) {
    API_BEGIN
    SmallNttScope small;
    auto c = ntt_call(false, Engine::TABLE_RNS, count, start);
    c.fixed_entry = (long) modulus_index;
    e->impl.ntt_call(U(out), U(in), c, nullptr, nullptr, S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_backward_scale(pfhe_engine *e, int table, uint64_t *out, const uint64_t *in, size_t count, size_t start,
                                      const uint64_t *scale, const uint64_t *scale_shoup, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(scale && scale_shoup, "null argument");
    e->impl.ntt_call(U(out), U(in), ntt_call(true, table, count, start), U(scale), U(scale_shoup), S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_backward_inplace_scale(pfhe_engine *e, int table, uint64_t *inout, size_t count, size_t start,
                                              const uint64_t *scale, const uint64_t *scale_shoup, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(scale && scale_shoup, "null argument");
    e->impl.ntt_call(U(inout), U(inout), ntt_call(true, table, count, start), U(scale), U(scale_shoup), S(stream));
    API_END
}
int pfhe_nwt_2d_radix8_backward_inplace_include_temp_mod_scale(pfhe_engine *e, int table, uint64_t *inout, size_t count,
                                                               size_t start, size_t total, const uint64_t *scale,
                                                               const uint64_t *scale_shoup, void *stream) {
    API_BEGIN
    SmallNttScope small;
    require(scale && scale_shoup, "null argument");
    auto c = ntt_call(true, table, count, start);
    c.remap = 2, c.a = total;
    e->impl.ntt_call(U(inout), U(inout), c, U(scale), U(scale_shoup), S(stream));
    API_END
}
int pfhe_bconv(pfhe_engine *e, int mode, const uint32_t *ibase, int ni, const uint32_t *obase, int no, uint64_t *dst,
               const uint64_t *src, void *stream) {
    API_BEGIN
    require(ibase && obase && ni > 0 && no > 0 && dst && src, "base is invalid");
    std::vector<int> in_rows, out_rows;
    for (int i = 0; i < ni; i++) in_rows.push_back(e->impl.table_row((int) (ibase[i] >> 16), ibase[i] & 0xffffu));
    for (int j = 0; j < no; j++) out_rows.push_back(e->impl.table_row((int) (obase[j] >> 16), obase[j] & 0xffffu));
    e->impl.bconv(mode, e->impl.converter(in_rows, out_rows), U(dst), U(src), S(stream));
    API_END
}
int pfhe_moddown(pfhe_engine *e, size_t chain_index, uint64_t *ct_i, uint64_t *cx_i, void *stream) {
    API_BEGIN
    e->impl.moddown_plain(e->impl.limbs_at(chain_index), U(ct_i), U(cx_i), S(stream));
    API_END
}
int pfhe_divide_and_round_q_last(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t size, uint64_t *dst,
                                 void *stream) {
    API_BEGIN
    e->impl.divide_round_q_last(e->impl.limbs_at(chain_index), U(dst), U(src), (int) size, 1, S(stream));
    API_END
}
int pfhe_divide_and_round_q_last_ntt(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t size, uint64_t *dst,
                                     void *stream) {
    API_BEGIN
    e->impl.rescale(e->impl.limbs_at(chain_index), U(dst), U(src), (int) size, S(stream));
    API_END
}
int pfhe_mod_t_and_divide_q_last_ntt(pfhe_engine *e, size_t chain_index, const uint64_t *src, size_t size, uint64_t *dst,
                                     void *stream) {
    API_BEGIN
    e->impl.divide_round_q_last(e->impl.limbs_at(chain_index), U(dst), U(src), (int) size, 2, S(stream));
    API_END
}
int pfhe_add_to_ct(pfhe_engine *e, uint64_t *ct, const uint64_t *cx, size_t size_Ql, void *stream) {
    API_BEGIN
    require(size_Ql >= 1 && size_Ql <= (size_t) e->impl.size_QP(), "coeff_mod_size is invalid");
    e->impl.elementwise(EW_ADD, U(ct), U(cx), U(ct), (int) size_Ql, S(stream));
    API_END
}

// ---- dyadic -----------------------------------------------------------------------------------------
int pfhe_tensor_prod_2x2(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t l, void *stream) {
    API_BEGIN
    require(l >= 1 && l <= (size_t) e->impl.size_QP(), "coeff_mod_size is invalid");
    e->impl.tensor_2x2(U(a), U(b), U(r), (int) l, S(stream));
    API_END
}
int pfhe_tensor_square_2x2(pfhe_engine *e, const uint64_t *a, uint64_t *r, size_t l, void *stream) {
    API_BEGIN
    require(l >= 1 && l <= (size_t) e->impl.size_QP(), "coeff_mod_size is invalid");
    e->impl.tensor_square(U(a), U(r), (int) l, S(stream));
    API_END
}
#define EW_API(NAME, OP)                                                                                  \
    int NAME(pfhe_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t l, void *stream) { \
        API_BEGIN                                                                                         \
        require(l >= 1 && l <= (size_t) e->impl.size_QP(), "coeff_mod_size is invalid");                  \
        e->impl.elementwise(OP, U(a), U(b), U(r), (int) l, S(stream));                                    \
        API_END                                                                                           \
    }
EW_API(pfhe_add_rns_poly, EW_ADD)
EW_API(pfhe_sub_rns_poly, EW_SUB)
EW_API(pfhe_multiply_rns_poly, EW_MUL)
int pfhe_negate_rns_poly(pfhe_engine *e, const uint64_t *a, uint64_t *r, size_t l, void *stream) {
    API_BEGIN
    require(l >= 1 && l <= (size_t) e->impl.size_QP(), "coeff_mod_size is invalid");
    e->impl.elementwise(EW_NEG, U(a), nullptr, U(r), (int) l, S(stream));
    API_END
}

// ---- key switching ------------------------------------------------------------------------------------
int pfhe_modup(pfhe_engine *e, size_t chain_index, uint64_t *dst, const uint64_t *cks, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.modup(l, U(dst), U(cks), e->impl.ws(S(stream)).t_cks.p, S(stream));
    API_END
}
int pfhe_key_switch_inner_prod(pfhe_engine *e, size_t chain_index, uint64_t *p_cx, const uint64_t *p_t_mod_up,
                               const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.inner_prod(l, U(p_cx), U(p_t_mod_up), K(rlk), S(stream));
    API_END
}
int pfhe_moddown_from_ntt(pfhe_engine *e, size_t chain_index, uint64_t *ct_i, uint64_t *cx_i, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    if (e->impl.scheme() == Scheme::ckks) e->impl.moddown(l, U(ct_i), U(cx_i), e->impl.ws(S(stream)).delta.p, 1, nullptr, 0u, S(stream));
    else e->impl.moddown_generic(l, U(ct_i), U(cx_i), 1, nullptr, 0u, S(stream));
    API_END
}
int pfhe_keyswitch_inplace(pfhe_engine *e, size_t chain_index, uint64_t *encrypted, const uint64_t *c2,
                           const uint64_t *const *relin_keys, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.keyswitch(l, U(encrypted), U(c2), K(relin_keys), U(encrypted), S(stream));
    API_END
}

// ---- scheme level -------------------------------------------------------------------------------------
int pfhe_multiply_and_relin_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct1, const uint64_t *ct2,
                                    const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.multiply_relin(l, U(ct1), U(ct1), U(ct2), K(rlk), S(stream));
    API_END
}
int pfhe_multiply_and_relin(pfhe_engine *e, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2,
                            uint64_t *dst, const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.multiply_relin(l, U(dst), U(ct1), U(ct2), K(rlk), S(stream));
    API_END
}
int pfhe_multiply(pfhe_engine *e, size_t chain_index, const uint64_t *ct1, const uint64_t *ct2, uint64_t *dst,
                  void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    if (e->impl.scheme() == Scheme::bfv) {   // bfv_multiply (evaluate.cu:803-818): BEHZ or HPS by mul_tech
        require(dst != ct1 && dst != ct2, "destination aliases an operand");
        e->impl.bfv_multiply(l, U(dst), U(ct1), U(ct2), S(stream));
        return PFHE_OK;
    }
    if (ct1 == ct2) e->impl.tensor_square(U(ct1), U(dst), l, S(stream));
    else e->impl.tensor_2x2(U(ct1), U(ct2), U(dst), l, S(stream));
    API_END
}
int pfhe_ckks_encode(pfhe_engine *e, size_t chain_index, const double *values, size_t count, double scale,
                     uint64_t *plain, void *stream) {
    API_BEGIN
    require(e && values && plain, "null pointer");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.ckks_encode(l, reinterpret_cast<const double2 *>(values), count, scale, U(plain), S(stream));
    API_END
}
int pfhe_ckks_decode(pfhe_engine *e, size_t chain_index, const uint64_t *plain, double scale, double *values,
                     void *stream) {
    API_BEGIN
    require(e && values && plain, "null pointer");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.ckks_decode(l, U(plain), scale, reinterpret_cast<double2 *>(values), S(stream));
    API_END
}
static pfhe::Seed seed_of(const uint8_t *bytes) {
    pfhe::Seed s;
    for (int i = 0; i < 16; i++)
        s.w[i] = (uint32_t) bytes[4 * i] | ((uint32_t) bytes[4 * i + 1] << 8) | ((uint32_t) bytes[4 * i + 2] << 16) |
                 ((uint32_t) bytes[4 * i + 3] << 24);
    return s;
}
int pfhe_sample_poly(pfhe_engine *e, int kind, size_t limbs, const uint8_t *seed, uint64_t *out, void *stream) {
    API_BEGIN
    require(e && seed && out, "null pointer");
    e->impl.sample_poly(kind, (int) limbs, seed_of(seed), U(out), S(stream));
    API_END
}
int pfhe_gen_secretkey(pfhe_engine *e, const uint8_t *seed, uint64_t *secret_key, void *stream) {
    API_BEGIN
    require(e && seed && secret_key, "null pointer");
    e->impl.gen_secret_key(seed_of(seed), U(secret_key), S(stream));
    API_END
}
int pfhe_encrypt_zero_symmetric(pfhe_engine *e, size_t chain_index, const uint64_t *secret_key, const uint8_t *seed_a,
                                const uint8_t *seed_e, uint64_t *ct, void *stream) {
    API_BEGIN
    require(e && secret_key && seed_a && seed_e && ct, "null pointer");
    const int limbs = chain_index == 0 ? e->impl.size_QP() : e->impl.limbs_at(chain_index);
    const bool ntt_form = chain_index == 0 || e->impl.scheme() != pfhe::Scheme::bfv;
    e->impl.encrypt_zero_symmetric(limbs, ntt_form, U(secret_key), seed_of(seed_a), seed_of(seed_e), U(ct), S(stream));
    API_END
}
int pfhe_encrypt_zero_asymmetric(pfhe_engine *e, size_t chain_index, const uint64_t *public_key, const uint8_t *seed_u,
                                 const uint8_t *seed_e, uint64_t *ct, void *stream) {
    API_BEGIN
    require(e && public_key && seed_u && seed_e && ct, "null pointer");
    require(chain_index == 1, "asymmetric encryption is built for the first data level");
    e->impl.encrypt_zero_asymmetric(U(public_key), seed_of(seed_u), seed_of(seed_e), U(ct), S(stream));
    API_END
}
int pfhe_gen_kswitch_key(pfhe_engine *e, const uint64_t *new_key, const uint64_t *secret_key, const uint8_t *seeds,
                         uint64_t *const *digits, void *stream) {
    API_BEGIN
    require(e && new_key && secret_key && seeds && digits, "null pointer");
    require(e->impl.size_P() > 0 && e->impl.size_Q() % e->impl.size_P() == 0, "size_Q must be a multiple of size_P");
    const int dnum = e->impl.size_Q() / e->impl.size_P();
    std::vector<pfhe::Seed> sd(2 * (size_t) dnum);
    for (int i = 0; i < 2 * dnum; i++) sd[i] = seed_of(seeds + 64 * (size_t) i);
    e->impl.kswitch_key(U(new_key), U(secret_key), sd.data(), reinterpret_cast<pfhe::u64 *const *>(digits), S(stream));
    API_END
}
int pfhe_galois_secret_key(pfhe_engine *e, const uint64_t *secret_key, uint32_t galois_elt, uint64_t *rotated, void *stream) {
    API_BEGIN
    require(e && secret_key && rotated, "null pointer");
    e->impl.galois_ntt(U(secret_key), e->impl.size_QP(), galois_elt, U(rotated), S(stream));
    API_END
}
int pfhe_apply_galois_ntt(pfhe_engine *e, const uint64_t *operand, size_t coeff_mod_size, uint32_t galois_elt, uint64_t *result,
                          void *stream) {
    API_BEGIN
    require(e && operand && result, "null pointer");
    e->impl.galois_ntt(U(operand), (int) coeff_mod_size, galois_elt, U(result), S(stream));
    API_END
}
int pfhe_apply_galois(pfhe_engine *e, const uint64_t *operand, size_t coeff_mod_size, uint32_t galois_elt, uint64_t *result,
                      void *stream) {
    API_BEGIN
    require(e && operand && result && operand != result, "null pointer or in-place call");
    require(coeff_mod_size >= 1 && coeff_mod_size <= (size_t) e->impl.size_QP(), "limb count out of range");
    require((galois_elt & 1) && galois_elt < 2 * e->impl.n(), "Galois element is not valid");
    e->impl.galois_coeff(U(result), U(operand), galois_elt, (int) coeff_mod_size, 1, S(stream));
    API_END
}
int pfhe_encrypt_add_plain(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, void *stream) {
    API_BEGIN
    require(e && ct && plain, "null pointer");
    e->impl.plain_add(e->impl.limbs_at(chain_index), U(ct), U(plain), false, 1, S(stream));
    API_END
}
int pfhe_add_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t correction_factor,
                           void *stream) {
    API_BEGIN
    require(e && ct && plain, "null pointer");
    e->impl.plain_add(e->impl.limbs_at(chain_index), U(ct), U(plain), false, correction_factor, S(stream));
    API_END
}
int pfhe_sub_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *plain, uint64_t correction_factor,
                           void *stream) {
    API_BEGIN
    require(e && ct && plain, "null pointer");
    e->impl.plain_add(e->impl.limbs_at(chain_index), U(ct), U(plain), true, correction_factor, S(stream));
    API_END
}
int pfhe_multiply_plain_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, size_t size, const uint64_t *plain,
                                void *stream) {
    API_BEGIN
    require(e && ct && plain, "null pointer");
    e->impl.plain_multiply(e->impl.limbs_at(chain_index), U(ct), (int) size, U(plain), S(stream));
    API_END
}
int pfhe_multiply_scalar_rns_poly(pfhe_engine *e, uint64_t *inout, size_t size, uint64_t scalar, size_t coeff_mod_size,
                                  void *stream) {
    API_BEGIN
    require(e && inout, "null pointer");
    e->impl.multiply_scalar((int) coeff_mod_size, U(inout), (int) size, scalar, S(stream));
    API_END
}
int pfhe_batch_encode(pfhe_engine *e, const uint64_t *values, size_t count, uint64_t *plain, void *stream) {
    API_BEGIN
    require(e && plain && (values || count == 0), "null pointer");
    e->impl.batch_encode(U(values), count, U(plain), S(stream));
    API_END
}
int pfhe_batch_decode(pfhe_engine *e, const uint64_t *plain, uint64_t *values, void *stream) {
    API_BEGIN
    require(e && plain && values, "null pointer");
    e->impl.batch_decode(U(plain), U(values), S(stream));
    API_END
}
int pfhe_decrypt(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, const uint64_t *secret_key_array,
                 uint64_t correction_factor, uint64_t *destination, void *stream) {
    API_BEGIN
    require(e && ct && destination && (size == 1 || secret_key_array), "null pointer");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.decrypt(l, U(ct), (int) size, U(secret_key_array), correction_factor, U(destination), S(stream));
    API_END
}
int pfhe_find_levels_to_drop(pfhe_engine *e, size_t multiplicative_depth, int is_key_switch, int is_asymmetric,
                             int *levels) {
    API_BEGIN
    require(e && levels, "null pointer");
    *levels = e->impl.find_levels_to_drop(multiplicative_depth, is_key_switch != 0, is_asymmetric != 0);
    API_END
}
static void require_leveled(pfhe_engine *e, int drop) {
    require(e != nullptr, "engine is null");
    require(e->impl.scheme() == Scheme::bfv && e->impl.mul_tech() == 4, "needs BFV with mul_tech hps_overq_leveled");
    require(drop >= 0 && drop < e->impl.size_Q(), "levels dropped out of range");
}
int pfhe_multiply_leveled(pfhe_engine *e, const uint64_t *ct1, const uint64_t *ct2, uint64_t *dst, int levels_dropped,
                          void *stream) {
    API_BEGIN
    require_leveled(e, levels_dropped);
    require(dst != ct1 && dst != ct2, "destination aliases an operand");
    e->impl.bfv_multiply(e->impl.size_Q(), U(dst), U(ct1), U(ct2), S(stream), levels_dropped);
    API_END
}
int pfhe_multiply_and_relin_leveled(pfhe_engine *e, const uint64_t *ct1, const uint64_t *ct2, uint64_t *dst,
                                    const uint64_t *const *rlk, int levels_dropped, void *stream) {
    API_BEGIN
    require_leveled(e, levels_dropped);
    e->impl.multiply_relin_leveled(e->impl.size_Q(), U(dst), U(ct1), U(ct2), K(rlk), levels_dropped, S(stream));
    API_END
}
int pfhe_keyswitch_leveled_inplace(pfhe_engine *e, uint64_t *ct, const uint64_t *c2, const uint64_t *const *keys,
                                   int levels_dropped, void *stream) {
    API_BEGIN
    require_leveled(e, levels_dropped);
    e->impl.keyswitch_leveled(U(ct), U(c2), K(keys), levels_dropped, false, S(stream));
    API_END
}
int pfhe_fnwt_1d(uint64_t *inout, const uint64_t *twiddles, const uint64_t *twiddles_shoup, const uint64_t *modulus,
                 size_t dim, size_t coeff_modulus_size, size_t start_modulus_idx, void *stream) {
    API_BEGIN
    require(inout && twiddles && twiddles_shoup && modulus, "null pointer");
    require(dim >= 2 && dim <= 2048 && (dim & (dim - 1)) == 0, "dim must be a power of two, at most 2048");
    PFHE_CUDA(ntt_1d(false, U(inout), U(twiddles), U(twiddles_shoup), reinterpret_cast<const Modulus *>(modulus), nullptr,
                     nullptr, dim, coeff_modulus_size, start_modulus_idx, S(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    API_END
}
int pfhe_inwt_1d(uint64_t *inout, const uint64_t *itwiddles, const uint64_t *itwiddles_shoup, const uint64_t *modulus,
                 const uint64_t *scalar, const uint64_t *scalar_shoup, size_t dim, size_t coeff_modulus_size,
                 size_t start_modulus_idx, void *stream) {
    API_BEGIN
    require(inout && itwiddles && itwiddles_shoup && modulus && scalar && scalar_shoup, "null pointer");
    require(dim >= 2 && dim <= 2048 && (dim & (dim - 1)) == 0, "dim must be a power of two, at most 2048");
    PFHE_CUDA(ntt_1d(true, U(inout), U(itwiddles), U(itwiddles_shoup), reinterpret_cast<const Modulus *>(modulus),
                     U(scalar), U(scalar_shoup), dim, coeff_modulus_size, start_modulus_idx, S(stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    API_END
}
int pfhe_multiply_sizes(pfhe_engine *e, size_t chain_index, const uint64_t *ct1, size_t size1, const uint64_t *ct2,
                        size_t size2, uint64_t *dst, void *stream) {
    API_BEGIN
    require(size1 >= 1 && size2 >= 1, "invalid ciphertext size");
    if (size1 == 2 && size2 == 2) return pfhe_multiply(e, chain_index, ct1, ct2, dst, stream);
    // BFV goes through BEHZ / HPS, which the reference restricts to dest_size 3 for HPS (evaluate.cu:665-666)
    require(e->impl.scheme() != Scheme::bfv, "dest_size must be 3 when computing BFV multiplication");
    require(dst != ct2, "destination aliases the second operand");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.tensor_mxn(U(ct1), (int) size1, U(ct2), (int) size2, U(dst), l, S(stream));
    API_END
}
int pfhe_relinearize_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const uint64_t *const *rlk,
                             void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    const size_t poly = (size_t) l * e->impl.n();
    e->impl.keyswitch(l, U(ct), U(ct) + 2 * poly, K(rlk), U(ct), S(stream));
    API_END
}
int pfhe_apply_galois_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, uint32_t elt,
                              const uint64_t *const *glk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    e->impl.apply_galois(l, U(ct), elt, K(glk), S(stream));
    API_END
}
int pfhe_rotate_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, int step, const uint64_t *const *glk,
                        void *stream) {
    uint32_t elt = 0;
    int rc = pfhe_galois_elt_from_step(step, e->impl.n(), &elt);
    if (rc != PFHE_OK) return rc;
    return pfhe_apply_galois_inplace(e, chain_index, ct, elt, glk, stream);
}
int pfhe_rotate_batch(pfhe_engine *e, size_t chain_index, uint64_t *const *cts, const int *steps,
                      const uint64_t *const *const *galois_keys, size_t count, void *stream) {
    API_BEGIN
    require(cts && steps && galois_keys, "null batch pointers");
    const int l = e->impl.limbs_at(chain_index);
    std::vector<uint32_t> elts(count);
    for (size_t i = 0; i < count; i++) {
        require(cts[i] && galois_keys[i], "null batch pointers");
        for (size_t j = 0; j < i; j++) require(cts[i] != cts[j], "the same ciphertext appears twice in the batch");
        if (pfhe_galois_elt_from_step(steps[i], e->impl.n(), &elts[i]) != PFHE_OK) throw std::invalid_argument(g_error);
    }
    e->impl.apply_galois_batch(l, reinterpret_cast<u64 *const *>(cts), elts.data(),
                               reinterpret_cast<const u64 *const *const *>(galois_keys), count, S(stream));
    API_END
}
int pfhe_hoisting_inplace(pfhe_engine *e, size_t chain_index, uint64_t *ct, const int *steps, size_t n_steps,
                          const uint64_t *const *const *galois_keys, void *stream) {
    API_BEGIN
    require(steps && galois_keys && n_steps > 0, "steps is empty");
    const int l = e->impl.limbs_at(chain_index);
    std::vector<uint32_t> elts(n_steps);
    std::vector<const u64 *const *> keys(n_steps);
    for (size_t i = 0; i < n_steps; i++) {
        if (pfhe_galois_elt_from_step(steps[i], e->impl.n(), &elts[i]) != PFHE_OK) throw std::invalid_argument(g_error);
        keys[i] = K(galois_keys[i]);
    }
    e->impl.hoisting(l, U(ct), elts, keys, S(stream));
    API_END
}
int pfhe_hoisting_leveled_inplace(pfhe_engine *e, uint64_t *ct, const int *steps, size_t n_steps,
                                  const uint64_t *const *const *galois_keys, int levels_dropped, void *stream) {
    API_BEGIN
    require(steps && galois_keys && n_steps > 0, "steps is empty");
    std::vector<uint32_t> elts(n_steps);
    std::vector<const u64 *const *> keys(n_steps);
    for (size_t i = 0; i < n_steps; i++) {
        if (pfhe_galois_elt_from_step(steps[i], e->impl.n(), &elts[i]) != PFHE_OK) throw std::invalid_argument(g_error);
        keys[i] = K(galois_keys[i]);
    }
    e->impl.hoisting(e->impl.size_Q(), U(ct), elts, keys, S(stream), levels_dropped);
    API_END
}
int pfhe_rescale_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *dst,
                         void *stream) {
    API_BEGIN
    require(e->impl.scheme() == Scheme::ckks, "unsupported scheme");
    require(size >= 1 && size <= 3, "encrypted size is invalid");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.rescale(l, U(dst), U(ct), (int) size, S(stream));
    API_END
}
int pfhe_mod_switch_to_next(pfhe_engine *e, size_t chain_index, const uint64_t *ct, size_t size, uint64_t *dst,
                            void *stream) {
    API_BEGIN
    require(size >= 1 && size <= 3, "encrypted size is invalid");
    const int l = e->impl.limbs_at(chain_index);
    if (e->impl.scheme() == Scheme::ckks) e->impl.mod_switch_drop(l, U(dst), U(ct), (int) size, S(stream));
    else e->impl.mod_switch_scale(l, U(dst), U(ct), (int) size, S(stream));   // BFV / BGV: with scaling
    API_END
}

// ---- host-buffer variants ---------------------------------------------------------------------------------
int pfhe_multiply_and_relin_host(pfhe_engine *e, size_t chain_index, const uint64_t *h1, const uint64_t *h2,
                                 uint64_t *hout, const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    const size_t words = (size_t) 2 * l * e->impl.n();
    auto &io = e->impl.host_io(2 * words, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(io.p, h1, words * 8, cudaMemcpyHostToDevice, S(stream)));
    PFHE_CUDA(cudaMemcpyAsync(io.p + words, h2, words * 8, cudaMemcpyHostToDevice, S(stream)));
    e->impl.multiply_relin(l, io.p, io.p, io.p + words, K(rlk), S(stream));
    PFHE_CUDA(cudaMemcpyAsync(hout, io.p, words * 8, cudaMemcpyDeviceToHost, S(stream)));
    API_END
}
int pfhe_multiply_and_relin_host_batch(pfhe_engine *e, size_t chain_index, const uint64_t *const *h1,
                                       const uint64_t *const *h2, uint64_t *const *hout, size_t count,
                                       const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    require(h1 && h2 && hout, "null batch pointers");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.multiply_relin_host_batch(l, reinterpret_cast<const u64 *const *>(h1),
                                      reinterpret_cast<const u64 *const *>(h2), reinterpret_cast<u64 *const *>(hout),
                                      count, K(rlk), S(stream));
    API_END
}
int pfhe_multiply_and_relin_batch(pfhe_engine *e, size_t chain_index, const uint64_t *const *ct1,
                                  const uint64_t *const *ct2, uint64_t *const *dst, size_t count,
                                  const uint64_t *const *rlk, void *stream) {
    API_BEGIN
    require(ct1 && ct2 && dst, "null batch pointers");
    for (size_t i = 0; i < count; i++)
        require(dst[i] && ct1[i] && ct2[i] && dst[i] != ct1[i] && dst[i] != ct2[i], "destination aliases an operand");
    const int l = e->impl.limbs_at(chain_index);
    e->impl.multiply_relin_batch(l, reinterpret_cast<const u64 *const *>(ct1), reinterpret_cast<const u64 *const *>(ct2),
                                 reinterpret_cast<u64 *const *>(dst), count, K(rlk), S(stream));
    API_END
}
int pfhe_engine_set_lanes(pfhe_engine *e, int lanes) {
    API_BEGIN
    e->impl.set_lanes(lanes);
    API_END
}
int pfhe_engine_lanes(const pfhe_engine *e) { return e ? e->impl.lanes() : 0; }
int pfhe_rotate_host(pfhe_engine *e, size_t chain_index, const uint64_t *h, int step, uint64_t *hout,
                     const uint64_t *const *glk, void *stream) {
    API_BEGIN
    const int l = e->impl.limbs_at(chain_index);
    const size_t words = (size_t) 2 * l * e->impl.n();
    uint32_t elt = 0;
    if (pfhe_galois_elt_from_step(step, e->impl.n(), &elt) != PFHE_OK) throw std::invalid_argument(g_error);
    auto &io = e->impl.host_io(words, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(io.p, h, words * 8, cudaMemcpyHostToDevice, S(stream)));
    e->impl.apply_galois(l, io.p, elt, K(glk), S(stream));
    PFHE_CUDA(cudaMemcpyAsync(hout, io.p, words * 8, cudaMemcpyDeviceToHost, S(stream)));
    API_END
}
int pfhe_rescale_host(pfhe_engine *e, size_t chain_index, const uint64_t *h, size_t size, uint64_t *hout,
                      void *stream) {
    API_BEGIN
    require(e->impl.scheme() == Scheme::ckks, "unsupported scheme");
    const int l = e->impl.limbs_at(chain_index);
    const size_t in_words = size * l * e->impl.n(), out_words = size * (l - 1) * e->impl.n();
    auto &io = e->impl.host_io(in_words + out_words, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(io.p, h, in_words * 8, cudaMemcpyHostToDevice, S(stream)));
    e->impl.rescale(l, io.p + in_words, io.p, (int) size, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(hout, io.p + in_words, out_words * 8, cudaMemcpyDeviceToHost, S(stream)));
    API_END
}
int pfhe_ntt_forward_host(pfhe_engine *e, const uint64_t *hin, uint64_t *hout, size_t count, size_t start,
                          void *stream) {
    API_BEGIN
    require(start + count <= (size_t) e->impl.size_QP(), "modulus index out of range");
    const size_t words = count * e->impl.n();
    auto &io = e->impl.host_io(words, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(io.p, hin, words * 8, cudaMemcpyHostToDevice, S(stream)));
    e->impl.ntt_fwd_rows_range(io.p, (int) count, (int) start, S(stream));
    PFHE_CUDA(cudaMemcpyAsync(hout, io.p, words * 8, cudaMemcpyDeviceToHost, S(stream)));
    API_END
}

} // extern "C"
