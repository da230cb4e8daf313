/*
 * phantom_on_pfhe.cu -- the reference-side binding of INTEGRATION.md section 3, compiled for real.
 *
 * Link-time substitution: this translation unit DEFINES, with the reference's own signatures, the kernel-level launchers
 * that the reference's evaluate.cu / rns.cu / rns_bconv.cu / secretkey.cu call --
 *
 *     nwt_2d_radix8_*                      (include/ntt.cuh:172-226, all twelve 2-D launchers)
 *     DRNSTool::modup                      (include/rns.cuh:156-158,  src/rns_bconv.cu:530-628)
 *     DRNSTool::moddown_from_NTT           (include/rns.cuh:163-165,  src/rns_bconv.cu:776-828)
 *     phantom::key_switch_inner_prod       (include/evaluate.cuh:25-27, src/eval_key_switch.cu:71-92)
 *
 * -- and forwards each to libpfhe_b200.so through the C-ABI of include/pfhe_b200.h.  oracle/Makefile.subst links it with the
 * reference's UNMODIFIED objects minus src/ntt/{fntt_2d,intt_2d,ntt_modup,ntt_moddown,ntt_keyswitch_old}.cu, the three
 * composite definitions above made weak in copies of rns_bconv.o / eval_key_switch.o (objcopy --weaken-symbol), into
 * oracle/_ref/libphantom_subst.so.  The reference's evaluate.cu, its examples and its benches then run on the sm_100a
 * kernels with no source change.  TEST INFRASTRUCTURE: nothing of the product links this file.
 *
 * Engines: one pfhe_engine per DNTTTable (its moduli are read back from the device once), one per (key-level table,
 * size_P, scheme, plain modulus) for the composites.  A table whose device address is reused by a later context with other
 * moduli is detected by re-reading the moduli on every call (PFHE_SUBST_TRUST_POINTERS=1 skips the check for benches).
 */
#define private public   /* DRNSTool::t_ has no accessor; layout is unchanged */
#include "phantom.h"
#undef private

#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "pfhe_b200.h"

using namespace phantom;

namespace {

void ok(int status) {
    if (status == PFHE_OK) return;
    const std::string msg = pfhe_last_error();
    if (status == PFHE_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (status == PFHE_ERR_LOGIC || status == PFHE_ERR_UNSUPPORTED) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

bool trust_pointers() {
    static const bool v = [] {
        const char *e = std::getenv("PFHE_SUBST_TRUST_POINTERS");
        return e && e[0] == '1';
    }();
    return v;
}

std::vector<uint64_t> read_moduli(const DModulus *dev, size_t count) {
    std::vector<DModulus> m(count);
    if (cudaMemcpy(m.data(), dev, count * sizeof(DModulus), cudaMemcpyDeviceToHost) != cudaSuccess)
        throw std::runtime_error("CUDA Runtime Error");
    std::vector<uint64_t> q(count);
    for (size_t i = 0; i < count; i++) q[i] = m[i].value();
    return q;
}

struct Entry {
    pfhe_engine *engine = nullptr;
    std::vector<uint64_t> moduli;
};
std::mutex g_mu;
// key: (device address of the moduli, size_P, scheme, plain modulus); NTT-only engines use size_P = 0, scheme ckks, t = 0
std::map<std::tuple<const void *, size_t, int, uint64_t>, Entry> g_engines;

pfhe_engine *engine_for(const DModulus *dev_moduli, size_t n, size_t count, size_t size_P, int scheme, uint64_t t) {
    std::lock_guard<std::mutex> g(g_mu);
    Entry &e = g_engines[{dev_moduli, size_P, scheme, t}];
    if (e.engine && trust_pointers()) return e.engine;
    std::vector<uint64_t> q = read_moduli(dev_moduli, count);
    if (e.engine && q == e.moduli) return e.engine;
    if (e.engine) pfhe_engine_destroy(e.engine), e.engine = nullptr;
    ok(pfhe_engine_create(&e.engine, scheme, n, q.data(), (int) count, (int) size_P, t, nullptr, 0));
    e.moduli = std::move(q);
    return e.engine;
}

pfhe_engine *ntt_engine(const DNTTTable &t) { return engine_for(t.modulus(), t.n(), t.size(), 0, 3, 0); }

struct ToolEngine {
    pfhe_engine *engine;
    size_t chain_index;
};
ToolEngine tool_engine(const DRNSTool &tool, const DModulus *key_moduli, const scheme_type &scheme) {
    const size_t size_QP = tool.size_QP(), size_P = tool.size_P(), l = tool.base_Ql().size();
    const uint64_t t = scheme == scheme_type::ckks ? 0 : tool.t_.value();
    pfhe_engine *e = engine_for(key_moduli, tool.n(), size_QP, size_P, (int) scheme, t);
    return {e, size_QP - size_P - l + 1};   // chain_index 1 = all of Q (context.cu:145-159)
}

// scheme of the last composite call per key-level table: key_switch_inner_prod carries no scheme argument
std::map<const void *, std::pair<int, const DModulus *>> g_tool_scheme;

}   // namespace

// ---- include/ntt.cuh:172-226 ------------------------------------------------------------------------------------------
void nwt_2d_radix8_forward_inplace(uint64_t *inout, const DNTTTable &t, size_t count, size_t start, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_forward_inplace(ntt_engine(t), PFHE_TABLE_RNS, inout, count, start, s));
}
void nwt_2d_radix8_forward_inplace_fuse_moddown(uint64_t *ct, const uint64_t *cx, const uint64_t *pinv, const uint64_t *pinv_shoup,
                                                uint64_t *delta, const DNTTTable &t, size_t count, size_t start,
                                                const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_forward_inplace_fuse_moddown(ntt_engine(t), ct, cx, pinv, pinv_shoup, delta, count, start, s));
}
void nwt_2d_radix8_forward_inplace_include_temp_mod(uint64_t *inout, const DNTTTable &t, size_t count, size_t start, size_t total,
                                                    const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_forward_inplace_include_temp_mod(ntt_engine(t), PFHE_TABLE_RNS, inout, count, start, total, s));
}
void nwt_2d_radix8_forward_inplace_include_special_mod(uint64_t *inout, const DNTTTable &t, size_t count, size_t start,
                                                       size_t size_QP, size_t size_P, const cudaStream_t &s) {
    ok(pfhe_ntt_forward_inplace_include_special_mod(ntt_engine(t), inout, count, start, size_QP, size_P, s));
}
void nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(uint64_t *inout, const DNTTTable &t, size_t count,
                                                                     size_t start, size_t size_QP, size_t size_P, size_t lo,
                                                                     size_t hi, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range(ntt_engine(t), inout, count, start, size_QP, size_P, lo, hi, s));
}
void nwt_2d_radix8_forward_modup_fuse(uint64_t *out, const uint64_t *in, size_t modulus_index, const DNTTTable &t, size_t count,
                                      size_t start, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_forward_modup_fuse(ntt_engine(t), out, in, modulus_index, count, start, s));
}
void nwt_2d_radix8_backward_inplace(uint64_t *inout, const DNTTTable &t, size_t count, size_t start, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_backward_inplace(ntt_engine(t), PFHE_TABLE_RNS, inout, count, start, s));
}
void nwt_2d_radix8_backward(uint64_t *out, const uint64_t *in, const DNTTTable &t, size_t count, size_t start, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_backward(ntt_engine(t), PFHE_TABLE_RNS, out, in, count, start, s));
}
void nwt_2d_radix8_backward_scale(uint64_t *out, const uint64_t *in, const DNTTTable &t, size_t count, size_t start,
                                  const uint64_t *scale, const uint64_t *scale_shoup, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_backward_scale(ntt_engine(t), PFHE_TABLE_RNS, out, in, count, start, scale, scale_shoup, s));
}
void nwt_2d_radix8_backward_inplace_scale(uint64_t *inout, const DNTTTable &t, size_t count, size_t start, const uint64_t *scale,
                                          const uint64_t *scale_shoup, const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_backward_inplace_scale(ntt_engine(t), PFHE_TABLE_RNS, inout, count, start, scale, scale_shoup, s));
}
void nwt_2d_radix8_backward_inplace_include_special_mod(uint64_t *inout, const DNTTTable &t, size_t count, size_t start,
                                                        size_t size_QP, size_t size_P, const cudaStream_t &s) {
    ok(pfhe_ntt_backward_inplace_include_special_mod(ntt_engine(t), inout, count, start, size_QP, size_P, s));
}
void nwt_2d_radix8_backward_inplace_include_temp_mod_scale(uint64_t *inout, const DNTTTable &t, size_t count, size_t start,
                                                           size_t total, const uint64_t *scale, const uint64_t *scale_shoup,
                                                           const cudaStream_t &s) {
    ok(pfhe_nwt_2d_radix8_backward_inplace_include_temp_mod_scale(ntt_engine(t), PFHE_TABLE_RNS, inout, count, start, total, scale,
                                                                  scale_shoup, s));
}

// ---- composites of the key switch ---------------------------------------------------------------------------------------
namespace phantom {

void DRNSTool::modup(uint64_t *dst, const uint64_t *cks, const DNTTTable &ntt_tables, const scheme_type &scheme,
                     const cudaStream_t &stream) const {
    const ToolEngine te = tool_engine(*this, ntt_tables.modulus(), scheme);
    {
        std::lock_guard<std::mutex> g(g_mu);
        g_tool_scheme[this] = {(int) scheme, ntt_tables.modulus()};
    }
    ok(pfhe_modup(te.engine, te.chain_index, dst, cks, stream));
}

void DRNSTool::moddown_from_NTT(uint64_t *ct_i, uint64_t *cx_i, const DNTTTable &ntt_tables, const scheme_type &scheme,
                                const cudaStream_t &stream) const {
    const ToolEngine te = tool_engine(*this, ntt_tables.modulus(), scheme);
    ok(pfhe_moddown_from_ntt(te.engine, te.chain_index, ct_i, cx_i, stream));
}

void key_switch_inner_prod(uint64_t *p_cx, const uint64_t *p_t_mod_up, const uint64_t *const *rlk, const DRNSTool &rns_tool,
                           const DModulus *modulus_QP, size_t, const cudaStream_t &stream) {
    int scheme = (int) scheme_type::ckks;
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_tool_scheme.find(&rns_tool);   // every caller runs modup on the same tool first (eval_key_switch.cu:147-160)
        if (it != g_tool_scheme.end()) scheme = it->second.first;
    }
    const ToolEngine te = tool_engine(rns_tool, modulus_QP, static_cast<scheme_type>(scheme));
    ok(pfhe_key_switch_inner_prod(te.engine, te.chain_index, p_cx, p_t_mod_up, rlk, stream));
}

}   // namespace phantom
