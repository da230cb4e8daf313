// engine.cu -- table generation, per-level constants and op sequencing of the B200 RNS engine.
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "hostmath.hpp"
#include "launch.hpp"
#include "poly_kernels.cuh"
#include "behz_kernels.cuh"
#include "sampler_kernels.cuh"

namespace pfhe {

namespace hm = pfhe::host;

std::atomic<unsigned long long> g_launches{0};

static int ceil_log2(int v) {
    int g = 0;
    while ((1 << g) < v) g++;
    return g;
}

// FP64 form of a base-conversion matrix entry for an output modulus p: {M, M/p} and {M * 2^30 mod p, that / p}
static void push_matf(std::vector<double2> &v, u64 M, u64 p) {
    const u64 M2 = hm::mulmod(M, ((u64) 1 << fp::SPLIT_BITS) % p, p);
    v.push_back(make_double2((double) M, (double) M / (double) p));
    v.push_back(make_double2((double) M2, (double) M2 / (double) p));
}

static Tw make_tw(u64 w, u64 q) { return make_ulonglong2(w, hm::shoup(w, q)); }
static Modulus host_modulus(u64 q) {
    const hm::BarrettRatio r = hm::barrett_ratio(q);
    return Modulus{q, r.lo, r.hi};
}

// launch helper: NTT lists longer than NTT_MAX_LIMBS are cut into chunks
struct LimbVec {
    std::vector<short> data, row, src;
    void push(int d, int r, int s = -1) {
        data.push_back((short) d), row.push_back((short) r), src.push_back((short) (s < 0 ? d : s));
    }
    size_t size() const { return data.size(); }
    LimbList chunk(size_t begin, size_t &taken, const std::vector<u64> &primes) const {
        LimbList ll{};
        taken = std::min<size_t>(NTT_MAX_LIMBS, data.size() - begin);
        ll.count = (int) taken;
        for (size_t i = 0; i < taken; i++) {
            ll.data[i] = data[begin + i], ll.row[i] = row[begin + i], ll.src[i] = src[begin + i];
            ll.q[i] = primes[row[begin + i]];
        }
        return ll;
    }
};

static LimbList single_list(const LimbVec &v, const std::vector<u64> &primes) {
    if (v.size() > NTT_MAX_LIMBS) throw std::logic_error("limb list too long");
    size_t taken;
    return v.chunk(0, taken, primes);
}

// ---------------------------------------------------------------------------------------------------
Engine::Engine(Scheme scheme, size_t n, const std::vector<u64> &primes, int size_P, u64 plain_modulus,
               const std::vector<uint32_t> &galois_elts)
        : scheme_(scheme), n_(n), size_QP_((int) primes.size()), size_P_(size_P), t_(plain_modulus), primes_(primes),
          galois_elts_(galois_elts) {
    logn_ = 0;
    while (((size_t) 1 << logn_) < n_) logn_++;
    if (((size_t) 1 << logn_) != n_ || logn_ < 12 || logn_ > 17)
        throw std::invalid_argument("poly_modulus_degree is invalid");   // engine covers the 2-D NTT range 2^12..2^17
    if (size_QP_ < 1 || size_QP_ > 16384 || size_P_ < 0 || size_P_ >= size_QP_)
        throw std::invalid_argument("coeff_modulus is invalid");
    size_Q_ = size_QP_ - size_P_;
    for (u64 q : primes_) {
        if (q >> 61 || q < 2 || (q - 1) % (2 * n_) != 0 || !hm::is_prime(q))
            throw std::invalid_argument("coeff_modulus primes must be NTT-friendly primes of at most 61 bits");
    }
    {
        const char *ov = std::getenv("PFHE_OVERLAP");   // 0: keep the whole key switch on the caller's stream
        overlap_ = !(ov && ov[0] == '0');
        const char *lz = std::getenv("PFHE_LAZY_T");   // 0: the fused key switch keeps its mod-up digits canonical
        lazy_t_ = !(lz && lz[0] == '0');
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char *sc = std::getenv("PFHE_SIDE_CTAS");   // CTAs per SM of the co-running inner product (0: plain grid)
        side_ctas_ = sms * (sc ? std::atoi(sc) : 2);
    }
    build_tables();
    levels_.resize(size_Q_ + 1);
    behz_.resize(size_Q_ + 1);

    // workspace sized for the top level
    const size_t alpha = std::max(size_P_, 1);
    const size_t beta_max = (size_Q_ + alpha - 1) / alpha;
    (void) beta_max;
    if (const char *e = std::getenv("PFHE_LANES")) set_lanes(std::atoi(e));

    // no Galois elements given: all power-of-two steps in both directions plus the conjugation element, which is what
    // the NAF decomposition of rotate_inplace needs (PhantomGaloisTool::get_elts_all, reference src/galois.cu:41-65,
    // substituted in include/galois.cuh:84-89)
    if (galois_elts_.empty()) {
        const uint32_t m = (uint32_t) (2 * n_);
        galois_elts_.push_back(m - 1);
        uint64_t pos = 5, neg = 1;
        while ((neg * 5) % m != 1) neg += 2;   // 5^-1 mod 2N
        for (int i = 0; i < logn_ - 1; i++) {
            galois_elts_.push_back((uint32_t) pos);
            pos = (pos * pos) & (m - 1);
            galois_elts_.push_back((uint32_t) neg);
            neg = (neg * neg) & (m - 1);
        }
    }
    // Galois permutation tables (reference include/galois.cuh:98-113)
    d_perm_.resize(galois_elts_.size());
    std::vector<uint32_t> table(n_);
    for (size_t g = 0; g < galois_elts_.size(); g++) {
        const uint32_t elt = galois_elts_[g];
        if (!(elt & 1) || elt >= 2 * n_) throw std::invalid_argument("Galois element is not valid");
        for (size_t i = 0; i < n_; i++) {
            const uint32_t rev = hm::bit_reverse((uint32_t) (i + n_), logn_ + 1);
            const u64 raw = (((u64) elt * rev) >> 1) & (n_ - 1);
            table[i] = hm::bit_reverse((uint32_t) raw, logn_);
        }
        d_perm_[g].upload(table);
    }
}

void Engine::alloc_workspace(Workspace &w) const {
    const size_t alpha = std::max(size_P_, 1);
    const size_t beta_max = (size_Q_ + alpha - 1) / alpha;
    w.t_cks.alloc((size_t) std::max(size_Q_, 2 * (size_P_ + 1)) * n_);   // also the compact [2][alpha + 1] buffer of the BGV mod-down
    w.t_mod_up.alloc(beta_max * size_QP_ * n_);
    w.cx.alloc((size_t) 2 * size_QP_ * n_);
    w.delta.alloc((size_t) 2 * (size_Q_ + 1) * n_);
    w.tmp.alloc((size_t) 3 * size_Q_ * n_);
}

void Engine::set_lanes(int k) {
    if (k < 1 || k > MAX_LANES) throw std::invalid_argument("lane count must be 1..4");
    n_lanes_ = k;
}

static std::pair<uintptr_t, std::thread::id> stream_key(cudaStream_t st) {
    // the legacy / per-thread handles name a different stream in every thread
    const bool implicit = st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
    return {reinterpret_cast<uintptr_t>(st), implicit ? std::this_thread::get_id() : std::thread::id()};
}

StreamCtx &Engine::sctx(cudaStream_t st) const {
    std::lock_guard<std::mutex> g(sctx_mu_);
    const auto key = stream_key(st);
    auto it = sctx_.find(key);
    if (it != sctx_.end()) return *it->second;
    // contexts live as long as the engine: other threads may be inside a call that holds a reference to theirs (a context
    // is ~110 MiB at the primary set; applications that churn through streams or threads should reuse them)
    auto &slot = sctx_[key];
    slot = std::make_unique<StreamCtx>();
    alloc_workspace(slot->ws);
    return *slot;
}

StreamCtx::~StreamCtx() {
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (s_side) cudaStreamDestroy(s_side);
    for (int i = 0; i < 2; i++) {
        if (ev_in[i]) cudaEventDestroy(ev_in[i]);
        if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
        if (ev_out[i]) cudaEventDestroy(ev_out[i]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
}

// independent ops round-robin over the lanes: lane 0 is the caller's stream, lane k > 0 an internal stream (which gets
// its own workspace like any other stream); forks from and joins back into `st`
void Engine::run_lanes(size_t count, cudaStream_t st, const std::function<void(size_t, cudaStream_t)> &op) {
    const int L = (int) std::min<size_t>((size_t) n_lanes_, count);
    if (L <= 1) {
        for (size_t i = 0; i < count; i++) op(i, st);
        return;
    }
    Lanes *ln;
    {
        std::lock_guard<std::mutex> g(sctx_mu_);
        ln = &lanes_[stream_key(st)];
    }
    if (!ln->ev_done[0]) PFHE_CUDA(cudaEventCreateWithFlags(&ln->ev_done[0], cudaEventDisableTiming));
    for (int k = 1; k < L; k++)
        if (!ln->stream[k]) {
            PFHE_CUDA(cudaStreamCreateWithFlags(&ln->stream[k], cudaStreamNonBlocking));
            PFHE_CUDA(cudaEventCreateWithFlags(&ln->ev_done[k], cudaEventDisableTiming));
        }
    PFHE_CUDA(cudaEventRecord(ln->ev_done[0], st));
    for (int k = 1; k < L; k++) PFHE_CUDA(cudaStreamWaitEvent(ln->stream[k], ln->ev_done[0], 0));
    for (size_t i = 0; i < count; i++) {
        const int k = (int) (i % (size_t) L);
        op(i, k == 0 ? st : ln->stream[k]);
    }
    for (int k = 1; k < L; k++) {
        PFHE_CUDA(cudaEventRecord(ln->ev_done[k], ln->stream[k]));
        PFHE_CUDA(cudaStreamWaitEvent(st, ln->ev_done[k], 0));
    }
}

void Engine::multiply_relin_batch(int l, const u64 *const *ct1, const u64 *const *ct2, u64 *const *out, size_t count,
                                  const u64 *const *rlk, cudaStream_t st) {
    run_lanes(count, st, [&](size_t i, cudaStream_t s) { multiply_relin(l, out[i], ct1[i], ct2[i], rlk, s); });
}

void Engine::apply_galois_batch(int l, u64 *const *ct, const uint32_t *elts, const u64 *const *const *glk, size_t count,
                                cudaStream_t st) {
    run_lanes(count, st, [&](size_t i, cudaStream_t s) { apply_galois(l, ct[i], elts[i], glk[i], s); });
}

Engine::~Engine() {
    cudaDeviceSynchronize();
    sctx_.clear();
    for (auto &kv : lanes_)
        for (int k = 0; k < 4; k++) {
            if (kv.second.ev_done[k]) cudaEventDestroy(kv.second.ev_done[k]);
            if (kv.second.stream[k]) cudaStreamDestroy(kv.second.stream[k]);
        }
}

Tw Engine::make_tw_row(int row, u64 w) const {
    const u64 q = rowq_[row];
    if (!is_fp_[row]) return make_tw(w, q);
    const double dw = (double) w, dq = (double) q;
    const double winv = dw / dq;
    u64 a, b;
    std::memcpy(&a, &dw, 8);
    std::memcpy(&b, &winv, 8);
    return make_ulonglong2(a, b);
}

void Engine::build_tables() {
    // FP64 butterflies need |values| < 2^51 over up to 17 lazy stages: q < 2^46.  PFHE_FP64_NTT=0
    // forces the integer path everywhere (A/B measurements).
    const char *env = std::getenv("PFHE_FP64_NTT");
    const bool allow_fp = !(env && env[0] == '0');
    rowq_ = primes_;
    if (t_ > 1) rowq_.push_back(t_);
    batching_ = t_ > 1 && (t_ - 1) % (2 * n_) == 0 && (t_ >> 61) == 0 && hm::is_prime(t_);
    row_aux_ = (int) rowq_.size();
    if (scheme_ == Scheme::bfv && t_ > 1) {
        // BEHZ auxiliary base (rns.cu:400-420): 61-bit primes downwards from 2^61, the first is m_sk, the next
        // base_B_size(l) of them are B at level l
        int tb = 64 - __builtin_clzll(t_), nB_max = 0;
        for (int l = 1; l <= size_Q_; l++) {
            std::vector<u64> ql(primes_.begin(), primes_.begin() + l);
            const int nB = l + ((32 + tb + hm::product_bits(ql) >= 61 * l + 61) ? 1 : 0);
            nB_max = std::max(nB_max, nB);
        }
        naux_ = nB_max + 1;
        auto aux = hm::create_primes(n_, std::vector<int>(naux_, 61));
        std::reverse(aux.begin(), aux.end());   // descending, as get_primes returns them
        for (u64 p : aux) {
            if (std::find(primes_.begin(), primes_.end(), p) != primes_.end())
                throw std::invalid_argument("coeff_modulus collides with the BEHZ auxiliary base");
            rowq_.push_back(p);
        }
    }
    row_R_ = (int) rowq_.size();
    if (scheme_ == Scheme::bfv && t_ > 1) {
        // HPS auxiliary base R (rns.cu:687-694): size_Q + 1 primes below the smallest prime of Q
        mul_tech_ = 2;
        const u64 qmin = *std::min_element(primes_.begin(), primes_.begin() + size_Q_);
        try {
            for (u64 p : hm::primes_below(n_, qmin, (size_t) size_Q_ + 1)) rowq_.push_back(p);
            nR_ = size_Q_ + 1;
        } catch (const std::logic_error &) {
            nR_ = 0;   // no room below min(q): HPS is reported as unavailable when it is asked for
            rowq_.resize(row_R_);
        }
    }
    mod_rows_ = (int) rowq_.size();
    is_fp_.assign(mod_rows_, 0);
    std::vector<double2> fpc(mod_rows_);
    for (int i = 0; i < mod_rows_; i++) {
        // BEHZ rows are 61-bit; Q, P and R rows below 2^46 run on the FP64 pipe; so does the plain-modulus row when it
        // carries NTT tables (batching), because the NTT kernels pick the arithmetic from the modulus alone
        is_fp_[i] = !(t_ > 1 && i == size_QP_ && !batching_) && allow_fp && (rowq_[i] >> 46) == 0;
        fpc[i] = make_double2((double) rowq_[i], 1.0 / (double) rowq_[i]);
    }
    for (int i = 0; i < size_QP_ && i < 128; i++)
        if (is_fp_[i]) fp_mask_[i >> 6] |= 1ull << (i & 63);
    d_is_fp_.upload(is_fp_);
    d_fpc_.upload(fpc);
    // twiddle rows exist for every table row; the row of t (no NTT) stays zero
    std::vector<Tw> tw((size_t) mod_rows_ * n_), itw((size_t) mod_rows_ * n_), fin((size_t) mod_rows_ * 2);
    std::vector<Modulus> mods(mod_rows_);
    h_ninv_.assign(mod_rows_, 0);
    h_itw1_.assign(mod_rows_, 0);
    for (int i = 0; i < mod_rows_; i++) {
        const u64 q = rowq_[i];
        const auto ratio = hm::barrett_ratio(q);
        mods[i] = Modulus{q, ratio.lo, ratio.hi};
        // the row of the plain modulus gets NTT tables only when t supports batching (prime, 1 mod 2N): the BFV / BGV batch
        // encoder transforms over it (gpu_plain_tables, reference src/context.cu); otherwise the row stays zero
        if (t_ > 1 && i == size_QP_ && !batching_) continue;
        const u64 psi = hm::minimal_primitive_root(2 * n_, q);
        const u64 ipsi = hm::invmod(psi, q);
        Tw *f = tw.data() + (size_t) i * n_, *b = itw.data() + (size_t) i * n_;
        f[0] = b[0] = make_tw_row(i, 1);
        u64 pw = psi, ipw = ipsi;
        for (size_t k = 1; k < n_; k++) {
            // standard position bitrev(k) = 2^s + B  ->  kernel-native position
            const uint32_t r = hm::bit_reverse((uint32_t) k, logn_);
            int s = 31 - __builtin_clz(r);
            const size_t B = r - ((size_t) 1 << s);
            const size_t pos = tw_native_index(logn_, s, B);
            f[pos] = make_tw_row(i, pw);
            b[pos] = make_tw_row(i, ipw);
            if (r == 1) h_itw1_[i] = ipw;
            pw = hm::mulmod(pw, psi, q);
            ipw = hm::mulmod(ipw, ipsi, q);
        }
        const u64 ninv = hm::invmod(n_ % q, q);
        h_ninv_[i] = ninv;
        fin[2 * i] = make_tw_row(i, ninv);
        fin[2 * i + 1] = make_tw_row(i, hm::mulmod(h_itw1_[i], ninv, q));
    }
    d_tw_.upload(tw);
    d_itw_.upload(itw);
    d_inv_fin_.upload(fin);
    {
        std::vector<u64> fi((size_t) 2 * mod_rows_);
        for (int i = 0; i < mod_rows_; i++)
            fi[2 * i] = h_ninv_[i], fi[2 * i + 1] = h_ninv_[i] ? hm::mulmod(h_itw1_[i], h_ninv_[i], rowq_[i]) : 0;
        d_fin_int_.upload(fi);
    }
    d_mod_.upload(mods);
    // single-word Barrett constants: growth class g covers values < 2^(2k+g), k = bit length of q
    std::vector<BarG> bars((size_t) 64 * mod_rows_);
    for (int g = 0; g < 64; g++)
        for (int i = 0; i < mod_rows_; i++) {
            const u64 q = mods[i].q;
            const int k = 64 - __builtin_clzll(q);
            BarG b{0, 0xffu, 0};
            if (k + g <= 63) {
                const int sh = std::max(0, 2 * k + g - 64);
                b.sh = (u32) sh;
                b.mu = (u64) ((((unsigned __int128) 1) << (64 + sh)) / q);   // < 2^64 because sh <= k - 1
            }
            bars[(size_t) g * mod_rows_ + i] = b;
        }
    d_bar_.upload(bars);
    const char *pfe = std::getenv("PFHE_EPI_PREFETCH");
    plan_ = NttPlan{logn_, d_tw_.p, d_itw_.p, d_mod_.p, d_inv_fin_.p, d_is_fp_.p, d_fpc_.p, allow_fp ? 1 : 0, (pfe && pfe[0] == '1') ? 1 : 0};
}

const BarG *Engine::bar(int terms, int extra_bits) const {
    int g = extra_bits;
    while ((1 << (g - extra_bits)) < terms) g++;
    if (g > 63) g = 63;   // classes that do not apply to a modulus fall back to the two-word Barrett
    return d_bar_.p + (size_t) g * mod_rows_;
}

int Engine::limbs_at(size_t chain_index) const {
    if (chain_index < 1 || chain_index > (size_t) size_Q_) throw std::invalid_argument("index is invalid!");
    return size_Q_ - (int) (chain_index - 1);
}

int Engine::galois_index(uint32_t elt) const {
    auto it = std::find(galois_elts_.begin(), galois_elts_.end(), elt);
    if (it == galois_elts_.end()) throw std::invalid_argument("Galois elt not present");
    return (int) (it - galois_elts_.begin());
}

const Level &Engine::level(int l) const {
    if (l < 1 || l > size_Q_) throw std::invalid_argument("index is invalid!");
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    auto &slot = const_cast<Engine *>(this)->levels_[l];
    if (!slot) const_cast<Engine *>(this)->build_level(l);
    return *slot;
}

void Engine::build_level(int l) {
    auto lv = std::make_unique<Level>();
    lv->l = l;
    lv->alpha = size_P_;
    lv->m = l + size_P_;
    const int alpha = size_P_;
    auto row_of = [&](int j) { return j < l ? j : size_Q_ + (j - l); };

    if (alpha > 0) {
        lv->beta = beta(l);
        std::vector<Tw> fin((size_t) l * 2), finc(l);
        std::vector<u64> mat;
        std::vector<double2> matf;
        std::vector<short> omod, olimb;
        LimbVec ntt_conv;
        for (int d = 0; d < lv->beta; d++) {
            const int start = alpha * d;
            const int size = d == lv->beta - 1 ? l - alpha * (lv->beta - 1) : alpha;
            std::vector<u64> ibase(rowq_.begin() + start, rowq_.begin() + start + size);
            for (int i = 0; i < size; i++) {
                const u64 q = ibase[i];
                const u64 hinv = hm::invmod(hm::product_mod(ibase, i, q), q);
                const u64 c = hm::mulmod(hinv, h_ninv_[start + i], q);
                fin[2 * (start + i)] = make_tw_row(start + i, c);
                fin[2 * (start + i) + 1] = make_tw_row(start + i, hm::mulmod(c, h_itw1_[start + i], q));
                finc[start + i] = make_tw(hinv, q);
            }
            lv->digit_start.push_back(start);
            lv->digit_size.push_back(size);
            lv->digit_off.push_back((int) omod.size());
            int no = 0;
            for (int j = 0; j < lv->m; j++) {
                if (j >= start && j < start + size) continue;
                const int row = row_of(j);
                for (int i = 0; i < size; i++) {
                    const u64 M = hm::product_mod(ibase, i, rowq_[row]);
                    mat.push_back(M);
                    push_matf(matf, M, rowq_[row]);
                }
                omod.push_back((short) row);
                olimb.push_back((short) j);
                ntt_conv.push(d * lv->m + j, row);
                no++;
            }
            lv->digit_no.push_back(no);
            unsigned big = 0;
            for (int i = 0; i < size; i++)
                if (ibase[i] >> fp::MAX_BITS) big |= 1u << i;
            lv->digit_big.push_back(big);
        }
        lv->modup_fin.upload(fin);
        lv->modup_fin_coeff.upload(finc);
        lv->modup_mat.upload(mat);
        lv->modup_matf.upload(matf);
        lv->modup_omod.upload(omod);
        lv->modup_olimb.upload(olimb);
        lv->modup_ntt_data = ntt_conv.data, lv->modup_ntt_row = ntt_conv.row;

        // mod-down
        std::vector<u64> pbase(rowq_.begin() + size_Q_, rowq_.begin() + size_QP_);
        std::vector<Tw> dfin((size_t) 2 * alpha * 2);
        for (int k = 0; k < 2; k++)
            for (int i = 0; i < alpha; i++) {
                const u64 p = pbase[i];
                const u64 hinv = hm::invmod(hm::product_mod(pbase, i, p), p);
                const u64 c = hm::mulmod(hinv, h_ninv_[size_Q_ + i], p);
                dfin[2 * (k * alpha + i)] = make_tw_row(size_Q_ + i, c);
                dfin[2 * (k * alpha + i) + 1] = make_tw_row(size_Q_ + i, hm::mulmod(c, h_itw1_[size_Q_ + i], p));
            }
        std::vector<u64> dmat((size_t) l * alpha);
        std::vector<double2> dmatf;
        for (int i = 0; i < alpha; i++)
            if (pbase[i] >> fp::MAX_BITS) lv->moddown_big |= 1u << i;
        std::vector<short> dmod(l), dlimb(l);
        std::vector<Tw> pinv((size_t) 2 * l);
        for (int j = 0; j < l; j++) {
            const u64 q = rowq_[j];
            for (int i = 0; i < alpha; i++) {
                dmat[(size_t) j * alpha + i] = hm::product_mod(pbase, i, q);
                push_matf(dmatf, dmat[(size_t) j * alpha + i], q);
            }
            dmod[j] = (short) j, dlimb[j] = (short) j;
            pinv[j] = pinv[l + j] = make_tw(hm::invmod(hm::product_mod(pbase, -1, q), q), q);
        }
        lv->moddown_fin.upload(dfin);
        {   // all limbs of cx[k]: plain n^-1 for the Q limbs, n^-1 * phat_i^-1 for the P limbs
            std::vector<Tw> fall((size_t) 2 * lv->m * 2);
            for (int k = 0; k < 2; k++)
                for (int j = 0; j < lv->m; j++) {
                    const int row = row_of(j);
                    const u64 q = rowq_[row];
                    u64 c = h_ninv_[row];
                    if (j >= l) c = hm::mulmod(c, hm::invmod(hm::product_mod(pbase, j - l, q), q), q);
                    fall[2 * ((size_t) k * lv->m + j)] = make_tw_row(row, c);
                    fall[2 * ((size_t) k * lv->m + j) + 1] = make_tw_row(row, hm::mulmod(c, h_itw1_[row], q));
                }
            lv->moddown_fin_all.upload(fall);
        }
        {
            std::vector<Tw> pmq(l);
            for (int j = 0; j < l; j++) pmq[j] = make_tw(hm::product_mod(pbase, -1, rowq_[j]), rowq_[j]);
            lv->P_mod_q.upload(pmq);
        }
        if (t_ > 1) {
            std::vector<u64> tm;
            std::vector<double2> tmf;
            std::vector<short> tomod, tolimb;
            for (int j = 0; j <= l; j++) {
                const u64 q = j < l ? rowq_[j] : t_;
                for (int i = 0; i < alpha; i++) {
                    const u64 M = hm::product_mod(pbase, i, q);
                    tm.push_back(M);
                    push_matf(tmf, M, q);
                }
                tomod.push_back((short) (j < l ? j : size_QP_));
                tolimb.push_back((short) j);
            }
            lv->moddown_mat_t.upload(tm);
            lv->moddown_matf_t.upload(tmf);
            {   // fused BGV mod-down: the plain-modulus correction enters the conversion as one more input limb c < t with
                // matrix entry -P mod q_j (bgv_moddown_kernel, rns_bconv.cu:636-652: delta' = delta - c * (P mod q_j))
                std::vector<u64> bm;
                std::vector<double2> bmf;
                for (int j = 0; j < l; j++) {
                    const u64 q = rowq_[j];
                    for (int i = 0; i < alpha; i++) {
                        bm.push_back(hm::product_mod(pbase, i, q));
                        push_matf(bmf, bm.back(), q);
                    }
                    const u64 pm = hm::product_mod(pbase, -1, q);
                    bm.push_back(pm ? q - pm : 0);
                    push_matf(bmf, bm.back(), q);
                }
                lv->moddown_mat_bgv.upload(bm);
                lv->moddown_matf_bgv.upload(bmf);
                lv->moddown_big_bgv = lv->moddown_big | ((t_ >> fp::MAX_BITS) ? 1u << alpha : 0u);
            }
            lv->moddown_omod_t.upload(tomod);
            lv->moddown_olimb_t.upload(tolimb);
            const u64 Pt = hm::product_mod(pbase, -1, t_);
            if (std::gcd(Pt, t_) == 1) lv->pinv_t = make_tw(hm::invmod(Pt, t_), t_);
        }
        lv->moddown_mat.upload(dmat);
        lv->moddown_matf.upload(dmatf);
        lv->moddown_omod.upload(dmod);
        lv->moddown_olimb.upload(dlimb);
        lv->pinv_slots.upload(pinv);
    }
    if (l >= 2) {
        std::vector<Tw> qli((size_t) 3 * (l - 1));
        const u64 qlast = rowq_[l - 1];
        for (int j = 0; j < l - 1; j++)
            qli[j] = qli[(l - 1) + j] = qli[2 * (l - 1) + j] =
                    make_tw(hm::invmod(qlast % rowq_[j], rowq_[j]), rowq_[j]);
        lv->qlast_inv_slots.upload(qli);
        std::vector<Tw> qlm(l - 1);
        for (int j = 0; j < l - 1; j++) qlm[j] = make_tw(qlast % rowq_[j], rowq_[j]);
        lv->qlast_mod_q.upload(qlm);
        if (t_ > 1 && std::gcd(qlast % t_, t_) == 1) lv->inv_qlast_t = make_tw(hm::invmod(qlast % t_, t_), t_);
    }
    levels_[l] = std::move(lv);
}

// ---------------------------------------------------------------------------------------------------
// kernel-level ops
// ---------------------------------------------------------------------------------------------------
static void check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw CudaError(e, what);
}

// The single-launch form for few limbs is a cooperative launch: it cannot overlap the kernels around it, which the
// pipelines of this engine rely on (measured: HMult+Relin 158 us with it against 143 us without).  It pays where a transform
// stands alone -- the kernel-level NTT entry points (1- and 4-limb forward NTT 17.0 / 15.2 us against 20.2 / 20.1 us for the
// launch pair) -- so only those ask for it (SmallNttScope in capi.cu).
thread_local int g_small_ntt = 0;

FusedSync *Engine::sync_of(cudaStream_t st) const {
    Workspace &w = ws(st);
    if (!w.sync.p) {
        w.sync.alloc(1);
        PFHE_CUDA(cudaMemset(w.sync.p, 0, sizeof(FusedSync)));
    }
    return w.sync.p;
}

void Engine::ntt_fwd_list(u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) const {
    Workspace &w = ws(st);
    FusedSync *sy = sync_of(st);
    static const bool fused_env = [] {
        const char *e = std::getenv("PFHE_NTT_FUSED");
        return e && e[0] == '1';
    }();
    if (!g_small_ntt && !fused_env) sy = nullptr;   // inside a pipeline: the launch pair
    g_launches.fetch_add(2, std::memory_order_relaxed);   // (one launch under PFHE_NTT_FUSED=1: counted as the pair it replaces)
    (void) w;
    if (plan_.logn == 16 && ntt_cluster_mode()) {   // opt-in experiment: cluster per limb, intermediate in distributed shared memory
        PFHE_CUDA(ntt_forward_cluster(plan_, dst, src, ll, st));
        return;
    }
    PFHE_CUDA(ntt_forward(plan_, dst, src, ll, st, sy));
}
void Engine::ntt_inv_list(u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                          cudaStream_t st, FusedSync *sync) const {
    g_launches.fetch_add(2, std::memory_order_relaxed);
    PFHE_CUDA(ntt_inverse(plan_, dst, src, ll, fin, by_slot, st, sync ? sync : (g_small_ntt ? sync_of(st) : nullptr)));
}

static void run_chunks(const LimbVec &v, const std::vector<u64> &primes,
                       const std::function<void(const LimbList &, size_t)> &fn) {
    for (size_t b = 0; b < v.size();) {
        size_t taken;
        LimbList ll = v.chunk(b, taken, primes);
        fn(ll, b);
        b += taken;
    }
}

void Engine::ntt_fwd_rows_range(u64 *inout, int count, int start_row, cudaStream_t st) const {
    LimbVec v;
    for (int i = 0; i < count; i++) v.push(i, start_row + i);
    run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(inout, inout, ll, st); });
}

void Engine::ntt_inv_rows_range(u64 *dst, const u64 *src, int count, int start_row, cudaStream_t st) const {
    LimbVec v;
    for (int i = 0; i < count; i++) v.push(i, start_row + i);
    run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_inv_list(dst, src, ll, nullptr, 0, st); });
}

void Engine::ntt_batch(u64 *inout, int n_poly, int count, int start_row, bool inverse, cudaStream_t st) const {
    LimbVec v;
    // limbs on the (slower) integer butterflies first: their tiles start in the first wave instead of stretching the last
    for (int pass = 0; pass < 2; pass++)
        for (int p = 0; p < n_poly; p++)
            for (int i = 0; i < count; i++)
                if ((!is_fp_[start_row + i]) == (pass == 0)) v.push(p * count + i, start_row + i);
    run_chunks(v, rowq_, [&](const LimbList &ll, size_t) {
        if (inverse) ntt_inv_list(inout, inout, ll, nullptr, 0, st);
        else ntt_fwd_list(inout, inout, ll, st);
    });
}

void Engine::ntt_special_range(u64 *inout, int count, int start, int size_Ql, bool inverse, cudaStream_t st) const {
    LimbVec v;
    for (int i = 0; i < count; i++) {
        const int t = start + i;
        v.push(t, t < size_Ql ? t : size_Q_ + (t - size_Ql));
    }
    run_chunks(v, rowq_, [&](const LimbList &ll, size_t) {
        if (inverse) ntt_inv_list(inout, inout, ll, nullptr, 0, st);
        else ntt_fwd_list(inout, inout, ll, st);
    });
}

void Engine::tensor_2x2(const u64 *a, const u64 *b, u64 *out, int l, cudaStream_t st) const {
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
    launch_pdl(k_tensor_2x2, grid, EW_THREADS, 0, st, a, b, out, d_mod_.p, bar(1, 2), RowArith{d_is_fp_.p, d_fpc_.p, fp_mask_[0], fp_mask_[1], size_QP_ <= 128}, n_, l);
    check_launch("k_tensor_2x2");
}

void Engine::tensor_square(const u64 *a, u64 *out, int l, cudaStream_t st) const {
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
    launch_pdl(k_tensor_square, grid, EW_THREADS, 0, st, a, out, d_mod_.p, bar(1, 0), n_, l);
    check_launch("k_tensor_square");
}

void Engine::tensor_mxn(const u64 *a, int sa, const u64 *b, int sb, u64 *out, int l, cudaStream_t st) const {
    if (sa < 1 || sb < 1 || sa > MXN_MAX || sb > MXN_MAX) throw std::invalid_argument("ciphertext size is not supported");
    dim3 grid((unsigned) (n_ / EW_THREADS), l);
    launch_pdl(k_tensor_mxn, grid, EW_THREADS, 0, st, a, sa, b, sb, out, d_mod_.p, n_, l);
    check_launch("k_tensor_mxn");
}

void Engine::elementwise(int op, const u64 *a, const u64 *b, u64 *out, int l, cudaStream_t st) const {
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
    switch (op) {
        case EW_ADD: launch_pdl(k_elementwise<EW_ADD>, grid, EW_THREADS, 0, st, a, b, out, d_mod_.p, bar(1, 0), n_); break;
        case EW_SUB: launch_pdl(k_elementwise<EW_SUB>, grid, EW_THREADS, 0, st, a, b, out, d_mod_.p, bar(1, 0), n_); break;
        case EW_MUL: launch_pdl(k_elementwise<EW_MUL>, grid, EW_THREADS, 0, st, a, b, out, d_mod_.p, bar(1, 0), n_); break;
        case EW_NEG: launch_pdl(k_elementwise<EW_NEG>, grid, EW_THREADS, 0, st, a, b, out, d_mod_.p, bar(1, 0), n_); break;
        default: throw std::invalid_argument("unknown elementwise op");
    }
    check_launch("k_elementwise");
}

static void launch_bconv(const BconvBatch &batch, int jobs, int ni, int no_max, const Modulus *mod, const BarG *bar,
                         RowArith ra, int size_QP, size_t n, cudaStream_t st) {
    dim3 grid((unsigned) (n / (2 * EW_THREADS)), jobs);
    const size_t smem = (size_t) no_max * ni * (8 + 2 * sizeof(double2)) +
                        (size_t) no_max * (sizeof(Modulus) + sizeof(BarG) + sizeof(double2) + sizeof(int)) + 16;
    switch (ni) {
        case 1: launch_pdl(k_bconv<1>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        case 2: launch_pdl(k_bconv<2>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        case 3: launch_pdl(k_bconv<3>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        case 4: launch_pdl(k_bconv<4>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        case 5: launch_pdl(k_bconv<5>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        case 6: launch_pdl(k_bconv<6>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
        default: launch_pdl(k_bconv<0>, grid, EW_THREADS, smem, st, batch, mod, bar, ra, size_QP, n); break;
    }
    check_launch("k_bconv");
}

// DRNSTool::modup (reference src/rns_bconv.cu:530-628), CKKS/BGV form: cks in NTT domain
void Engine::modup(int l, u64 *t_mod_up, const u64 *cks, u64 *t_cks, cudaStream_t st) const {
    const Level &lv = level(l);
    if (lv.alpha == 0) throw std::logic_error("key switching needs special primes");
    const bool bfv = scheme_ == Scheme::bfv;
    if (!bfv) {
        // 1. inverse NTT fused with the n^-1 * qhat_i^-1 scaling (iNTT+scale, rns_bconv.cu:558)
        LimbVec v;
        for (int i = 0; i < l; i++) v.push(i, i);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t b) { ntt_inv_list(t_cks, cks, ll, lv.modup_fin.p + 2 * b, 1, st); });
    } else {
        // BFV: cks is already in coefficient form, only the qhat_i^-1 scaling (bconv_mult_kernel, :598)
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
        launch_pdl(k_scale_limbs, grid, EW_THREADS, 0, st, t_cks, cks, lv.modup_fin_coeff.p, d_mod_.p, n_);
        check_launch("k_scale_limbs");
    }
    // 2. each digit: own limbs copied (modup_copy_partQl_kernel :522-528), other limbs converted (:455-485)
    for (int d = 0; d < lv.beta;) {
        BconvBatch batch{};
        int jobs = 0, no_max = 0;
        const int ni = lv.digit_size[d];
        while (d < lv.beta && jobs < BCONV_MAX_JOBS && lv.digit_size[d] == ni) {
            const int start = lv.digit_start[d], off = lv.digit_off[d];
            u64 *dst = t_mod_up + (size_t) d * lv.m * n_;
            PFHE_CUDA(cudaMemcpyAsync(dst + (size_t) start * n_, cks + (size_t) start * n_, (size_t) ni * n_ * 8,
                                      cudaMemcpyDeviceToDevice, st));
            // matrix offset: digits before d contributed digit_no * digit_size entries each
            size_t moff = 0;
            for (int e = 0; e < d; e++) moff += (size_t) lv.digit_no[e] * lv.digit_size[e];
            int kin = 0;
            for (int i = 0; i < ni; i++) kin = std::max(kin, 64 - __builtin_clzll(rowq_[start + i]));
            batch.job[jobs] = BconvJob{t_cks + (size_t) start * n_, dst, lv.modup_mat.p + moff, lv.modup_omod.p + off,
                                       lv.modup_olimb.p + off, ni, lv.digit_no[d], kin + ceil_log2(ni),
                                       lv.modup_matf.p + 2 * moff, lv.digit_big[d]};
            no_max = std::max(no_max, lv.digit_no[d]);
            jobs++, d++;
        }
        launch_bconv(batch, jobs, ni, no_max, d_mod_.p, d_bar_.p, RowArith{d_is_fp_.p, d_fpc_.p, fp_mask_[0], fp_mask_[1], mod_rows_ <= 128}, mod_rows_, n_, st);
    }
    // 3. forward NTT of the converted limbs only (..._exclude_range, rns_bconv.cu:618); BFV: of every limb (:622)
    {
        LimbVec v;
        if (!bfv) {
            v.data = lv.modup_ntt_data, v.row = lv.modup_ntt_row, v.src = lv.modup_ntt_data;
        } else {
            for (int d = 0; d < lv.beta; d++)
                for (int j = 0; j < lv.m; j++) v.push(d * lv.m + j, j < l ? j : size_Q_ + (j - l));
        }
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(t_mod_up, t_mod_up, ll, st); });
    }
}

void Engine::inner_prod(int l, u64 *cx, const u64 *t_mod_up, const u64 *const *evk, cudaStream_t st,
                        const u64 *own_c2, const TensorSrc *ts, const uint32_t *perm, bool accumulate, int j_begin,
                        int j_count, int persist_ctas, bool t_lazy) const {
    const Level &lv = level(l);
    if (j_count < 0) j_count = lv.m - j_begin;
    if (j_count == 0) return;
    OwnSrc os{nullptr, nullptr, nullptr, 0};
    if (own_c2) os = OwnSrc{own_c2, nullptr, nullptr, lv.alpha};
    else if (ts) os = OwnSrc{nullptr, ts->a + (size_t) ts->l * n_, ts->b + (size_t) ts->l * n_, lv.alpha};
    const InnerProdArgs A{cx, t_mod_up, evk, d_mod_.p, bar(lv.beta), bar(1, 0),
                          RowArith{d_is_fp_.p, d_fpc_.p, fp_mask_[0], fp_mask_[1], size_QP_ <= 128}, os, perm,
                          accumulate ? 1 : 0, t_lazy ? 1 : 0, n_, l, lv.m, size_Q_, size_QP_, lv.beta, j_begin, j_count};
    // few limbs (the P limbs on the critical path of the fused key switch): one coefficient pair per thread
    const bool thin = persist_ctas == 0 && lv.beta == 4 && !perm && !accumulate && (size_t) j_count * n_ / IP_TILE <= 296;
    const unsigned tiles = (unsigned) (n_ / (thin ? IP_TILE / IP_PAIRS : IP_TILE)) * (unsigned) j_count;
    const dim3 grid(persist_ctas > 0 ? std::min((unsigned) persist_ctas, tiles) : tiles);
    const bool plain = !perm && !accumulate;
    auto go = [&](auto kern) { launch_pdl(kern, grid, EW_THREADS, 0, st, A); };
    if (!plain) go(k_inner_prod<0, false>);
    else if (thin) go(k_inner_prod<4, true, 1>);
    else if (lv.beta == 1) go(k_inner_prod<1, true>);
    else if (lv.beta == 2) go(k_inner_prod<2, true>);
    else if (lv.beta == 3) go(k_inner_prod<3, true>);
    else if (lv.beta == 4) go(k_inner_prod<4, true>);
    else go(k_inner_prod<0, true>);
    check_launch("k_inner_prod");
}

// DRNSTool::moddown_from_NTT (reference src/rns_bconv.cu:776-828), CKKS form, for npoly polynomials laid
// out as cx[k] = [m][n]:  out[k][j] = (cx[k][j] - NTT(bconv_{P->q_j}(iNTT(cx[k][P])))) * P^-1  (+ addend[k][j])
void Engine::moddown(int l, u64 *out, u64 *cx, u64 *delta, int npoly, const u64 *addend, unsigned add_mask,
                     cudaStream_t st) const {
    const Level &lv = level(l);
    const int alpha = lv.alpha, m = lv.m;
    if (npoly < 1 || npoly > 2) throw std::invalid_argument("moddown handles 1 or 2 polynomials");
    // 1. inverse NTT of the P limbs, fused with n^-1 * phat_i^-1 (iNTT :788 + bconv_mult :40-60)
    {
        LimbVec v;
        for (int k = 0; k < npoly; k++)
            for (int i = 0; i < alpha; i++) v.push(k * m + l + i, size_Q_ + i);
        ntt_inv_list(cx, cx, single_list(v, rowq_), lv.moddown_fin.p, 1, st);
    }
    // 2. P -> Ql conversion (bConv_BEHZ matmul :143-168 / single-P :691-707)
    {
        int pbits = 0;
        for (int i = 0; i < alpha; i++) pbits = std::max(pbits, 64 - __builtin_clzll(rowq_[size_Q_ + i]));
        BconvBatch batch{};
        for (int k = 0; k < npoly; k++)
            batch.job[k] = BconvJob{cx + ((size_t) k * m + l) * n_, delta + (size_t) k * l * n_, lv.moddown_mat.p,
                                    lv.moddown_omod.p, lv.moddown_olimb.p, alpha, l, pbits + ceil_log2(alpha),
                                    lv.moddown_matf.p, lv.moddown_big};
        launch_bconv(batch, npoly, alpha, l, d_mod_.p, d_bar_.p, RowArith{d_is_fp_.p, d_fpc_.p, fp_mask_[0], fp_mask_[1], mod_rows_ <= 128}, mod_rows_, n_, st);
    }
    // 3. forward NTT of delta with the fused (cx - delta) * P^-1 (+ ct) epilogue (:820, ntt_moddown.cu:106-216)
    {
        LimbVec v;
        for (int k = 0; k < npoly; k++)
            for (int j = 0; j < l; j++) v.push(k * l + j, j);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t b) {
            EpiArgs ea{};
            ea.sub_base = cx, ea.out_base = out, ea.add_base = addend, ea.mulc = lv.pinv_slots.p + b;
            for (int s = 0; s < ll.count; s++) {
                const int k = (int) ((b + s) / l), j = (int) ((b + s) % l);
                ea.sub[s] = (short) (k * m + j);
                ea.out[s] = (short) (k * l + j);
                ea.add[s] = (short) ((addend && ((add_mask >> k) & 1)) ? k * l + j : -1);
            }
            g_launches.fetch_add(2, std::memory_order_relaxed);
            PFHE_CUDA(ntt_forward_epilogue(plan_, delta, ll, ea, st));
        });
    }
}

// Fused key switch (same result as keyswitch(): modup -> inner product -> moddown -> add):
//   inverse NTT (+ digit scaling, optionally of a1*b1)  ->  [bconv | forward NTT] of the converted limbs
//   ->  inner product (own-digit limbs straight from c2 / a1*b1)  ->  inverse NTT of the P limbs
//   ->  [bconv | forward NTT | (cx - delta) P^-1 + addend]          9 kernels, no copies, no t_cks->t_mod_up pass
void Engine::keyswitch_fused(int l, u64 *out, const u64 *c2, const TensorSrc *ts, const u64 *const *evk,
                             const u64 *addend, unsigned add_mask, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    const Level &lv = level(l);
    const int alpha = lv.alpha, m = lv.m;
    if (alpha == 0) throw std::logic_error("key switching needs special primes");
    const bool bgv = scheme_ == Scheme::bgv;
    if (bgv && (t_ <= 1 || lv.pinv_t.x == 0)) throw std::logic_error("invalid rns bases when computing pjInv_mod_t");
    if (alpha + (bgv ? 1 : 0) > FUSE_MAX_IN || m > NTT_MAX_LIMBS || 2 * l > NTT_MAX_LIMBS || lv.beta * m > 32767) {
        // shapes outside the fused kernels' limits take the modular path
        const u64 *src = c2;
        if (ts) {
            tensor_2x2(ts->a, ts->b, ws_.tmp.p, l, st);
            src = ws_.tmp.p + (size_t) 2 * l * n_;
            addend = ws_.tmp.p, add_mask = 3u;
        }
        modup(l, ws_.t_mod_up.p, src, ws_.t_cks.p, st);
        inner_prod(l, ws_.cx.p, ws_.t_mod_up.p, evk, st);
        if (scheme_ != Scheme::ckks) moddown_generic(l, out, ws_.cx.p, 2, addend, add_mask, st);
        else moddown(l, out, ws_.cx.p, ws_.delta.p, 2, addend, add_mask, st);
        return;
    }
    u64 *t_cks = ws_.t_cks.p, *t_mod_up = ws_.t_mod_up.p, *cx = ws_.cx.p, *delta = ws_.delta.p;
    // 1. inverse NTT fused with n^-1 * qhat_i^-1 (and with the a1*b1 product for HMult).  BFV: c2 is already in coefficient
    //    form -- only the qhat_i^-1 scaling (bconv_mult_kernel, rns_bconv.cu:598), and the digits' own limbs, which the inner
    //    product takes in NTT form, are transformed once into `delta` (free until the mod-down)
    const bool bfv = scheme_ == Scheme::bfv;
    if (bfv) {
        if (ts) throw std::logic_error("the BFV product is not a limb-wise tensor product");
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
        launch_pdl(k_scale_limbs, grid, EW_THREADS, 0, st, t_cks, c2, (const Tw *) lv.modup_fin_coeff.p, (const Modulus *) d_mod_.p, n_);
        check_launch("k_scale_limbs");
        LimbVec v;
        for (int i = 0; i < l; i++) v.push(i, i);
        ntt_fwd_list(delta, c2, single_list(v, rowq_), st);
    } else {
        LimbVec v;
        for (int i = 0; i < l; i++) v.push(i, i);
        const LimbList ll = single_list(v, rowq_);
        g_launches.fetch_add(2, std::memory_order_relaxed);
        if (ts) PFHE_CUDA(ntt_inverse_mul(plan_, t_cks, *ts, bar(1, 0), ll, lv.modup_fin.p, 1, st));
        else PFHE_CUDA(ntt_inverse(plan_, t_cks, c2, ll, lv.modup_fin.p, 1, st));
    }
    // 2. mod-up: convert + forward NTT, one launch pair per group of digits of equal size (and of at most NTT_MAX_LIMBS
    //    converted limbs: larger parameter sets, e.g. the 36..43-prime sets of benchmark/ckks_bench.cu, take several)
    for (int d0 = 0; d0 < lv.beta;) {
        const int ni = lv.digit_size[d0];
        const int per_digit = m - ni;
        int d1 = d0;
        while (d1 < lv.beta && lv.digit_size[d1] == ni && (d1 - d0 + 1) * per_digit <= NTT_MAX_LIMBS) d1++;
        if (d1 == d0) throw std::logic_error("digit wider than a launch");   // excluded by the shape test above
        LimbList ll{};
        BconvLoad bl{};
        size_t moff = 0;
        for (int e = 0; e < d0; e++) moff += (size_t) lv.digit_no[e] * lv.digit_size[e];
        bl.in_base = t_cks, bl.mat = lv.modup_mat.p + moff, bl.matf = lv.modup_matf.p + 2 * moff;
        bl.bar = d_bar_.p, bl.size_QP = mod_rows_, bl.ni = ni;
        int kin = 0, cnt = 0;
        // slots whose modulus takes the (slower) integer path are issued first: longest tiles start earliest
        for (int pass = 0; pass < 2; pass++)
            for (int d = d0; d < d1; d++) {
                const int start = lv.digit_start[d];
                for (int i = 0; i < ni; i++) kin = std::max(kin, 64 - __builtin_clzll(rowq_[start + i]));
                int jo = 0;
                for (int j = 0; j < m; j++) {
                    if (j >= start && j < start + ni) continue;
                    const int row = j < l ? j : size_Q_ + (j - l);
                    const bool slow = !is_fp_[row];
                    if (slow == (pass == 0)) {
                        ll.data[cnt] = ll.src[cnt] = (short) (d * m + j);
                        ll.row[cnt] = (short) row;
                        ll.q[cnt] = rowq_[row];
                        bl.in_limb[cnt] = (short) start;
                        bl.mat_row[cnt] = (short) (lv.digit_off[d] - lv.digit_off[d0] + jo);
                        bl.in_big[cnt] = (unsigned char) lv.digit_big[d];
                        cnt++;
                    }
                    jo++;
                }
            }
        ll.count = cnt;
        bl.xbits = kin + ceil_log2(ni);
        g_launches.fetch_add(2, std::memory_order_relaxed);
        PFHE_CUDA(ntt_forward_bconv(plan_, t_mod_up, ll, bl, nullptr, nullptr, nullptr, st, lazy_t_ ? 3 : 0));
        d0 = d1;
    }
    // 3. inner product; the digit's own limbs are read from c2 (or formed as a1*b1).  Only the P limbs of cx feed
    //    the mod-down chain (steps 4, 5a); the Q limbs are needed by the epilogue (5b) alone.  So: P limbs first, then
    //    fork -- the latency-bound chain 4 -> 5a goes to a high-priority side stream, the bandwidth-bound Q-limb half of
    //    the inner product stays on the caller's stream and fills the SMs the small launches leave idle -- and join
    //    before the epilogue.
    if (bfv) {   // coefficient-domain mod-down (moddown_from_NTT, BFV branch): every limb of cx goes back, no epilogue to fuse
        inner_prod(l, cx, t_mod_up, evk, st, delta, nullptr, nullptr, false, 0, -1, 0, lazy_t_);
        moddown_generic(l, out, cx, 2, addend, add_mask, st);
        return;
    }
    const bool fork = overlap_;
    cudaStream_t sc = st;   // stream of the P-limb chain
    cudaEvent_t ev_join = nullptr;
    inner_prod(l, cx, t_mod_up, evk, st, ts ? nullptr : c2, ts, nullptr, false, l, alpha, 0, lazy_t_);
    if (fork) {
        StreamCtx &fj = sctx(st);
        if (!fj.s_side) {
            int least = 0, greatest = 0;
            PFHE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            PFHE_CUDA(cudaStreamCreateWithPriority(&fj.s_side, cudaStreamNonBlocking, greatest));
            PFHE_CUDA(cudaEventCreateWithFlags(&fj.ev_fork, cudaEventDisableTiming));
            PFHE_CUDA(cudaEventCreateWithFlags(&fj.ev_join, cudaEventDisableTiming));
        }
        ev_join = fj.ev_join;
        PFHE_CUDA(cudaEventRecord(fj.ev_fork, st));
        PFHE_CUDA(cudaStreamWaitEvent(fj.s_side, fj.ev_fork, 0));
        sc = fj.s_side;
    } else {
        inner_prod(l, cx, t_mod_up, evk, st, ts ? nullptr : c2, ts, nullptr, false, 0, l, 0, lazy_t_);
    }
    // 4. inverse NTT of the P limbs fused with n^-1 * phat_i^-1.  BGV: out of place into a compact [2][alpha + 1][n] buffer
    //    whose last limb receives the plain-modulus correction c = [(sum_k y_k (phat_k mod t)) P^-1]_t
    const int nin = alpha + (bgv ? 1 : 0);
    u64 *pin = bgv ? t_cks : cx;   // where the conversion of step 5 finds its inputs
    {
        LimbVec v;
        for (int k = 0; k < 2; k++)
            for (int i = 0; i < alpha; i++) {
                if (bgv) v.push(k * nin + i, size_Q_ + i, k * m + l + i);
                else v.push(k * m + l + i, size_Q_ + i);
            }
        ntt_inv_list(pin, cx, single_list(v, rowq_), lv.moddown_fin.p, 1, sc);
        if (bgv) {
            launch_pdl(k_bgv_corr, dim3((unsigned) (n_ / EW_THREADS), 2), EW_THREADS, 0, sc, pin, (const u64 *) (lv.moddown_mat_t.p + (size_t) l * alpha),
                       alpha, lv.pinv_t, host_modulus(t_), n_);
            check_launch("k_bgv_corr");
        }
    }
    // 5. mod-down: convert P -> q_j, forward NTT, (cx - delta) * P^-1 + addend
    {
        LimbList ll{};
        BconvLoad bl{};
        EpiArgs ea{};
        int pbits = 0;
        for (int i = 0; i < alpha; i++) pbits = std::max(pbits, 64 - __builtin_clzll(rowq_[size_Q_ + i]));
        if (bgv) pbits = std::max(pbits, 64 - __builtin_clzll(t_));
        bl.in_base = pin, bl.bar = d_bar_.p;
        bl.mat = bgv ? lv.moddown_mat_bgv.p : lv.moddown_mat.p, bl.matf = bgv ? lv.moddown_matf_bgv.p : lv.moddown_matf.p;
        bl.size_QP = mod_rows_, bl.ni = nin, bl.xbits = pbits + ceil_log2(nin);
        ea.sub_base = cx, ea.out_base = out, ea.add_base = addend, ea.mulc = lv.pinv_slots.p;
        int cnt = 0;
        for (int k = 0; k < 2; k++)   // natural slot order: mulc (P^-1 per slot) is indexed by k * l + j
            for (int j = 0; j < l; j++) {
                ll.data[cnt] = ll.src[cnt] = (short) (k * l + j);
                ll.row[cnt] = (short) j;
                ll.q[cnt] = rowq_[j];
                bl.in_limb[cnt] = (short) (bgv ? k * nin : k * m + l);
                bl.mat_row[cnt] = (short) j;
                bl.in_big[cnt] = (unsigned char) (bgv ? lv.moddown_big_bgv : lv.moddown_big);
                ea.sub[cnt] = (short) (k * m + j);
                ea.out[cnt] = (short) (k * l + j);
                ea.add[cnt] = ts ? (short) k : (short) ((addend && ((add_mask >> k) & 1)) ? k * l + j : -1);
                cnt++;
            }
        ll.count = cnt;
        g_launches.fetch_add(2, std::memory_order_relaxed);
        if (!fork) {
            PFHE_CUDA(ntt_forward_bconv(plan_, delta, ll, bl, &ea, ts, bar(2, 0), st));
        } else {
            PFHE_CUDA(ntt_forward_bconv(plan_, delta, ll, bl, &ea, ts, bar(2, 0), sc, 1));   // 5a: column pass
            PFHE_CUDA(cudaEventRecord(ev_join, sc));
            inner_prod(l, cx, t_mod_up, evk, st, ts ? nullptr : c2, ts, nullptr, false, 0, l, side_ctas_, lazy_t_);   // Q limbs
            PFHE_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
            PFHE_CUDA(ntt_forward_bconv(plan_, delta, ll, bl, &ea, ts, bar(2, 0), st, 2));   // 5b: row pass + epilogue
        }
    }
}

// keyswitch_inplace (reference src/eval_key_switch.cu:95-182): out[2][l][n] = addend + moddown(<modup(c2), evk>)
void Engine::keyswitch(int l, u64 *out, const u64 *c2, const u64 *const *evk, const u64 *addend, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    // CKKS; BGV with the plain-modulus correction folded into the mod-down conversion; BFV with the fused mod-up and the
    // coefficient-domain mod-down
    (void) ws_;
    keyswitch_fused(l, out, c2, nullptr, evk, addend, addend ? 3u : 0u, st);
}

void Engine::moddown_generic(int l, u64 *out, u64 *cx, int npoly, const u64 *addend, unsigned add_mask,
                             cudaStream_t st) {
    Workspace &ws_ = ws(st);
    const Level &lv = level(l);
    const int alpha = lv.alpha, m = lv.m;
    const bool bgv = scheme_ == Scheme::bgv;
    if (bgv && (t_ <= 1 || lv.pinv_t.x == 0)) throw std::logic_error("invalid rns bases when computing pjInv_mod_t");
    // 1. every limb of cx back to coefficient form (rns_bconv.cu:790-794), P limbs also scaled by phat_i^-1
    {
        LimbVec v;
        for (int k = 0; k < npoly; k++)
            for (int j = 0; j < m; j++) v.push(k * m + j, j < l ? j : size_Q_ + (j - l));
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t b) { ntt_inv_list(cx, cx, ll, lv.moddown_fin_all.p + 2 * b, 1, st); });
    }
    static const bool one_pass = [] {
        const char *e = std::getenv("PFHE_BFV_MODDOWN_FUSED");
        return !(e && e[0] == '0');
    }();
    if (!bgv && one_pass) {
        // BFV: conversion, (cx - delta) * P^-1 and the addition in one pass; the result stays in the coefficient domain
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l, npoly);
        launch_pdl(k_moddown_coeff_conv, grid, EW_THREADS, 0, st, out, (const u64 *) cx, (const u64 *) lv.moddown_mat.p, alpha, m,
                   (const Tw *) lv.pinv_slots.p, addend, add_mask, (const Modulus *) d_mod_.p, n_);
        check_launch("k_moddown_coeff_conv");
        return;
    }
    // 2. P -> Ql (and, BGV, P -> t) conversion
    u64 *delta = ws_.delta.p;
    const int no = bgv ? l + 1 : l;
    {
        int pbits = 0;
        for (int i = 0; i < alpha; i++) pbits = std::max(pbits, 64 - __builtin_clzll(rowq_[size_Q_ + i]));
        BconvBatch batch{};
        for (int k = 0; k < npoly; k++) {
            if (bgv)
                batch.job[k] = BconvJob{cx + ((size_t) k * m + l) * n_, delta + (size_t) k * (l + 1) * n_,
                                        lv.moddown_mat_t.p, lv.moddown_omod_t.p, lv.moddown_olimb_t.p, alpha, no,
                                        pbits + ceil_log2(alpha), lv.moddown_matf_t.p, lv.moddown_big};
            else
                batch.job[k] = BconvJob{cx + ((size_t) k * m + l) * n_, delta + (size_t) k * (l + 1) * n_,
                                        lv.moddown_mat.p, lv.moddown_omod.p, lv.moddown_olimb.p, alpha, no,
                                        pbits + ceil_log2(alpha), lv.moddown_matf.p, lv.moddown_big};
        }
        launch_bconv(batch, npoly, alpha, no, d_mod_.p, d_bar_.p,
                     RowArith{d_is_fp_.p, d_fpc_.p, fp_mask_[0], fp_mask_[1], mod_rows_ <= 128}, mod_rows_, n_, st);
    }
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
    for (int k = 0; k < npoly; k++) {
        const u64 *add = (addend && ((add_mask >> k) & 1)) ? addend + (size_t) k * l * n_ : nullptr;
        u64 *dk = delta + (size_t) k * (l + 1) * n_;
        u64 *ok = out + (size_t) k * l * n_;
        if (!bgv) {
            // BFV: result stays in the coefficient domain (moddown_kernel :680-689 + add_to_ct :763-769)
            launch_pdl(k_moddown_coeff, grid, EW_THREADS, 0, st, ok, cx + (size_t) k * m * n_, dk, lv.pinv_slots.p, add,
                       d_mod_.p, n_);
            check_launch("k_moddown_coeff");
        } else {
            // BGV: plain-modulus correction, back to NTT form, then add (bgv_moddown_kernel :636-652, :810-817)
            u64 *tmp = ws_.t_cks.p;   // [l][n]
            launch_pdl(k_bgv_moddown, grid, EW_THREADS, 0, st, tmp, cx + (size_t) k * m * n_, dk, dk + (size_t) l * n_,
                       lv.P_mod_q.p, lv.pinv_slots.p, lv.pinv_t, t_, d_mod_.p, n_);
            check_launch("k_bgv_moddown");
            ntt_fwd_rows_range(tmp, l, 0, st);
            if (add) elementwise(EW_ADD, tmp, add, ok, l, st);
            else PFHE_CUDA(cudaMemcpyAsync(ok, tmp, (size_t) l * n_ * 8, cudaMemcpyDeviceToDevice, st));
        }
    }
}

void Engine::divide_round_q_last(int l, u64 *out, const u64 *in, int size, int mode, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (mode != 1 && mode != 2) throw std::invalid_argument("unsupported scheme");
    if (l < 2) throw std::invalid_argument("end of modulus switching chain reached");
    const Level &lv = level(l);
    const int nl = l - 1;
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), nl);
    for (int s = 0; s < size; s++) {
        const u64 *ci = in + (size_t) s * l * n_;
        u64 *co = out + (size_t) s * nl * n_;
        if (mode == 1) {
            launch_pdl(k_divide_round_last, grid, EW_THREADS, 0, st, co, ci, lv.qlast_inv_slots.p, d_mod_.p, n_, nl);
            check_launch("k_divide_round_last");
        } else {   // BGV
            if (t_ <= 1 || lv.inv_qlast_t.x == 0) throw std::logic_error("invalid rns bases");
            u64 *tmp = ws_.tmp.p;   // coefficient form of all l limbs
            ntt_inv_rows_range(tmp, ci, l, 0, st);
            launch_pdl(k_bgv_mod_t_divide, grid, EW_THREADS, 0, st, co, tmp, tmp + (size_t) nl * n_, lv.qlast_mod_q.p,
                       lv.qlast_inv_slots.p, lv.inv_qlast_t, Modulus{t_, 0, hm::barrett_ratio(t_).hi}, d_mod_.p, n_);
            check_launch("k_bgv_mod_t_divide");
            ntt_fwd_rows_range(co, nl, 0, st);
        }
    }
}

void Engine::galois_coeff(u64 *dst, const u64 *src, uint32_t elt, int l, int npoly, cudaStream_t st) const {
    dim3 grid((unsigned) (n_ / EW_THREADS), npoly * l);
    launch_pdl(k_galois_coeff, grid, EW_THREADS, 0, st, dst, src, d_mod_.p, elt, n_, l);
    check_launch("k_galois_coeff");
}

// ---------------------------------------------------------------------------------------------------
// BFV multiplication, BEHZ variant
// ---------------------------------------------------------------------------------------------------
const Behz &Engine::behz(int l) {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if (scheme_ != Scheme::bfv || naux_ == 0) throw std::invalid_argument("unsupported scheme");
    if (l < 1 || l > size_Q_) throw std::invalid_argument("index is invalid!");
    if (behz_[l]) return *behz_[l];
    auto b = std::make_unique<Behz>();
    const std::vector<u64> Q(primes_.begin(), primes_.begin() + l);
    const int tb = 64 - __builtin_clzll(t_);
    const int nB = l + ((32 + tb + hm::product_bits(Q) >= 61 * l + 61) ? 1 : 0);   // rns.cu:400-406
    const int nbsk = nB + 1;
    if (nbsk > naux_ || nbsk > BEHZ_MAX_LIMBS || l > BEHZ_MAX_LIMBS) throw std::logic_error("BEHZ base too large");
    b->l = l, b->nB = nB, b->nbsk = nbsk;
    const u64 msk = rowq_[row_aux_];
    const std::vector<u64> B(rowq_.begin() + row_aux_ + 1, rowq_.begin() + row_aux_ + 1 + nB);
    const u64 mt = (u64) 1 << 32;
    auto bsk = [&](int j) { return rowq_[row_aux_ + j]; };   // [m_sk, B_0, ...]

    std::vector<Tw> q_mt_hinv(l), q_hinv(l);
    std::vector<u32> q_to_mt(l);
    for (int i = 0; i < l; i++) {
        const u64 q = Q[i], hinv = hm::invmod(hm::product_mod(Q, i, q), q);
        q_hinv[i] = make_tw(hinv, q);
        q_mt_hinv[i] = make_tw(hm::mulmod(mt % q, hinv, q), q);
        q_to_mt[i] = (u32) hm::product_mod(Q, i, mt);
    }
    std::vector<u64> q_to_bsk((size_t) nbsk * l), q_mod_bsk(nbsk);
    std::vector<Tw> inv_mt(nbsk), inv_q(nbsk);
    for (int j = 0; j < nbsk; j++) {
        const u64 p = bsk(j);
        for (int i = 0; i < l; i++) q_to_bsk[(size_t) j * l + i] = hm::product_mod(Q, i, p);
        q_mod_bsk[j] = hm::product_mod(Q, -1, p);
        inv_mt[j] = make_tw(hm::invmod(mt % p, p), p);
        inv_q[j] = make_tw(hm::invmod(q_mod_bsk[j], p), p);
    }
    b->neg_inv_q_mt = (u32) ((mt - hm::invmod(hm::product_mod(Q, -1, mt), mt)) % mt);
    std::vector<Tw> b_hinv(nB);
    std::vector<u64> b_to_q((size_t) l * nB), b_to_msk(nB), B_mod_q(l);
    for (int i = 0; i < nB; i++) {
        b_hinv[i] = make_tw(hm::invmod(hm::product_mod(B, i, B[i]), B[i]), B[i]);
        b_to_msk[i] = hm::product_mod(B, i, msk);
        for (int k = 0; k < l; k++) b_to_q[(size_t) k * nB + i] = hm::product_mod(B, i, Q[k]);
    }
    for (int k = 0; k < l; k++) B_mod_q[k] = hm::product_mod(B, -1, Q[k]);
    b->inv_B_msk = make_tw(hm::invmod(hm::product_mod(B, -1, msk), msk), msk);
    // last inverse-NTT stage constants with the multiplication by t folded in (evaluate.cu:520-531)
    std::vector<Tw> fin((size_t) 3 * (l + nbsk) * 2);
    for (int p = 0; p < 3; p++) {
        for (int j = 0; j < l + nbsk; j++) {
            const int slot = j < l ? p * l + j : 3 * l + p * nbsk + (j - l);
            const int row = j < l ? j : row_aux_ + (j - l);
            const u64 q = rowq_[row], c = hm::mulmod(h_ninv_[row], t_ % q, q);
            fin[2 * slot] = make_tw_row(row, c);
            fin[2 * slot + 1] = make_tw_row(row, hm::mulmod(c, h_itw1_[row], q));
        }
    }
    b->q_mt_hinv.upload(q_mt_hinv), b->q_hinv.upload(q_hinv), b->q_to_bsk.upload(q_to_bsk), b->q_to_mt.upload(q_to_mt);
    b->q_mod_bsk.upload(q_mod_bsk), b->inv_mt_bsk.upload(inv_mt), b->inv_q_bsk.upload(inv_q);
    b->b_hinv.upload(b_hinv), b->b_to_q.upload(b_to_q), b->b_to_msk.upload(b_to_msk), b->B_mod_q.upload(B_mod_q);
    b->fin_t.upload(fin);
    behz_[l] = std::move(b);
    return *behz_[l];
}

void Engine::bfv_multiply_behz(int l, u64 *out3, const u64 *ct1, const u64 *ct2, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    const Behz &b = behz(l);
    const int nbsk = b.nbsk;
    const size_t need = (size_t) 7 * (l + nbsk) * n_;
    if (ws_.behz.count < need) ws_.behz.alloc((size_t) 7 * (size_Q_ + naux_) * n_);
    // workspace: operands in q [4][l], tensor result in q [3][l] and in Bsk [3][nbsk] (adjacent: one inverse-NTT
    // list), operands in Bsk [4][nbsk]
    u64 *eq = ws_.behz.p, *dq = eq + (size_t) 4 * l * n_, *db = dq + (size_t) 3 * l * n_,
        *eb = db + (size_t) 3 * nbsk * n_;
    // 1. operands to NTT form over q (BEHZ_mul_1, evaluate.cu:404-438)
    {
        LimbVec v;
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < l; i++) v.push(p * l + i, i);
        for (int s = 0; s < 2; s++) {
            u64 *dst = eq + (size_t) s * 2 * l * n_;
            const u64 *src = s ? ct2 : ct1;
            run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(dst, src, ll, st); });
        }
    }
    // 2. q -> Bsk with the m_tilde correction (fastbconv_m_tilde + sm_mrq), then NTT over Bsk
    {
        BehzLiftArgs a{ct1, ct2, eb, b.q_mt_hinv.p, b.q_to_bsk.p, b.q_to_mt.p, b.q_mod_bsk.p, b.inv_mt_bsk.p,
                       d_mod_.p, d_mod_.p + row_aux_, b.neg_inv_q_mt, l, nbsk, n_};
        launch_pdl(k_behz_lift, dim3((unsigned) (n_ / BEHZ_THREADS), 4), BEHZ_THREADS, 0, st, a);
        check_launch("k_behz_lift");
        LimbVec v;
        for (int p = 0; p < 4; p++)
            for (int j = 0; j < nbsk; j++) v.push(p * nbsk + j, row_aux_ + j);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(eb, eb, ll, st); });
    }
    // 3. dyadic tensor products in both bases (evaluate.cu:479-500)
    tensor_2x2(eq, eq + (size_t) 2 * l * n_, dq, l, st);
    {
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), nbsk);
        launch_pdl(k_tensor_2x2, grid, EW_THREADS, 0, st, (const u64 *) eb, (const u64 *) (eb + (size_t) 2 * nbsk * n_), db,
                   (const Modulus *) (d_mod_.p + row_aux_), bar(1, 2) + row_aux_,
                   RowArith{d_is_fp_.p + row_aux_, d_fpc_.p + row_aux_, 0, 0, 0}, n_, nbsk);
        check_launch("k_tensor_2x2");
    }
    // 4. back to coefficient form, times t (folded into the last inverse stage)
    {
        LimbVec v;
        for (int p = 0; p < 3; p++)
            for (int i = 0; i < l; i++) v.push(p * l + i, i);
        for (int p = 0; p < 3; p++)
            for (int j = 0; j < nbsk; j++) v.push(3 * l + p * nbsk + j, row_aux_ + j);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t bgn) { ntt_inv_list(dq, dq, ll, b.fin_t.p + 2 * bgn, 1, st); });
    }
    // 5. floor(t * x / Q) in Bsk and Shenoy-Kumaresan conversion back to q (fast_floor + fastbconv_sk)
    {
        BehzFloorArgs a{dq, db, out3, b.q_hinv.p, b.q_to_bsk.p, b.inv_q_bsk.p, b.b_hinv.p, b.b_to_q.p, b.b_to_msk.p,
                        b.B_mod_q.p, b.inv_B_msk, d_mod_.p, d_mod_.p + row_aux_, l, nbsk, n_};
        launch_pdl(k_behz_floor_sk, dim3((unsigned) (n_ / BEHZ_THREADS), 3), BEHZ_THREADS, 0, st, a);
        check_launch("k_behz_floor_sk");
    }
}

// ---------------------------------------------------------------------------------------------------
// BFV multiplication, HPS variant
// ---------------------------------------------------------------------------------------------------
void Engine::set_mul_tech(int m) {
    if (scheme_ != Scheme::bfv) throw std::invalid_argument("mul_tech selection is only supported for BFV");
    if (m < 1 || m > 4) throw std::invalid_argument("unsupported multiplication technique for BFV");
    mul_tech_ = m;
}

const Hps &Engine::hps() {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if (scheme_ != Scheme::bfv || nR_ == 0) throw std::invalid_argument("unsupported scheme");
    if (hps_) return *hps_;
    auto h = std::make_unique<Hps>();
    const int l = size_Q_, nR = nR_;
    h->l = l, h->nR = nR;
    const std::vector<u64> Q(primes_.begin(), primes_.begin() + l);
    const std::vector<u64> R(rowq_.begin() + row_R_, rowq_.begin() + row_R_ + nR);
    std::vector<u64> S(Q);
    S.insert(S.end(), R.begin(), R.end());
    std::vector<Tw> q_hinv(l), r_hinv(nR);
    std::vector<double> q_inv(l), r_inv(nR), sr_frac(l);
    std::vector<u64> q_to_r((size_t) nR * l), Q_mod_r(nR), r_to_q((size_t) l * nR), R_mod_q(l), sr_tab((size_t) nR * (l + 1));
    for (int i = 0; i < l; i++) {
        q_hinv[i] = make_tw(hm::invmod(hm::product_mod(Q, i, Q[i]), Q[i]), Q[i]);
        q_inv[i] = 1.0 / (double) Q[i];   // host/rns.cu:319-324
        R_mod_q[i] = hm::product_mod(R, -1, Q[i]);
        for (int j = 0; j < nR; j++) r_to_q[(size_t) i * nR + j] = hm::product_mod(R, j, Q[i]);
    }
    for (int j = 0; j < nR; j++) {
        r_hinv[j] = make_tw(hm::invmod(hm::product_mod(R, j, R[j]), R[j]), R[j]);
        r_inv[j] = 1.0 / (double) R[j];
        Q_mod_r[j] = hm::product_mod(Q, -1, R[j]);
        for (int i = 0; i < l; i++) q_to_r[(size_t) j * l + i] = hm::product_mod(Q, i, R[j]);
    }
    // scale-and-round tables (rns.cu:727-789): A_i = t * R * (Shat_i^-1 mod s_i) over S = Q u R
    for (int i = 0; i < l + nR; i++) {
        hm::BigUint A;
        for (u64 r : R) A.mul_word(r);
        A.mul_word(t_);
        A.mul_word(hm::invmod(hm::product_mod(S, i, S[i]), S[i]));
        const u64 rem = A.divmod_word(S[i]);
        if (i < l) {
            sr_frac[i] = (double) rem / (double) S[i];
            for (int j = 0; j < nR; j++) sr_tab[(size_t) j * (l + 1) + i] = A.mod_word(R[j]);
        } else {
            sr_tab[(size_t) (i - l) * (l + 1) + l] = A.mod_word(R[i - l]);
        }
    }
    h->q_hinv.upload(q_hinv), h->q_inv.upload(q_inv), h->q_to_r.upload(q_to_r), h->Q_mod_r.upload(Q_mod_r);
    h->sr_frac.upload(sr_frac), h->sr_tab.upload(sr_tab);
    h->r_hinv.upload(r_hinv), h->r_inv.upload(r_inv), h->r_to_q.upload(r_to_q), h->R_mod_q.upload(R_mod_q);
    hps_ = std::move(h);
    return *hps_;
}

void Engine::bfv_multiply_hps(int l, u64 *out3, const u64 *ct1, const u64 *ct2, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    // the reference always takes the constants of the first data level here (evaluate.cu:672-696)
    if (l != size_Q_) throw std::invalid_argument("HPS multiplication is defined at the first data level only");
    const Hps &h = hps();
    const int nR = h.nR;
    const size_t need = (size_t) 7 * (l + nR) * n_;
    if (ws_.behz.count < need) ws_.behz.alloc(std::max(need, (size_t) 7 * (size_Q_ + naux_) * n_));
    // operands over Q [4][l], tensor result over Q [3][l] and over R [3][nR] (adjacent), operands over R [4][nR]
    u64 *eq = ws_.behz.p, *dq = eq + (size_t) 4 * l * n_, *dr = dq + (size_t) 3 * l * n_,
        *er = dr + (size_t) 3 * nR * n_;
    {   // Q limbs: NTT straight from the ciphertexts (the reference's D2D copy is the out-of-place store)
        LimbVec v;
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < l; i++) v.push(p * l + i, i);
        for (int s = 0; s < 2; s++) {
            u64 *dst = eq + (size_t) s * 2 * l * n_;
            const u64 *src = s ? ct2 : ct1;
            run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(dst, src, ll, st); });
        }
    }
    {   // Q -> R (bConv_HPS), then NTT over R
        HpsLiftArgs a{ct1, ct2, er, h.q_hinv.p, h.q_inv.p, h.q_to_r.p, h.Q_mod_r.p, d_mod_.p, d_mod_.p + row_R_, l, nR, n_};
        launch_pdl(k_hps_lift, dim3((unsigned) (n_ / BEHZ_THREADS), 4), BEHZ_THREADS, 0, st, a);
        check_launch("k_hps_lift");
        LimbVec v;
        for (int p = 0; p < 4; p++)
            for (int j = 0; j < nR; j++) v.push(p * nR + j, row_R_ + j);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(er, er, ll, st); });
    }
    tensor_2x2(eq, eq + (size_t) 2 * l * n_, dq, l, st);
    {
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), nR);
        launch_pdl(k_tensor_2x2, grid, EW_THREADS, 0, st, (const u64 *) er, (const u64 *) (er + (size_t) 2 * nR * n_), dr,
                   (const Modulus *) (d_mod_.p + row_R_), bar(1, 2) + row_R_,
                   RowArith{d_is_fp_.p + row_R_, d_fpc_.p + row_R_, 0, 0, 0}, n_, nR);
        check_launch("k_tensor_2x2");
    }
    {
        LimbVec v;
        for (int p = 0; p < 3; p++)
            for (int i = 0; i < l; i++) v.push(p * l + i, i);
        for (int p = 0; p < 3; p++)
            for (int j = 0; j < nR; j++) v.push(3 * l + p * nR + j, row_R_ + j);
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_inv_list(dq, dq, ll, nullptr, 0, st); });
    }
    {   // t/Q scale-and-round QR -> R, then R -> Q (scaleAndRound_HPS_QR_R + bConv_HPS)
        HpsScaleArgs a{dq, dr, out3, h.sr_frac.p, h.sr_tab.p, h.r_hinv.p, h.r_inv.p, h.r_to_q.p, h.R_mod_q.p,
                       d_mod_.p, d_mod_.p + row_R_, l, nR, n_};
        launch_pdl(k_hps_scale_round, dim3((unsigned) (n_ / BEHZ_THREADS), 3), BEHZ_THREADS, 0, st, a);
        check_launch("k_hps_scale_round");
    }
}

// constants of HPS over Q with `drop` dropped levels (reference rns.cu:794-975; host/rns.cu:470-495 for the var1 form)
const HpsQ &Engine::hpsq(int drop) {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if (scheme_ != Scheme::bfv || nR_ == 0) throw std::invalid_argument("unsupported scheme");
    if (drop < 0 || drop >= size_Q_) throw std::invalid_argument("levels dropped out of range");
    if ((int) hpsq_.size() <= drop) hpsq_.resize(drop + 1);
    if (hpsq_[drop]) return *hpsq_[drop];
    auto h = std::make_unique<HpsQ>();
    const int ll = size_Q_ - drop;
    h->ll = ll, h->drop = drop;
    const std::vector<u64> Qall(primes_.begin(), primes_.begin() + size_Q_);
    const std::vector<u64> Ql(primes_.begin(), primes_.begin() + ll);
    const std::vector<u64> Qd(primes_.begin() + ll, primes_.begin() + size_Q_);
    const std::vector<u64> R(rowq_.begin() + row_R_, rowq_.begin() + row_R_ + ll);   // Rl: as many primes as Ql
    // bConv_HPS Ql -> Rl and Rl -> Ql
    std::vector<Tw> q_hinv(ll), r_hinv(ll);
    std::vector<double> q_inv(ll), r_inv(ll);
    std::vector<u64> q_to_r((size_t) ll * ll), Ql_mod_r(ll), r_to_q((size_t) ll * ll), Rl_mod_q(ll);
    for (int i = 0; i < ll; i++) {
        q_hinv[i] = make_tw(hm::invmod(hm::product_mod(Ql, i, Ql[i]), Ql[i]), Ql[i]);
        q_inv[i] = 1.0 / (double) Ql[i];
        r_hinv[i] = make_tw(hm::invmod(hm::product_mod(R, i, R[i]), R[i]), R[i]);
        r_inv[i] = 1.0 / (double) R[i];
        Ql_mod_r[i] = hm::product_mod(Ql, -1, R[i]);
        Rl_mod_q[i] = hm::product_mod(R, -1, Ql[i]);
        for (int j = 0; j < ll; j++) {
            q_to_r[(size_t) i * ll + j] = hm::product_mod(Ql, j, R[i]);   // row = output r_i, column = input q_j
            r_to_q[(size_t) i * ll + j] = hm::product_mod(R, j, Ql[i]);
        }
    }
    // bConv_BEHZ_var1 from Q (levels dropped) or Ql to Rl
    const std::vector<u64> &I = drop ? Qall : Ql;
    const int ni = (int) I.size();
    std::vector<Tw> v1_c(ni);
    std::vector<u64> v1_mat((size_t) ll * ni);
    for (int i = 0; i < ni; i++) {
        const u64 qi = I[i];
        const u64 pq = hm::mulmod(hm::product_mod(R, -1, qi), hm::invmod(hm::product_mod(I, i, qi), qi), qi);
        v1_c[i] = make_ulonglong2(qi - pq, 0);   // may equal q_i (host/rns.cu:481-482): used through the general product
        for (int j = 0; j < ll; j++) v1_mat[(size_t) j * ni + i] = hm::invmod(qi % R[j], R[j]);
    }
    // scale-and-round tables: W_i = mult * prod(A) * (Shat_i^-1 mod s_i) over S = A u B
    auto tables = [&](const std::vector<u64> &A, const std::vector<u64> &B, u64 mult, std::vector<u64> &tab,
                      std::vector<double> &frac) {
        const int na = (int) A.size(), nb = (int) B.size();
        std::vector<u64> S(A);
        S.insert(S.end(), B.begin(), B.end());
        tab.assign((size_t) na * (nb + 1), 0);
        frac.assign(nb, 0.0);
        for (int i = 0; i < na + nb; i++) {
            hm::BigUint W;
            for (u64 a : A) W.mul_word(a);
            W.mul_word(mult);
            W.mul_word(hm::invmod(hm::product_mod(S, i, S[i]), S[i]));
            const u64 rem = W.divmod_word(S[i]);
            if (i >= na) {
                frac[i - na] = (double) rem / (double) S[i];
                for (int a = 0; a < na; a++) tab[(size_t) a * (nb + 1) + (i - na)] = W.mod_word(A[a]);
            } else {
                tab[(size_t) i * (nb + 1) + nb] = W.mod_word(A[i]);
            }
        }
    };
    std::vector<u64> sr_tab, dr_tab;
    std::vector<double> sr_frac, dr_frac;
    tables(Ql, R, t_, sr_tab, sr_frac);
    h->q_hinv.upload(q_hinv), h->q_inv.upload(q_inv), h->q_to_r.upload(q_to_r), h->Ql_mod_r.upload(Ql_mod_r);
    h->v1_c.upload(v1_c), h->v1_mat.upload(v1_mat);
    h->r_hinv.upload(r_hinv), h->r_inv.upload(r_inv), h->r_to_q.upload(r_to_q), h->Rl_mod_q.upload(Rl_mod_q);
    h->sr_tab.upload(sr_tab), h->sr_frac.upload(sr_frac);
    if (drop) {
        tables(Ql, Qd, 1, dr_tab, dr_frac);
        std::vector<Tw> expand(ll);
        for (int i = 0; i < ll; i++) expand[i] = make_tw(hm::product_mod(Qd, -1, Ql[i]), Ql[i]);
        h->dr_tab.upload(dr_tab), h->dr_frac.upload(dr_frac), h->expand.upload(expand);
    }
    hpsq_[drop] = std::move(h);
    return *hpsq_[drop];
}

// bfv_multiply_hps with mul_tech hps_overq / hps_overq_leveled (evaluate.cu:647-801): ct1 keeps its Ql residues (scaled
// down from Q when levels are dropped) and is lifted to Rl exactly; ct2 goes to Rl through the var1 conversion and
// comes back to Ql from there; tensor product over Ql u Rl; t/Rl scale-and-round straight to Ql; expansion back to Q.
void Engine::bfv_multiply_hps_overq(int l, u64 *out3, const u64 *ct1, const u64 *ct2, int drop, cudaStream_t st,
                                    bool keep_c2_low) {
    Workspace &ws_ = ws(st);
    if (l != size_Q_) throw std::invalid_argument("HPS multiplication is defined at the first data level only");
    const HpsQ &h = hpsq(drop);
    const int lq = size_Q_, ll = h.ll;
    if (2 * ll > BEHZ_MAX_LIMBS || lq > BEHZ_MAX_LIMBS) throw std::invalid_argument("too many limbs for the HPS kernels");
    const size_t need = (size_t) 14 * ll * n_;
    if (ws_.behz.count < need) ws_.behz.alloc(std::max(need, (size_t) 7 * (size_Q_ + naux_) * n_));
    const size_t pl = (size_t) ll * n_, pq = (size_t) lq * n_;
    // operands over Ql [4][ll], tensor result over Ql [3][ll] and over Rl [3][ll] (adjacent), operands over Rl [4][ll]
    u64 *eq = ws_.behz.p, *dq = eq + 4 * pl, *dr = dq + 3 * pl, *er = dr + 3 * pl;
    const Modulus *mod_q = d_mod_.p, *mod_r = d_mod_.p + row_R_;
    const dim3 g2((unsigned) (n_ / BEHZ_THREADS), 2), g3((unsigned) (n_ / BEHZ_THREADS), 3);
    u64 *c1 = const_cast<u64 *>(ct1), *c2 = const_cast<u64 *>(ct2);
    // ct1: Ql part (scaleAndRound_HPS_Q_Ql when levels were dropped), then the exact lift Ql -> Rl
    PolyView ct1_ql{c1, pq};
    if (drop) {
        ScaleRoundArgs a{PolyView{c1, pq}, PolyView{c1 + pl, pq}, PolyView{eq, pl}, h.dr_tab.p, h.dr_frac.p, nullptr, mod_q,
                         ll, drop, 0, n_};
        launch_pdl(k_scale_round, g2, BEHZ_THREADS, 0, st, a);
        check_launch("k_scale_round");
        ct1_ql = PolyView{eq, pl};
    }
    {
        BconvHpsArgs a{ct1_ql, PolyView{er, pl}, h.q_hinv.p, h.q_inv.p, h.q_to_r.p, h.Ql_mod_r.p, mod_q, mod_r, ll, ll, n_};
        launch_pdl(k_bconv_hps, g2, BEHZ_THREADS, 0, st, a);
        check_launch("k_bconv_hps");
    }
    // ct2: Q (or Ql) -> Rl by bConv_BEHZ_var1, then Rl -> Ql by bConv_HPS
    {
        BconvVar1Args a{PolyView{c2, pq}, PolyView{er + 2 * pl, pl}, h.v1_c.p, h.v1_mat.p, mod_q, mod_r, drop ? lq : ll, ll, n_};
        launch_pdl(k_bconv_var1, g2, BEHZ_THREADS, 0, st, a);
        check_launch("k_bconv_var1");
        BconvHpsArgs b{PolyView{er + 2 * pl, pl}, PolyView{eq + 2 * pl, pl}, h.r_hinv.p, h.r_inv.p, h.r_to_q.p, h.Rl_mod_q.p,
                       mod_r, mod_q, ll, ll, n_};
        launch_pdl(k_bconv_hps, g2, BEHZ_THREADS, 0, st, b);
        check_launch("k_bconv_hps");
    }
    {   // forward NTTs: ct1's Ql limbs come straight from the ciphertext when nothing was dropped
        LimbVec v1, v2, vr;
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < ll; i++) {
                v1.push(p * ll + i, i, drop ? p * ll + i : p * lq + i);
                v2.push(2 * ll + p * ll + i, i);
            }
        for (int p = 0; p < 4; p++)
            for (int j = 0; j < ll; j++) vr.push(p * ll + j, row_R_ + j);
        run_chunks(v1, rowq_, [&](const LimbList &lst, size_t) { ntt_fwd_list(eq, drop ? eq : ct1, lst, st); });
        run_chunks(v2, rowq_, [&](const LimbList &lst, size_t) { ntt_fwd_list(eq, eq, lst, st); });
        run_chunks(vr, rowq_, [&](const LimbList &lst, size_t) { ntt_fwd_list(er, er, lst, st); });
    }
    tensor_2x2(eq, eq + 2 * pl, dq, ll, st);
    {
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), ll);
        launch_pdl(k_tensor_2x2, grid, EW_THREADS, 0, st, (const u64 *) er, (const u64 *) (er + 2 * pl), dr, mod_r,
                   bar(1, 2) + row_R_, RowArith{d_is_fp_.p + row_R_, d_fpc_.p + row_R_, 0, 0, 0}, n_, ll);
        check_launch("k_tensor_2x2");
    }
    {
        LimbVec v;
        for (int p = 0; p < 3; p++)
            for (int i = 0; i < ll; i++) v.push(p * ll + i, i);
        for (int p = 0; p < 3; p++)
            for (int j = 0; j < ll; j++) v.push(3 * ll + p * ll + j, row_R_ + j);
        run_chunks(v, rowq_, [&](const LimbList &lst, size_t) { ntt_inv_list(dq, dq, lst, nullptr, 0, st); });
    }
    {   // scaleAndRound_HPS_QlRl_Ql (+ ExpandCRTBasis_Ql_Q when levels were dropped)
        ScaleRoundArgs a{PolyView{dq, pl}, PolyView{dr, pl}, PolyView{out3, pq}, h.sr_tab.p, h.sr_frac.p,
                         drop ? h.expand.p : nullptr, mod_q, ll, ll, drop, n_};
        if (!(drop && keep_c2_low)) {
            launch_pdl(k_scale_round, g3, BEHZ_THREADS, 0, st, a);
        } else {   // bfv_mul_relin_hps (evaluate.cu:954-956): only c0 and c1 are expanded
            launch_pdl(k_scale_round, g2, BEHZ_THREADS, 0, st, a);
            ScaleRoundArgs b = a;
            b.xa.base += 2 * pl, b.xb.base += 2 * pl, b.out.base += 2 * pq, b.expand = nullptr, b.zero = 0;
            launch_pdl(k_scale_round, dim3(g2.x, 1), BEHZ_THREADS, 0, st, b);
        }
        check_launch("k_scale_round");
    }
}

void Engine::keyswitch_leveled(u64 *ct, const u64 *c2, const u64 *const *evk, int drop, bool c2_low, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (scheme_ != Scheme::bfv) throw std::invalid_argument("unsupported scheme");
    const int lq = size_Q_;
    if (drop == 0) {
        keyswitch(lq, ct, c2, evk, ct, st);
        return;
    }
    const HpsQ &h = hpsq(drop);
    const int ll = h.ll;
    const size_t pl = (size_t) ll * n_, pq = (size_t) lq * n_;
    if (ws_.behz.count < 3 * pl) ws_.behz.alloc(std::max(3 * pl, (size_t) 7 * (size_Q_ + naux_) * n_));
    u64 *ks = ws_.behz.p, *low = ks + 2 * pl;   // [2][ll][n] key-switch result, [ll][n] c2 at Ql
    const dim3 g1((unsigned) (n_ / BEHZ_THREADS), 1);
    u64 *c2m = const_cast<u64 *>(c2);
    if (!c2_low) {   // scaleAndRound_HPS_Q_Ql (eval_key_switch.cu:139-144)
        ScaleRoundArgs a{PolyView{c2m, pq}, PolyView{c2m + pl, pq}, PolyView{low, pl}, h.dr_tab.p, h.dr_frac.p, nullptr, d_mod_.p,
                         ll, drop, 0, n_};
        launch_pdl(k_scale_round, g1, BEHZ_THREADS, 0, st, a);
        check_launch("k_scale_round");
    } else {
        PFHE_CUDA(cudaMemcpyAsync(low, c2, pl * 8, cudaMemcpyDeviceToDevice, st));
    }
    keyswitch(ll, ks, low, evk, nullptr, st);
    launch_pdl(k_expand_add, dim3((unsigned) (n_ / (2 * BEHZ_THREADS)), 2 * ll), BEHZ_THREADS, 0, st, PolyView{ct, pq},
               PolyView{ks, pl}, (const Tw *) h.expand.p, (const Modulus *) d_mod_.p, ll, n_);
    check_launch("k_expand_add");
}

const Decrypt &Engine::decrypt_tables(int l) {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if (l < 1 || l > size_Q_ || t_ <= 1) throw std::invalid_argument("decryption needs a plain modulus");
    if ((int) dec_.size() <= size_Q_) dec_.resize(size_Q_ + 1);
    if (dec_[l]) return *dec_[l];
    auto d = std::make_unique<Decrypt>();
    d->l = l;
    const std::vector<u64> Q(primes_.begin(), primes_.begin() + l);
    const u64 t = t_;
    std::vector<Tw> q_hinv(l), tg(l);
    std::vector<u64> q_to_t(l), mt(l), mtB(l), q_to_tg((size_t) 2 * l);
    std::vector<double> fr(l), frB(l);
    u64 qmax = 0;
    for (int i = 0; i < size_Q_; i++) qmax = std::max(qmax, primes_[i]);
    auto bits = [](u64 v) { return v ? 64 - __builtin_clzll(v) : 0; };
    const int qMSB = bits(qmax), sizeQMSB = bits((u64) l), tMSB = bits(t);
    d->hf = qMSB >> 1;
    d->large = qMSB + sizeQMSB >= 52;
    d->lazy = d->large ? (d->hf + tMSB + sizeQMSB) < 52 : (qMSB + tMSB + sizeQMSB) < 52;
    if (scheme_ == Scheme::bfv) {
        d->gamma = hm::create_primes(n_, std::vector<int>{61}).front();   // get_primes(n, 61, 1)[0] (rns.cu:333-334)
        const u64 g = d->gamma;
        d->ninv_t = (t - hm::invmod(hm::product_mod(Q, -1, t), t)) % t;
        d->ninv_g = (g - hm::invmod(hm::product_mod(Q, -1, g), g)) % g;
        d->inv_gamma_t = hm::invmod(g % t, t);
    }
    for (int i = 0; i < l; i++) {
        const u64 qi = Q[i], hinv = hm::invmod(hm::product_mod(Q, i, qi), qi);
        q_hinv[i] = make_tw(hinv, qi);
        q_to_t[i] = hm::product_mod(Q, i, t);
        unsigned __int128 w = (unsigned __int128) t * hinv;
        mt[i] = (u64) ((w / qi) % t);
        fr[i] = (double) (u64) (w % qi) / (double) qi;
        const u64 hb = (u64) (((unsigned __int128) hinv << d->hf) % qi);
        w = (unsigned __int128) t * hb;
        mtB[i] = (u64) ((w / qi) % t);
        frB[i] = (double) (u64) (w % qi) / (double) qi;
        if (d->gamma) {
            tg[i] = make_tw(hm::mulmod(t % qi, d->gamma % qi, qi), qi);
            q_to_tg[i] = q_to_t[i];
            q_to_tg[(size_t) l + i] = hm::product_mod(Q, i, d->gamma);
        }
    }
    d->q_mod_t = hm::product_mod(Q, -1, t);
    d->q_hinv.upload(q_hinv), d->q_to_t.upload(q_to_t);
    d->mt.upload(mt), d->mtB.upload(mtB), d->fr.upload(fr), d->frB.upload(frB);
    if (d->gamma) d->tg_mod_q.upload(tg), d->q_to_tg.upload(q_to_tg);
    dec_[l] = std::move(d);
    return *dec_[l];
}


// PhantomSecretKey::decrypt (reference src/secretkey.cu:533-691)
void Engine::decrypt(int l, const u64 *ct, int size, const u64 *sk_pow, u64 correction_factor, u64 *out, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (size < 1 || size > MXN_MAX + 1) throw std::invalid_argument("ciphertext size is not supported");
    const size_t pl = (size_t) l * n_, pk = (size_t) size_QP_ * n_;
    const dim3 gl((unsigned) (n_ / (2 * BEHZ_THREADS)), l), g1((unsigned) (n_ / BEHZ_THREADS));
    if (scheme_ == Scheme::ckks) {   // c_0 + sum c_k s^k, NTT form (ckks_decrypt :533-569)
        launch_pdl(k_decrypt_inner, gl, BEHZ_THREADS, 0, st, out, ct, ct + pl, pl, sk_pow, pk, size - 1,
                   (const Modulus *) d_mod_.p, n_);
        check_launch("k_decrypt_inner");
        return;
    }
    const Decrypt &d = decrypt_tables(l);
    const Modulus tm = host_modulus(t_);
    u64 *acc = ws_.tmp.p;   // [l][n]
    LimbVec v;
    for (int i = 0; i < l; i++) v.push(i, i);
    if (scheme_ == Scheme::bgv) {   // bgv_decrypt :638-691
        launch_pdl(k_decrypt_inner, gl, BEHZ_THREADS, 0, st, acc, ct, ct + pl, pl, sk_pow, pk, size - 1,
                   (const Modulus *) d_mod_.p, n_);
        check_launch("k_decrypt_inner");
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_inv_list(acc, acc, ll, nullptr, 0, st); });
        u64 fix = 1;
        if (correction_factor != 1) {
            if (std::gcd(correction_factor % t_, t_) != 1) throw std::logic_error("invalid correction factor");
            fix = hm::invmod(correction_factor % t_, t_);
        }
        ExactConvertArgs a{acc, out, d.q_hinv.p, d.q_to_t.p, d_mod_.p, tm, d.q_mod_t, fix, l, n_};
        launch_pdl(k_exact_convert_t, g1, BEHZ_THREADS, 0, st, a);
        check_launch("k_exact_convert_t");
        return;
    }
    // bfv_decrypt :571-636: c_k to NTT form, inner product with the key powers, back, + c_0, scale and round
    u64 *cn = ws_.t_mod_up.p;   // [size-1][l][n] transformed copies
    if (size > 1) {
        if ((size_t) (size - 1) * pl > ws_.t_mod_up.count) throw std::invalid_argument("ciphertext size is not supported");
        LimbVec vv;
        for (int k = 0; k < size - 1; k++)
            for (int i = 0; i < l; i++) vv.push(k * l + i, i);
        run_chunks(vv, rowq_, [&](const LimbList &ll, size_t) { ntt_fwd_list(cn, ct + pl, ll, st); });
        launch_pdl(k_decrypt_inner, gl, BEHZ_THREADS, 0, st, acc, (const u64 *) nullptr, (const u64 *) cn, pl, sk_pow, pk,
                   size - 1, (const Modulus *) d_mod_.p, n_);
        check_launch("k_decrypt_inner");
        run_chunks(v, rowq_, [&](const LimbList &ll, size_t) { ntt_inv_list(acc, acc, ll, nullptr, 0, st); });
    } else {
        PFHE_CUDA(cudaMemsetAsync(acc, 0, pl * 8, st));
    }
    if (mul_tech_ == 1) {
        const Modulus gm = host_modulus(d.gamma);
        BehzDecryptArgs a{acc, ct, out, d.tg_mod_q.p, d.q_hinv.p, d.q_to_tg.p, d_mod_.p, tm, gm,
                          d.ninv_t, d.ninv_g, d.inv_gamma_t, l, n_};
        launch_pdl(k_behz_decrypt, g1, BEHZ_THREADS, 0, st, a);
        check_launch("k_behz_decrypt");
    } else {
        HpsDecryptArgs a{acc, ct, out, d.mt.p, d.mtB.p, d.fr.p, d.frB.p, d_mod_.p, tm, l, d.large, d.lazy, d.hf, n_};
        launch_pdl(k_hps_decrypt, g1, BEHZ_THREADS, 0, st, a);
        check_launch("k_hps_decrypt");
    }
}

void Engine::ckks_tables() {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    const size_t slots = n_ >> 1;
    const uint32_t M = (uint32_t) (n_ << 1);
    if (!d_ckks_roots_.p) {
        // ComplexRoots (src/fft.cu:13-43): an eighth of the circle from polar(), the rest by symmetry
        const double PI_ = 3.1415926535897932384626433832795028842;
        std::vector<double2> eighth(M / 8 + 1), roots(M + 1);
        for (size_t i = 0; i <= M / 8; i++) {
            const double th = 2 * PI_ * (double) i / (double) M;
            eighth[i] = make_double2(std::cos(th), std::sin(th));
        }
        std::function<double2(size_t)> root = [&](size_t i) -> double2 {
            i &= M - 1;
            if (i <= M / 8) return eighth[i];
            if (i <= M / 4) {
                const double2 r = eighth[M / 4 - i];
                return make_double2(r.y, r.x);
            }
            if (i <= M / 2) {
                const double2 r = root(M / 2 - i);
                return make_double2(0.0 - r.x, 0.0 - (-r.y));
            }
            if (i <= 3 * (size_t) M / 4) {
                const double2 r = root(i - M / 2);
                return make_double2(0.0 - r.x, 0.0 - r.y);
            }
            const double2 r = root(M - i);
            return make_double2(r.x, -r.y);
        };
        for (size_t i = 0; i <= M; i++) roots[i] = root(i);   // entry M = entry 0: the inverse reads tw[M - psi]
        std::vector<uint32_t> group(std::max<size_t>(slots / 2, 1));
        uint32_t pos = 1;
        for (size_t i = 0; i < slots / 2; i++) group[i] = pos, pos = (pos * 5) & (M - 1);
        d_ckks_roots_.upload(roots), d_ckks_group_.upload(group);
        d_ckks_x_.alloc(slots), d_ckks_max_.alloc(1);
    }
}

// PhantomCKKSEncoder::decode_internal (reference src/ckks.cu:137-190)
void Engine::ckks_decode(int l, const u64 *plain, double scale, double2 *out, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (scheme_ != Scheme::ckks) throw std::invalid_argument("unsupported scheme");
    if (l < 1 || l > size_Q_) throw std::invalid_argument("index is invalid!");
    if (l > CKKS_MAX_WORDS) throw std::invalid_argument("too many limbs for the CKKS decoder");
    std::vector<u64> ql(primes_.begin(), primes_.begin() + l);
    if (scale <= 0 || (int) std::log2(scale) >= hm::product_bits(ql)) throw std::invalid_argument("scale out of bounds");
    ckks_tables();
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if ((int) ckks_dec_.size() <= size_Q_) ckks_dec_.resize(size_Q_ + 1);
    if (!ckks_dec_[l]) {
        auto d = std::make_unique<CkksDec>();
        auto words = [&](hm::BigUint b) {
            b.w.resize(l, 0);
            return b.w;
        };
        hm::BigUint Qb;
        for (u64 q : ql) Qb.mul_word(q);
        std::vector<u64> Qw = words(Qb), thr(l), hat((size_t) l * l);
        std::vector<Tw> hinv(l);
        {   // (Q + 1) >> 1
            std::vector<u64> t(Qw);
            for (int k = 0; k < l; k++)
                if (++t[k] != 0) break;
            for (int k = 0; k < l; k++) thr[k] = (t[k] >> 1) | (k + 1 < l ? t[k + 1] << 63 : 0);
        }
        for (int i = 0; i < l; i++) {
            hm::BigUint h;
            for (int j = 0; j < l; j++)
                if (j != i) h.mul_word(ql[j]);
            const std::vector<u64> hw = words(h);
            std::copy(hw.begin(), hw.end(), hat.begin() + (size_t) i * l);
            hinv[i] = make_tw(hm::invmod(hm::product_mod(ql, i, ql[i]), ql[i]), ql[i]);
        }
        d->hat.upload(hat), d->Qw.upload(Qw), d->thr.upload(thr), d->hinv.upload(hinv);
        ckks_dec_[l] = std::move(d);
    }
    const CkksDec &d = *ckks_dec_[l];
    const size_t slots = n_ >> 1;
    const uint32_t M = (uint32_t) (n_ << 1);
    const int logs = logn_ - 1;
    u64 *w = ws_.tmp.p;
    ntt_inv_rows_range(w, plain, l, 0, st);
    double2 *x = d_ckks_x_.p;
    CkksComposeArgs a{x, w, d.hat.p, d.Qw.p, d.thr.p, d.hinv.p, d_mod_.p, 1.0 / scale, l, n_};
    launch_pdl(k_ckks_compose, dim3((unsigned) (n_ / EW_THREADS)), EW_THREADS, 0, st, a);
    check_launch("k_ckks_compose");
    const int blk_log = std::min(logs, CKKS_FFT_LOG_BLOCK);
    const int iter_begin = logs - blk_log;
    for (int iter = 0; iter < iter_begin; iter++) {
        launch_pdl(k_ckks_fft_stage, dim3((unsigned) (slots / 2 / EW_THREADS)), EW_THREADS, 0, st, x,
                   (const double2 *) d_ckks_roots_.p, (const uint32_t *) d_ckks_group_.p, logs, iter, M);
        check_launch("k_ckks_fft_stage");
    }
    launch_pdl(k_ckks_fft_block, dim3((unsigned) (slots >> blk_log)), dim3(1u << (blk_log - 1)),
               ((size_t) 1 << blk_log) * sizeof(double2), st, x, (const double2 *) d_ckks_roots_.p,
               (const uint32_t *) d_ckks_group_.p, logs, iter_begin, M);
    check_launch("k_ckks_fft_block");
    launch_pdl(k_ckks_unplace, dim3((unsigned) (slots / EW_THREADS)), EW_THREADS, 0, st, out, (const double2 *) x, logs);
    check_launch("k_ckks_unplace");
}

// PhantomCKKSEncoder::encode_internal (reference src/ckks.cu:66-135)
void Engine::ckks_encode(int l, const double2 *values, size_t count, double scale, u64 *out, cudaStream_t st) {
    if (scheme_ != Scheme::ckks) throw std::invalid_argument("unsupported scheme");
    const size_t slots = n_ >> 1;
    const uint32_t M = (uint32_t) (n_ << 1);
    const int logs = logn_ - 1;
    if (count == 0) throw std::invalid_argument("Input vector is empty");
    if (count > slots) throw std::invalid_argument("Input vector exceeds max slots");
    std::vector<u64> ql(primes_.begin(), primes_.begin() + l);
    const int qbits = hm::product_bits(ql);
    if (scale <= 0 || (int) std::log2(scale) + 1 >= qbits) throw std::invalid_argument("scale out of bounds");
    ckks_tables();
    double2 *x = d_ckks_x_.p;
    launch_pdl(k_ckks_place, dim3((unsigned) (slots / EW_THREADS)), EW_THREADS, 0, st, x, values, count, logs);
    check_launch("k_ckks_place");
    const double fix = scale / (double) slots;
    // stages logs-1 .. iter_end inside blocks, the rest one launch each
    const int blk_log = std::min(logs, CKKS_FFT_LOG_BLOCK);
    const int iter_end = logs - blk_log;
    const unsigned threads = 1u << (blk_log - 1);
    launch_pdl(k_ckks_ifft_block, dim3((unsigned) (slots >> blk_log)), dim3(threads), ((size_t) 1 << blk_log) * sizeof(double2), st,
               x, (const double2 *) d_ckks_roots_.p, (const uint32_t *) d_ckks_group_.p, logs, iter_end, M, fix);
    check_launch("k_ckks_ifft_block");
    for (int iter = iter_end - 1; iter >= 0; iter--) {
        launch_pdl(k_ckks_ifft_stage, dim3((unsigned) (slots / 2 / EW_THREADS)), EW_THREADS, 0, st, x,
                   (const double2 *) d_ckks_roots_.p, (const uint32_t *) d_ckks_group_.p, logs, iter, M, fix);
        check_launch("k_ckks_ifft_stage");
    }
    // size of the encoded coefficients (the reference copies the vector to the host for this, ckks.cu:107-124)
    PFHE_CUDA(cudaMemsetAsync(d_ckks_max_.p, 0, sizeof(unsigned long long), st));
    launch_pdl(k_ckks_absmax, dim3((unsigned) (slots / EW_THREADS)), EW_THREADS, 0, st, (const double2 *) x, d_ckks_max_.p);
    check_launch("k_ckks_absmax");
    unsigned long long mb = 0;
    PFHE_CUDA(cudaMemcpyAsync(&mb, d_ckks_max_.p, sizeof(mb), cudaMemcpyDeviceToHost, st));
    PFHE_CUDA(cudaStreamSynchronize(st));
    double max_coeff;
    std::memcpy(&max_coeff, &mb, sizeof(double));
    const int bits = (int) std::ceil(std::log2(std::max(max_coeff, 1.0))) + 1;
    if (bits >= qbits) throw std::invalid_argument("encoded values are too large");
    if (bits > 128) throw std::invalid_argument("encoded values need more than 128 bits: not supported");
    launch_pdl(k_ckks_decompose, dim3((unsigned) (n_ / EW_THREADS), l), EW_THREADS, 0, st, out, (const double2 *) x,
               (const Modulus *) d_mod_.p, n_, bits > 64 ? 1 : 0);
    check_launch("k_ckks_decompose");
    ntt_fwd_rows_range(out, l, 0, st);
}

// PhantomBatchEncoder (reference src/batchencoder.cu): slots <-> plaintext polynomial mod t
void Engine::batch_encode(const u64 *values, size_t count, u64 *plain, cudaStream_t st) {
    if (scheme_ == Scheme::ckks) throw std::invalid_argument("PhantomBatchEncoder only supports BFV/BGV scheme");
    if (!batching_) throw std::invalid_argument("the plain modulus does not support batching");
    if (count > n_) throw std::logic_error("values_matrix size is too large");
    ensure_batch_map();
    launch_pdl(k_batch_encode, dim3((unsigned) (n_ / EW_THREADS)), EW_THREADS, 0, st, plain, values, count,
               (const uint32_t *) d_batch_map_.p, t_);
    check_launch("k_batch_encode");
    LimbVec v;
    v.push(0, size_QP_);   // the row of t
    ntt_inv_list(plain, plain, single_list(v, rowq_), nullptr, 0, st);
}

void Engine::ensure_batch_map() {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    if (!d_batch_map_.p) {
        std::vector<uint32_t> map(n_);
        const size_t row = n_ >> 1, m = n_ << 1;
        u64 pos = 1;
        for (size_t i = 0; i < row; i++) {   // populate_matrix_reps_index_map, batchencoder.cu:26-49
            map[i] = hm::bit_reverse((uint32_t) ((pos - 1) >> 1), logn_);
            map[row | i] = hm::bit_reverse((uint32_t) ((m - pos - 1) >> 1), logn_);
            pos = (pos * 5) & (m - 1);
        }
        d_batch_map_.upload(map);
    }
}

void Engine::batch_decode(const u64 *plain, u64 *values, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (scheme_ == Scheme::ckks) throw std::invalid_argument("PhantomBatchEncoder only supports BFV/BGV scheme");
    if (!batching_) throw std::invalid_argument("the plain modulus does not support batching");
    ensure_batch_map();
    u64 *tmp = ws_.tmp.p;
    LimbVec v;
    v.push(0, size_QP_);
    ntt_fwd_list(tmp, plain, single_list(v, rowq_), st);
    launch_pdl(k_batch_decode, dim3((unsigned) (n_ / EW_THREADS)), EW_THREADS, 0, st, values, (const u64 *) tmp,
               (const uint32_t *) d_batch_map_.p);
    check_launch("k_batch_decode");
}

// FindLevelsToDrop (reference src/evaluate.cu:550-643), same double-precision formulas in the same order
int Engine::find_levels_to_drop(size_t multiplicativeDepth, bool isKeySwitch, bool isAsymmetric) {
    if (scheme_ != Scheme::bfv || mul_tech_ != 4)
        throw std::invalid_argument("FindLevelsToDrop is only used in HPS over Q Leveled");
    const uint32_t n = (uint32_t) n_;
    const double sigma = 3.2f, alpha = 36.0f;   // distributionParameter, assuranceMeasure (host/hestdparms.h:152-153)
    const double p = (double) t_;
    const uint32_t k = (uint32_t) size_P_;
    const uint32_t numPartQ = (uint32_t) beta(size_Q_);
    const double Bkey = 1.0;
    u64 qmax = 0;
    for (int i = 0; i < size_Q_; i++) qmax = std::max(qmax, primes_[i]);
    const double dcrtBits = (double) (64 - __builtin_clzll(qmax));   // qMSB (rns.cu:587)
    const double Berr = sigma * sqrt(alpha);
    auto delta = [](uint32_t nn) -> double { return (2. * sqrt(nn)); };
    auto Vnorm = [&](uint32_t nn) -> double {
        if (isAsymmetric) return (1. + delta(nn) * Bkey) / 2.;
        return Berr * (1. + 2. * delta(nn) * Bkey);
    };
    auto noiseKS = [&](uint32_t nn) -> double { return k * (numPartQ * delta(nn) * Berr + delta(nn) * Bkey + 1.0) / 2; };
    auto C1 = [&](uint32_t nn) -> double { return delta(nn) * delta(nn) * p * Bkey; };
    auto C2 = [&](uint32_t nn) -> double { return delta(nn) * delta(nn) * Bkey * Bkey / 2.0 + noiseKS(nn); };
    auto logqBFV = [&](uint32_t nn) -> double {
        if (multiplicativeDepth > 0)
            return log(4 * p) + (multiplicativeDepth - 1) * log(C1(nn)) +
                   log(C1(nn) * Vnorm(nn) + multiplicativeDepth * C2(nn));
        return log(p * (4 * (Vnorm(nn))));
    };
    double logqPrev = 6. * log(10);
    double logq = logqBFV(n);
    while (fabs(logq - logqPrev) > log(1.001)) {
        logqPrev = logq;
        logq = logqBFV(n);
    }
    const double loge = logq / log(2) - 2 - log2(p);
    const double logExtra = isKeySwitch ? log2(noiseKS(n)) : log2(delta(n));
    int32_t levels = (int32_t) std::floor((loge - 2 * multiplicativeDepth - 16 - logExtra) / dcrtBits);
    if (levels < 0) levels = 0;
    else if (levels > size_Q_ - 1) levels = size_Q_ - 1;
    return levels;
}

void Engine::bfv_multiply(int l, u64 *out3, const u64 *ct1, const u64 *ct2, cudaStream_t st, int drop) {
    if (scheme_ != Scheme::bfv) throw std::invalid_argument("unsupported scheme");
    if (mul_tech_ == 1) bfv_multiply_behz(l, out3, ct1, ct2, st);
    else if (mul_tech_ == 2) bfv_multiply_hps(l, out3, ct1, ct2, st);
    else if (mul_tech_ == 3) bfv_multiply_hps_overq(l, out3, ct1, ct2, 0, st);
    else bfv_multiply_hps_overq(l, out3, ct1, ct2, drop, st);   // hps_overq_leveled: the caller supplies the levels
}

// multiply_inplace + relinearize_inplace (reference src/evaluate.cu:345-397,451-548,819-1026,1342-1374)
void Engine::multiply_relin(int l, u64 *out, const u64 *ct1, const u64 *ct2, const u64 *const *rlk, cudaStream_t st,
                            int leveled_drop) {
    Workspace &ws_ = ws(st);
    if (scheme_ == Scheme::bfv) {
        u64 *d = ws_.tmp.p;
        const int drop = mul_tech_ == 4 ? leveled_drop : 0;   // hps_overq_leveled: the caller supplies the levels
        if (drop) {   // bfv_mul_relin_hps with levels dropped (evaluate.cu:819-1026): c2 stays at Ql and is switched there
            bfv_multiply_hps_overq(l, d, ct1, ct2, drop, st, true);
            keyswitch_leveled(d, d + (size_t) 2 * l * n_, rlk, drop, true, st);
            PFHE_CUDA(cudaMemcpyAsync(out, d, (size_t) 2 * l * n_ * 8, cudaMemcpyDeviceToDevice, st));
        } else {
            bfv_multiply(l, d, ct1, ct2, st);
            keyswitch(l, out, d + (size_t) 2 * l * n_, rlk, d, st);   // out = (d0, d1) + keyswitch(d2): no copy
        }
        return;
    }
    if (out == ct1 || out == ct2) {
        // the fused epilogue reads a0, a1, b0, b1 while writing out: keep the operands alive in the workspace
        const size_t words = (size_t) 2 * l * n_;
        u64 *keep = ws_.tmp.p;   // [2][l][n] copy of whichever operand aliases the output
        if (out == ct1) {
            PFHE_CUDA(cudaMemcpyAsync(keep, ct1, words * 8, cudaMemcpyDeviceToDevice, st));
            const TensorSrc ts{keep, ct2 == ct1 ? keep : ct2, l};
            keyswitch_fused(l, out, nullptr, &ts, rlk, nullptr, 0u, st);
        } else {
            PFHE_CUDA(cudaMemcpyAsync(keep, ct2, words * 8, cudaMemcpyDeviceToDevice, st));
            const TensorSrc ts{ct1, keep, l};
            keyswitch_fused(l, out, nullptr, &ts, rlk, nullptr, 0u, st);
        }
        return;
    }
    const TensorSrc ts{ct1, ct2, l};
    keyswitch_fused(l, out, nullptr, &ts, rlk, nullptr, 0u, st);
}

void Engine::multiply_relin_host_batch(int l, const u64 *const *h1, const u64 *const *h2, u64 *const *hout,
                                       size_t count, const u64 *const *rlk, cudaStream_t st) {
    const size_t words = (size_t) 2 * l * n_;
    StreamCtx &c = sctx(st);
    if (!c.s_in) {
        PFHE_CUDA(cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking));
        PFHE_CUDA(cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            PFHE_CUDA(cudaEventCreateWithFlags(&c.ev_in[i], cudaEventDisableTiming));
            PFHE_CUDA(cudaEventCreateWithFlags(&c.ev_comp[i], cudaEventDisableTiming));
            PFHE_CUDA(cudaEventCreateWithFlags(&c.ev_out[i], cudaEventDisableTiming));
        }
    }
    for (int i = 0; i < 2; i++) {
        if (c.pipe_in[i].count < 2 * words) c.pipe_in[i].alloc(2 * words);
        if (c.pipe_out[i].count < words) c.pipe_out[i].alloc(words);
    }
    // side streams start after whatever the caller already queued on `st`
    PFHE_CUDA(cudaEventRecord(c.ev_comp[0], st));
    PFHE_CUDA(cudaStreamWaitEvent(c.s_in, c.ev_comp[0], 0));
    PFHE_CUDA(cudaStreamWaitEvent(c.s_out, c.ev_comp[0], 0));
    for (size_t i = 0; i < count; i++) {
        const int b = (int) (i & 1);
        if (i >= 2) PFHE_CUDA(cudaStreamWaitEvent(c.s_in, c.ev_comp[b], 0));   // input buffer b consumed
        PFHE_CUDA(cudaMemcpyAsync(c.pipe_in[b].p, h1[i], words * 8, cudaMemcpyHostToDevice, c.s_in));
        PFHE_CUDA(cudaMemcpyAsync(c.pipe_in[b].p + words, h2[i], words * 8, cudaMemcpyHostToDevice, c.s_in));
        PFHE_CUDA(cudaEventRecord(c.ev_in[b], c.s_in));
        PFHE_CUDA(cudaStreamWaitEvent(st, c.ev_in[b], 0));
        if (i >= 2) PFHE_CUDA(cudaStreamWaitEvent(st, c.ev_out[b], 0));       // output buffer b drained
        multiply_relin(l, c.pipe_out[b].p, c.pipe_in[b].p, c.pipe_in[b].p + words, rlk, st);
        PFHE_CUDA(cudaEventRecord(c.ev_comp[b], st));
        PFHE_CUDA(cudaStreamWaitEvent(c.s_out, c.ev_comp[b], 0));
        PFHE_CUDA(cudaMemcpyAsync(hout[i], c.pipe_out[b].p, words * 8, cudaMemcpyDeviceToHost, c.s_out));
        PFHE_CUDA(cudaEventRecord(c.ev_out[b], c.s_out));
    }
    for (int b = 0; b < 2 && (size_t) b < count; b++) PFHE_CUDA(cudaStreamWaitEvent(st, c.ev_out[b], 0));
}

// apply_galois_inplace for CKKS/BGV (reference src/evaluate.cu:1567-1630)
void Engine::apply_galois(int l, u64 *ct, uint32_t galois_elt, const u64 *const *glk, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    const int gi = galois_index(galois_elt);
    u64 *tmp = ws_.tmp.p;   // [2][l][n]: permuted c0, c1
    if (scheme_ == Scheme::bfv) {
        galois_coeff(tmp, ct, galois_elt, l, 2, st);
    } else {
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), 2 * l);
        launch_pdl(k_galois_ntt, grid, EW_THREADS, 0, st, tmp, ct, d_perm_[gi].p, n_);
        check_launch("k_galois_ntt");
    }
    if (scheme_ != Scheme::ckks) {
        // ct0 = perm(c0) + ks0, ct1 = ks1
        PFHE_CUDA(cudaMemcpyAsync(ct, tmp, (size_t) l * n_ * 8, cudaMemcpyDeviceToDevice, st));
        PFHE_CUDA(cudaMemsetAsync(ct + (size_t) l * n_, 0, (size_t) l * n_ * 8, st));
        // tmp[1] (the key-switch input) must survive modup's use of the workspace: it only touches t_cks/t_mod_up/cx
        keyswitch(l, ct, tmp + (size_t) l * n_, glk, ct, st);
        return;
    }
    // ct0 = perm(c0) + ks0, ct1 = 0 + ks1: the "wipe c1" memset of the reference is folded away by
    // pointing poly 1 at a zero addend, i.e. no addend at all.
    keyswitch_fused(l, ct, tmp + (size_t) l * n_, nullptr, glk, tmp, 1u, st);
}

// hoisting_inplace (reference src/evaluate.cu:1670-1865): one mod-up of c1 shared by all rotations, the automorphism of the
// digits folded into the inner product, one mod-down.  BFV: coefficient-form ends (apply_galois on c0, mod-up from and
// mod-down to coefficient form); under hps_overq_leveled with `drop` levels dropped the ciphertext is first scaled from Q
// to Ql (scaleAndRound_HPS_Q_Ql, :1731-1733,1761-1763) and the result expanded back (ExpandCRTBasis_Ql_Q, :1847-1862).
void Engine::hoisting(int l, u64 *ct, const std::vector<uint32_t> &elts, const std::vector<const u64 *const *> &keys,
                      cudaStream_t st, int drop) {
    Workspace &ws_ = ws(st);
    if (elts.empty() || elts.size() != keys.size()) throw std::invalid_argument("steps / keys mismatch");
    const bool bfv = scheme_ == Scheme::bfv;
    if (drop && (!bfv || mul_tech_ != 4 || l != size_Q_)) throw std::invalid_argument("levels can be dropped under hps_overq_leveled only");
    if (drop < 0 || drop >= l) throw std::invalid_argument("levels dropped is out of range");
    const int ll = l - drop;
    const Level &lv = level(ll);
    (void) lv;
    const size_t poly = (size_t) ll * n_, pq = (size_t) l * n_;
    u64 *acc_c0 = ws_.tmp.p;                 // [ll][n]
    u64 *c0 = ws_.tmp.p + poly, *c1 = ws_.tmp.p + 2 * poly;   // private copies: ct is overwritten at the end
    if (drop) {
        const HpsQ &h = hpsq(drop);
        ScaleRoundArgs a{PolyView{ct, pq}, PolyView{ct + poly, pq}, PolyView{c0, poly}, h.dr_tab.p, h.dr_frac.p, nullptr, d_mod_.p,
                         ll, drop, 0, n_};
        launch_pdl(k_scale_round, dim3((unsigned) (n_ / BEHZ_THREADS), 2), BEHZ_THREADS, 0, st, a);
        check_launch("k_scale_round");
    } else {
        PFHE_CUDA(cudaMemcpyAsync(c0, ct, 2 * poly * 8, cudaMemcpyDeviceToDevice, st));
    }
    // one mod-up of c1 for all rotations (:1769)
    modup(ll, ws_.t_mod_up.p, c1, ws_.t_cks.p, st);
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), ll);
    for (size_t i = 0; i < elts.size(); i++) {
        int gi;
        try {
            gi = galois_index(elts[i]);
        } catch (const std::invalid_argument &) { throw std::logic_error("Galois key not present in hoisting"); }
        if (bfv) {   // apply_galois on the coefficient form (:1745-1747, 1808-1810)
            if (i == 0) {
                galois_coeff(acc_c0, c0, elts[i], ll, 1, st);
            } else {
                u64 *t = ws_.delta.p;
                galois_coeff(t, c0, elts[i], ll, 1, st);
                elementwise(EW_ADD, acc_c0, t, acc_c0, ll, st);
            }
        } else {
            if (i == 0) launch_pdl(k_galois_ntt, grid, EW_THREADS, 0, st, acc_c0, c0, d_perm_[gi].p, n_);
            else launch_pdl(k_galois_ntt_acc, grid, EW_THREADS, 0, st, acc_c0, c0, d_perm_[gi].p, d_mod_.p, n_);
            check_launch("k_galois_ntt");
        }
        // automorphism of the digits + inner product + accumulation in one kernel
        inner_prod(ll, ws_.cx.p, ws_.t_mod_up.p, keys[i], st, nullptr, nullptr, d_perm_[gi].p, i > 0);
    }
    // one mod-down per polynomial; new c0 = acc_c0 + moddown(cx0), new c1 = moddown(cx1)
    u64 *out = drop ? c0 : ct;
    if (scheme_ == Scheme::ckks) moddown(ll, out, ws_.cx.p, ws_.delta.p, 2, acc_c0, 1u, st);
    else moddown_generic(ll, out, ws_.cx.p, 2, acc_c0, 1u, st);
    if (drop) {
        const HpsQ &h = hpsq(drop);
        PFHE_CUDA(cudaMemsetAsync(ct, 0, 2 * pq * 8, st));
        launch_pdl(k_expand_add, dim3((unsigned) (n_ / (2 * BEHZ_THREADS)), 2 * ll), BEHZ_THREADS, 0, st, PolyView{ct, pq},
                   PolyView{c0, poly}, (const Tw *) h.expand.p, (const Modulus *) d_mod_.p, ll, n_);
        check_launch("k_expand_add");
    }
}

// rescale_to_next for CKKS (reference src/evaluate.cu:1376-1427 + divide_and_round_q_last_ntt rns.cu:1160-1184)
void Engine::rescale(int l, u64 *out, const u64 *in, int size, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (l < 2) throw std::invalid_argument("end of modulus switching chain reached");
    const Level &lv = level(l);
    const int nl = l - 1;
    u64 *last = ws_.t_cks.p;   // [size][n] coefficient form of the dropped limb
    {
        LimbVec v;
        for (int s = 0; s < size; s++) v.push(s, l - 1, s * l + (l - 1));
        ntt_inv_list(last, in, single_list(v, rowq_), nullptr, 0, st);
    }
    for (int s = 0; s < size; s++) {
        dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), nl);
        launch_pdl(k_reduce_last, grid, EW_THREADS, 0, st, out + (size_t) s * nl * n_, last + (size_t) s * n_, d_mod_.p, n_);
        check_launch("k_reduce_last");
    }
    LimbVec v;
    for (int s = 0; s < size; s++)
        for (int j = 0; j < nl; j++) v.push(s * nl + j, j);
    run_chunks(v, rowq_, [&](const LimbList &ll, size_t b) {
        EpiArgs ea{};
        ea.sub_base = in, ea.out_base = out, ea.add_base = nullptr, ea.mulc = lv.qlast_inv_slots.p + b;
        for (int k = 0; k < ll.count; k++) {
            const int s = (int) ((b + k) / nl), j = (int) ((b + k) % nl);
            ea.sub[k] = (short) (s * l + j);
            ea.out[k] = (short) (s * nl + j);
            ea.add[k] = -1;
        }
        g_launches.fetch_add(2, std::memory_order_relaxed);
            PFHE_CUDA(ntt_forward_epilogue(plan_, out, ll, ea, st));
    });
}

// mod_switch_drop_to_next for CKKS (reference src/evaluate.cu:1429-1472): keep the first l-1 limbs
void Engine::mod_switch_drop(int l, u64 *out, const u64 *in, int size, cudaStream_t st) const {
    if (l < 2) throw std::invalid_argument("end of modulus switching chain reached");
    for (int s = 0; s < size; s++)
        PFHE_CUDA(cudaMemcpyAsync(out + (size_t) s * (l - 1) * n_, in + (size_t) s * l * n_, (size_t) (l - 1) * n_ * 8,
                                  cudaMemcpyDeviceToDevice, st));
}

// ---- samplers, key generation, encryption ----------------------------------------------------

// sample_ternary_poly / sample_error_poly / sample_uniform_poly (reference src/prng.cu:142-244): out = [limbs][n] over the
// modulus rows 0 .. limbs-1 (limbs = size_QP at the key level, l at a data level)
void Engine::sample_poly(int kind, int limbs, const Seed &seed, u64 *out, cudaStream_t st) const {
    if (limbs < 1 || limbs > size_QP_) throw std::invalid_argument("limb count out of range");
    const dim3 per_coeff((unsigned) (n_ / EW_THREADS));
    switch (kind) {
        case SAMPLE_TERNARY:
            launch_pdl(k_sample_small<SAMPLE_TERNARY>, per_coeff, EW_THREADS, 0, st, out, seed, (const Modulus *) d_mod_.p, n_, limbs);
            break;
        case SAMPLE_ERROR:
            launch_pdl(k_sample_small<SAMPLE_ERROR>, per_coeff, EW_THREADS, 0, st, out, seed, (const Modulus *) d_mod_.p, n_, limbs);
            break;
        case SAMPLE_UNIFORM:
            launch_pdl(k_sample_uniform, dim3((unsigned) ((n_ / 8 + EW_THREADS - 1) / EW_THREADS), limbs), EW_THREADS, 0, st, out,
                       seed, (const Modulus *) d_mod_.p, n_, limbs);
            break;
        default: throw std::invalid_argument("unknown sampler");
    }
    check_launch("k_sample");
}

// PhantomSecretKey::gen_secretkey (secretkey.cu:345-378): ternary secret over every key prime, NTT form, [size_QP][n]
void Engine::gen_secret_key(const Seed &seed, u64 *sk, cudaStream_t st) const {
    sample_poly(SAMPLE_TERNARY, size_QP_, seed, sk, st);
    ntt_fwd_rows_range(sk, size_QP_, 0, st);
}

// PhantomSecretKey::encrypt_zero_symmetric (secretkey.cu:232-295): out = [2][limbs][n] = (-(a s + e), a), e scaled by t
// for BGV; NTT form, or (BFV) coefficient form.  sk = first power of the key, NTT form, key-level layout
void Engine::encrypt_zero_symmetric(int limbs, bool ntt_form, const u64 *sk, const Seed &seed_a, const Seed &seed_e, u64 *out,
                                    cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (limbs < 1 || limbs > size_QP_) throw std::invalid_argument("limb count out of range");
    u64 *c0 = out, *c1 = out + (size_t) limbs * n_, *e = ws_.cx.p;
    const dim3 grid((unsigned) (n_ / EW_THREADS), limbs);
    sample_poly(SAMPLE_ERROR, limbs, seed_e, e, st);
    sample_poly(SAMPLE_UNIFORM, limbs, seed_a, c1, st);
    if (ntt_form) {
        if (scheme_ == Scheme::bgv) {
            launch_pdl(k_scale_by, grid, EW_THREADS, 0, st, e, t_, (const Modulus *) d_mod_.p, n_);
            check_launch("k_scale_by");
        }
        ntt_fwd_rows_range(e, limbs, 0, st);
        launch_pdl(k_enc_fma<true>, grid, EW_THREADS, 0, st, c0, (const u64 *) c1, sk, (const u64 *) e, (const Modulus *) d_mod_.p, n_);
        check_launch("k_enc_fma");
    } else {
        // the uniform sample is read as the NTT form of a: c0 = -(intt(a s) + e), c1 = intt(a).  In NTT form first
        // (-(a s + ntt(e))), then both polynomials back: the same residues, one batched inverse transform
        ntt_fwd_rows_range(e, limbs, 0, st);
        launch_pdl(k_enc_fma<true>, grid, EW_THREADS, 0, st, c0, (const u64 *) c1, sk, (const u64 *) e, (const Modulus *) d_mod_.p, n_);
        check_launch("k_enc_fma");
        ntt_batch(out, 2, limbs, 0, true, st);
    }
}

// PhantomPublicKey::encrypt_zero_asymmetric_internal (secretkey.cu:10-128) at the first data level: an encryption of zero
// under the public key at the key level, (u pk_i + e) with e scaled by t for BGV, brought down to Q by dividing by P
// (DRNSTool::moddown, rns_bconv.cu:712-761).  out = [2][size_Q][n], NTT form (CKKS, BGV) or coefficient form (BFV).
// Like the reference, both polynomials get the same error polynomial (one seed, nonces restart at zero).
void Engine::encrypt_zero_asymmetric(const u64 *pk, const Seed &seed_u, const Seed &seed_e, u64 *out, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (size_P_ < 1) throw std::invalid_argument("asymmetric encryption needs a special modulus");
    if (3 * size_Q_ < size_QP_) throw std::invalid_argument("special modulus larger than the workspace allows");
    const int m = size_QP_, l = size_Q_;
    u64 *u = ws_.t_mod_up.p, *e = ws_.tmp.p, *cx = ws_.cx.p;
    const dim3 grid((unsigned) (n_ / EW_THREADS), m);
    sample_poly(SAMPLE_TERNARY, m, seed_u, u, st);
    ntt_fwd_rows_range(u, m, 0, st);
    sample_poly(SAMPLE_ERROR, m, seed_e, e, st);
    if (scheme_ == Scheme::bgv) {
        launch_pdl(k_scale_by, grid, EW_THREADS, 0, st, e, t_, (const Modulus *) d_mod_.p, n_);
        check_launch("k_scale_by");
    }
    ntt_fwd_rows_range(e, m, 0, st);
    for (int i = 0; i < 2; i++) {
        launch_pdl(k_enc_fma<false>, grid, EW_THREADS, 0, st, cx + (size_t) i * m * n_, (const u64 *) u, pk + (size_t) i * m * n_,
                   (const u64 *) e, (const Modulus *) d_mod_.p, n_);
        check_launch("k_enc_fma");
    }
    // BFV: the reference goes to coefficient form before the division; dividing from NTT form gives the same residues
    if (scheme_ == Scheme::ckks) moddown(l, out, cx, ws_.delta.p, 2, nullptr, 0, st);
    else moddown_generic(l, out, cx, 2, nullptr, 0, st);
}

// PhantomSecretKey::generate_one_kswitch_key (secretkey.cu:297-343): digit d = encryption of zero at the key level with
// P * new_key added to the limbs of digit d of its first polynomial.  digits = host array of dnum device buffers
// [2][size_QP][n]; seeds = dnum pairs (a, e).  Like the reference dnum = size_Q / size_P: size_Q must be a multiple of it.
void Engine::kswitch_key(const u64 *new_key, const u64 *sk, const Seed *seeds, u64 *const *digits, cudaStream_t st) {
    if (size_P_ < 1 || size_Q_ % size_P_) throw std::invalid_argument("size_Q must be a multiple of size_P");
    const int dnum = size_Q_ / size_P_;
    if (!d_p_mod_q_.p) {
        std::vector<u64> v(size_Q_);
        for (int j = 0; j < size_Q_; j++) {
            u64 r = 1;
            for (int i = 0; i < size_P_; i++) r = hm::mulmod(r, primes_[size_Q_ + i] % primes_[j], primes_[j]);
            v[j] = r;
        }
        d_p_mod_q_.upload(v);
    }
    for (int d = 0; d < dnum; d++) encrypt_zero_symmetric(size_QP_, true, sk, seeds[2 * d], seeds[2 * d + 1], digits[d], st);
    DevBuf<u64 *> &ptrs = d_digit_ptrs_;
    if (ptrs.count < (size_t) dnum) ptrs.alloc(dnum);
    PFHE_CUDA(cudaMemcpyAsync(ptrs.p, digits, sizeof(u64 *) * dnum, cudaMemcpyHostToDevice, st));
    PFHE_CUDA(cudaStreamSynchronize(st));   // `digits` is the caller's (pageable) array
    launch_pdl(k_kswitch_target, dim3((unsigned) (n_ / EW_THREADS), size_Q_), EW_THREADS, 0, st, (u64 *const *) ptrs.p, new_key,
               (const u64 *) d_p_mod_q_.p, (const Modulus *) d_mod_.p, n_, size_P_);
    check_launch("k_kswitch_target");
}

// PhantomGaloisTool::apply_galois_ntt (galois.cu:86-102): result[limb][i] = operand[limb][perm[i]] over `limbs` limbs; the
// index permutation is the same for every modulus.  With limbs = size_QP on the secret key this is the key-switching
// target of a Galois key (create_galois_keys, secretkey.cu:443-451)
void Engine::galois_ntt(const u64 *operand, int limbs, uint32_t galois_elt, u64 *result, cudaStream_t st) const {
    if (limbs < 1) throw std::invalid_argument("limb count out of range");
    if (operand == result) throw std::invalid_argument("the permutation does not work in place");
    const int gi = galois_index(galois_elt);
    launch_pdl(k_galois_ntt, dim3((unsigned) (n_ / (2 * EW_THREADS)), limbs), EW_THREADS, 0, st, result, operand,
               (const uint32_t *) d_perm_[gi].p, n_);
    check_launch("k_galois_ntt");
}

// add_plain_inplace / sub_plain_inplace (evaluate.cu:1106-1224) and the last step of encrypt_symmetric / encrypt_asymmetric
// (secretkey.cu:130-190, 463-530): c0 +-= plaintext.
//   BFV   multiply_add / multiply_sub_plain_with_scaling_variant (scalingvariant.cu:10-60): plain = [n] mod t, scaled by Q_l / t
//   CKKS  plain = [l][n] in NTT form
//   BGV   plain = [n] mod t, lifted to every limb, transformed, times the ciphertext's correction factor
void Engine::plain_add(int l, u64 *ct0, const u64 *plain, bool sub, u64 correction_factor, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (l < 1 || l > size_Q_) throw std::invalid_argument("index is invalid!");
    const dim3 grid((unsigned) (n_ / EW_THREADS), l);
    if (scheme_ == Scheme::ckks) {
        elementwise(sub ? EW_SUB : EW_ADD, ct0, plain, ct0, l, st);
        return;
    }
    if (scheme_ == Scheme::bfv) {
        if (!d_tinv_mod_q_.p) {
            std::vector<u64> v(size_Q_);
            for (int j = 0; j < size_Q_; j++) v[j] = hm::invmod(t_ % primes_[j], primes_[j]);
            d_tinv_mod_q_.upload(v);
        }
        u64 q_mod_t = 1 % t_;
        for (int j = 0; j < l; j++) q_mod_t = hm::mulmod(q_mod_t, primes_[j] % t_, t_);
        const u64 neg = (t_ - q_mod_t) % t_;
        if (sub) launch_pdl(k_bfv_add_plain<true>, grid, EW_THREADS, 0, st, ct0, plain, neg, t_, (const u64 *) d_tinv_mod_q_.p, (const Modulus *) d_mod_.p, n_);
        else launch_pdl(k_bfv_add_plain<false>, grid, EW_THREADS, 0, st, ct0, plain, neg, t_, (const u64 *) d_tinv_mod_q_.p, (const Modulus *) d_mod_.p, n_);
        check_launch("k_bfv_add_plain");
        return;
    }
    u64 *lifted = ws_.tmp.p;
    launch_pdl(k_lift_plain<false>, grid, EW_THREADS, 0, st, lifted, plain, (const Modulus *) d_mod_.p, n_, t_);
    check_launch("k_lift_plain");
    ntt_fwd_rows_range(lifted, l, 0, st);
    if (sub) launch_pdl(k_axpy<true>, grid, EW_THREADS, 0, st, ct0, (const u64 *) ct0, (const u64 *) lifted, correction_factor, (const Modulus *) d_mod_.p, n_);
    else launch_pdl(k_axpy<false>, grid, EW_THREADS, 0, st, ct0, (const u64 *) ct0, (const u64 *) lifted, correction_factor, (const Modulus *) d_mod_.p, n_);
    check_launch("k_axpy");
}

// multiply_plain_inplace (evaluate.cu:1226-1340): every polynomial of ct times the plaintext.  CKKS: plain = [l][n] NTT form;
// BGV: plain lifted and transformed; BFV (multiply_plain_normal): plain lifted with the upper half moved to negative
// residues, ciphertext to NTT form and back.  ct = [size][l][n]
void Engine::plain_multiply(int l, u64 *ct, int size, const u64 *plain, cudaStream_t st) {
    Workspace &ws_ = ws(st);
    if (l < 1 || l > size_Q_) throw std::invalid_argument("index is invalid!");
    if (size < 1) throw std::invalid_argument("ciphertext is empty");
    const dim3 grid((unsigned) (n_ / EW_THREADS), l);
    const u64 *factor = plain;
    if (scheme_ != Scheme::ckks) {
        u64 *lifted = ws_.tmp.p;
        if (scheme_ == Scheme::bfv) launch_pdl(k_lift_plain<true>, grid, EW_THREADS, 0, st, lifted, plain, (const Modulus *) d_mod_.p, n_, t_);
        else launch_pdl(k_lift_plain<false>, grid, EW_THREADS, 0, st, lifted, plain, (const Modulus *) d_mod_.p, n_, t_);
        check_launch("k_lift_plain");
        ntt_fwd_rows_range(lifted, l, 0, st);
        factor = lifted;
    }
    if (scheme_ == Scheme::bfv) ntt_batch(ct, size, l, 0, false, st);
    for (int k = 0; k < size; k++) elementwise(EW_MUL, ct + (size_t) k * l * n_, factor, ct + (size_t) k * l * n_, l, st);
    if (scheme_ == Scheme::bfv) ntt_batch(ct, size, l, 0, true, st);
}

// multiply_scalar_rns_poly (polymath.cu:210-228) over `count` limbs starting at modulus row 0 of each polynomial:
// the BGV correction-factor balancing of add / sub (evaluate.cu:148-165)
void Engine::multiply_scalar(int l, u64 *inout, int size, u64 scalar, cudaStream_t st) const {
    if (l < 1 || l > size_QP_) throw std::invalid_argument("limb count out of range");
    for (int k = 0; k < size; k++) {
        u64 *pk = inout + (size_t) k * l * n_;
        launch_pdl(k_axpy<false>, dim3((unsigned) (n_ / EW_THREADS), l), EW_THREADS, 0, st, pk, (const u64 *) nullptr, (const u64 *) pk, scalar,
                   (const Modulus *) d_mod_.p, n_);
        check_launch("k_axpy");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The reference's kernel-level launchers with the reference's own addressing (cut line 2): thin forms over
// the LimbList machinery above, for a maintainer who keeps evaluate.cu / rns.cu untouched and swaps the launchers only.
// ---------------------------------------------------------------------------------------------------------------------
// `table` = family | chain_index << 8 (chain_index 0 = the first data level): the Bsk and QlRl tables belong to a level's
// DRNSTool in the reference (rns.cu:435-448, 700-714)
int Engine::table_size(int table) const {
    const int fam = table & 0xff, ci = table >> 8;
    const int l = limbs_at(ci ? (size_t) ci : 1);
    switch (fam) {
        case TABLE_RNS: return size_QP_;
        case TABLE_BSK: {
            if (!naux_) return 0;
            std::vector<u64> ql(primes_.begin(), primes_.begin() + l);
            const int tb = 64 - __builtin_clzll(t_);
            return l + ((32 + tb + hm::product_bits(ql) >= 61 * l + 61) ? 1 : 0) + 1;   // base_B_size + m_sk (rns.cu:400-406)
        }
        case TABLE_QLRL: return nR_ ? l + std::min(nR_, l + 1) : 0;
        case TABLE_PLAIN: return (t_ > 1 && batching_) ? 1 : 0;
        default: throw std::invalid_argument("unknown table family");
    }
}

int Engine::table_row(int table, size_t idx) const {
    const int size = table_size(table);
    if ((long) idx >= (long) size) throw std::invalid_argument("modulus index out of range");
    const int fam = table & 0xff, ci = table >> 8;
    const int l = limbs_at(ci ? (size_t) ci : 1);
    switch (fam) {
        case TABLE_RNS: return (int) idx;
        // gpu_Bsk_tables: B_0 .. B_{nB-1}, then m_sk; stored here as [m_sk, B...]
        case TABLE_BSK: return (int) idx == size - 1 ? row_aux_ : row_aux_ + 1 + (int) idx;
        // gpu_QlRl_tables: the level's Q rows, then R rows
        case TABLE_QLRL: return (int) idx < l ? (int) idx : row_R_ + ((int) idx - l);
        default: return size_QP_;   // the plain modulus
    }
}

void Engine::ntt_call(u64 *out, const u64 *in, const NttCall &c, const u64 *scale, const u64 *scale_shoup, cudaStream_t st) {
    if (c.count == 0) return;
    if ((scale == nullptr) != (scale_shoup == nullptr)) throw std::invalid_argument("scale and scale_shoup go together");
    if (scale && !c.inverse) throw std::invalid_argument("only the inverse transforms take a scale");
    LimbVec v;
    std::vector<int> sidx;   // index into scale[] per slot
    for (size_t i = 0; i < c.count; i++) {
        const size_t idx = c.start + i;
        if (idx >= c.excl_lo && idx < c.excl_hi) continue;
        size_t entry = idx;
        if (c.fixed_entry >= 0) entry = (size_t) c.fixed_entry;
        else if (c.remap == 1) {   // fntt_2d.cu:433-436
            if (c.b > c.count || c.start + c.count > c.a + c.count) throw std::invalid_argument("modulus index out of range");
            if (idx >= c.start + c.count - c.b) entry = c.a - (c.start + c.count - idx);
        } else if (c.remap == 2) {   // fntt_2d.cu:225-226: the last limb of the window uses the last table entry
            if (idx == c.count - 1) entry = c.a - 1;
        }
        if (idx > 32000) throw std::invalid_argument("modulus index out of range");
        v.push((int) idx, table_row(c.table, entry));
        sidx.push_back((int) (c.remap == 2 ? entry : idx));
    }
    for (size_t b = 0; b < v.size();) {
        size_t taken;
        const LimbList ll = v.chunk(b, taken, rowq_);
        if (!c.inverse) {
            ntt_fwd_list(out, in, ll, st);
        } else if (!scale) {
            ntt_inv_list(out, in, ll, nullptr, 0, st);
        } else {
            // constants of the last inverse stage from the caller's device arrays: {n^-1 s, itw[1] n^-1 s} per slot
            Workspace &w = ws(st);
            if (w.fin.count < 2 * NTT_MAX_LIMBS) w.fin.alloc(2 * NTT_MAX_LIMBS);
            FinList fl{};
            fl.count = ll.count;
            for (int k = 0; k < ll.count; k++) fl.row[k] = ll.row[k], fl.idx[k] = (short) sidx[b + k];
            launch_pdl(k_make_fin, dim3(1), dim3(NTT_MAX_LIMBS), 0, st, w.fin.p, scale, fl, (const u64 *) d_fin_int_.p,
                       (const Modulus *) d_mod_.p, (const unsigned char *) d_is_fp_.p, plan_.fp_enabled);
            check_launch("k_make_fin");
            ntt_inv_list(out, in, ll, w.fin.p, 1, st);
        }
        b += taken;
    }
}

void Engine::ntt_fuse_moddown(u64 *ct, const u64 *cx, const u64 *pinv, const u64 *pinv_shoup, u64 *delta, size_t count,
                              size_t start, cudaStream_t st) {
    if (count == 0) return;
    if (start + count > (size_t) size_QP_) throw std::invalid_argument("modulus index out of range");
    LimbVec v;
    for (size_t i = 0; i < count; i++) v.push((int) (start + i), (int) (start + i));
    Workspace &w = ws(st);
    if (w.fin.count < 2 * NTT_MAX_LIMBS) w.fin.alloc(2 * NTT_MAX_LIMBS);
    for (size_t b = 0; b < v.size();) {
        size_t taken;
        const LimbList ll = v.chunk(b, taken, rowq_);
        FinList fl{};
        fl.count = ll.count;
        for (int k = 0; k < ll.count; k++) fl.row[k] = ll.row[k], fl.idx[k] = ll.data[k];
        launch_pdl(k_zip_tw, dim3(1), dim3(NTT_MAX_LIMBS), 0, st, w.fin.p, pinv, pinv_shoup, fl);
        check_launch("k_zip_tw");
        EpiArgs ea{};
        ea.sub_base = cx, ea.out_base = ct, ea.add_base = nullptr, ea.mulc = w.fin.p;
        for (int k = 0; k < ll.count; k++) ea.sub[k] = ea.out[k] = ll.data[k], ea.add[k] = -1;
        g_launches.fetch_add(2, std::memory_order_relaxed);
        PFHE_CUDA(ntt_forward_epilogue(plan_, delta, ll, ea, st));
        b += taken;
    }
}

const Engine::Converter &Engine::converter(const std::vector<int> &in_rows, const std::vector<int> &out_rows) {
    std::lock_guard<std::recursive_mutex> g(table_mu_);
    auto &slot = converters_[{in_rows, out_rows}];
    if (slot) return *slot;
    const int ni = (int) in_rows.size(), no = (int) out_rows.size();
    if (ni < 1 || no < 1 || ni > BEHZ_MAX_LIMBS || no > BEHZ_MAX_LIMBS) throw std::invalid_argument("base size is not supported");
    std::vector<u64> ib, ob;
    for (int r : in_rows) {
        if (r < 0 || r >= mod_rows_) throw std::invalid_argument("modulus index out of range");
        ib.push_back(rowq_[r]);
    }
    for (int r : out_rows) {
        if (r < 0 || r >= mod_rows_) throw std::invalid_argument("modulus index out of range");
        ob.push_back(rowq_[r]);
    }
    auto cv = std::make_unique<Converter>();
    cv->in_rows = in_rows, cv->out_rows = out_rows;
    std::vector<Modulus> mi, mo;
    for (u64 q : ib) mi.push_back(host_modulus(q));
    for (u64 q : ob) mo.push_back(host_modulus(q));
    std::vector<Tw> hinv(ni), v1c(ni);
    std::vector<double> inv(ni);
    std::vector<u64> mat((size_t) no * ni), Imod(no), v1m((size_t) no * ni);
    for (int i = 0; i < ni; i++) {
        const u64 q = ib[i];
        const u64 hat_inv = hm::invmod(hm::product_mod(ib, i, q), q);
        hinv[i] = make_tw(hat_inv, q);
        inv[i] = 1.0 / (double) q;
        // negPQHatInvModq (host/rns.cu:478-483): -(O mod q_i) * ihat_i^-1 mod q_i, stored unreduced as q - x (may equal q)
        const u64 PQ = hm::mulmod(hm::product_mod(ob, -1, q), hat_inv, q);
        v1c[i] = make_ulonglong2(q - PQ, 0);
    }
    for (int j = 0; j < no; j++) {
        const u64 p = ob[j];
        Imod[j] = hm::product_mod(ib, -1, p);
        for (int i = 0; i < ni; i++) {
            mat[(size_t) j * ni + i] = hm::product_mod(ib, i, p);
            v1m[(size_t) j * ni + i] = (ib[i] % p) ? hm::invmod(ib[i] % p, p) : 0;
        }
    }
    cv->mod_in.upload(mi), cv->mod_out.upload(mo), cv->hinv.upload(hinv), cv->inv.upload(inv), cv->mat.upload(mat);
    cv->Imod.upload(Imod), cv->v1_c.upload(v1c), cv->v1_mat.upload(v1m);
    slot = std::move(cv);
    return *slot;
}

void Engine::bconv(int mode, const Converter &cv, u64 *dst, const u64 *src, cudaStream_t st) {
    const int ni = (int) cv.in_rows.size(), no = (int) cv.out_rows.size();
    const dim3 grid((unsigned) (n_ / BEHZ_THREADS), 1);
    const PolyView in{const_cast<u64 *>(src), 0}, out{dst, 0};
    if (mode == 0) {          // bConv_BEHZ: y_i = x_i ihat_i^-1, out_j = sum y_i (ihat_i mod o_j)
        BconvVar1Args a{in, out, cv.hinv.p, cv.mat.p, cv.mod_in.p, cv.mod_out.p, ni, no, n_};
        launch_pdl(k_bconv_var1, grid, BEHZ_THREADS, 0, st, a);
    } else if (mode == 1) {   // bConv_BEHZ_var1
        BconvVar1Args a{in, out, cv.v1_c.p, cv.v1_mat.p, cv.mod_in.p, cv.mod_out.p, ni, no, n_};
        launch_pdl(k_bconv_var1, grid, BEHZ_THREADS, 0, st, a);
    } else if (mode == 2) {   // bConv_HPS
        BconvHpsArgs a{in, out, cv.hinv.p, cv.inv.p, cv.mat.p, cv.Imod.p, cv.mod_in.p, cv.mod_out.p, ni, no, n_};
        launch_pdl(k_bconv_hps, grid, BEHZ_THREADS, 0, st, a);
    } else {
        throw std::invalid_argument("unknown base conversion");
    }
    check_launch("k_bconv (boundary)");
}

void Engine::moddown_plain(int l, u64 *ct_i, u64 *cx_i, cudaStream_t st) {
    if (scheme_ != Scheme::bfv) {   // CKKS / BGV: the same result as moddown_from_NTT (rns_bconv.cu:720-753)
        if (scheme_ == Scheme::ckks) moddown(l, ct_i, cx_i, ws(st).delta.p, 1, nullptr, 0u, st);
        else moddown_generic(l, ct_i, cx_i, 1, nullptr, 0u, st);
        return;
    }
    const Level &lv = level(l);
    if (lv.alpha == 0) throw std::logic_error("key switching needs special primes");
    std::vector<int> pr, qr;
    for (int i = 0; i < lv.alpha; i++) pr.push_back(size_Q_ + i);
    for (int j = 0; j < l; j++) qr.push_back(j);
    u64 *delta = ws(st).delta.p;
    bconv(0, converter(pr, qr), delta, cx_i + (size_t) l * n_, st);
    dim3 grid((unsigned) (n_ / (2 * EW_THREADS)), l);
    launch_pdl(k_moddown_coeff, grid, EW_THREADS, 0, st, ct_i, (const u64 *) cx_i, (const u64 *) delta, (const Tw *) lv.pinv_slots.p,
               (const u64 *) nullptr, (const Modulus *) d_mod_.p, n_);
    check_launch("k_moddown_coeff");
}

} // namespace pfhe
