// ntt_kernels.cu -- kernels and launchers built on the pass drivers of ntt.cuh.
#include "ntt_api.cuh"
#include "launch.hpp"

namespace pfhe {

// every kernel picks the arithmetic of its limb (CTA-uniform): FP64 butterflies for q < 2^46, integer otherwise
#define PFHE_ARITH_DISPATCH(ROW, ...)                                                                     \
    const u64 q = ll.q[slot];                                                                             \
    if (p.fp_enabled && (q >> fp::MAX_BITS) == 0) {                                                       \
        using A = FpArith;                                                                                \
        __VA_ARGS__                                                                                       \
    } else {                                                                                              \
        using A = IntArith;                                                                               \
        __VA_ARGS__                                                                                       \
    }

// dynamic shared memory of every NTT kernel: [exchange tile | staged twiddles | mbarrier]
constexpr size_t NTT_DYN_SMEM = NTT_TILE * sizeof(u64) + NTT_STW_ENTRIES * sizeof(Tw) + 16;

#define PFHE_NTT_SMEM(P, ROWS, TABLE)                                                                     \
    extern __shared__ __align__(128) unsigned char dyn_smem[];                                            \
    u64 *smem = reinterpret_cast<u64 *>(dyn_smem);                                                        \
    Tw *stw = reinterpret_cast<Tw *>(smem + NTT_TILE);                                                    \
    uint64_t *bar = reinterpret_cast<uint64_t *>(stw + NTT_STW_ENTRIES);                                  \
    if (threadIdx.x == 0) mbar_init(bar, 1);                                                              \
    __syncthreads();                                                                                      \
    pdl_launch_dependents();                                                                              \
    if (threadIdx.x == 0) stage_twiddles<P, ROWS, LOGN>(stw, (TABLE) + ((size_t) row << LOGN), (int) blockIdx.x, bar); \
    pdl_wait();

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, 4) k_fwd_cols(u64 *dst, const u64 *src, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM(ntt_p1(LOGN), false, p.tw)
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                smem, cx, [&](size_t i) { return A::load(s[i], c); }, [&](size_t i, typename A::T v) { d[i] = A::raw(v); });
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, 4) k_fwd_rows(u64 *data, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM(ntt_p2(LOGN), true, p.tw)
    u64 *d = data + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                smem, cx, [&](size_t i) { return A::from_raw(d[i]); },
                [&](size_t i, typename A::T v) { d[i] = A::canon_fwd(v, c); });
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, 4) k_fwd_rows_epi(u64 *data, LimbList ll, NttPlan p, EpiArgs ea) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM(ntt_p2(LOGN), true, p.tw)
    const u64 *d = data + ((size_t) ll.data[slot] << LOGN);
    const u64 *sub = ea.sub_base + ((size_t) ea.sub[slot] << LOGN);
    u64 *out = ea.out_base + ((size_t) ea.out[slot] << LOGN);
    const int addl = ea.add[slot];
    const u64 *add = addl >= 0 ? ea.add_base + ((size_t) addl << LOGN) : nullptr;
    const Tw k = ea.mulc[slot];
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                smem, cx, [&](size_t i) { return A::from_raw(d[i]); },
                [&](size_t i, typename A::T v) {
                    const u64 t = A::canon_fwd(v, c);
                    u64 r = mul_shoup(sub[i] + q - t, k, q);
                    if (add) r = add_mod(r, add[i], q);
                    out[i] = r;
                });
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, 4) k_inv_rows(u64 *dst, const u64 *src, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM(ntt_p2(LOGN), true, p.itw)
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        inverse_pass<A, ntt_p2(LOGN), true, LOGN, false>(
                smem, cx, [&](size_t i) { return A::load(s[i], c); }, [&](size_t i, typename A::T v) { d[i] = A::raw(v); });
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, 4) k_inv_cols(u64 *data, LimbList ll, NttPlan p, const Tw *fin,
                                                              int fin_by_slot) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM(ntt_p1(LOGN), false, p.itw)
    const int f = fin_by_slot ? slot : row;
    u64 *d = data + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, fin[2 * f], fin[2 * f + 1]};
        inverse_pass<A, ntt_p1(LOGN), false, LOGN, true>(
                smem, cx, [&](size_t i) { return A::from_raw(d[i]); },
                [&](size_t i, typename A::T v) { d[i] = A::canon_inv(v, c); });
    })
}

template<class K>
static void opt_in_smem(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) NTT_DYN_SMEM);
}
template<int LOGN>
static void opt_in_all() {
    static bool done = false;
    if (done) return;
    opt_in_smem(k_fwd_cols<LOGN>);
    opt_in_smem(k_fwd_rows<LOGN>);
    opt_in_smem(k_fwd_rows_epi<LOGN>);
    opt_in_smem(k_inv_rows<LOGN>);
    opt_in_smem(k_inv_cols<LOGN>);
    done = true;
}

template<int LOGN>
static void fwd_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_fwd_cols<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, dst, src, ll, p);
    launch_pdl(k_fwd_rows<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, dst, ll, p);
}

template<int LOGN>
static void fwd_epi_impl(const NttPlan &p, u64 *data, const LimbList &ll, const EpiArgs &ea, cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_fwd_cols<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, data, data, ll, p);
    launch_pdl(k_fwd_rows_epi<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, data, ll, p, ea);
}

template<int LOGN>
static void inv_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                     cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_inv_rows<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, dst, src, ll, p);
    launch_pdl(k_inv_cols<LOGN>, grid, NTT_THREADS, NTT_DYN_SMEM, st, dst, ll, p, fin ? fin : p.inv_fin, fin ? by_slot : 0);
}

#define PFHE_DISPATCH_LOGN(FN, ...)                                                                       \
    switch (p.logn) {                                                                                     \
        case 12: FN<12>(__VA_ARGS__); break;                                                              \
        case 13: FN<13>(__VA_ARGS__); break;                                                              \
        case 14: FN<14>(__VA_ARGS__); break;                                                              \
        case 15: FN<15>(__VA_ARGS__); break;                                                              \
        case 16: FN<16>(__VA_ARGS__); break;                                                              \
        case 17: FN<17>(__VA_ARGS__); break;                                                              \
        default: return cudaErrorInvalidValue;                                                            \
    }

cudaError_t ntt_forward(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) {
    if (ll.count == 0) return cudaSuccess;
    PFHE_DISPATCH_LOGN(fwd_impl, p, dst, src, ll, st)
    return cudaGetLastError();
}

cudaError_t ntt_forward_epilogue(const NttPlan &p, u64 *data, const LimbList &ll, const EpiArgs &ea, cudaStream_t st) {
    if (ll.count == 0) return cudaSuccess;
    PFHE_DISPATCH_LOGN(fwd_epi_impl, p, data, ll, ea, st)
    return cudaGetLastError();
}

cudaError_t ntt_inverse(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                        cudaStream_t st) {
    if (ll.count == 0) return cudaSuccess;
    PFHE_DISPATCH_LOGN(inv_impl, p, dst, src, ll, fin, by_slot, st)
    return cudaGetLastError();
}

} // namespace pfhe
