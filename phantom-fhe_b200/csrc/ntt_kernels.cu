// ntt_kernels.cu -- kernels and launchers built on the pass drivers of ntt.cuh.
#include <algorithm>
#include <cstdlib>

#include "ntt_api.cuh"
#include "launch.hpp"

#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 4
#endif

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

// dynamic shared memory: column passes [exchange tile | staged twiddles | mbarrier], row passes [exchange tile]
constexpr size_t NTT_SMEM_COLS = NTT_TILE * sizeof(u64) + NTT_STW_ENTRIES * sizeof(Tw) + 16;
constexpr size_t NTT_SMEM_ROWS = NTT_TILE * sizeof(u64);

#define PFHE_NTT_SMEM_COLS(P, TABLE)                                                                      \
    extern __shared__ __align__(128) unsigned char dyn_smem[];                                            \
    u64 *smem = reinterpret_cast<u64 *>(dyn_smem);                                                        \
    Tw *stw = reinterpret_cast<Tw *>(smem + NTT_TILE);                                                    \
    uint64_t *bar = reinterpret_cast<uint64_t *>(stw + NTT_STW_ENTRIES);                                  \
    if (threadIdx.x == 0) mbar_init(bar, 1);                                                              \
    __syncthreads();                                                                                      \
    pdl_launch_dependents();                                                                              \
    if (threadIdx.x == 0) stage_twiddles<P>(stw, (TABLE) + ((size_t) row << LOGN), bar);                  \
    pdl_wait();

#define PFHE_NTT_SMEM_ROWS(TABLE)                                                                         \
    extern __shared__ __align__(128) unsigned char dyn_smem[];                                            \
    u64 *smem = reinterpret_cast<u64 *>(dyn_smem);                                                        \
    const Tw *stw = (TABLE) + ((size_t) row << LOGN);                                                     \
    uint64_t *bar = nullptr;                                                                              \
    pdl_launch_dependents();                                                                              \
    pdl_wait();

// epilogue of the fused forward row pass, all operand loads of a thread's eight outputs issued up front (16-byte
// loads: the row pass hands over runs of consecutive coefficients).  out = (sub - NTT) * k (+ add) mod q.
// FP64 limbs stay in FP64: v = sub - x (x is the lazy transform value), one error-free product by the constant, the
// addend, one reduction -- the same canonical residue as the integer form at half the instructions.
template<class A>
__device__ __forceinline__ u64 epi_one(u64 s, typename A::T x, bool has_add, u64 a, Tw k, double kd, double ki,
                                       const typename A::Consts &c, u64 q) {
    if constexpr (std::is_same<A, FpArith>::value) {
        double r = fp::mulmod_c(fp::from_u64(s) - x, kd, ki, c.q);
        if (has_add) r = fp::reduce(r + fp::from_u64(a), c.q, c.qinv);
        return fp::canon(r, c.q);
    } else {
        const u64 t = A::canon_fwd(x, c);
        u64 r = mul_shoup(s + q - t, k, q);
        if (has_add) r = add_mod(r, a, q);
        return r;
    }
}

template<class A>
struct EpiStore {
    const u64 *sub, *add;
    u64 *out;
    Tw k;
    typename A::Consts c;
    u64 q;
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const typename A::T (&x)[NTT_EPT]) const {
        static_assert(RUN >= 2, "the epilogue belongs to the row pass");
        u64 s[NTT_EPT], a[NTT_EPT], r[NTT_EPT];
        ld_runs<RUN>(sub, idx, s);
        if (add) ld_runs<RUN>(add, idx, a);
        double kd = 0, ki = 0;
        if constexpr (std::is_same<A, FpArith>::value) kd = (double) k.x, ki = kd * c.qinv;
#pragma unroll
        for (int i = 0; i < NTT_EPT; i++) r[i] = epi_one<A>(s[i], x[i], add != nullptr, add ? a[i] : 0, k, kd, ki, c, q);
        st_runs<RUN>(out, idx, r);
    }
};

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_cols(u64 *dst, const u64 *src, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_COLS(ntt_p1(LOGN), p.tw)
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::load(s[i], c); }),
                per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
    })
}

// LAZY: FP64 limbs are stored as they leave the butterflies (exact integers in FP64 form, |v| < 12 q), for consumers
// inside the engine that take them in that form (the key-switch inner product).  Integer limbs are always canonical.
template<int LOGN, bool LAZY>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_rows(u64 *data, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_ROWS(p.tw)
    u64 *d = data + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(d[i]); }),
                vec_store<typename A::T>(d, [&](typename A::T v) {
                    if constexpr (LAZY && std::is_same<A, FpArith>::value) return A::raw(v);
                    else return A::canon_fwd(v, c);
                }));
    })
}

// ---- both passes in one persistent launch -------------------------------------------------------------------------
// Work items in ticket order: column tiles of slot k, then row tiles of slot k - delay (so that by the time a row tile is
// drawn the column tiles it depends on were drawn ~delay limbs earlier and are normally finished).  A row tile waits on
// ready[slot] == tiles-per-pass; it can only wait for lower tickets, which running CTAs hold: no deadlock whatever the
// residency.  The intermediate goes through L2: column tiles publish with fence + atomic, row tiles read with ld.cg (the
// transform is in place: the SM's L1 may still hold the pre-transform words of the same addresses).
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void publish_tile(unsigned *ready, int slot) {   // thread 0, after a CTA barrier
    __threadfence();   // release: cumulative over the stores the barrier made visible to this thread
    atomicAdd(ready + slot, 1u);
}

// stores of a column tile; just before them thread 0 publishes the PREVIOUS column tile of this CTA: by now those stores
// have long been acknowledged, so the fence costs nothing -- publishing right after a tile's own stores would stall the
// whole CTA for the store round trip, which a non-persistent CTA never waits for
template<class A>
struct ColStorePublish {
    u64 *d;
    unsigned *ready;
    int pending;
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const typename A::T (&x)[NTT_EPT]) const {
        if (threadIdx.x == 0 && pending >= 0) publish_tile(ready, pending);
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) d[idx[k]] = A::raw(x[k]);
    }
};

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_fused(u64 *dst, const u64 *src, LimbList ll, NttPlan p,
                                                                           FusedSync *sy, int delay) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    u64 *smem = reinterpret_cast<u64 *>(dyn_smem);
    Tw *stw = reinterpret_cast<Tw *>(smem + NTT_TILE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stw + NTT_STW_ENTRIES);
    __shared__ unsigned s_item;
    constexpr unsigned T = 1u << (LOGN - NTT_LOG_TILE);
    const unsigned count = (unsigned) ll.count, D = min((unsigned) delay, count);
    const unsigned total = count * 2 * T, head = D * T, mid = (count - D) * 2 * T;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    uint32_t parity = 0;
    // thread 0 draws the NEXT ticket while the current tile is being worked on: the atomic's round trip is off the path
    unsigned next = 0;
    int pending = -1;   // slot of this CTA's last column tile, not yet counted in ready[]
    if (threadIdx.x == 0) next = atomicAdd(&sy->ticket, 1u);
    for (;;) {
        if (threadIdx.x == 0) s_item = next;
        __syncthreads();
        const unsigned w = s_item;
        if (w >= total) break;
        if (threadIdx.x == 0) next = atomicAdd(&sy->ticket, 1u);
        unsigned slot, tile;
        bool rows;
        if (w < head) {
            rows = false, slot = w / T, tile = w % T;
        } else if (w < head + mid) {
            const unsigned v = w - head, k = D + v / (2 * T), r = v % (2 * T);
            rows = r >= T, slot = rows ? k - D : k, tile = rows ? r - T : r;
        } else {
            const unsigned v = w - head - mid;
            rows = true, slot = count - D + v / T, tile = v % T;
        }
        const int row = ll.row[slot];
        const u64 q = ll.q[slot];
        const bool fp = p.fp_enabled && (q >> fp::MAX_BITS) == 0;
        u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
        if (!rows) {
            if (threadIdx.x == 0) stage_twiddles<ntt_p1(LOGN)>(stw, p.tw + ((size_t) row << LOGN), bar);
            const u64 *sp = src + ((size_t) ll.src[slot] << LOGN);
            if (fp) {
                using A = FpArith;
                const A::Consts c = A::consts(q);
                PassCtx<A> cx{stw, bar, c, (int) tile, {}, {}, parity};
                forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                        smem, cx, per_elem_load<A::T>([&](size_t i) { return A::load(sp[i], c); }),
                        ColStorePublish<A>{d, sy->ready, pending});
            } else {
                using A = IntArith;
                const A::Consts c = A::consts(q);
                PassCtx<A> cx{stw, bar, c, (int) tile, {}, {}, parity};
                forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                        smem, cx, per_elem_load<A::T>([&](size_t i) { return A::load(sp[i], c); }),
                        ColStorePublish<A>{d, sy->ready, pending});
            }
            parity ^= 1u;
            pending = (int) slot;
            __syncthreads();   // every thread's stores are ordered before whatever thread 0 publishes later
        } else {
            if (threadIdx.x == 0) {
                if (pending >= 0) publish_tile(sy->ready, pending);   // never wait on a count this CTA still holds back
                while (ld_acquire_u32(&sy->ready[slot]) < T) __nanosleep(64);
            }
            pending = -1;
            __syncthreads();
            const Tw *tw = p.tw + ((size_t) row << LOGN);
            if (fp) {
                using A = FpArith;
                const A::Consts c = A::consts(q);
                PassCtx<A> cx{tw, nullptr, c, (int) tile, {}, {}};
                forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                        smem, cx, per_elem_load<A::T>([&](size_t i) { return A::from_raw(__ldcg(d + i)); }),
                        vec_store<A::T>(d, [&](A::T v) { return A::canon_fwd(v, c); }));
            } else {
                using A = IntArith;
                const A::Consts c = A::consts(q);
                PassCtx<A> cx{tw, nullptr, c, (int) tile, {}, {}};
                forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                        smem, cx, per_elem_load<A::T>([&](size_t i) { return A::from_raw(__ldcg(d + i)); }),
                        vec_store<A::T>(d, [&](A::T v) { return A::canon_fwd(v, c); }));
            }
            __syncthreads();   // the exchange tile and s_item are free again
        }
    }
    // the last CTA to leave restores the counters for the next launch on this stream
    if (threadIdx.x == 0) {
        if (pending >= 0) publish_tile(sy->ready, pending);
        __threadfence();
        if (atomicAdd(&sy->done, 1u) == gridDim.x - 1) {
            for (unsigned i = 0; i < count; i++) sy->ready[i] = 0;
            sy->ticket = 0;
            sy->done = 0;
            __threadfence();
        }
    }
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_rows_epi(u64 *data, LimbList ll, NttPlan p, EpiArgs ea) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_ROWS(p.tw)
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
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(d[i]); }),
                EpiStore<A>{sub, add, out, k, c, q});
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_inv_rows(u64 *dst, const u64 *src, LimbList ll, NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_ROWS(p.itw)
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        inverse_pass<A, ntt_p2(LOGN), true, LOGN, false>(
                smem, cx, vec_load<typename A::T>(s, [&](u64 v) { return A::load(v, c); }),
                per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
    })
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_inv_cols(u64 *data, LimbList ll, NttPlan p, const Tw *fin,
                                                              int fin_by_slot) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_COLS(ntt_p1(LOGN), p.itw)
    const int f = fin_by_slot ? slot : row;
    u64 *d = data + ((size_t) ll.data[slot] << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, fin[2 * f], fin[2 * f + 1]};
        inverse_pass<A, ntt_p1(LOGN), false, LOGN, true>(
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(d[i]); }),
                per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::canon_inv(v, c); }));
    })
}

// ---- fused variants -----------------------------------------------------------------------------------
// a * b mod q in the representation of arithmetic A (value entering an inverse pass)
template<class A>
__device__ __forceinline__ typename A::T mul_in(u64 x, u64 y, const typename A::Consts &c, const BarG &bg, u64 q) {
    if constexpr (std::is_same<A, FpArith>::value) {
        return fp::mulmod_v(fp::from_u64(x), fp::from_u64(y), c.q, c.qinv);
    } else {
        const Modulus m{q, 0, 0};
        return mul_mod_g(x, y, bg, m);
    }
}

template<class A>
struct MulLoad {
    const u64 *a1, *b1;
    typename A::Consts c;
    BarG bg;
    u64 q;
    template<int RUN>
    __device__ __forceinline__ void gather(const size_t (&idx)[NTT_EPT], typename A::T (&x)[NTT_EPT]) const {
        static_assert(RUN >= 2, "the product load belongs to the row pass");
        u64 u[NTT_EPT], v[NTT_EPT];
        ld_runs<RUN>(a1, idx, u);
        ld_runs<RUN>(b1, idx, v);
#pragma unroll
        for (int i = 0; i < NTT_EPT; i++) x[i] = mul_in<A>(u[i], v[i], c, bg, q);
    }
};

// first inverse pass of HMult+Relin: loads a1 * b1 (the d2 component of the tensor product)
template<class A, int LOGN>
__device__ __forceinline__ void inv_rows_mul_body(u64 *smem, const Tw *stw, uint64_t *bar, u64 q, const u64 *a1, const u64 *b1,
                                                  u64 *d, const BarG bg) {
    const typename A::Consts c = A::consts(q);
    PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
    inverse_pass<A, ntt_p2(LOGN), true, LOGN, false>(
            smem, cx, MulLoad<A>{a1, b1, c, bg, q},
            per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_inv_rows_mul(u64 *dst, TensorSrc ts, const BarG *bar0, LimbList ll,
                                                                  NttPlan p) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_ROWS(p.itw)
    const size_t poly = (size_t) ts.l << LOGN;
    const u64 *a1 = ts.a + poly + ((size_t) ll.src[slot] << LOGN);
    const u64 *b1 = ts.b + poly + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    const BarG bg = bar0[row];
    const u64 q = ll.q[slot];
    if (p.fp_enabled && (q >> fp::MAX_BITS) == 0) inv_rows_mul_body<FpArith, LOGN>(smem, stw, bar, q, a1, b1, d, bg);
    else inv_rows_mul_body<IntArith, LOGN>(smem, stw, bar, q, a1, b1, d, bg);
}

// ---- few limbs: both passes of a transform in ONE launch of limbs x tiles CTAs ---------------------------------------------
// A transform of a few limbs is two single-wave launches: its time is the latency chain of one tile twice plus the gap
// between the launches.  Here CTA (tile t, slot s) runs its first-pass tile, counts itself into ready[s], waits until the
// N / 2048 tiles of ITS limb are in, and runs second-pass tile t -- no launch boundary, no grid-wide barrier.  The CTAs wait
// for each other, so the launch is cooperative (the runtime guarantees that all of them are resident, also next to another
// such launch of a different lane).  The last CTA to leave restores the counters.
__device__ __forceinline__ void small_sync(FusedSync *sy, int slot, unsigned tiles) {
    __syncthreads();   // every thread's first-pass stores are ordered before thread 0's release
    if (threadIdx.x == 0) {
        publish_tile(sy->ready, slot);
        while (ld_acquire_u32(&sy->ready[slot]) < tiles) __nanosleep(32);
    }
    __syncthreads();
}
__device__ __forceinline__ void small_exit(FusedSync *sy, unsigned count) {
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sy->done, 1u) == gridDim.x * gridDim.y - 1) {
            for (unsigned i = 0; i < count; i++) sy->ready[i] = 0;
            sy->done = 0;
            __threadfence();
        }
    }
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_small(u64 *dst, const u64 *src, LimbList ll, NttPlan p,
                                                                           FusedSync *sy) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_COLS(ntt_p1(LOGN), p.tw)
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    const Tw *tw = p.tw + ((size_t) row << LOGN);
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::load(s[i], c); }),
                per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
        small_sync(sy, slot, gridDim.x);
        PassCtx<A> cr{tw, nullptr, c, (int) blockIdx.x, {}, {}};
        forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
                smem, cr, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(__ldcg(d + i)); }),
                vec_store<typename A::T>(d, [&](typename A::T v) { return A::canon_fwd(v, c); }));
    })
    small_exit(sy, gridDim.y);
}

// inverse: row pass (optionally of the product a1 * b1, MUL) then column pass with the folded constants of the last stage
template<int LOGN, bool MUL>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_inv_small(u64 *dst, const u64 *src, TensorSrc ts, const BarG *bar0,
                                                                           LimbList ll, NttPlan p, const Tw *fin, int fin_by_slot,
                                                                           FusedSync *sy) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_COLS(ntt_p1(LOGN), p.itw)
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    const Tw *tw = p.itw + ((size_t) row << LOGN);
    const int f = fin_by_slot ? slot : row;
    PFHE_ARITH_DISPATCH(row, {
        const typename A::Consts c = A::consts(q);
        PassCtx<A> cr{tw, nullptr, c, (int) blockIdx.x, {}, {}};
        if constexpr (MUL) {
            const size_t poly = (size_t) ts.l << LOGN;
            const u64 *a1 = ts.a + poly + ((size_t) ll.src[slot] << LOGN);
            const u64 *b1 = ts.b + poly + ((size_t) ll.src[slot] << LOGN);
            inverse_pass<A, ntt_p2(LOGN), true, LOGN, false>(
                    smem, cr, MulLoad<A>{a1, b1, c, bar0[row], q},
                    per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
        } else {
            const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
            inverse_pass<A, ntt_p2(LOGN), true, LOGN, false>(
                    smem, cr, vec_load<typename A::T>(s, [&](u64 v) { return A::load(v, c); }),
                    per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::raw(v); }));
        }
        small_sync(sy, slot, gridDim.x);
        PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, fin[2 * f], fin[2 * f + 1]};
        inverse_pass<A, ntt_p1(LOGN), false, LOGN, true>(
                smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(__ldcg(d + i)); }),
                per_elem_store<typename A::T>([&](size_t i, typename A::T v) { d[i] = A::canon_inv(v, c); }));
    })
    small_exit(sy, gridDim.y);
}

// base conversion as the gather of a column pass: inputs outer, elements inner, so the eight loads of one input
// limb are in flight together and the uniform "split this input" branch sits outside the element loop.
//
// FP-limb outputs (p < 2^46) split the work over two pipes: with terms a_t (an input residue, or its 30-bit halves
// when it comes from a modulus >= 2^46) and matrix entries M_t < p, the quotient K = rint(sum a_t * (M_t / p)) is
// formed on the FP64 pipe (one conversion and one FMA per term; absolute error < 0.2, proof in DESIGN.md 4.2), the
// remainder r = sum a_t * M_t - K * p, |r| < 0.7 p, as the low 64 bits of that expression on the integer pipe
// (3 IMADs per 64-bit term, 2 per half).  r enters the butterflies as a signed FP64 integer; the canonical residues
// stored at the end of the transform are the same as for any other exact evaluation of bconv_matmul.
#ifndef PFHE_BCONV_DUAL
#define PFHE_BCONV_DUAL 0
#endif
struct Lo64 {   // low 64 bits of a sum of products; the cross terms go straight into the upper word
    u32 lo, hi;
    __device__ __forceinline__ void mac(u64 a, u64 b) {   // full 64-bit a
        const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
        u64 acc = ((u64) hi << 32) | lo;
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a0), "r"(b0));
        lo = (u32) acc;
        hi = a1 * b0 + (a0 * b1 + (u32) (acc >> 32));
    }
    __device__ __forceinline__ void mac32(u32 a0, u64 b) {   // a < 2^32
        const u32 b0 = (u32) b, b1 = (u32) (b >> 32);
        u64 acc = ((u64) hi << 32) | lo;
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a0), "r"(b0));
        lo = (u32) acc;
        hi = a0 * b1 + (u32) (acc >> 32);
    }
    __device__ __forceinline__ u64 value() const { return ((u64) hi << 32) | lo; }
};

template<int LOGN, int MAXIN>
struct BconvGatherFp {
    const u64 *in;
    const double2 *mf;   // [ni][2], global memory: read where used (CTA-uniform, L1 broadcast) to keep registers free
    int ni;
    unsigned big;
    double q;
    template<int RUN>
    __device__ __forceinline__ void gather(const size_t (&idx)[NTT_EPT], double (&x)[NTT_EPT]) const {
#if PFHE_BCONV_DUAL
        constexpr int HB = NTT_EPT / 2;   // two half-batches: accumulators of four elements live at a time
        const u64 np = 0 - (u64) q;
        // bits(2^52 + K) carry 0x43300000 in the upper word: its product with the low word of -p is taken out here;
        // the accumulator starts at bits(1.5 * 2^52) so that the signed remainder converts with one subtraction
        const u32 h0 = 0x43380000u - 0x43300000u * (u32) np;
#pragma unroll
        for (int h = 0; h < NTT_EPT; h += HB) {
            double s[HB];
            Lo64 r[HB];
#pragma unroll
            for (int k = 0; k < HB; k++) s[k] = 0.0, r[k].lo = 0u, r[k].hi = h0;
#pragma unroll
            for (int i = 0; i < MAXIN; i++) {
                if (i < ni) {
                    u64 y[HB];
#pragma unroll
                    for (int k = 0; k < HB; k++) y[k] = in[((size_t) i << LOGN) + idx[h + k]];
                    const double2 m0 = __ldg(&mf[2 * i]);
                    const u64 M0 = (u64) m0.x;
                    if ((big >> i) & 1) {
                        const double2 m1 = __ldg(&mf[2 * i + 1]);
                        const u64 M1 = (u64) m1.x;
                        const u32 msk = (1u << fp::SPLIT_BITS) - 1;
#pragma unroll
                        for (int k = 0; k < HB; k++) {
                            const u32 ylo = (u32) y[k] & msk, yhi = (u32) (y[k] >> fp::SPLIT_BITS);
                            s[k] = __fma_rn(fp::from_u64(ylo), m0.y, s[k]);
                            s[k] = __fma_rn(fp::from_u64(yhi), m1.y, s[k]);
                            r[k].mac32(ylo, M0);
                            r[k].mac32(yhi, M1);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < HB; k++) {
                            s[k] = __fma_rn(fp::from_u64(y[k]), m0.y, s[k]);
                            r[k].mac(y[k], M0);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < HB; k++) {
                const u64 kb = (u64) __double_as_longlong(s[k] + fp::TWO52);   // 2^52 + K, exact integer
                r[k].mac(kb, np);
                x[h + k] = __longlong_as_double((long long) r[k].value()) - fp::MAGIC;
            }
        }
#else
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) x[k] = 0.0;
#pragma unroll
        for (int i = 0; i < MAXIN; i++) {
            if (i < ni) {
                u64 y[NTT_EPT];
#pragma unroll
                for (int k = 0; k < NTT_EPT; k++) y[k] = in[((size_t) i << LOGN) + idx[k]];
                const double2 m0 = __ldg(&mf[2 * i]), m1 = __ldg(&mf[2 * i + 1]);
                if ((big >> i) & 1) {
                    const u64 msk = (1ull << fp::SPLIT_BITS) - 1;
#pragma unroll
                    for (int k = 0; k < NTT_EPT; k++) {
                        x[k] += fp::mulmod_c(fp::from_u64(y[k] & msk), m0.x, m0.y, q);
                        x[k] += fp::mulmod_c(fp::from_u64(y[k] >> fp::SPLIT_BITS), m1.x, m1.y, q);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NTT_EPT; k++) x[k] += fp::mulmod_c(fp::from_u64(y[k]), m0.x, m0.y, q);
                }
            }
        }
#endif
    }
};
template<int LOGN, int MAXIN>
struct BconvGatherInt {
    const u64 *in;
    const u64 *mi;       // [ni] (<= MAXIN)
    int ni;
    BarG bg;
    Modulus m;
    template<int RUN>
    __device__ __forceinline__ void gather(const size_t (&idx)[NTT_EPT], u64 (&x)[NTT_EPT]) const {
        Acc128 acc[NTT_EPT];
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) acc[k] = Acc128{0, 0};
#pragma unroll
        for (int i = 0; i < MAXIN; i++) {
            if (i < ni) {
                u64 y[NTT_EPT];
#pragma unroll
                for (int k = 0; k < NTT_EPT; k++) y[k] = in[((size_t) i << LOGN) + idx[k]];
#pragma unroll
                for (int k = 0; k < NTT_EPT; k++) acc[k].mac(y[k], mi[i]);
            }
        }
#pragma unroll
        for (int k = 0; k < NTT_EPT; k++) x[k] = barrett_g(acc[k].lo, acc[k].hi, bg, m);
    }
};

// forward column pass whose input is produced by the fast base conversion of `ni` coefficient-form limbs
template<class A, int LOGN, int MAXIN>
__device__ __forceinline__ void fwd_cols_bconv_body(u64 *smem, const Tw *stw, uint64_t *bar, u64 q, int row, int slot,
                                                    const u64 *in, u64 *d, const BconvLoad &bl, const Modulus *mod) {
    const typename A::Consts c = A::consts(q);
    PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
    const int ni = bl.ni;
    const unsigned big = bl.in_big[slot];
    if constexpr (std::is_same<A, FpArith>::value) {
        const double2 *mf = bl.matf + 2 * ((size_t) bl.mat_row[slot] * ni);
        forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                smem, cx, BconvGatherFp<LOGN, MAXIN>{in, mf, ni, big, c.q},
                per_elem_store<double>([&](size_t i, double v) { d[i] = A::raw(v); }));
    } else {
        u64 mi[MAXIN];
#pragma unroll
        for (int i = 0; i < MAXIN; i++)
            if (i < ni) mi[i] = bl.mat[(size_t) bl.mat_row[slot] * ni + i];
        const int cls = max(0, bl.xbits - (64 - __clzll((long long) q)));
        const BarG bg = bl.bar[(size_t) min(cls, 63) * bl.size_QP + row];
        const Modulus m = mod[row];
        forward_pass<A, ntt_p1(LOGN), false, LOGN, 0>(
                smem, cx, BconvGatherInt<LOGN, MAXIN>{in, mi, ni, bg, m},
                per_elem_store<u64>([&](size_t i, u64 v) { d[i] = v; }));
    }
}

// MAXIN: unroll bound of the conversion's input loop (4 covers the common digit widths without register spills, 6 the rest)
template<int LOGN, int MAXIN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_cols_bconv(u64 *dst, LimbList ll, NttPlan p, BconvLoad bl) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_COLS(ntt_p1(LOGN), p.tw)
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    const u64 *in = bl.in_base + ((size_t) bl.in_limb[slot] << LOGN);
    const u64 q = ll.q[slot];
    if (p.fp_enabled && (q >> fp::MAX_BITS) == 0)
        fwd_cols_bconv_body<FpArith, LOGN, MAXIN>(smem, stw, bar, q, row, slot, in, d, bl, p.mod);
    else fwd_cols_bconv_body<IntArith, LOGN, MAXIN>(smem, stw, bar, q, row, slot, in, d, bl, p.mod);
}

// last forward pass of HMult+Relin: out = (cx - NTT(delta)) * P^-1 + d_k with d_0 = a0 b0, d_1 = a0 b1 + a1 b0
struct EpiTensorPtrs {
    const u64 *d, *sub, *a0, *a1, *b0, *b1;
    u64 *out;
    int kpoly;
    Tw k;
    BarG bg;
    Modulus m;
};

template<class A>
struct EpiTensorStore {
    EpiTensorPtrs e;
    typename A::Consts c;
    u64 q;
    // one output: (sub - NTT) * k + d_k with d_0 = x0 y0, d_1 = x0 y1 + x1 y0.  FP64 limbs: everything stays in FP64 (the
    // lazy transform value, the product by the constant, the tensor terms) and is reduced once
    __device__ __forceinline__ u64 one(u64 s, typename A::T xv, u64 x0, u64 y0, u64 x1, u64 y1, double kd, double ki) const {
        if constexpr (std::is_same<A, FpArith>::value) {
            double r = fp::mulmod_c(fp::from_u64(s) - xv, kd, ki, c.q);
            const double a0 = fp::from_u64(x0), b0 = fp::from_u64(y0);
            if (e.kpoly == 0) {
                r += fp::mulmod_v(a0, b0, c.q, c.qinv);
            } else {
                r += fp::mulmod_v(a0, fp::from_u64(y1), c.q, c.qinv);
                r += fp::mulmod_v(fp::from_u64(x1), b0, c.q, c.qinv);
            }
            return fp::canon(fp::reduce(r, c.q, c.qinv), c.q);
        } else {
            const u64 t = A::canon_fwd(xv, c);
            const u64 r = mul_shoup(s + q - t, e.k, q);
            u64 dk;
            if (e.kpoly == 0) {
                dk = mul_mod_g(x0, y0, e.bg, e.m);
            } else {
                Acc128 acc{0, 0};
                acc.mac(x0, y1);
                acc.mac(x1, y0);
                dk = barrett_g(acc.lo, acc.hi, e.bg, e.m);
            }
            return add_mod(r, dk, q);
        }
    }
    // two half-batches of four outputs: all operand loads of a half (16-byte, runs of consecutive coefficients) are
    // issued before the arithmetic
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const typename A::T (&x)[NTT_EPT]) const {
        static_assert(RUN >= 2, "the epilogue belongs to the row pass");
        double kd = 0, ki = 0;
        if constexpr (std::is_same<A, FpArith>::value) kd = (double) e.k.x, ki = kd * c.qinv;
        constexpr int HB = RUN >= 4 ? 4 : 2;   // outputs per half-batch = one run
#pragma unroll
        for (int h = 0; h < NTT_EPT; h += HB) {
            u64 s[HB], x0[HB], y0[HB], x1[HB], y1[HB], r[HB];
            const size_t j = idx[h];
            ld_run<HB>(e.sub + j, s);
            ld_run<HB>(e.a0 + j, x0);
            ld_run<HB>(e.b0 + j, y0);
            if (e.kpoly != 0) {
                ld_run<HB>(e.a1 + j, x1);
                ld_run<HB>(e.b1 + j, y1);
            } else {
#pragma unroll
                for (int i = 0; i < HB; i++) x1[i] = y1[i] = 0;
            }
#pragma unroll
            for (int i = 0; i < HB; i++) r[i] = one(s[i], x[h + i], x0[i], y0[i], x1[i], y1[i], kd, ki);
            st_run<HB>(e.out + j, r);
        }
    }
};

template<class A, int LOGN>
__device__ __forceinline__ void fwd_rows_epi_tensor_body(u64 *smem, const Tw *stw, uint64_t *bar, u64 q,
                                                         const EpiTensorPtrs &e) {
    const typename A::Consts c = A::consts(q);
    PassCtx<A> cx{stw, bar, c, (int) blockIdx.x, {}, {}};
    forward_pass<A, ntt_p2(LOGN), true, LOGN, ntt_p1(LOGN)>(
            smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(e.d[i]); }),
            EpiTensorStore<A>{e, c, q});
}

template<int LOGN>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_fwd_rows_epi_tensor(u64 *data, LimbList ll, NttPlan p, EpiArgs ea,
                                                                         TensorSrc ts, const BarG *bar1) {
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    PFHE_NTT_SMEM_ROWS(p.tw)
    const size_t poly = (size_t) ts.l << LOGN;
    const size_t off = (size_t) row << LOGN;   // limb j of the data level == row j
    EpiTensorPtrs e;
    e.d = data + ((size_t) ll.data[slot] << LOGN);
    e.sub = ea.sub_base + ((size_t) ea.sub[slot] << LOGN);
    e.out = ea.out_base + ((size_t) ea.out[slot] << LOGN);
    e.kpoly = ea.add[slot];   // here: which polynomial (0 / 1) this slot produces
    e.a0 = ts.a + off, e.a1 = ts.a + poly + off, e.b0 = ts.b + off, e.b1 = ts.b + poly + off;
    e.k = ea.mulc[slot];
    e.bg = bar1[row];   // growth class 1: sum of two products
    e.m = p.mod[row];
    if (p.epi_prefetch && threadIdx.x < NTT_TILE * 8 / 128) {
        // the epilogue's operands were last touched a whole key switch ago: start them on their way from HBM to L2 now,
        // the transform hides the latency (one 128-byte line per thread and operand; the tile is 16 KiB of each)
        const size_t o = ((size_t) blockIdx.x << NTT_LOG_TILE) + (size_t) threadIdx.x * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(e.a0 + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(e.b0 + o));
        if (e.kpoly != 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(e.a1 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(e.b1 + o));
        }
    }
    const u64 q = ll.q[slot];
    if (p.fp_enabled && (q >> fp::MAX_BITS) == 0) fwd_rows_epi_tensor_body<FpArith, LOGN>(smem, stw, bar, q, e);
    else fwd_rows_epi_tensor_body<IntArith, LOGN>(smem, stw, bar, q, e);
}

template<class K>
static void opt_in_smem(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) NTT_SMEM_COLS);
}
template<int LOGN>
static void opt_in_all() {
    static bool done = false;
    if (done) return;
    opt_in_smem(k_fwd_cols<LOGN>);
    opt_in_smem(k_fwd_rows<LOGN, false>);
    opt_in_smem(k_fwd_rows<LOGN, true>);
    opt_in_smem(k_fwd_rows_epi<LOGN>);
    opt_in_smem(k_inv_rows<LOGN>);
    opt_in_smem(k_inv_cols<LOGN>);
    opt_in_smem(k_inv_rows_mul<LOGN>);
    opt_in_smem(k_fwd_cols_bconv<LOGN, 4>);
    opt_in_smem(k_fwd_cols_bconv<LOGN, FUSE_MAX_IN>);
    opt_in_smem(k_fwd_rows_epi_tensor<LOGN>);
    opt_in_smem(k_fwd_fused<LOGN>);
    opt_in_smem(k_fwd_small<LOGN>);
    opt_in_smem(k_inv_small<LOGN, false>);
    opt_in_smem(k_inv_small<LOGN, true>);
    done = true;
}

template<int LOGN>
static void fwd_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_fwd_cols<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, src, ll, p);
    launch_pdl(k_fwd_rows<LOGN, false>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ll, p);
}

// persistent single-launch form: as many CTAs as the device holds at once (a smaller grid only lowers the parallelism,
// a larger one is harmless: tickets decide who works)
static int fused_grid_limit() {
    static int limit = 0;
    if (!limit) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        limit = sms * NTT_MIN_BLOCKS;
    }
    return limit;
}
static int fused_delay() {
    static int d = -1;
    if (d < 0) {
        const char *e = std::getenv("PFHE_NTT_FUSED_DELAY");
        d = e ? std::atoi(e) : NTT_MAX_LIMBS;   // all column tiles first: a row tile then never finds its limb unfinished
    }
    return d;
}
static bool fused_enabled() {
    static int on = -1;
    if (on < 0) {
        // opt-in: measured slower than the launch pair on B200 (37.4 vs 33.2 us for 64 limbs of N = 2^16, DESIGN.md 4.1:
        // +7 % instructions and 3x the barrier stalls of the persistent loop outweigh the filled tail, which programmatic
        // dependent launch already half-hides for the pair).  Kept for the record and for A/B runs.
        const char *e = std::getenv("PFHE_NTT_FUSED");
        on = e && e[0] == '1';
    }
    return on;
}

// single launch for few limbs: every CTA must be resident at once (cooperative launch; on refusal the caller falls back)
static bool small_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = std::getenv("PFHE_NTT_SMALL");
        on = !(e && e[0] == '0');
    }
    return on;
}
static bool small_fits(const NttPlan &p, const LimbList &ll, const FusedSync *sync) {
    return sync && small_enabled() && (ll.count << (p.logn - NTT_LOG_TILE)) <= fused_grid_limit();
}
template<int LOGN>
static cudaError_t fwd_small_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st, FusedSync *sync) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    return launch_coop_pdl(k_fwd_small<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, src, ll, p, sync);
}
template<int LOGN>
static cudaError_t inv_small_impl(const NttPlan &p, u64 *dst, const u64 *src, const TensorSrc *ts, const BarG *bar0,
                                  const LimbList &ll, const Tw *fin, int by_slot, cudaStream_t st, FusedSync *sync) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    const Tw *f = fin ? fin : p.inv_fin;
    const int bs = fin ? by_slot : 0;
    if (ts) return launch_coop_pdl(k_inv_small<LOGN, true>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, src, *ts, bar0, ll, p, f, bs, sync);
    return launch_coop_pdl(k_inv_small<LOGN, false>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, src, TensorSrc{}, bar0, ll, p, f, bs, sync);
}
#define PFHE_DISPATCH_LOGN_RC(RC, FN, ...)                                                                \
    switch (p.logn) {                                                                                     \
        case 12: RC = FN<12>(__VA_ARGS__); break;                                                         \
        case 13: RC = FN<13>(__VA_ARGS__); break;                                                         \
        case 14: RC = FN<14>(__VA_ARGS__); break;                                                         \
        case 15: RC = FN<15>(__VA_ARGS__); break;                                                         \
        case 16: RC = FN<16>(__VA_ARGS__); break;                                                         \
        case 17: RC = FN<17>(__VA_ARGS__); break;                                                         \
        default: return cudaErrorInvalidValue;                                                            \
    }

template<int LOGN>
static void fwd_fused_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st, FusedSync *sync) {
    opt_in_all<LOGN>();
    const int total = ll.count * 2 * (1 << (LOGN - NTT_LOG_TILE));
    dim3 grid((unsigned) std::min(total, fused_grid_limit()));
    launch_pdl(k_fwd_fused<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, src, ll, p, sync, fused_delay());
}

template<int LOGN>
static void fwd_epi_impl(const NttPlan &p, u64 *data, const LimbList &ll, const EpiArgs &ea, cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_fwd_cols<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, data, data, ll, p);
    launch_pdl(k_fwd_rows_epi<LOGN>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, data, ll, p, ea);
}

template<int LOGN>
static void inv_impl(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                     cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_inv_rows<LOGN>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, src, ll, p);
    launch_pdl(k_inv_cols<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, ll, p, fin ? fin : p.inv_fin, fin ? by_slot : 0);
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

// the row passes move runs of four coefficients with 256-bit accesses: every polynomial buffer must be 32-byte aligned
// (limb offsets are multiples of 8 N bytes, so this is a condition on the base pointers only)
template<class... Ptr>
static bool aligned32(Ptr... p) {
    return ((... | reinterpret_cast<uintptr_t>(p)) & 31u) == 0;
}

cudaError_t ntt_forward(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st, FusedSync *sync) {
    if (ll.count == 0) return cudaSuccess;
    if (!aligned32(dst, src)) return cudaErrorMisalignedAddress;
    if (small_fits(p, ll, sync)) {   // few limbs: one cooperative launch for both passes
        cudaError_t rc = cudaSuccess;
        PFHE_DISPATCH_LOGN_RC(rc, fwd_small_impl, p, dst, src, ll, st, sync)
        if (rc == cudaSuccess) return cudaGetLastError();
        cudaGetLastError();   // refused (does not fit next to what is running, or not supported): the launch pair
    }
    // one persistent launch when the tiles of one pass fill the device at least once, else the launch pair
    if (sync && fused_enabled() && (ll.count << (p.logn - NTT_LOG_TILE)) >= fused_grid_limit()) {
        PFHE_DISPATCH_LOGN(fwd_fused_impl, p, dst, src, ll, st, sync)
        return cudaGetLastError();
    }
    PFHE_DISPATCH_LOGN(fwd_impl, p, dst, src, ll, st)
    return cudaGetLastError();
}

cudaError_t ntt_forward_epilogue(const NttPlan &p, u64 *data, const LimbList &ll, const EpiArgs &ea, cudaStream_t st) {
    if (ll.count == 0) return cudaSuccess;
    if (!aligned32(data, ea.sub_base, ea.out_base, ea.add_base)) return cudaErrorMisalignedAddress;
    PFHE_DISPATCH_LOGN(fwd_epi_impl, p, data, ll, ea, st)
    return cudaGetLastError();
}

template<int LOGN>
static void inv_mul_impl(const NttPlan &p, u64 *dst, const TensorSrc &ts, const BarG *bar0, const LimbList &ll,
                         const Tw *fin, int by_slot, cudaStream_t st) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    launch_pdl(k_inv_rows_mul<LOGN>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ts, bar0, ll, p);
    launch_pdl(k_inv_cols<LOGN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, ll, p, fin ? fin : p.inv_fin, fin ? by_slot : 0);
}

template<int LOGN>
static void fwd_bconv_impl(const NttPlan &p, u64 *dst, const LimbList &ll, const BconvLoad &bl, const EpiArgs *ea,
                           const TensorSrc *ts, const BarG *bar1, cudaStream_t st, int phase) {
    opt_in_all<LOGN>();
    dim3 grid(1 << (LOGN - NTT_LOG_TILE), ll.count);
    if (phase != 2) {
        if (bl.ni <= 4) launch_pdl(k_fwd_cols_bconv<LOGN, 4>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, ll, p, bl);
        else launch_pdl(k_fwd_cols_bconv<LOGN, FUSE_MAX_IN>, grid, NTT_THREADS, NTT_SMEM_COLS, st, dst, ll, p, bl);
    }
    if (phase == 1) return;
    if (!ea && phase == 3) launch_pdl(k_fwd_rows<LOGN, true>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ll, p);
    else if (!ea) launch_pdl(k_fwd_rows<LOGN, false>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ll, p);
    else if (!ts) launch_pdl(k_fwd_rows_epi<LOGN>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ll, p, *ea);
    else launch_pdl(k_fwd_rows_epi_tensor<LOGN>, grid, NTT_THREADS, NTT_SMEM_ROWS, st, dst, ll, p, *ea, *ts, bar1);
}

cudaError_t ntt_inverse_mul(const NttPlan &p, u64 *dst, const TensorSrc &ts, const BarG *bar0, const LimbList &ll,
                            const Tw *fin, int by_slot, cudaStream_t st, FusedSync *sync) {
    if (ll.count == 0) return cudaSuccess;
    if (!aligned32(dst, ts.a, ts.b)) return cudaErrorMisalignedAddress;
    if (small_fits(p, ll, sync)) {
        cudaError_t rc = cudaSuccess;
        PFHE_DISPATCH_LOGN_RC(rc, inv_small_impl, p, dst, nullptr, &ts, bar0, ll, fin, by_slot, st, sync)
        if (rc == cudaSuccess) return cudaGetLastError();
        cudaGetLastError();
    }
    PFHE_DISPATCH_LOGN(inv_mul_impl, p, dst, ts, bar0, ll, fin, by_slot, st)
    return cudaGetLastError();
}

cudaError_t ntt_forward_bconv(const NttPlan &p, u64 *dst, const LimbList &ll, const BconvLoad &bl, const EpiArgs *ea,
                              const TensorSrc *ts, const BarG *bar1, cudaStream_t st, int phase) {
    if (ll.count == 0) return cudaSuccess;
    if (!aligned32(dst, bl.in_base) || (ea && !aligned32(ea->sub_base, ea->out_base, ea->add_base)) ||
        (ts && !aligned32(ts->a, ts->b)))
        return cudaErrorMisalignedAddress;
    PFHE_DISPATCH_LOGN(fwd_bconv_impl, p, dst, ll, bl, ea, ts, bar1, st, phase)
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------
// single-CTA transforms for dim <= 2048 on caller-supplied tables (fnwt_1d[_opt] / inwt_1d[_opt], reference
// src/ntt/ntt_1d.cu:146-292; tables in the reference's own order: tw[bitrev(i)] = psi^i, Shoup companions apart).
// One CTA per limb, one butterfly per thread and stage, the limb lives in shared memory between the first load and
// the last store.  Limb i of the launch is absolute index start + i in every array, like the reference.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_fnwt_1d(u64 *inout, const u64 *tw, const u64 *tws, const Modulus *mod, int dim,
                                                  int start) {
    extern __shared__ __align__(16) u64 s1d[];
    pdl_launch_dependents();
    pdl_wait();
    const size_t limb = (size_t) (blockIdx.x + start);
    const u64 q = mod[limb].q, q2 = 2 * q;
    u64 *d = inout + limb * dim;
    const u64 *w = tw + limb * dim, *ws = tws + limb * dim;
    const int t = threadIdx.x;
    for (int m = 1, gap = dim >> 1; m < dim; m <<= 1, gap >>= 1) {
        const int i = t / gap, j = t - i * gap, idx = 2 * i * gap + j;
        const u64 X = m == 1 ? d[idx] : s1d[idx], Y = m == 1 ? d[idx + gap] : s1d[idx + gap];
        // Harvey butterfly, values stay below 4q (reference butterfly.cuh:10-26)
        const u64 Xr = csub(X, q2);
        const u64 T = mul_shoup_lazy(Y, w[m + i], ws[m + i], q);
        const u64 a = Xr + T, b = Xr + q2 - T;
        if (gap == 1) {
            d[idx] = csub(csub(a, q2), q);
            d[idx + 1] = csub(csub(b, q2), q);
        } else {
            __syncthreads();   // every thread has read this stage's operands
            s1d[idx] = a, s1d[idx + gap] = b;
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024) k_inwt_1d(u64 *inout, const u64 *itw, const u64 *itws, const Modulus *mod,
                                                  const u64 *scalar, const u64 *scalar_shoup, int dim, int start) {
    extern __shared__ __align__(16) u64 s1d[];
    pdl_launch_dependents();
    pdl_wait();
    const size_t limb = (size_t) (blockIdx.x + start);
    const u64 q = mod[limb].q, q2 = 2 * q;
    u64 *d = inout + limb * dim;
    const u64 *w = itw + limb * dim, *ws = itws + limb * dim;
    const int t = threadIdx.x;
    for (int m = dim >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1) {
        const int i = t / gap, j = t - i * gap, idx = 2 * i * gap + j;
        const bool first = gap == 1;
        const u64 X = first ? d[idx] : s1d[idx], Y = first ? d[idx + gap] : s1d[idx + gap];
        // Gentleman-Sande butterfly on [0, 2q) (butterfly.cuh:28-37); canonical inputs are inside that range
        const u64 S = csub(X + Y, q2);
        const u64 D = mul_shoup_lazy(X + q2 - Y, w[m + i], ws[m + i], q);
        if (m == 1) {
            // lower half times scalar[limb]; the upper half carries only what the caller folded into itw[1]
            d[idx] = mul_shoup(S, scalar[limb], scalar_shoup[limb], q);
            d[idx + gap] = csub(D, q);
        } else {
            __syncthreads();
            s1d[idx] = S, s1d[idx + gap] = D;
            __syncthreads();
        }
    }
}

cudaError_t ntt_1d(bool inverse, u64 *inout, const u64 *tw, const u64 *tws, const Modulus *mod, const u64 *scalar,
                   const u64 *scalar_shoup, size_t dim, size_t count, size_t start, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    if (dim < 2 || dim > 2048 || (dim & (dim - 1))) return cudaErrorInvalidValue;
    const dim3 grid((unsigned) count), block((unsigned) (dim / 2));
    const size_t smem = dim * sizeof(u64);
    if (inverse) launch_pdl(k_inwt_1d, grid, block, smem, st, inout, tw, tws, mod, scalar, scalar_shoup, (int) dim, (int) start);
    else launch_pdl(k_fnwt_1d, grid, block, smem, st, inout, tw, tws, mod, (int) dim, (int) start);
    return cudaGetLastError();
}

cudaError_t ntt_inverse(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                        cudaStream_t st, FusedSync *sync) {
    if (ll.count == 0) return cudaSuccess;
    if (!aligned32(dst, src)) return cudaErrorMisalignedAddress;
    if (small_fits(p, ll, sync)) {
        cudaError_t rc = cudaSuccess;
        PFHE_DISPATCH_LOGN_RC(rc, inv_small_impl, p, dst, src, nullptr, nullptr, ll, fin, by_slot, st, sync)
        if (rc == cudaSuccess) return cudaGetLastError();
        cudaGetLastError();
    }
    PFHE_DISPATCH_LOGN(inv_impl, p, dst, src, ll, fin, by_slot, st)
    return cudaGetLastError();
}

} // namespace pfhe
