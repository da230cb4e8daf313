// ntt_cluster.cu -- EXPERIMENT (opt-in, PFHE_NTT_CLUSTER=8 | 16): the N = 2^16 forward transform in ONE launch with the
// intermediate in distributed shared memory instead of L2.
//
// A thread-block cluster of CL CTAs owns one limb.  The limb's 2^16 words are cut into CL contiguous slabs of 2^16 / CL words,
// slab r lives in the shared memory of CTA r (64 KiB at CL = 8, 32 KiB at CL = 16).  CTA r runs the column pass of column tiles
// r * TPC .. r * TPC + TPC - 1 (TPC = 32 / CL; loads from global memory as in k_fwd_cols) and stores every result word straight
// into the slab of the CTA that owns its address (st.shared::cluster through mapa); after a cluster barrier it runs the row
// pass of the row tiles r * TPC .. -- whose rows are exactly its own slab -- and stores canonical residues to global memory as
// k_fwd_rows does.  Against the launch pair this removes one global store + load of every word, the second launch and the
// second ramp; it costs 36 / 68 KiB more shared memory per CTA and the barrier.  Pass drivers, arithmetic and twiddle tables
// are those of ntt.cuh, so the words are the same bit for bit.  Measured in DESIGN.md section 4.1.
#include <cstdlib>

#include "ntt_api.cuh"
#include "launch.hpp"

namespace pfhe {

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, u64 v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// column-pass results into the slabs: word i of the limb goes to CTA i >> LOG_SLAB, position i & (2^LOG_SLAB - 1).
// The last column round hands a thread two groups of four rows 4 hi .. 4 hi + 3 (RoundMap<8, 2, false>): one owner per group.
template<class A, int LOG_SLAB>
struct SlabStore {
    uint32_t slab_saddr;   // this CTA's slab in the shared window; the same offset in every CTA of the cluster
    template<int RUN>
    __device__ __forceinline__ void scatter(const size_t (&idx)[NTT_EPT], const typename A::T (&x)[NTT_EPT]) const {
        static_assert(NTT_EPT == 8, "written for radix-8 tiles");
#pragma unroll
        for (int g = 0; g < NTT_EPT; g += 4) {
            const uint32_t base = map_to_rank(slab_saddr, (uint32_t) (idx[g] >> LOG_SLAB));
#pragma unroll
            for (int k = 0; k < 4; k++)
                st_cluster(base + (((uint32_t) idx[g + k] & ((1u << LOG_SLAB) - 1u)) << 3), A::raw(x[g + k]));
        }
    }
};

template<class T>
struct ArithTag {
    using type = T;
};

constexpr size_t cluster_smem_bytes(int cl) {
    return NTT_TILE * sizeof(u64) + NTT_STW_ENTRIES * sizeof(Tw) + 16 + ((size_t) 65536 / cl) * sizeof(u64);
}

template<int CL>
__global__ void __launch_bounds__(NTT_THREADS, CL == 16 ? 4 : 2) k_fwd_cluster(u64 *dst, const u64 *src, LimbList ll, NttPlan p) {
    constexpr int LOGN = 16, P1 = ntt_p1(LOGN), P2 = ntt_p2(LOGN);
    constexpr int TILES = 1 << (LOGN - NTT_LOG_TILE), TPC = TILES / CL;
    constexpr int LOG_SLAB = LOGN - (CL == 16 ? 4 : 3);
    static_assert(CL == 8 || CL == 16, "cluster sizes tried");
    static_assert(NTT_LOG_TILE == 11 && P1 == 8, "slab ownership below assumes 2048-word tiles of 256 x 8");
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    u64 *smem = reinterpret_cast<u64 *>(dyn_smem);
    Tw *stw = reinterpret_cast<Tw *>(smem + NTT_TILE);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stw + NTT_STW_ENTRIES);
    u64 *slab = reinterpret_cast<u64 *>(bar + 2);
    const int slot = blockIdx.y;
    const int row = ll.row[slot];
    const u64 q = ll.q[slot];
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    cluster_arrive();   // nobody writes into a slab before its CTA runs; the wait comes just before the first tile's stores...
    pdl_launch_dependents();
    if (threadIdx.x == 0) stage_twiddles<P1>(stw, p.tw + ((size_t) row << LOGN), bar);   // one staging for all column tiles
    pdl_wait();
    cluster_wait();     // ...but the pass driver has no hook there, and the arrival of a started CTA is immediate
    const u64 *s = src + ((size_t) ll.src[slot] << LOGN);
    u64 *d = dst + ((size_t) ll.data[slot] << LOGN);
    const Tw *tw = p.tw + ((size_t) row << LOGN);
    const uint32_t slab_saddr = smem_u32(slab);
    auto body = [&](auto tag) {
        using A = typename decltype(tag)::type;
        const typename A::Consts c = A::consts(q);
#pragma unroll 1
        for (int j = 0; j < TPC; j++) {
            PassCtx<A> cx{stw, bar, c, (int) rank * TPC + j, {}, {}, 0};
            forward_pass<A, P1, false, LOGN, 0>(
                    smem, cx, per_elem_load<typename A::T>([&](size_t i) { return A::load(s[i], c); }),
                    SlabStore<A, LOG_SLAB>{slab_saddr});
            __syncthreads();   // the exchange tile is free again
        }
        cluster_arrive();
        cluster_wait();        // every slab is complete and visible
#pragma unroll 1
        for (int j = 0; j < TPC; j++) {
            PassCtx<A> cx{tw, nullptr, c, (int) rank * TPC + j, {}, {}};
            forward_pass<A, P2, true, LOGN, P1>(
                    smem, cx,
                    per_elem_load<typename A::T>([&](size_t i) { return A::from_raw(slab[i & ((1u << LOG_SLAB) - 1u)]); }),
                    vec_store<typename A::T>(d, [&](typename A::T v) { return A::canon_fwd(v, c); }));
            __syncthreads();
        }
    };
    if (p.fp_enabled && (q >> fp::MAX_BITS) == 0) body(ArithTag<FpArith>{});
    else body(ArithTag<IntArith>{});
}

template<int CL>
static cudaError_t launch_cluster(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) {
    static bool ready = false;
    if (!ready) {
        cudaError_t rc = cudaFuncSetAttribute(k_fwd_cluster<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int) cluster_smem_bytes(CL));
        if (rc != cudaSuccess) return rc;
        if (CL > 8) {
            rc = cudaFuncSetAttribute(k_fwd_cluster<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (rc != cudaSuccess) return rc;
        }
        ready = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL, ll.count);
    cfg.blockDim = dim3(NTT_THREADS);
    cfg.dynamicSmemBytes = cluster_smem_bytes(CL);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, k_fwd_cluster<CL>, dst, src, ll, p);
}

// 0 = off (default), 8 or 16 = cluster size
int ntt_cluster_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = std::getenv("PFHE_NTT_CLUSTER");
        const int v = e ? std::atoi(e) : 0;
        mode = (v == 8 || v == 16) ? v : 0;
    }
    return mode;
}

cudaError_t ntt_forward_cluster(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st) {
    if (p.logn != 16) return cudaErrorInvalidValue;
    if (ll.count == 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 31u) return cudaErrorMisalignedAddress;
    const cudaError_t rc = ntt_cluster_mode() == 16 ? launch_cluster<16>(p, dst, src, ll, st) : launch_cluster<8>(p, dst, src, ll, st);
    return rc != cudaSuccess ? rc : cudaGetLastError();
}

} // namespace pfhe
