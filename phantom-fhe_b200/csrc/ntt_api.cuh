// ntt_api.cuh -- host-callable NTT launchers.
#pragma once
#include "ntt.cuh"

namespace pfhe {

struct NttPlan {
    int logn;
    const Tw *tw;        // [size_QP][N] forward twiddles, kernel-native order (tw_native_index)
    const Tw *itw;       // [size_QP][N] inverse twiddles, same order
    const Modulus *mod;  // [size_QP]
    const Tw *inv_fin;   // [size_QP][2]: {n^-1, itw[1] * n^-1} for the last inverse stage
    // rows with q < 2^46 run the FP64 butterflies: their table entries hold (double(w), double(w)/double(q))
    const unsigned char *is_fp;   // [size_QP]
    const double2 *fpc;           // [size_QP] {double(q), 1/double(q)}
    int fp_enabled;               // 0: integer butterflies everywhere (PFHE_FP64_NTT=0)
    int epi_prefetch;             // fused epilogues prefetch their operands into L2 at tile start (PFHE_EPI_PREFETCH=1; measured +1.3 % time on B200, off by default)
};

// per-stream state of the single-launch transform (k_fwd_fused): work-item ticket, per-slot counts of finished column
// tiles, exit count.  Zero-initialised once; every launch leaves it zeroed.
struct FusedSync {
    unsigned ticket, done;
    unsigned ready[NTT_MAX_LIMBS];
};

// forward negacyclic NTT of the limbs in `ll` (replaces nwt_2d_radix8_forward_inplace and its
// include_special_mod / include_temp_mod / exclude_range variants, reference include/ntt.cuh:172-201).
// With `sync` (and enough limbs to fill the GPU) both passes run in ONE persistent launch: CTAs draw (limb, pass, tile)
// work items from a ticket counter, a row tile starts as soon as the column tiles of ITS limb are done (per-limb
// counters, release / acquire through L2) instead of waiting for the whole column-pass grid.
cudaError_t ntt_forward(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st,
                        FusedSync *sync = nullptr);

// experiment (ntt_cluster.cu, PFHE_NTT_CLUSTER=8 | 16, N = 2^16 only): one launch, a thread-block cluster per limb, the
// intermediate in distributed shared memory.  ntt_cluster_mode() = 0 unless the variable asks for it.
int ntt_cluster_mode();
cudaError_t ntt_forward_cluster(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, cudaStream_t st);

// forward NTT of `data` (in place, ll.src must equal ll.data) whose row pass ends in the EpiArgs epilogue
cudaError_t ntt_forward_epilogue(const NttPlan &p, u64 *data, const LimbList &ll, const EpiArgs &ea, cudaStream_t st);

// inverse NTT whose input is the limb-wise product a1 * b1 (HMult fused into the key switch)
cudaError_t ntt_inverse_mul(const NttPlan &p, u64 *dst, const TensorSrc &ts, const BarG *bar0, const LimbList &ll,
                            const Tw *fin, int by_slot, cudaStream_t st, FusedSync *sync = nullptr);

// forward NTT of base-converted inputs (BconvLoad) written to `dst`; optional epilogue (ea) and tensor addend (ts)
// phase: 0 = column pass (with the conversion) then row pass; 1 = column pass only; 2 = row pass only (the two
// halves may then be enqueued on different streams); 3 = both passes, FP64 limbs left in lazy FP64 form (no epilogue)
cudaError_t ntt_forward_bconv(const NttPlan &p, u64 *dst, const LimbList &ll, const BconvLoad &bl, const EpiArgs *ea,
                              const TensorSrc *ts, const BarG *bar1, cudaStream_t st, int phase = 0);

// inverse NTT incl. n^-1; `fin` (optional) = per-slot or per-row {c, itw1*c} pairs with c = n^-1 * scalar
// (replaces nwt_2d_radix8_backward[_inplace][_scale] and variants, include/ntt.cuh:206-226)
// (with `sync` and few enough limbs for all CTAs to be resident: one cooperative launch for both passes, as ntt_forward)
cudaError_t ntt_inverse(const NttPlan &p, u64 *dst, const u64 *src, const LimbList &ll, const Tw *fin, int by_slot,
                        cudaStream_t st, FusedSync *sync = nullptr);

// single-CTA transforms for dim <= 2048 on caller-supplied reference-order tables (fnwt_1d / inwt_1d,
// reference include/ntt.cuh:157-170); limb i of the call is absolute index start + i in every array
cudaError_t ntt_1d(bool inverse, u64 *inout, const u64 *tw, const u64 *tws, const Modulus *mod, const u64 *scalar,
                   const u64 *scalar_shoup, size_t dim, size_t count, size_t start, cudaStream_t st);

} // namespace pfhe
