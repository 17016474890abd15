// ntt_kernels.cu — batched forward / inverse NTT kernels (sm_100a).
//
// Replaces ntt()/invntt() (ref_ntt.cpp:28-47, :59-87) and ntt2x2_ref()/invntt2x2_ref()
// (ref_ntt2x2.cpp:37-145) for batches of polynomials resident in HBM.
//
// Execution model: persistent CTAs, one polynomial per warp per iteration.  Every warp
// runs its own TMA pipeline: lane 0 issues 1 KiB cp.async.bulk loads STAGES-1 polynomials
// ahead into the warp's private ring of shared-memory slots (completion on one mbarrier
// per slot), the warp transforms a polynomial entirely in registers + its private scratch
// (ntt_core.cuh), writes the canonical result back into the slot and lane 0 issues a
// cp.async.bulk store.  There is no CTA-wide barrier anywhere; HBM traffic is exactly
// 1 KiB in + 1 KiB out per polynomial, always as whole aligned 1 KiB bursts.
#include <cuda_runtime.h>

#include <cstdlib>

#include "kernels.h"
#include "ntt_core.cuh"
#include "tma.cuh"

namespace dil {

constexpr int POLY_BYTES = N * 4;

template <int WARPS, int STAGES>
struct NttSmem {
    alignas(128) uint32_t slot[WARPS][STAGES][N];
    alignas(16) uint32_t scratch[WARPS][SCRATCH_WORDS];
    alignas(8) uint64_t full[WARPS][STAGES];
};

template <int WARPS, int STAGES, int MIN_CTAS, bool INVERSE>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
ntt_tma_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, uint32_t n_polys) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sm = *reinterpret_cast<NttSmem<WARPS, STAGES>*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * WARPS + warp;
    const uint32_t stride = gridDim.x * WARPS;
    if (gwarp >= n_polys) return;  // warp-uniform; no CTA-wide sync below
    const uint32_t n_my = (n_polys - gwarp + stride - 1) / stride;

    FwdTw ftw;
    InvTw itw;
    if constexpr (INVERSE) load_inv_tw(itw, &TW_INV, lane);
    else load_fwd_tw(ftw, &TW_FWD, lane);

    uint32_t* scr = sm.scratch[warp];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&sm.full[warp][s]), 1);
        fence_mbar_init();
        // prologue: prefetch STAGES-1 polynomials
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if ((uint32_t)s < n_my) {
                uint32_t bar = smem_u32(&sm.full[warp][s]);
                mbar_expect_tx(bar, POLY_BYTES);
                bulk_g2s(smem_u32(sm.slot[warp][s]), src + (size_t)(gwarp + s * stride) * N, POLY_BYTES, bar);
            }
        }
    }
    __syncwarp();

    int s = 0;
    uint32_t parity = 0;
    for (uint32_t k = 0; k < n_my; k++) {
        uint32_t* slot = sm.slot[warp][s];
        mbar_wait(smem_u32(&sm.full[warp][s]), parity);
        uint32_t x[8];
        if constexpr (!INVERSE) {
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = slot[32 * r + lane];
            ntt_fwd_warp(x, scr, ftw, lane);
            uint4* o = reinterpret_cast<uint4*>(slot) + lane;
            o[0] = make_uint4(x[0], x[1], x[2], x[3]);
            o[32] = make_uint4(x[4], x[5], x[6], x[7]);
        } else {
            const uint4* in = reinterpret_cast<const uint4*>(slot) + lane;
            uint4 lo = in[0], hi = in[32];
            x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
            x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
            ntt_inv_warp(x, scr, itw, lane);
#pragma unroll
            for (int r = 0; r < 8; r++) slot[32 * r + lane] = x[r];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(dst + (size_t)(gwarp + k * stride) * N, smem_u32(slot), POLY_BYTES);
            bulk_commit();
            // refill the slot used one iteration ago (its store has had a full iteration)
            uint32_t kn = k + STAGES - 1;
            if (kn < n_my) {
                int sn = s == 0 ? STAGES - 1 : s - 1;
                bulk_wait_read<1>();
                uint32_t bar = smem_u32(&sm.full[warp][sn]);
                mbar_expect_tx(bar, POLY_BYTES);
                bulk_g2s(smem_u32(sm.slot[warp][sn]), src + (size_t)(gwarp + kn * stride) * N, POLY_BYTES, bar);
            }
        }
        if (++s == STAGES) {
            s = 0;
            parity ^= 1;
        }
    }
    if (lane == 0) bulk_wait_all<0>();  // smem must outlive the last stores
}

// ---------------------------------------------------------------------------------------
// Pair variant: a warp owns PAIRS of consecutive polynomials (2 KiB bulk copies), which halves the
// per-polynomial cost of issuing TMA operations, mbarrier waits, proxy fences and loop control
// (about 45 of 304 warp instructions per polynomial in the single-polynomial kernel are issued by
// lane 0 alone for that bookkeeping).  3 slots of 2 KiB per warp, 24 warps, one CTA per SM.
// ---------------------------------------------------------------------------------------
template <int WARPS, int STAGES>
struct NttPairSmem {
    alignas(128) uint32_t slot[WARPS][STAGES][2 * N];
    alignas(16) uint32_t scratch[WARPS][SCRATCH_WORDS];
    alignas(8) uint64_t full[WARPS][STAGES];
};

template <int WARPS, int STAGES, bool INVERSE>
__global__ void __launch_bounds__(WARPS * 32, 1)
ntt_pair_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, uint32_t n_polys) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sm = *reinterpret_cast<NttPairSmem<WARPS, STAGES>*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t n_pairs = (n_polys + 1) >> 1;
    const uint32_t gwarp = blockIdx.x * WARPS + warp;
    const uint32_t stride = gridDim.x * WARPS;
    if (gwarp >= n_pairs) return;
    const uint32_t n_my = (n_pairs - gwarp + stride - 1) / stride;

    FwdTw ftw;
    InvTw itw;
    if constexpr (INVERSE) load_inv_tw(itw, &TW_INV, lane);
    else load_fwd_tw(ftw, &TW_FWD, lane);

    uint32_t* scr = sm.scratch[warp];
    // bytes of pair index q (the last pair of an odd batch holds one polynomial)
    auto pair_bytes = [&](uint32_t q) -> uint32_t { return (2 * q + 1 < n_polys) ? 2 * POLY_BYTES : POLY_BYTES; };
    const size_t step = (size_t)stride * 2 * N;                 // elements between a warp's consecutive pairs
    const int32_t* ld_ptr = src + (size_t)gwarp * 2 * N;        // next pair to load
    int32_t* st_ptr = dst + (size_t)gwarp * 2 * N;              // next pair to store
    uint32_t ld_q = gwarp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&sm.full[warp][s]), 1);
        fence_mbar_init();
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if ((uint32_t)s < n_my) {
                uint32_t bar = smem_u32(&sm.full[warp][s]);
                uint32_t bytes = pair_bytes(ld_q);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(sm.slot[warp][s]), ld_ptr, bytes, bar);
                ld_ptr += step;
                ld_q += stride;
            }
        }
    }
    __syncwarp();

    int s = 0;
    uint32_t parity = 0;
    uint32_t q = gwarp;
    for (uint32_t k = 0; k < n_my; k++, q += stride) {
        uint32_t* slot = sm.slot[warp][s];
        mbar_wait(smem_u32(&sm.full[warp][s]), parity);
        const int np = (2 * q + 1 < n_polys) ? 2 : 1;
        for (int p = 0; p < np; p++) {
            uint32_t* poly = slot + p * N;
            uint32_t x[8];
            if constexpr (!INVERSE) {
#pragma unroll
                for (int r = 0; r < 8; r++) x[r] = poly[32 * r + lane];
                ntt_fwd_warp(x, scr, ftw, lane);
                uint4* o = reinterpret_cast<uint4*>(poly) + lane;
                o[0] = make_uint4(x[0], x[1], x[2], x[3]);
                o[32] = make_uint4(x[4], x[5], x[6], x[7]);
            } else {
                const uint4* in = reinterpret_cast<const uint4*>(poly) + lane;
                uint4 lo = in[0], hi = in[32];
                x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
                x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
                ntt_inv_warp(x, scr, itw, lane);
#pragma unroll
                for (int r = 0; r < 8; r++) poly[32 * r + lane] = x[r];
            }
            __syncwarp();   // scratch is reused by the next polynomial
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(st_ptr, smem_u32(slot), np * POLY_BYTES);
            bulk_commit();
            uint32_t kn = k + STAGES - 1;
            if (kn < n_my) {
                int sn = s == 0 ? STAGES - 1 : s - 1;
                bulk_wait_read<1>();
                uint32_t bar = smem_u32(&sm.full[warp][sn]);
                uint32_t bytes = pair_bytes(ld_q);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(sm.slot[warp][sn]), ld_ptr, bytes, bar);
                ld_ptr += step;
                ld_q += stride;
            }
        }
        st_ptr += step;
        if (++s == STAGES) {
            s = 0;
            parity ^= 1;
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

template <int WARPS, bool INVERSE>
static cudaError_t launch_ntt_pair(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st) {
    using Smem = NttPairSmem<WARPS, 3>;
    auto kern = ntt_pair_kernel<WARPS, 3, INVERSE>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(sizeof(Smem)), configured); e != cudaSuccess) return e;
    size_t n_pairs = (n_polys + 1) / 2;
    size_t want = (n_pairs + WARPS - 1) / WARPS;
    size_t cap = (size_t)sm_count;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid, WARPS * 32, sizeof(Smem), st>>>(dst, src, (uint32_t)n_polys);
    return cudaGetLastError();
}

// ---- launch configuration ----
// Default: ONE 32-warp CTA per SM, 3 stages (32 resident polynomials per SM).  Measured at 2^18 polys
// (fraction of the 6482.7 GB/s HBM peak, forward / inverse): 32w x 1 CTA 0.846 / 0.801, 16w x 2 0.816 / 0.781,
// 8w x 4 0.785 / 0.757, 4w x 8 0.771 / 0.747; 4 stages change nothing, forcing <= 48 registers for 5-6 CTAs
// of 8 warps loses 10-18 %.  DIL_NTT_CFG=<n> selects the alternative shapes (development knob).
template <int WARPS, int STAGES, int CTAS, bool INVERSE>
static cudaError_t launch_ntt_cfg(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st) {
    using Smem = NttSmem<WARPS, STAGES>;
    auto kern = ntt_tma_kernel<WARPS, STAGES, CTAS, INVERSE>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(sizeof(Smem)), configured); e != cudaSuccess) return e;
    size_t want = (n_polys + WARPS - 1) / WARPS;
    size_t cap = (size_t)sm_count * CTAS;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid, WARPS * 32, sizeof(Smem), st>>>(dst, src, (uint32_t)n_polys);
    return cudaGetLastError();
}

static int ntt_cfg() {   // development knob, read once (thread-safe static initialisation)
    static const int cfg = [] {
        const char* e = std::getenv("DIL_NTT_CFG");
        return e ? std::atoi(e) : 0;
    }();
    return cfg;
}

template <bool INVERSE>
static cudaError_t launch_ntt(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    switch (ntt_cfg()) {
        case 1: return launch_ntt_cfg<32, 3, 1, INVERSE>(dst, src, n_polys, sm_count, st);
        case 2: return launch_ntt_cfg<4, 3, 8, INVERSE>(dst, src, n_polys, sm_count, st);
        case 3: return launch_ntt_cfg<8, 4, 4, INVERSE>(dst, src, n_polys, sm_count, st);
        case 4: return launch_ntt_cfg<16, 4, 2, INVERSE>(dst, src, n_polys, sm_count, st);
        case 5: return launch_ntt_cfg<16, 3, 2, INVERSE>(dst, src, n_polys, sm_count, st);
        case 6: return launch_ntt_cfg<32, 4, 1, INVERSE>(dst, src, n_polys, sm_count, st);
        case 7: return launch_ntt_cfg<16, 5, 2, INVERSE>(dst, src, n_polys, sm_count, st);
        case 8: return launch_ntt_cfg<8, 3, 4, INVERSE>(dst, src, n_polys, sm_count, st);
        case 9: return launch_ntt_pair<24, INVERSE>(dst, src, n_polys, sm_count, st);
        case 10: return launch_ntt_pair<16, INVERSE>(dst, src, n_polys, sm_count, st);
        case 11: return launch_ntt_pair<28, INVERSE>(dst, src, n_polys, sm_count, st);
        case 12: return launch_ntt_cfg<32, 3, 1, INVERSE>(dst, src, n_polys, sm_count, st);
        default:
            // forward: one polynomial per warp iteration; inverse: the pair variant measured 2 % faster
            if (INVERSE) return launch_ntt_pair<24, INVERSE>(dst, src, n_polys, sm_count, st);
            return launch_ntt_cfg<32, 3, 1, INVERSE>(dst, src, n_polys, sm_count, st);
    }
}

cudaError_t launch_ntt_fwd(int32_t* dst, const int32_t* src, size_t n, int sms, cudaStream_t st) {
    return launch_ntt<false>(dst, src, n, sms, st);
}
cudaError_t launch_ntt_inv(int32_t* dst, const int32_t* src, size_t n, int sms, cudaStream_t st) {
    return launch_ntt<true>(dst, src, n, sms, st);
}

}  // namespace dil
