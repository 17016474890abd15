// matvec_kernels.cu — ExpandA and the fused sign/verify core:
//     w = [INTT] ( A_hat * [NTT] v ),   A_hat = ExpandA(rho) or a pre-expanded shared matrix.
//
// Reference data flow being replaced (rtl_src/combined_top.v): NTT_Y :1850-1874 ->
// MULT_A_Y :1875-1913 -> NTTI_W :1914-1933 for signing, :921-988 for keygen, :1347-1469 for
// verification, with A produced by gen_a_ext/sampler_a_ext/rejection_a.  The FPGA writes A
// to BRAM_0 (combined_top.v:795-803); here A_hat lives only in shared memory:
//   * shared-rho / shared-A mode: every persistent CTA holds the whole k*l matrix in shared
//     memory (16/30/56 KiB) for its lifetime (expanded in the prologue by k*l threads, one
//     Keccak state per thread, or copied once from a pre-expanded buffer); each warp then
//     streams batch items: l forward NTTs in registers, k accumulate-reduce-INTT rounds.
//   * per-item-rho mode: one small CTA per item; k*l threads expand the item's matrix into
//     shared memory, then warp 0 runs the same per-item core.
// HBM traffic per item is (l + k) KiB (+32 B of rho); A never touches HBM.
#include <cuda_runtime.h>

#include "dilithium_b200.h"
#include "keccak.cuh"
#include "kernels.h"
#include "matvec_core.cuh"

namespace dil {

// ---- materialising ExpandA (keys, tests): one thread per polynomial ----
__global__ void __launch_bounds__(64) expand_a_kernel(int32_t* __restrict__ a_hat, const uint8_t* __restrict__ rho,
                                                      int k, int l, size_t n_polys) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_polys) return;
    int kl = k * l;
    size_t r = t / kl;
    int ij = (int)(t % kl);
    int32_t* out = a_hat + t * N;
    expand_a_poly(rho + r * 32, ij / l, ij % l, [&](int idx, uint32_t val) { out[idx] = (int32_t)val; });
}

cudaError_t launch_expand_a(int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l, int sm_count, cudaStream_t st) {
    size_t n_polys = n_rho * (size_t)(k * l);
    if (n_polys == 0) return cudaSuccess;
    unsigned grid = (unsigned)((n_polys + 63) / 64);
    expand_a_kernel<<<grid, 64, 0, st>>>(a_hat, rho, k, l, n_polys);
    return cudaGetLastError();
}

// ---- shared-A persistent kernel ----
// register budget: 2 CTAs/SM (<= 128 registers) for every level's shape, 1 for the 8 x 8 verification core.
// (3 CTAs/SM at 80 registers was measured 5-9 % slower for levels 2/3: the kernel is bound by the
// integer-multiply pipe, not by latency.)
constexpr int shared_min_ctas(int k, int l, int warps = 8) { return warps > 8 ? 1 : (k * l <= 56 ? 2 : 1); }

template <int K, int L, int WARPS, bool EXPAND, bool NTT_IN, bool INTT_OUT, bool W1 = false, int ZBITS = 0, int BETA = 0>
__global__ void __launch_bounds__(WARPS * 32, shared_min_ctas(K, L, WARPS)) matvec_shared_kernel(int32_t* __restrict__ w, const int32_t* __restrict__ a_hat,
                                                                   const uint8_t* __restrict__ rho,
                                                                   const int32_t* __restrict__ v, uint32_t batch,
                                                                   uint32_t* __restrict__ work_ctr,
                                                                   uint8_t* __restrict__ w1p = nullptr,
                                                                   const uint32_t* __restrict__ batch_dev = nullptr,
                                                                   const uint32_t* __restrict__ hmask = nullptr,
                                                                   const uint8_t* __restrict__ zp = nullptr,
                                                                   uint32_t* __restrict__ bad_flags = nullptr) {
    // hmask != nullptr (verification, W1): the packed output is w1' = UseHint(h, w) and w is not stored (matvec_core.cuh)
    // ZBITS > 0 (verification): the first L - 1 inputs are read from the packed z of the signature (zp), only the challenge
    // polynomial comes from v; items violating ||z|| < gamma1 - beta are flagged in bad_flags
    if (batch_dev != nullptr) batch = *batch_dev;   // round loop of batched signing: the size lives on the device
    constexpr int W1_ROW = K * (K == 4 ? 192 : 128);   // packed w1 bytes per item
    extern __shared__ __align__(16) uint32_t smem_u32v[];
    uint32_t* a_sm = smem_u32v;                               // K*L*A_STRIDE
    uint32_t* scr_all = smem_u32v + K * L * A_STRIDE;         // WARPS*SCRATCH_WORDS
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // When the output is inverse-transformed the matrix is stored pre-multiplied by 256^-1 (once per CTA),
    // which removes the scaling multiplications from every inverse transform (ntt_inv_warp<true>).
    auto scale = [](uint32_t v) -> uint32_t { return INTT_OUT ? mul_full(v, INV256) : v; };
    if (work_ctr != nullptr) {   // a CTA that starts when every item is already claimed leaves at once (uniform decision)
        __shared__ uint32_t late;
        if (threadIdx.x == 0) late = *reinterpret_cast<volatile uint32_t*>(work_ctr) >= batch;
        __syncthreads();
        if (late) return;
    }
    if constexpr (EXPAND) {
        for (int t = threadIdx.x; t < K * L; t += blockDim.x) {
            uint32_t* out = a_sm + t * A_STRIDE;
            expand_a_poly(rho, t / L, t % L, [&](int idx, uint32_t val) { out[idx] = scale(val); });
        }
    } else {
        for (int t = threadIdx.x; t < K * L * (N / 4); t += blockDim.x) {
            int p = t >> 6, c = t & 63;
            int4 q = __ldg(reinterpret_cast<const int4*>(a_hat) + t);
            reinterpret_cast<uint4*>(a_sm + p * A_STRIDE)[c] =
                make_uint4(scale(canon_signed(q.x)), scale(canon_signed(q.y)), scale(canon_signed(q.z)), scale(canon_signed(q.w)));
        }
    }
    __syncthreads();
    uint32_t* scr = scr_all + warp * SCRATCH_WORDS;
    // With a work counter (zeroed by the caller before the launch) items are claimed dynamically, one atomic
    // per item, issued before the current item is processed so its latency is hidden.  Dynamic claiming keeps
    // the kernel balanced when SMs run at different speeds or are partly taken by other streams' kernels (a
    // CTA that starts late finds the counter exhausted and leaves).  Without a counter the distribution is
    // the static warp-stride one.
    if (work_ctr != nullptr) {
        // items are claimed TWO ahead, so that the next item's inputs can be pulled into L2 (register-free prefetches) while
        // the current one is transformed
        uint32_t claim = 0, next = 0;
        if (lane == 0) {
            next = atomicAdd(work_ctr, 2u);
            claim = next + 1;
        }
        uint32_t item = __shfl_sync(0xffffffffu, next, 0);
        next = __shfl_sync(0xffffffffu, claim, 0);
        while (item < batch) {
            if (lane == 0) claim = atomicAdd(work_ctr, 1u);
            if (next < batch) {
                const char* vl = reinterpret_cast<const char*>(v + (size_t)next * L * N) + 128 * lane;
#pragma unroll
                for (int q = 0; q < (L * N * 4 + 4095) / 4096; q++)
                    if (4096 * q + 128 * lane < L * N * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(vl + 4096 * q));
            }
            item_core<K, L, NTT_IN, INTT_OUT, false, W1>(w + (size_t)item * K * N, v + (size_t)item * L * N, a_sm, scr, lane, nullptr,
                                                         W1 ? w1p + (size_t)item * W1_ROW : nullptr);
            item = next;
            next = __shfl_sync(0xffffffffu, claim, 0);
        }
    } else {
        for (uint32_t item = blockIdx.x * WARPS + warp; item < batch; item += gridDim.x * WARPS) {
            if constexpr (ZBITS > 0) {
                static_assert(NTT_IN && INTT_OUT && !EXPAND, "packed-z inputs belong to the verification core");
                uint32_t yh[L][8];
                const bool bad = item_inputs_zpacked<L - 1, ZBITS, BETA>(yh, zp + (size_t)item * (L - 1) * (32 * ZBITS),
                                                                         v + (size_t)item * L * N + (L - 1) * N, scr, lane);
                if (bad && lane == 0) bad_flags[item] = 1;
                item_rows<K, L, INTT_OUT, false, W1>(w + (size_t)item * K * N, yh, a_sm, scr, lane, nullptr,
                                                     W1 ? w1p + (size_t)item * W1_ROW : nullptr, 0, K, 0,
                                                     W1 && hmask != nullptr ? hmask + (size_t)item * K * 8 : nullptr);
            } else {
                item_core<K, L, NTT_IN, INTT_OUT, false, W1>(w + (size_t)item * K * N, v + (size_t)item * L * N, a_sm, scr, lane, nullptr,
                                                             W1 ? w1p + (size_t)item * W1_ROW : nullptr, nullptr, 0,
                                                             W1 && hmask != nullptr ? hmask + (size_t)item * K * 8 : nullptr);
            }
        }
    }
}

// ---- per-item-rho kernel: G items per CTA ----
// G*K*L threads expand the G matrices (one Keccak state per thread; G is chosen so that the Keccak
// phase fills whole warps: level 2 has only 16 polynomials per item), then each warp of the CTA runs
// the per-item core for its share of the G items.
template <int K, int L>
constexpr int item_group() { return K * L == 16 ? 4 : 1; }
// level 5 (k*l = 56 Keccak threads = two warps for one item): both warps share the item's core
template <int K, int L, int G, bool NTT_IN>
__host__ __device__ constexpr bool item_split() { return G == 1 && (K * L + 31) / 32 == 2 && K % 2 == 0 && NTT_IN; }
template <int K, int L, int G, bool NTT_IN, bool EXTRA>
constexpr size_t item_smem_bytes() {
    constexpr int NW = (G * K * L + 31) / 32;
    return (size_t)(G * K * L * A_STRIDE + NW * SCRATCH_WORDS + (item_split<K, L, G, NTT_IN>() ? (L + (EXTRA ? 1 : 0)) * N : 0)) * 4;
}

template <int K, int L, int G, bool NTT_IN, bool INTT_OUT, bool EXTRA = false, bool W1 = false>
__global__ void __launch_bounds__(((G * K * L + 31) / 32) * 32) matvec_item_kernel(int32_t* __restrict__ w,
                                                                                 const uint8_t* __restrict__ rho,
                                                                                 const int32_t* __restrict__ v, uint32_t batch,
                                                                                 const int32_t* __restrict__ extra = nullptr,
                                                                                 uint8_t* __restrict__ w1p = nullptr,
                                                                                 const uint32_t* __restrict__ hmask = nullptr) {
    // W1 (per-key verification): the output is the packed w1' = UseHint(h, w) of every item, w is not stored
    constexpr int W1_ROW = K * (K == 4 ? 192 : 128);
    extern __shared__ __align__(16) uint32_t smem_u32v[];
    constexpr int NW = (G * K * L + 31) / 32;
    constexpr bool SPLIT = item_split<K, L, G, NTT_IN>();         // two warps share one item's core
    uint32_t* a_sm = smem_u32v;                                   // G * K*L * A_STRIDE
    uint32_t* scr_all = smem_u32v + G * K * L * A_STRIDE;         // NW * SCRATCH_WORDS
    uint32_t* yh_sm = scr_all + NW * SCRATCH_WORDS;               // SPLIT: (L [+1]) * 256 transformed inputs
    const int t = threadIdx.x;
    const size_t item0 = (size_t)blockIdx.x * G;
    if (t < G * K * L) {
        const int g = t / (K * L), ij = t % (K * L);
        if (item0 + g < batch) {
            uint32_t* out = a_sm + t * A_STRIDE;
            expand_a_poly(rho + (item0 + g) * 32, ij / L, ij % L, [&](int idx, uint32_t val) { out[idx] = val; });   // unscaled
        }
    }
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    if constexpr (SPLIT) {
        if (item0 < batch)   // uniform for the CTA: both warps enter (the core synchronises the CTA once)
            item_core<K, L, NTT_IN, INTT_OUT, EXTRA, W1, true, false>(w + item0 * K * N, v + item0 * (L + (EXTRA ? 1 : 0)) * N, a_sm,
                                                                 scr_all + warp * SCRATCH_WORDS, lane,
                                                                 EXTRA ? extra + item0 * K * N : nullptr, W1 ? w1p + item0 * W1_ROW : nullptr,
                                                                 yh_sm, warp, W1 ? hmask + item0 * K * 8 : nullptr);
    } else {
    for (int g = warp; g < G; g += NW) {
        const size_t item = item0 + g;
        if (item < batch)
            item_core<K, L, NTT_IN, INTT_OUT, EXTRA, W1, false, false>(w + item * K * N, v + item * (L + (EXTRA ? 1 : 0)) * N,
                                                     a_sm + g * K * L * A_STRIDE, scr_all + warp * SCRATCH_WORDS, lane,
                                                     EXTRA ? extra + item * K * N : nullptr, W1 ? w1p + item * W1_ROW : nullptr, nullptr, 0,
                                                     W1 ? hmask + item * K * 8 : nullptr);
    }
    }
}

// ---- per-item-rho kernel, row-streamed: G = 8 items per CTA, A generated R rows at a time ----
// The per-item matrix is the Keccak-bound part of key generation and per-key verification (>= 5 permutations for each
// of the k*l polynomials of EVERY item, gen_a_ext.v:100-116 / rejection_a.v:67-112).  One item's whole matrix in
// shared memory (matvec_item_kernel above) costs 16-56 KiB per item and leaves 6-7 warps per SM, most of them waiting
// for their CTA's few Keccak threads.  Here a CTA owns 8 items and streams their matrices through shared memory R rows
// at a time: G*l*R = 128 / 120 / 112 threads (levels 2 / 3 / 5) - four warps, one per SM sub-partition, every lane a
// Keccak state - expand rows [i, i+R) of all 8 items, then the CTA's 8 warps (one per item, its transformed inputs in
// registers) multiply-accumulate those rows, reduce, inverse-transform and store them, and the next R rows follow.
// A never exists anywhere but in this R-row window; the Keccak phase, which is the bound, runs with full lanes on all
// four ALU pipes for ~80 % of the CTA's life instead of ~50 %.
template <int K, int L>
__host__ __device__ constexpr int rows_per_step() { return K * L == 16 ? 4 : (K * L == 30 ? 3 : 2); }
constexpr int ROWS_G = 8;   // items (= consumer warps) per CTA

template <int K, int L, bool NTT_IN, bool INTT_OUT, bool EXTRA>
__global__ void __launch_bounds__(ROWS_G * 32, 1) matvec_item_rows_kernel(int32_t* __restrict__ w, const uint8_t* __restrict__ rho,
                                                                          const int32_t* __restrict__ v, uint32_t batch,
                                                                          const int32_t* __restrict__ extra) {
    constexpr int R = rows_per_step<K, L>();
    constexpr int LI = L + (EXTRA ? 1 : 0);                      // input polynomials per item
    static_assert(K % R == 0 && ROWS_G * L * R <= ROWS_G * 32, "row window must tile the matrix and fit the CTA's threads");
    extern __shared__ __align__(16) uint32_t smem_u32v[];
    uint32_t* rows_sm = smem_u32v;                               // [R][G][L] polynomials, stride A_STRIDE
    uint32_t* scr_all = smem_u32v + R * ROWS_G * L * A_STRIDE;   // G * SCRATCH_WORDS
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint32_t* scr = scr_all + warp * SCRATCH_WORDS;
    for (size_t item0 = (size_t)blockIdx.x * ROWS_G; item0 < batch; item0 += (size_t)gridDim.x * ROWS_G) {
        const size_t item = item0 + warp;
        const bool live = item < batch;                          // warp-uniform
        uint32_t yh[LI][8];
        if (live) item_inputs<LI, NTT_IN>(yh, v + item * LI * N, scr, lane);
#pragma unroll 1
        for (int i0 = 0; i0 < K; i0 += R) {
            __syncthreads();                                     // the previous window has been consumed
            if (t < R * ROWS_G * L) {
                const int r = t / (ROWS_G * L), g = (t / L) % ROWS_G, j = t % L;
                if (item0 + g < batch) {
                    uint32_t* out = rows_sm + t * A_STRIDE;      // t == (r * G + g) * L + j
                    expand_a_poly(rho + (item0 + g) * 32, i0 + r, j, [&](int idx, uint32_t val) { out[idx] = val; });   // unscaled
                }
            }
            __syncthreads();
            if (live) {
#pragma unroll 1
                for (int r = 0; r < R; r++)
                    item_rows<K, L, INTT_OUT, EXTRA, false, false>(w + item * K * N, yh, rows_sm + (size_t)(r * ROWS_G + warp) * L * A_STRIDE, scr, lane,
                                                           EXTRA ? extra + item * K * N : nullptr, nullptr, i0 + r, i0 + r + 1, i0 + r);
            }
        }
    }
}

template <int K, int L, bool NTT_IN, bool INTT_OUT, bool EXTRA>
static cudaError_t launch_item_rows_t(int32_t* w, const uint8_t* rho, const int32_t* v, const int32_t* extra, size_t batch, int sm_count,
                                      cudaStream_t st) {
    constexpr int R = rows_per_step<K, L>();
    constexpr size_t smem = (size_t)(R * ROWS_G * L * A_STRIDE + ROWS_G * SCRATCH_WORDS) * 4;
    static_assert(smem <= 227 * 1024, "row window does not fit in shared memory");
    auto kern = matvec_item_rows_kernel<K, L, NTT_IN, INTT_OUT, EXTRA>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, configured); e != cudaSuccess) return e;
    const size_t want = (batch + ROWS_G - 1) / ROWS_G, cap = (size_t)sm_count * 8;
    kern<<<(unsigned)(want < cap ? want : cap), ROWS_G * 32, smem, st>>>(w, rho, v, (uint32_t)batch, extra);
    return cudaGetLastError();
}

template <int K, int L>
constexpr size_t shared_smem_bytes(int warps) {
    return (size_t)(K * L * A_STRIDE + warps * SCRATCH_WORDS) * 4;
}

template <int K, int L, int WARPS, bool EXPAND, bool NTT_IN, bool INTT_OUT, int MAX_CTAS = 8, bool W1 = false, int ZBITS = 0, int BETA = 0>
static cudaError_t launch_shared_t(int32_t* w, const int32_t* a_hat, const uint8_t* rho, const int32_t* v, size_t batch,
                                   int sm_count, cudaStream_t st, uint32_t* work_ctr = nullptr, uint8_t* w1p = nullptr,
                                   const uint32_t* batch_dev = nullptr, const uint32_t* hmask = nullptr, const uint8_t* zp = nullptr,
                                   uint32_t* bad_flags = nullptr) {
    auto kern = matvec_shared_kernel<K, L, WARPS, EXPAND, NTT_IN, INTT_OUT, W1, ZBITS, BETA>;
    constexpr size_t smem = shared_smem_bytes<K, L>(WARPS);
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    int ctas_per_sm = (int)((220 * 1024) / smem);
    if (ctas_per_sm > shared_min_ctas(K, L, WARPS)) ctas_per_sm = shared_min_ctas(K, L, WARPS);
    if (ctas_per_sm > MAX_CTAS) ctas_per_sm = MAX_CTAS;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    size_t want = (batch + WARPS - 1) / WARPS;
    size_t cap = (size_t)sm_count * ctas_per_sm;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid, WARPS * 32, smem, st>>>(w, a_hat, rho, v, (uint32_t)batch, work_ctr, w1p, batch_dev, hmask, zp, bad_flags);
    return cudaGetLastError();
}

template <int K, int L, bool EXPAND>
static cudaError_t launch_shared_flags(int32_t* w, const int32_t* a_hat, const uint8_t* rho, const int32_t* v, size_t batch,
                                       bool ntt_in, bool intt_out, int sm_count, cudaStream_t st,
                                       uint32_t* work_ctr = nullptr) {
    constexpr int WARPS = 8;
    // one 16-warp CTA per SM measured 3-5 % faster than two 8-warp CTAs for the fully fused shape
    if (ntt_in && intt_out) return launch_shared_t<K, L, 16, EXPAND, true, true>(w, a_hat, rho, v, batch, sm_count, st, work_ctr);
    if (ntt_in) return launch_shared_t<K, L, WARPS, EXPAND, true, false>(w, a_hat, rho, v, batch, sm_count, st);
    if (intt_out) return launch_shared_t<K, L, WARPS, EXPAND, false, true>(w, a_hat, rho, v, batch, sm_count, st);
    return launch_shared_t<K, L, WARPS, EXPAND, false, false>(w, a_hat, rho, v, batch, sm_count, st);
}

template <int K, int L, bool NTT_IN, bool INTT_OUT>
static cudaError_t launch_item_t(int32_t* w, const uint8_t* rho, const int32_t* v, size_t batch, cudaStream_t st) {
    constexpr int G = item_group<K, L>();
    constexpr int NW = (G * K * L + 31) / 32;
    auto kern = matvec_item_kernel<K, L, G, NTT_IN, INTT_OUT>;
    constexpr size_t smem = item_smem_bytes<K, L, G, NTT_IN, false>();
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    constexpr int threads = NW * 32;
    kern<<<(unsigned)((batch + G - 1) / G), threads, smem, st>>>(w, rho, v, (uint32_t)batch, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}

// Batches of at least this many items use the row-streamed per-item kernel.  Measured on B200 (profiles/r2b_per_item_ab.txt)
// the one-CTA-per-item kernel is 5-10 % faster at every level, so the row-streamed variant is off unless
// dil_diag_item_rows_threshold selects it (A/B measurements).
static std::atomic<size_t> g_item_rows_min{~(size_t)0};
void set_item_rows_threshold(size_t n) { g_item_rows_min.store(n); }

template <int K, int L>
static cudaError_t launch_item_flags(int32_t* w, const uint8_t* rho, const int32_t* v, size_t batch, bool ntt_in,
                                     bool intt_out, int sm_count, cudaStream_t st) {
    // the fully fused shape (key generation, cfg3 mode P) streams A row windows; small batches keep one CTA per item
    if (ntt_in && intt_out && batch >= g_item_rows_min.load()) return launch_item_rows_t<K, L, true, true, false>(w, rho, v, nullptr, batch, sm_count, st);
    if (ntt_in && intt_out) return launch_item_t<K, L, true, true>(w, rho, v, batch, st);
    if (ntt_in) return launch_item_t<K, L, true, false>(w, rho, v, batch, st);
    if (intt_out) return launch_item_t<K, L, false, true>(w, rho, v, batch, st);
    return launch_item_t<K, L, false, false>(w, rho, v, batch, st);
}

cudaError_t launch_matvec_expand(int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch,
                                 unsigned flags, int sm_count, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    const bool per_item = flags & DIL_RHO_PER_ITEM, ni = flags & DIL_NTT_INPUT, io = flags & DIL_INTT_OUTPUT;
    if (per_item) {
        if (k == 4 && l == 4) return launch_item_flags<4, 4>(w, rho, v, batch, ni, io, sm_count, st);
        if (k == 6 && l == 5) return launch_item_flags<6, 5>(w, rho, v, batch, ni, io, sm_count, st);
        if (k == 8 && l == 7) return launch_item_flags<8, 7>(w, rho, v, batch, ni, io, sm_count, st);
    } else {
        if (k == 4 && l == 4) return launch_shared_flags<4, 4, true>(w, nullptr, rho, v, batch, ni, io, sm_count, st);
        if (k == 6 && l == 5) return launch_shared_flags<6, 5, true>(w, nullptr, rho, v, batch, ni, io, sm_count, st);
        if (k == 8 && l == 7) return launch_shared_flags<8, 7, true>(w, nullptr, rho, v, batch, ni, io, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// w1p != nullptr: also emit the packed w1 = HighBits(w) per item (signing)
template <int K, int L>
static cudaError_t launch_signcore_t(int32_t* w, const int32_t* a_hat, const int32_t* y, size_t batch, int sm_count, cudaStream_t st,
                                     uint32_t* work_ctr, uint8_t* w1p, const uint32_t* batch_dev) {
    if (w1p != nullptr)
        return launch_shared_t<K, L, 16, false, true, true, 8, true>(w, a_hat, nullptr, y, batch, sm_count, st, work_ctr, w1p, batch_dev);
    return launch_shared_t<K, L, 16, false, true, true>(w, a_hat, nullptr, y, batch, sm_count, st, work_ctr, nullptr, batch_dev);
}

cudaError_t launch_signcore(int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch, int sm_count,
                            cudaStream_t st, uint32_t* work_ctr, uint8_t* w1p, const uint32_t* batch_dev) {
    if (batch == 0) return cudaSuccess;
    if (k == 4 && l == 4) return launch_signcore_t<4, 4>(w, a_hat, y, batch, sm_count, st, work_ctr, w1p, batch_dev);
    if (k == 6 && l == 5) return launch_signcore_t<6, 5>(w, a_hat, y, batch, sm_count, st, work_ctr, w1p, batch_dev);
    if (k == 8 && l == 7) return launch_signcore_t<8, 7>(w, a_hat, y, batch, sm_count, st, work_ctr, w1p, batch_dev);
    return cudaErrorInvalidValue;
}


// verification core with per-item public keys: A from rho[item] on chip, extra column t1neg_hat[item] from HBM
template <int K, int L>
static cudaError_t launch_verify_item_t(int32_t* w, const uint8_t* rho, const int32_t* v, const int32_t* extra, size_t batch,
                                        cudaStream_t st, uint8_t* w1p, const uint32_t* hmask) {
    constexpr int G = item_group<K, L>();
    constexpr int NW = (G * K * L + 31) / 32;
    constexpr size_t smem = item_smem_bytes<K, L, G, true, true>();
    static std::atomic<uint64_t> configured{0}, configured_w1{0};   // one bit per device: the attribute is per device
    if (w1p != nullptr && hmask != nullptr) {   // fused UseHint + w1 packing: w is not written
        auto kern = matvec_item_kernel<K, L, G, true, true, true, true>;
        if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured_w1); e != cudaSuccess) return e;
        kern<<<(unsigned)((batch + G - 1) / G), NW * 32, smem, st>>>(w, rho, v, (uint32_t)batch, extra, w1p, hmask);
        return cudaGetLastError();
    }
    auto kern = matvec_item_kernel<K, L, G, true, true, true>;
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    kern<<<(unsigned)((batch + G - 1) / G), NW * 32, smem, st>>>(w, rho, v, (uint32_t)batch, extra, nullptr, nullptr);
    return cudaGetLastError();
}
// w1p / hmask non-null: the core also applies the hints and emits the packed w1' (no usehint_pack pass, w is not written);
// *fused tells the caller whether that happened (the row-streamed A/B kernel keeps the separate pass)
cudaError_t launch_verify_core_item(int32_t* w, const uint8_t* rho, const int32_t* v, const int32_t* extra, int level,
                                    size_t batch, cudaStream_t st, uint8_t* w1p, const uint32_t* hmask, bool* fused) {
    if (fused) *fused = false;
    if (batch == 0) return cudaSuccess;
    if (batch >= g_item_rows_min.load()) {   // row-streamed kernel (8 items per CTA); tiny batches keep one CTA per item
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        switch (level) {
            case 2: return launch_item_rows_t<4, 4, true, true, true>(w, rho, v, extra, batch, sms, st);
            case 3: return launch_item_rows_t<6, 5, true, true, true>(w, rho, v, extra, batch, sms, st);
            case 5: return launch_item_rows_t<8, 7, true, true, true>(w, rho, v, extra, batch, sms, st);
        }
        return cudaErrorInvalidValue;
    }
    if (fused) *fused = w1p != nullptr && hmask != nullptr;
    switch (level) {
        case 2: return launch_verify_item_t<4, 4>(w, rho, v, extra, batch, st, w1p, hmask);
        case 3: return launch_verify_item_t<6, 5>(w, rho, v, extra, batch, st, w1p, hmask);
        case 5: return launch_verify_item_t<8, 7>(w, rho, v, extra, batch, st, w1p, hmask);
    }
    return cudaErrorInvalidValue;
}

// verification core: k x (l+1) matrix [A_hat | -t1_hat*2^13], inputs [z_0..z_{l-1}, c] in the time domain
cudaError_t launch_verify_core(int32_t* w, const int32_t* a_ext, const int32_t* v, int level, size_t batch, int sm_count,
                               cudaStream_t st, uint8_t* w1p, const uint32_t* hmask, const uint8_t* zp, uint32_t* bad_flags) {
    if (batch == 0) return cudaSuccess;
    if (w1p != nullptr && hmask != nullptr && zp != nullptr && bad_flags != nullptr) {   // packed z in, packed w1' out (levels 2, 3)
        switch (level) {
            case 2: return launch_shared_t<4, 5, 8, false, true, true, 8, true, 18, 78>(w, a_ext, nullptr, v, batch, sm_count, st, nullptr, w1p, nullptr, hmask, zp, bad_flags);
            case 3: return launch_shared_t<6, 6, 8, false, true, true, 8, true, 20, 196>(w, a_ext, nullptr, v, batch, sm_count, st, nullptr, w1p, nullptr, hmask, zp, bad_flags);
        }
        return cudaErrorInvalidValue;
    }
    if (w1p != nullptr && hmask != nullptr) {   // fused UseHint + w1 packing: w is not written
        switch (level) {
            case 2: return launch_shared_t<4, 5, 8, false, true, true, 8, true>(w, a_ext, nullptr, v, batch, sm_count, st, nullptr, w1p, nullptr, hmask);
            case 3: return launch_shared_t<6, 6, 8, false, true, true, 8, true>(w, a_ext, nullptr, v, batch, sm_count, st, nullptr, w1p, nullptr, hmask);
            case 5: return launch_shared_t<8, 8, 8, false, true, true, 8, true>(w, a_ext, nullptr, v, batch, sm_count, st, nullptr, w1p, nullptr, hmask);
        }
        return cudaErrorInvalidValue;
    }
    switch (level) {
        case 2: return launch_shared_t<4, 5, 8, false, true, true>(w, a_ext, nullptr, v, batch, sm_count, st);
        case 3: return launch_shared_t<6, 6, 8, false, true, true>(w, a_ext, nullptr, v, batch, sm_count, st);
        case 5: return launch_shared_t<8, 8, 16, false, true, true>(w, a_ext, nullptr, v, batch, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace dil
