// mask_core.cu — ExpandMask fused in front of the signing core:
//     y = ExpandMask(rho', kappa)   ->   w = INTT(A_hat * NTT(y)),  w1 = HighBits(w)
// in ONE persistent kernel.  Reference data flow: GENY (expandmask_ext.v:98, :284-294, sampler_y_ext.v:101-131,
// rejection_y.v:76-99) -> NTT_Y -> MULT_A_Y -> NTTI_W (combined_top.v:1830-1933).  The FPGA overlaps its Keccak core
// with the butterfly datapath the same way: the sampler fills a FIFO while the NTT unit drains it.
//
// Why fuse: ExpandMask is pure Keccak (integer ALU pipe: LOP3 / SHF), the core is bound by the integer-multiply
// pipe (IMAD on the FMA pipe) and leaves half of the ALU pipe idle.  As separate kernels the two pipes take turns;
// here every SM runs both at once, and the squeezed mask bytes go from the Keccak threads to the transform warps
// through shared memory instead of a 4 KiB-per-attempt round trip through HBM.
//
// Per CTA (16 warps, one CTA per SM) a small scheduler in shared memory hands out two kinds of tasks:
//   PRODUCE  one warp, one Keccak state per lane: the l polynomials of G = 32/l consecutive attempt slots (5 SHAKE-256
//            blocks each) squeezed into a ring buffer in shared memory, still bit-packed (18 / 20 bits per coefficient);
//   CONSUME  one warp, one slot: unpack the slot's l polynomials (y also goes to HBM: the tail needs it for
//            z = y + c*s1), l forward transforms, k rows of {multiply-accumulate against A_hat in shared memory,
//            Barrett, inverse transform, store w, pack HighBits(w)}.
// Any warp takes either kind (producing is preferred while a buffer is free and fewer than MAXP warps produce), so the
// mix adapts to the level and to the other kernels sharing the GPU; slot groups are claimed from the round's work
// counter, so CTAs balance dynamically as in the unfused kernels.
#include <cuda_runtime.h>

#include "keccak.cuh"
#include "kernels.h"
#include "matvec_core.cuh"

namespace dil {

namespace {

enum : int { BUF_FREE = 0, BUF_PRODUCING = 1, BUF_READY = 2 };
enum : int { TASK_WAIT = 0, TASK_PRODUCE = 1, TASK_CONSUME = 2, TASK_EXIT = 3 };

template <int NBUF>
struct MaskSched {
    int state[NBUF];       // BUF_*
    int next[NBUF];        // next unclaimed slot of a READY buffer
    int remaining[NBUF];   // slots of the buffer not yet finished
    int count[NBUF];       // slots in the buffer
    uint32_t base[NBUF];   // first attempt slot of the buffer's group
    int producing;         // warps currently in a PRODUCE task
    int all_claimed;       // the round's work counter is exhausted
};

// 8 consecutive coefficients (BITS bytes) of a squeezed mask polynomial: y = gamma1 - r  (rejection_y.v:76-99)
template <int BITS>
__device__ __forceinline__ void unpack_mask8(const unsigned char* __restrict__ src, int32_t (&v)[8]) {
    constexpr int32_t G1 = 1 << (BITS - 1);
    uint64_t lo, mid;
    uint32_t hi;
    if constexpr (BITS == 18) {  // 18 bytes, 2-byte aligned
        const uint16_t* h = reinterpret_cast<const uint16_t*>(src);
        lo = (uint64_t)h[0] | ((uint64_t)h[1] << 16) | ((uint64_t)h[2] << 32) | ((uint64_t)h[3] << 48);
        mid = (uint64_t)h[4] | ((uint64_t)h[5] << 16) | ((uint64_t)h[6] << 32) | ((uint64_t)h[7] << 48);
        hi = h[8];
    } else {                     // 20 bytes, 4-byte aligned
        const uint32_t* h = reinterpret_cast<const uint32_t*>(src);
        lo = (uint64_t)h[0] | ((uint64_t)h[1] << 32);
        mid = (uint64_t)h[2] | ((uint64_t)h[3] << 32);
        hi = h[4];
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int pos = c * BITS;
        uint64_t bits;
        if (pos + BITS <= 64) bits = lo >> pos;
        else if (pos < 64) bits = (lo >> pos) | (mid << (64 - pos));
        else if (pos + BITS <= 128) bits = mid >> (pos - 64);
        else if (pos < 128) bits = (mid >> (pos - 64)) | ((uint64_t)hi << (128 - pos));
        else bits = hi >> (pos - 128);
        v[c] = G1 - (int32_t)((uint32_t)bits & ((1u << BITS) - 1));
    }
}

}  // namespace

template <int K, int L, int GAMMA1_BITS, int WARPS, int NBUF>
__global__ void __launch_bounds__(WARPS * 32, 1) mask_core_kernel(int32_t* __restrict__ w, uint8_t* __restrict__ w1p, int32_t* __restrict__ y,
                                                                const int32_t* __restrict__ a_hat, const uint64_t* __restrict__ rhop,
                                                                const uint16_t* __restrict__ kappa, const uint32_t* __restrict__ active0,
                                                                const uint32_t* __restrict__ active1, RoundCtl* __restrict__ ctl,
                                                                const int MAXP) {
    constexpr int BITS = GAMMA1_BITS + 1;                // 18 / 20 bits per coefficient
    constexpr int ZB = 32 * BITS;                         // squeezed bytes per polynomial: 576 / 640
    constexpr int ROW = ZB + 16;                          // row stride in the ring (skews banks, keeps 16-byte alignment)
    constexpr int LANES = ZB / 8;                         // 64-bit words per polynomial: 72 / 80
    constexpr int G = 32 / L;                             // slots per group: 8 / 6 / 4
    constexpr int BUF_BYTES = 32 * ROW;
    constexpr int W1_ROW = K * (K == 4 ? 192 : 128);      // packed w1 bytes per slot
    extern __shared__ __align__(16) uint32_t smem_u32v[];
    uint32_t* a_sm = smem_u32v;                                                   // K*L*A_STRIDE
    uint32_t* scr_all = a_sm + K * L * A_STRIDE;                                  // WARPS*SCRATCH_WORDS
    unsigned char* ring = reinterpret_cast<unsigned char*>(scr_all + WARPS * SCRATCH_WORDS);   // NBUF * BUF_BYTES
    __shared__ MaskSched<NBUF> sc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_slots = ctl->n_slots, spec = ctl->spec;
    const uint32_t n_groups = (n_slots + G - 1) / G;
    const uint32_t* __restrict__ active = ctl->cur ? active1 : active0;
    uint32_t* const work_ctr = &ctl->ctr_core;
    {   // a CTA that starts when every group is already claimed (or in a round without work) leaves at once
        __shared__ uint32_t late;
        if (threadIdx.x == 0) late = *reinterpret_cast<volatile uint32_t*>(work_ctr) >= n_groups;
        __syncthreads();
        if (late) return;
    }
    if (threadIdx.x < NBUF) {
        sc.state[threadIdx.x] = BUF_FREE;
        sc.next[threadIdx.x] = 0; sc.remaining[threadIdx.x] = 0; sc.count[threadIdx.x] = 0; sc.base[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) { sc.producing = 0; sc.all_claimed = 0; }
    // A_hat into shared memory, pre-multiplied by 256^-1 (the inverse transforms skip their scaling, ntt_inv_warp<true>)
    for (int t = threadIdx.x; t < K * L * (N / 4); t += blockDim.x) {
        const int p = t >> 6, c = t & 63;
        const int4 q = __ldg(reinterpret_cast<const int4*>(a_hat) + t);
        reinterpret_cast<uint4*>(a_sm + p * A_STRIDE)[c] =
            make_uint4(mul_full(canon_signed(q.x), INV256), mul_full(canon_signed(q.y), INV256), mul_full(canon_signed(q.z), INV256),
                       mul_full(canon_signed(q.w), INV256));
    }
    __syncthreads();
    uint32_t* scr = scr_all + warp * SCRATCH_WORDS;
    volatile int* vstate = sc.state;
    volatile int* vnext = sc.next;
    volatile int* vcount = sc.count;

    for (;;) {
        // ---- lane 0 picks the warp's next task ----
        int task = TASK_WAIT, buf = 0, sl = 0;
        uint32_t grp = 0;
        if (lane == 0) {
            if (!*reinterpret_cast<volatile int*>(&sc.all_claimed) && *reinterpret_cast<volatile int*>(&sc.producing) < MAXP) {
                if (atomicAdd(&sc.producing, 1) < MAXP) {
                    for (int b = 0; b < NBUF && task == TASK_WAIT; b++) {
                        if (vstate[b] == BUF_FREE && atomicCAS(&sc.state[b], BUF_FREE, BUF_PRODUCING) == BUF_FREE) {
                            grp = atomicAdd(work_ctr, 1u);
                            if (grp < n_groups) { task = TASK_PRODUCE; buf = b; }
                            else {
                                *reinterpret_cast<volatile int*>(&sc.all_claimed) = 1;
                                __threadfence_block();
                                vstate[b] = BUF_FREE;
                                break;
                            }
                        }
                    }
                }
                if (task != TASK_PRODUCE) atomicSub(&sc.producing, 1);
            }
            if (task == TASK_WAIT) {
                for (int t = 0; t < NBUF; t++) {
                    const int b = (warp + t) % NBUF;
                    if (vstate[b] == BUF_READY && vnext[b] < vcount[b]) {
                        const int s = atomicAdd(&sc.next[b], 1);
                        if (s < vcount[b]) { task = TASK_CONSUME; buf = b; sl = s; break; }
                    }
                }
            }
            if (task == TASK_WAIT && *reinterpret_cast<volatile int*>(&sc.all_claimed)) {
                bool busy = false;
                for (int b = 0; b < NBUF; b++) {
                    const int st = vstate[b];
                    busy |= st == BUF_PRODUCING || (st == BUF_READY && vnext[b] < vcount[b]);
                }
                if (!busy) task = TASK_EXIT;
            }
        }
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task == TASK_EXIT) break;
        if (task == TASK_WAIT) { __nanosleep(200); continue; }
        buf = __shfl_sync(0xffffffffu, buf, 0);
        unsigned char* rows = ring + (size_t)buf * BUF_BYTES;

        if (task == TASK_PRODUCE) {
            // ---- one Keccak state per lane: polynomial j of slot slot0 + s ----
            grp = __shfl_sync(0xffffffffu, grp, 0);
            const uint32_t slot0 = grp * G;
            const uint32_t cnt = n_slots - slot0 < (uint32_t)G ? n_slots - slot0 : (uint32_t)G;
            const uint32_t s = lane / L, j = lane % L;
            if (s < cnt) {
                const uint32_t a = slot0 + s;
                const uint32_t item = active[a / spec];
                const uint32_t nonce = (uint32_t)L * (kappa[item] + a % spec) + j;
                uint64_t A[25];
#pragma unroll
                for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) A[i] = rhop[(size_t)item * 8 + i];
                A[8] = (uint64_t)(nonce & 0xFFFF) | (0x1FULL << 16);
                A[16] = 0x80ULL << 56;
                uint64_t* row = reinterpret_cast<uint64_t*>(rows + (size_t)lane * ROW);
#pragma unroll
                for (int blk = 0; blk < (LANES + 16) / 17; blk++) {
                    keccak_f1600(A);
#pragma unroll
                    for (int i = 0; i < 17; i++)
                        if (blk * 17 + i < LANES) row[blk * 17 + i] = A[i];
                }
            }
            __syncwarp();
            if (lane == 0) {
                sc.base[buf] = slot0;
                sc.count[buf] = (int)cnt;
                sc.remaining[buf] = (int)cnt;
                sc.next[buf] = 0;
                __threadfence_block();
                vstate[buf] = BUF_READY;
                atomicSub(&sc.producing, 1);
            }
            continue;
        }

        // ---- CONSUME: slot `sl` of buffer `buf` ----
        sl = __shfl_sync(0xffffffffu, sl, 0);
        __threadfence_block();
        const uint32_t a = *reinterpret_cast<volatile uint32_t*>(&sc.base[buf]) + (uint32_t)sl;
        uint32_t yh[L][8];
        {
            FwdTw ftw;
            {   // 2 KiB table, L1 resident; loaded per slot (opaque pointer defeats hoisting) to keep registers low
                const TwTable* tab = &TW_FWD;
                asm volatile("" : "+l"(tab));
                load_fwd_tw(ftw, tab, lane);
            }
#pragma unroll
            for (int j = 0; j < L; j++) {
                int32_t v[8];
                unpack_mask8<BITS>(rows + (size_t)(sl * L + j) * ROW + lane * BITS, v);
                // y leaves for HBM in natural order (the tail adds c*s1 to it): 512 contiguous bytes per warp store
                int4* dst = reinterpret_cast<int4*>(y + ((size_t)a * L + j) * N) + 2 * lane;
                dst[0] = make_int4(v[0], v[1], v[2], v[3]);
                dst[1] = make_int4(v[4], v[5], v[6], v[7]);
                // natural order -> layout A (coefficient 32 r + lane) through the warp's scratch: coefficient i lives at
                // word 36 (i >> 5) + (i & 31), the addressing the transform itself uses between its phases
                uint4* sp = reinterpret_cast<uint4*>(scr + 36 * (lane >> 2) + 8 * (lane & 3));
                sp[0] = make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]);
                sp[1] = make_uint4((uint32_t)v[4], (uint32_t)v[5], (uint32_t)v[6], (uint32_t)v[7]);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 8; r++) yh[j][r] = scr[36 * r + lane];
                ntt_fwd_warp(yh[j], scr, ftw, lane);
                __syncwarp();
            }
        }
        // the packed rows of this slot are no longer needed: hand the buffer back as soon as its last slot got here
        __threadfence_block();
        if (lane == 0 && atomicSub(&sc.remaining[buf], 1) == 1) vstate[buf] = BUF_FREE;
        item_rows<K, L, true, false, true>(w + (size_t)a * K * N, yh, a_sm, scr, lane, nullptr, w1p + (size_t)a * W1_ROW, 0, K);
    }
}

template <int K, int L, int G1B, int NBUF>
static cudaError_t launch_mask_core_t(const SignBufs& b, const int32_t* a_hat, uint32_t cap_slots, int sm_count, cudaStream_t st, int maxp) {
    constexpr int WARPS = 16;
    constexpr int ROW = 32 * (G1B + 1) + 16;
    constexpr size_t smem = (size_t)(K * L * A_STRIDE + WARPS * SCRATCH_WORDS) * 4 + (size_t)NBUF * 32 * ROW;
    static_assert(smem + 1024 <= 227 * 1024, "mask core: shared memory budget");
    auto kern = mask_core_kernel<K, L, G1B, WARPS, NBUF>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, configured); e != cudaSuccess) return e;
    constexpr int G = 32 / L;
    const unsigned groups = (cap_slots + G - 1) / G;
    const unsigned want = (groups + 3) / 4;   // a CTA is worth launching for a handful of groups (MAXP of them are squeezed at once)
    const unsigned grid = want < (unsigned)sm_count ? (want ? want : 1) : (unsigned)sm_count;
    kern<<<grid, WARPS * 32, smem, st>>>(b.w, reinterpret_cast<uint8_t*>(b.w1p), b.y, a_hat, b.rhop, b.kappa, b.active[0], b.active[1], b.ctl,
                                         maxp > 0 ? maxp : NBUF);
    return cudaGetLastError();
}

// maxp: how many warps of a CTA may squeeze masks at the same time (0 = as many as there are ring buffers)
cudaError_t launch_mask_core(int level, const SignBufs& b, const int32_t* a_hat, uint32_t cap_slots, int sm_count, cudaStream_t st, int maxp) {
    if (cap_slots == 0) return cudaSuccess;
    switch (level) {
        case 2: return launch_mask_core_t<4, 4, 17, 8>(b, a_hat, cap_slots, sm_count, st, maxp);
        case 3: return launch_mask_core_t<6, 5, 19, 8>(b, a_hat, cap_slots, sm_count, st, maxp);
        case 5: return launch_mask_core_t<8, 7, 19, 6>(b, a_hat, cap_slots, sm_count, st, maxp);
    }
    return cudaErrorInvalidValue;
}

}  // namespace dil
