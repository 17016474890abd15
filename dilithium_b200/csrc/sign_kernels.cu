// sign_kernels.cu — device side of batched deterministic signing (Dilithium round 3.1), the
// caller of the polynomial-arithmetic hot path.
//
// Reference data flow (rtl_src/combined_top.v, mode 2): LOAD_MU (expandmask_ext.v:131-185:
// mu = SHAKE256(tr||M), rho' = SHAKE256(K||mu)) -> per attempt kappa: GENY (expandmask_ext.v:98,
// :284-294, rejection_y.v:76-99) -> NTT_Y -> MULT_A_Y -> NTTI_W (:1850-1933) -> DECOMP
// (decomp_map1.v, coeff_decomposer.v:80-88) + c~ = SHAKE256(mu||w1) (gen_c.v:163-191) -> GEN_C /
// SampleInBall (gen_c.v:192-222) -> NTT_C, c*s1, c*s2, c*t0, INTTs, norm checks
// (:2011-2179, norm_check.v:84-105), MAKEHINT (makehint.v:99-102), restart on reject (:2217-2228).
//
// The FPGA runs one signature at a time with speculative pipelining; here every kernel
// processes all still-active signatures of the batch for one attempt ("round"), and rejected
// items are compacted into the next round's active list.
#include <cuda_runtime.h>

#include "dil_params.h"
#include "keccak.cuh"
#include "kernels.h"
#include "ntt_core.cuh"
#include "rounding.cuh"

namespace dil {

// ---------------------------------------------------------------------------------------
// S0: mu = SHAKE256(tr || M)[0:64], rho' = SHAKE256(K || mu)[0:64]; one thread per item
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sign_init_kernel(uint64_t* __restrict__ mu, uint64_t* __restrict__ rhop,
                                                        uint16_t* __restrict__ kappa, const uint8_t* __restrict__ tr,
                                                        const uint8_t* __restrict__ key, size_t key_stride,
                                                        const uint8_t* __restrict__ msgs,
                                                        const uint64_t* __restrict__ offsets, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    tr += (size_t)t * key_stride;     // stride 0: one key for the batch; otherwise one (tr, K) record per item
    key += (size_t)t * key_stride;
    const uint8_t* m = msgs + offsets[t];
    // the host entry points validate the offsets; a non-monotonic pair in device memory reads as an empty message
    const size_t mlen = offsets[t + 1] >= offsets[t] ? (size_t)(offsets[t + 1] - offsets[t]) : 0;
    uint64_t A[25];
    shake256_absorb_lanes(A, 32 + mlen, [&](size_t idx) -> uint64_t {
        if (idx < 4) return load_lane_bytes(tr, idx * 8, 32);
        return load_lane_bytes(m, (idx - 4) * 8, mlen);
    });
    uint64_t muv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        muv[i] = A[i];
        mu[(size_t)t * 8 + i] = A[i];
    }
    // K (4 lanes) || mu (8 lanes) = 96 bytes: one block
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) A[i] = load_lane_bytes(key, i * 8, 32);
#pragma unroll
    for (int i = 0; i < 8; i++) A[4 + i] = muv[i];
    A[12] = 0x1F;
    A[16] = 0x80ULL << 56;
    keccak_f1600(A);
#pragma unroll
    for (int i = 0; i < 8; i++) rhop[(size_t)t * 8 + i] = A[i];
    kappa[t] = 0;
}

// ---------------------------------------------------------------------------------------
// Slots: in a round every active item owns `spec` consecutive slots a = idx*spec + s that try the
// attempts kappa+s speculatively (spec = 1 while the GPU is saturated by distinct items, up to 16
// in the straggler rounds); the smallest accepted kappa wins, exactly as sequential signing.
//
// S1: ExpandMask.  One thread per polynomial y[a][j] = SHAKE256(rho' || u16le(l*kappa + j)),
// squeezed into a per-warp shared staging area, then the warp unpacks the 32 polynomials
// cooperatively so that HBM sees coalesced 1 KiB polynomial writes.
// ---------------------------------------------------------------------------------------
template <int L, int GAMMA1_BITS>
__global__ void __launch_bounds__(128) expand_mask_kernel(int32_t* __restrict__ y, const uint64_t* __restrict__ rhop,
                                                          const uint16_t* __restrict__ kappa, const uint32_t* __restrict__ active0,
                                                          const uint32_t* __restrict__ active1, const RoundCtl* __restrict__ ctl) {
    const uint32_t n_slots = ctl->n_slots, spec = ctl->spec;
    const uint32_t* __restrict__ active = ctl->cur ? active1 : active0;
    constexpr int ZB = 32 * (GAMMA1_BITS + 1);       // packed bytes per poly: 576 / 640
    constexpr int ROW = ZB + 16;                      // row stride in shared memory
    constexpr int LANES = ZB / 8;                     // 72 / 80 lanes of output needed
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = sm_raw + (size_t)warp * 32 * ROW;
    const uint32_t n_polys = n_slots * L;
    const uint32_t wstride = gridDim.x * (blockDim.x >> 5) * 32;
    // grid-stride over 32-polynomial chunks so that the launch can be capped to a fixed number of CTAs per SM
    for (uint32_t wbase = (blockIdx.x * (blockDim.x >> 5) + warp) * 32; wbase < n_polys; wbase += wstride) {
    const uint32_t gid = wbase + lane;
    if (gid < n_polys) {
        const uint32_t a = gid / L, j = gid % L;
        const uint32_t item = active[a / spec];
        const uint32_t nonce = (uint32_t)L * (kappa[item] + a % spec) + j;
        uint64_t A[25];
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) A[i] = rhop[(size_t)item * 8 + i];
        A[8] = (uint64_t)(nonce & 0xFFFF) | (0x1FULL << 16);
        A[16] = 0x80ULL << 56;
        uint64_t* row = reinterpret_cast<uint64_t*>(stage + (size_t)lane * ROW);
#pragma unroll
        for (int blk = 0; blk < (LANES + 16) / 17; blk++) {
            keccak_f1600(A);
#pragma unroll
            for (int i = 0; i < 17; i++)
                if (blk * 17 + i < LANES) row[blk * 17 + i] = A[i];
        }
    }
    __syncwarp();
    // cooperative unpack: lane handles coefficients 8*lane .. 8*lane+7 of polynomial p
    constexpr int BITS = GAMMA1_BITS + 1;             // 18 / 20
    constexpr int CHUNK = BITS;                       // bytes per 8 coefficients
    constexpr int32_t G1 = 1 << GAMMA1_BITS;
    for (int p = 0; p < 32 && wbase + p < n_polys; p++) {
        // the lane's 18 / 20 bytes as five aligned 32-bit words (18-byte chunks of odd lanes start two bytes into a word:
        // one funnel shift per word lines them up), then one funnel shift, one mask and one subtraction per coefficient
        const unsigned off = lane * CHUNK;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(stage + (size_t)p * ROW + (off & ~3u));
        uint32_t wd[5];
#pragma unroll
        for (int i = 0; i < 5; i++) wd[i] = src[i];
        if constexpr (BITS == 18) {
            const unsigned s = (off & 2u) * 8;
#pragma unroll
            for (int i = 0; i < 4; i++) wd[i] = __funnelshift_r(wd[i], wd[i + 1], s);
            wd[4] >>= s;
        }
        int32_t v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int pos = c * BITS, wi = pos >> 5, sh = pos & 31;
            uint32_t x;
            if (sh + BITS <= 32) x = wd[wi] >> sh;
            else x = __funnelshift_r(wd[wi], wd[wi + 1], sh);
            if (sh + BITS != 32) x &= (1u << BITS) - 1;
            v[c] = G1 - (int32_t)x;
        }
        int4* dst = reinterpret_cast<int4*>(y + (size_t)(wbase + p) * N) + 2 * lane;
        dst[0] = make_int4(v[0], v[1], v[2], v[3]);
        dst[1] = make_int4(v[4], v[5], v[6], v[7]);
    }
    __syncwarp();   // staging area is reused by the next chunk
    }
    // the masks are secret (y together with z reveals s1): nothing of them stays behind in shared memory
    for (int i = lane; i < 32 * ROW / 16; i += 32) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------
// S4: c~ = SHAKE256(mu || w1_packed)[0:32]; c = SampleInBall(c~) (gen_c.v:163-222, :317-343).
// One thread per active item; c is written as 256 int8 in {-1,0,1}.
// ---------------------------------------------------------------------------------------
// SampleInBall walks the squeezed bytes and the challenge polynomial with data-dependent indices; both live in
// shared memory (a private row per thread) - as per-thread local arrays they cost an L1/L2 round trip per
// access and dominated the kernel.  The finished polynomials leave as coalesced 16-byte vectors.
constexpr int CH_THREADS = 64;
constexpr int CH_C_STRIDE = 272;     // 256 coefficient bytes + 16: rows stay 16-byte aligned, banks are skewed
constexpr int CH_BUF_STRIDE = 144;   // 136 squeezed bytes + 8
template <int K, int W1_BYTES, int TAU>
__global__ void __launch_bounds__(CH_THREADS) challenge_kernel(int8_t* __restrict__ c_out, uint64_t* __restrict__ ctilde,
                                                               const uint64_t* __restrict__ mu, const uint64_t* __restrict__ w1p,
                                                               const uint32_t* __restrict__ active0, const uint32_t* __restrict__ active1,
                                                               const RoundCtl* __restrict__ ctl) {
    const uint32_t n_slots = ctl->n_slots, spec = ctl->spec;
    const uint32_t* __restrict__ active = ctl->cur ? active1 : active0;
    __shared__ __align__(16) uint8_t c_sm[CH_THREADS * CH_C_STRIDE];
    __shared__ __align__(16) uint8_t buf_sm[CH_THREADS * CH_BUF_STRIDE];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* c = c_sm + threadIdx.x * CH_C_STRIDE;
    uint8_t* buf = buf_sm + threadIdx.x * CH_BUF_STRIDE;
    for (uint32_t base = blockIdx.x * blockDim.x + warp * 32; base < n_slots; base += gridDim.x * blockDim.x) {
        const uint32_t a = base + lane;
        if (a < n_slots) {
            const uint32_t item = active[a / spec];
            constexpr int W1_LANES = K * W1_BYTES / 8;
            const uint64_t* m = mu + (size_t)item * 8;
            const uint64_t* w = w1p + (size_t)a * W1_LANES;
            uint64_t A[25];
            shake256_absorb_lanes(A, 64 + K * W1_BYTES, [&](size_t idx) -> uint64_t { return idx < 8 ? m[idx] : w[idx - 8]; });
            uint64_t ct[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ct[i] = A[i];
                ctilde[(size_t)a * 4 + i] = A[i];
            }
            // SampleInBall: SHAKE256(c~): first 8 bytes = sign bits, then rejection bytes
#pragma unroll
            for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) A[i] = ct[i];
            A[4] = 0x1F;
            A[16] = 0x80ULL << 56;
            keccak_f1600(A);
            uint64_t signs = A[0];
#pragma unroll
            for (int i = 0; i < 17; i++) reinterpret_cast<uint64_t*>(buf)[i] = A[i];
#pragma unroll
            for (int i = 0; i < N / 16; i++) reinterpret_cast<uint4*>(c)[i] = make_uint4(0, 0, 0, 0);
            int pos = 8;
            for (int i = N - TAU; i < N; i++) {
                int b;
                do {
                    if (pos == 136) {
                        keccak_f1600(A);
#pragma unroll
                        for (int q = 0; q < 17; q++) reinterpret_cast<uint64_t*>(buf)[q] = A[q];
                        pos = 0;
                    }
                    b = buf[pos++];
                } while (b > i);
                c[i] = c[b];
                c[b] = (signs & 1) ? (uint8_t)0xFF : (uint8_t)1;
                signs >>= 1;
            }
        }
        __syncwarp();
        // the warp's 32 polynomials (256 bytes each, consecutive slots) leave as 16-byte vectors
        const uint32_t rows = n_slots - base < 32 ? n_slots - base : 32;
        uint4* dst = reinterpret_cast<uint4*>(c_out + (size_t)base * N);
        const uint8_t* src = c_sm + (size_t)warp * 32 * CH_C_STRIDE;
        for (uint32_t t = lane; t < rows * (N / 16); t += 32)
            dst[t] = *reinterpret_cast<const uint4*>(src + (t >> 4) * CH_C_STRIDE + (t & 15) * 16);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// Accepted slot `a` of `item` (its s-th speculative attempt) becomes the signature: z is packed
// (encoder.v:96-133: gamma1 - z, 18 or 20 bits) through a per-warp shared staging row and leaves as
// 16-byte vectors, h and c~ are copied, the attempt count is recorded and the item joins the
// completion-ordered done list (RoundCtl::done counts finished items over the whole batch; the host path
// drains each round's finished signatures while later rounds still sign).  Executed by one warp.
// ---------------------------------------------------------------------------------------
template <int L, int GAMMA1_BITS, int HB, bool DIRECT = false>
__device__ __forceinline__ void resolve_finish(uint8_t* __restrict__ zp, uint8_t* __restrict__ h_out, uint64_t* __restrict__ ct_out,
                                               uint32_t* __restrict__ attempts, const uint16_t* __restrict__ kappa,
                                               uint32_t* __restrict__ done_ctr, const int32_t* zslot,
                                               const uint8_t* h_slot, const uint64_t* __restrict__ ct_slot,
                                               uint32_t* __restrict__ done_list, uint32_t item, uint32_t a, uint32_t s,
                                               uint8_t* dstp, int lane) {
    constexpr int BITS = GAMMA1_BITS + 1;
    constexpr int32_t G1 = 1 << GAMMA1_BITS;
    constexpr int ZB = L * 32 * BITS;
    const int4* src = reinterpret_cast<const int4*>(zslot + (size_t)a * L * N);
    for (int g = lane; g < L * 32; g += 32) {
        int4 va = src[2 * g], vb = src[2 * g + 1];
        uint32_t v[8] = {(uint32_t)(G1 - va.x), (uint32_t)(G1 - va.y), (uint32_t)(G1 - va.z), (uint32_t)(G1 - va.w),
                         (uint32_t)(G1 - vb.x), (uint32_t)(G1 - vb.y), (uint32_t)(G1 - vb.z), (uint32_t)(G1 - vb.w)};
        uint64_t lo = 0, mid = 0;
        uint32_t hi = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int pos = c * BITS;
            uint64_t x = v[c] & ((1u << BITS) - 1);
            if (pos < 64) {
                lo |= x << pos;
                if (pos + BITS > 64) mid |= x >> (64 - pos);
            } else if (pos < 128) {
                mid |= x << (pos - 64);
                if (pos + BITS > 128) hi |= (uint32_t)(x >> (128 - pos));
            } else {
                hi |= (uint32_t)x << (pos - 128);
            }
        }
        // DIRECT: no staging row, the 18/20-byte pieces go straight to the signature (2-/4-byte stores)
        uint8_t* dst = (DIRECT ? zp + (size_t)item * ZB : dstp) + (size_t)g * BITS;
        if constexpr (BITS == 18) {
            uint16_t* d = reinterpret_cast<uint16_t*>(dst);
#pragma unroll
            for (int q = 0; q < 4; q++) d[q] = (uint16_t)(lo >> (16 * q));
#pragma unroll
            for (int q = 0; q < 4; q++) d[4 + q] = (uint16_t)(mid >> (16 * q));
            d[8] = (uint16_t)hi;
        } else {
            uint32_t* d = reinterpret_cast<uint32_t*>(dst);
            d[0] = (uint32_t)lo; d[1] = (uint32_t)(lo >> 32); d[2] = (uint32_t)mid; d[3] = (uint32_t)(mid >> 32); d[4] = hi;
        }
    }
    __syncwarp();
    if constexpr (!DIRECT) {
        uint4* out = reinterpret_cast<uint4*>(zp + (size_t)item * ZB);
        const uint4* st = reinterpret_cast<const uint4*>(dstp);
        for (int t = lane; t < ZB / 16; t += 32) out[t] = st[t];
    }
    for (int t = lane; t < HB; t += 32) h_out[(size_t)item * HB + t] = h_slot[(size_t)a * HB + t];
    if (lane < 4) ct_out[(size_t)item * 4 + lane] = ct_slot[(size_t)a * 4 + lane];
    if (lane == 0) {
        attempts[item] = (uint32_t)kappa[item] + s + 1;
        if (done_list) done_list[atomicAdd(done_ctr, 1u)] = item;
    }
    __syncwarp();
}

// Kernel argument of the tail kernels: where signatures go and how items are re-queued.  In rounds with one slot
// per item (ctl->spec == 1) the tail finishes or re-queues its item itself and the resolve pass has nothing to do.
struct ResolveArgs {
    uint8_t* zp = nullptr;
    uint8_t* h_out = nullptr;
    uint64_t* ct_out = nullptr;
    uint32_t* attempts = nullptr;
    uint16_t* kappa = nullptr;
    uint32_t* active0 = nullptr;
    uint32_t* active1 = nullptr;
    RoundCtl* ctl = nullptr;
    const uint64_t* ct_slot = nullptr;
    uint32_t* done_list = nullptr;
};
// the same, resolved against the round state at kernel start
struct ResolveCtx {
    uint8_t* zp;
    uint8_t* h_out;
    uint64_t* ct_out;
    uint32_t* attempts;
    uint16_t* kappa;
    uint32_t* next_active;
    uint32_t* next_count;
    uint32_t* done_ctr;
    const uint64_t* ct_slot;
    const uint32_t* active;
    uint32_t* done_list;
    bool fused;
};
__device__ __forceinline__ ResolveCtx resolve_ctx(const ResolveArgs& ra) {
    const bool cur = ra.ctl->cur != 0;
    return ResolveCtx{ra.zp, ra.h_out, ra.ct_out, ra.attempts, ra.kappa, cur ? ra.active0 : ra.active1, &ra.ctl->next_count, &ra.ctl->done,
                      ra.ct_slot, cur ? ra.active1 : ra.active0, ra.done_list, ra.ctl->spec == 1};
}

// Survivors of the r0 and z checks: ct0_i = INTT(c_hat o t0_hat_i), ||ct0|| < gamma2, hint bits of
// (r0 + ct0, HighBits(w)) as 32-bit ballot masks hm[i*8 + r] (bit = lane, coefficient lane + 32 r), their count in nh.
// wi points at this lane's parked "r0 * 2 + (HighBits(w) != 0)" words (coefficient lane + 32 r at wi[i*N + 32 r]);
// t0_sm holds t0_hat pre-multiplied by 256^-1.  Returns true when a bound is violated.  Executed by one warp.
template <int K, int32_t GAMMA2>
__device__ __forceinline__ bool tail_ct0_hints(const uint32_t (&ch)[8], const uint32_t* __restrict__ t0_sm, const int32_t* wi,
                                               uint32_t* __restrict__ scr, const InvTw& itw, uint32_t* __restrict__ hm, int lane,
                                               uint32_t& nh) {
    bool bad = false;
#pragma unroll 1
    for (int i = 0; i < K && !bad; i++) {
        int32_t rv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) rv[r] = wi[i * N + 32 * r];
        uint32_t x[8];
        const uint4* kp = reinterpret_cast<const uint4*>(t0_sm + i * N) + lane;
        const uint4 lo = kp[0], hi = kp[32];
        x[0] = mul_full(ch[0], lo.x); x[1] = mul_full(ch[1], lo.y); x[2] = mul_full(ch[2], lo.z); x[3] = mul_full(ch[3], lo.w);
        x[4] = mul_full(ch[4], hi.x); x[5] = mul_full(ch[5], hi.y); x[6] = mul_full(ch[6], hi.z); x[7] = mul_full(ch[7], hi.w);
        ntt_inv_warp<true>(x, scr, itw, lane);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int32_t ct0 = centre(x[r]);
            bad |= (ct0 >= GAMMA2) || (ct0 <= -GAMMA2);
            const int32_t v = (rv[r] >> 1) + ct0;
            const bool hint = (v > GAMMA2) || (v < -GAMMA2) || (v == -GAMMA2 && (rv[r] & 1));
            const uint32_t mask = __ballot_sync(0xffffffffu, hint);
            nh += __popc(mask);
            if (lane == 0) hm[i * 8 + r] = mask;
        }
        bad = __any_sync(0xffffffffu, bad);
    }
    return bad;
}

// End of a slot: hint bytes of an accepted slot (omega position bytes, ascending inside each polynomial, then k
// running counts), the accept flag, and - in rounds with one slot per item (ra.fused) - the resolve step:
// the accepted slot becomes the signature, a rejected item advances kappa and joins the next round.
template <int K, int L, int G1BITS, int OMEGA, bool DIRECT>
__device__ __forceinline__ void tail_finish(bool bad, uint32_t a, const int32_t* y, uint8_t* h_slot, uint8_t* __restrict__ accepted,
                                            const uint32_t* __restrict__ hm, const ResolveCtx& ra, uint8_t* zstage, int lane) {
    __syncwarp();
    if (!bad) {
        uint8_t* ho = h_slot + (size_t)a * (OMEGA + K);
        for (int t = lane; t < OMEGA + K; t += 32) ho[t] = 0;
        __syncwarp();
        uint32_t run = 0;
        for (int i = 0; i < K; i++) {
            for (int r = 0; r < 8; r++) {
                const uint32_t mask = hm[i * 8 + r];
                if ((mask >> lane) & 1u) ho[run + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)(32 * r + lane);
                run += __popc(mask);
            }
            if (lane == 0) ho[OMEGA + i] = (uint8_t)run;
        }
    }
    if (lane == 0) accepted[a] = bad ? 0 : 1;
    __syncwarp();
    if (ra.fused) {
        const uint32_t item = ra.active[a];
        if (bad) {
            if (lane == 0) {
                ra.kappa[item] += 1;
                ra.next_active[atomicAdd(ra.next_count, 1u)] = item;
            }
        } else {
            resolve_finish<L, G1BITS, OMEGA + K, DIRECT>(ra.zp, ra.h_out, ra.ct_out, ra.attempts, ra.kappa, ra.done_ctr, y, h_slot,
                                                         ra.ct_slot, ra.done_list, item, a, 0u, zstage, lane);
        }
    }
}


// ---------------------------------------------------------------------------------------
// S5: signature tail, one warp per active item.
//   c_hat = NTT(c); z = y + INTT(c_hat o s1_hat); ||z|| < gamma1 - beta
//   r0 = LowBits(w) - INTT(c_hat o s2_hat); ||r0|| < gamma2 - beta
//   ct0 = INTT(c_hat o t0_hat); ||ct0|| < gamma2; h = MakeHint(r0 + ct0, HighBits(w)); #h <= omega
// Key polynomials (s1_hat | s2_hat | t0_hat, NTT domain) live in shared memory per persistent CTA.
// Rejected items are appended to the next round's active list.
// ---------------------------------------------------------------------------------------

// PER_ITEM = true (one key per signature, dil_sign_multi_*): key_hat holds one (s1_hat | s2_hat | t0_hat) record per
// ITEM, canonical and already multiplied by 256^-1, and is read from global memory instead of shared memory.
template <int K, int L, int32_t GAMMA1, int32_t GAMMA2, int BETA, int OMEGA, int WARPS, int CTAS, bool PER_ITEM = false>
__global__ void __launch_bounds__(WARPS * 32, CTAS) sign_tail_kernel(
    int32_t* __restrict__ y /* in: y, out: z (in place) */, uint8_t* __restrict__ h_slot, uint8_t* __restrict__ accepted,
    const int32_t* __restrict__ key_hat, int32_t* __restrict__ w /* in: w; scratch afterwards */, const int8_t* __restrict__ c,
    const ResolveArgs rargs) {
    const uint32_t n_slots = rargs.ctl->n_slots;
    uint32_t* const work_ctr = &rargs.ctl->ctr_tail;
    const ResolveCtx ra = resolve_ctx(rargs);
    extern __shared__ __align__(16) uint32_t sm_words[];
    constexpr int NKEY = L + 2 * K;
    constexpr int G1BITS = GAMMA1 == (1 << 17) ? 17 : 19;
    constexpr int ZB = L * 32 * (G1BITS + 1);          // packed z bytes per signature
    uint32_t* key_sm = sm_words;                       // NKEY * 256 (shared key only)
    uint32_t* scr_all = sm_words + (PER_ITEM ? 0 : NKEY * N);   // WARPS * SCRATCH_WORDS
    uint32_t* hm_all = scr_all + WARPS * SCRATCH_WORDS;  // WARPS * K * 8 hint masks
    uint8_t* zstage_all = reinterpret_cast<uint8_t*>(hm_all + WARPS * K * 8);   // WARPS * ZB (fused resolve)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // a CTA that starts when every slot is already claimed (or in a round without work) leaves at once (uniform decision)
        __shared__ uint32_t late;
        if (threadIdx.x == 0) late = *reinterpret_cast<volatile uint32_t*>(work_ctr) >= n_slots;
        __syncthreads();
        if (late) return;
    }
    // key polynomials are kept pre-multiplied by 256^-1 so that every inverse transform below can skip
    // its scaling multiplications (ntt_inv_warp<true>)
    if constexpr (!PER_ITEM) {
        for (int t = threadIdx.x; t < NKEY * (N / 4); t += blockDim.x) {
            int4 q = __ldg(reinterpret_cast<const int4*>(key_hat) + t);
            reinterpret_cast<uint4*>(key_sm)[t] = make_uint4(mul_full(canon_signed(q.x), INV256), mul_full(canon_signed(q.y), INV256),
                                                             mul_full(canon_signed(q.z), INV256), mul_full(canon_signed(q.w), INV256));
        }
        __syncthreads();
    }
    const uint32_t spec = rargs.ctl->spec;
    uint32_t* scr = scr_all + warp * SCRATCH_WORDS;
    uint32_t* hm = hm_all + warp * K * 8;
    FwdTw ftw;
    InvTw itw;
    load_inv_tw(itw, &TW_INV, lane);

    // Slots are claimed dynamically from the round's work counter: the work per slot varies by an order of
    // magnitude (early exits below), and SMs may run at different speeds or be partly taken by other streams'
    // kernels; the claim for the next slot is issued before the current one is processed.
    uint32_t claim = 0;
    if (lane == 0) claim = atomicAdd(work_ctr, 1u);
    for (uint32_t a = __shfl_sync(0xffffffffu, claim, 0); a < n_slots; a = __shfl_sync(0xffffffffu, claim, 0)) {
        if (lane == 0) claim = atomicAdd(work_ctr, 1u);
        // c_hat in layout C
        uint32_t ch[8];
        {
            const int8_t* cp = c + (size_t)a * N + lane;
#pragma unroll
            for (int r = 0; r < 8; r++) ch[r] = (uint32_t)(int32_t)cp[32 * r];
            load_fwd_tw(ftw, &TW_FWD, lane);
            ntt_fwd_warp(ch, scr, ftw, lane);
            __syncwarp();
        }
        // this slot's key record: the CTA's shared-memory copy, or the item's record in global memory
        const uint32_t* kb = key_sm;
        if constexpr (PER_ITEM) kb = reinterpret_cast<const uint32_t*>(key_hat) + (size_t)ra.active[a / spec] * NKEY * N;
        auto mul_inv = [&](uint32_t (&x)[8], int p) {   // x = INTT(c_hat o key[p]) in layout A
            const uint4* kp = reinterpret_cast<const uint4*>(kb + p * N) + lane;
            uint4 lo = kp[0], hi = kp[32];
            x[0] = mul_full(ch[0], lo.x); x[1] = mul_full(ch[1], lo.y); x[2] = mul_full(ch[2], lo.z); x[3] = mul_full(ch[3], lo.w);
            x[4] = mul_full(ch[4], hi.x); x[5] = mul_full(ch[5], hi.y); x[6] = mul_full(ch[6], hi.z); x[7] = mul_full(ch[7], hi.w);
            ntt_inv_warp<true>(x, scr, itw, lane);
            __syncwarp();
        };
        // The three norm checks are independent tests of equal cost per polynomial, so they run most-selective
        // first (the accept/reject outcome does not depend on the order, combined_top.v:2098-2228 / SURVEY A.3):
        // r0 = LowBits(w) - c*s2 rejects ~19 % per polynomial at level 2, z ~14 %, c*t0 < 1 %.  r0 and the
        // "HighBits(w) != 0" flag needed by MakeHint are parked over w (dead after this kernel) for the survivors.
        bool bad = false;
        int32_t* wi = w + (size_t)a * K * N + lane;
#pragma unroll 1
        for (int i = 0; i < K && !bad; i++) {
            int32_t wv[8];
#pragma unroll
            for (int r = 0; r < 8; r++) wv[r] = wi[i * N + 32 * r];   // issued before the transform: latency hidden
            uint32_t x[8];
            mul_inv(x, L + i);                      // c*s2_i
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int32_t a0, w1;
                decompose<GAMMA2>(wv[r], w1, a0);
                const int32_t r0 = a0 - centre(x[r]);
                bad |= (r0 >= GAMMA2 - BETA) || (r0 <= -(GAMMA2 - BETA));
                wv[r] = r0 * 2 + (w1 != 0 ? 1 : 0);
            }
            bad = __any_sync(0xffffffffu, bad);
            if (!bad) {
#pragma unroll
                for (int r = 0; r < 8; r++) wi[i * N + 32 * r] = wv[r];
            }
        }
        // z = y + c*s1, written over y; stop at the first polynomial that violates the bound
        int32_t* yi = y + (size_t)a * L * N + lane;
#pragma unroll 1
        for (int j = 0; j < L && !bad; j++) {
            int32_t yv[8];
#pragma unroll
            for (int r = 0; r < 8; r++) yv[r] = yi[j * N + 32 * r];
            uint32_t x[8];
            mul_inv(x, j);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int32_t z = yv[r] + centre(x[r]);
                bad |= (z >= GAMMA1 - BETA) || (z <= -(GAMMA1 - BETA));
                yi[j * N + 32 * r] = z;
            }
            bad = __any_sync(0xffffffffu, bad);
        }
        uint32_t nh = 0;
        if (!bad) bad = tail_ct0_hints<K, GAMMA2>(ch, kb + (L + K) * N, wi, scr, itw, hm, lane, nh) || nh > OMEGA;
        tail_finish<K, L, G1BITS, OMEGA, false>(bad, a, y, h_slot, accepted, hm, ra, zstage_all + (size_t)warp * ZB, lane);
    }
}

// ---------------------------------------------------------------------------------------
// S5 (default): signature tail with SPARSE challenge products.  c has only TAU coefficients +-1 and s1, s2
// have coefficients in [-eta, eta], so c*s1 and c*s2 are TAU signed negacyclic shifts of a small polynomial:
//     (c*s)[n] = sum_t sign_t * e[256 + n - pos_t],   e[256 + i] = s[i], e[i] = -s[i]   (0 <= i < 256)
// Per key polynomial the CTA keeps e in shared memory as biased NIBBLES (eta + e in 0..2*eta), once per sign and
// once per alignment (pos_t mod 8), so a lane adds 8 consecutive coefficients of one term with ONE aligned
// 4-byte shared load and one packed integer add; every GROUP terms (as many as fit in a nibble) the nibble
// sums are spread into byte sums.  No NTT(c), no inverse transform and no multiplier work for the two
// selective checks (r0, then z); shared-memory bandwidth (128 B/clk per SM) is the limit instead, which is why
// the elements are nibbles and not bytes.  Only survivors (~24 %) pay NTT(c) and the c*t0 products through
// the transform path.  Results are the same integers the transform path produces (|c*s| <= TAU*eta << Q/2),
// so signatures stay bit-exact.  Used for eta = 2 (levels 2 and 5); level 3 (eta = 4) keeps the transform tail.
// ---------------------------------------------------------------------------------------
template <int K, int L, int32_t GAMMA1, int32_t GAMMA2, int BETA, int OMEGA, int TAU, int ETA, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) sign_tail_sparse_kernel(
    int32_t* __restrict__ y /* in: y, out: z (in place) */, uint8_t* __restrict__ h_slot, uint8_t* __restrict__ accepted,
    const int32_t* __restrict__ key_hat, const int8_t* __restrict__ key_small, int32_t* __restrict__ w /* in: w; scratch afterwards */,
    const int8_t* __restrict__ c, const ResolveArgs rargs) {
    const uint32_t n_slots = rargs.ctl->n_slots;
    uint32_t* const work_ctr = &rargs.ctl->ctr_tail;
    const ResolveCtx ra = resolve_ctx(rargs);
    extern __shared__ __align__(16) uint32_t sm_words[];
    constexpr int NP = L + K;                              // small key polynomials: s1 | s2
    // eta = 2: elements are biased NIBBLES (0..4), three terms fit a nibble sum, one 4-byte load covers a lane's 8 coefficients.
    // eta = 4: biased elements (0..8) need BYTES (one 8-byte load per term), 31 terms fit a byte sum, then 16-bit sums.
    constexpr bool BYTES = ETA > 3;
    constexpr int GROUP = BYTES ? 255 / (2 * ETA) : 15 / (2 * ETA);   // terms that can be added without a carry
    constexpr int TABB = BYTES ? 2 * 8 * 512 : 2 * 8 * 256;           // table bytes per polynomial: sign x alignment x 512 elements
    static_assert(GROUP >= 1 && (BYTES ? 2 * ETA * TAU < 65536 : 2 * ETA * TAU < 256), "partial sums must not carry");
    constexpr int G1BITS = GAMMA1 == (1 << 17) ? 17 : 19;
    uint32_t* t0_sm = sm_words;                                   // K * 256: t0_hat * 256^-1
    uint32_t* scr_all = t0_sm + K * N;                            // WARPS * SCRATCH_WORDS
    uint32_t* hm_all = scr_all + WARPS * SCRATCH_WORDS;           // WARPS * K * 8 hint masks
    uint16_t* terms_all = reinterpret_cast<uint16_t*>(hm_all + WARPS * K * 8);   // WARPS * 64 term offsets
    uint8_t* tabs = reinterpret_cast<uint8_t*>(terms_all + WARPS * 64);          // NP * TABB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // a CTA that starts when every slot is already claimed (or in a round without work) leaves at once (uniform decision)
        __shared__ uint32_t late;
        if (threadIdx.x == 0) late = *reinterpret_cast<volatile uint32_t*>(work_ctr) >= n_slots;
        __syncthreads();
        if (late) return;
    }
    for (int t = threadIdx.x; t < K * (N / 4); t += blockDim.x) {
        int4 q = __ldg(reinterpret_cast<const int4*>(key_hat + (size_t)(L + K) * N) + t);
        reinterpret_cast<uint4*>(t0_sm)[t] = make_uint4(mul_full(canon_signed(q.x), INV256), mul_full(canon_signed(q.y), INV256),
                                                        mul_full(canon_signed(q.z), INV256), mul_full(canon_signed(q.w), INV256));
    }
    // tables: T[p][sign][a][m] = eta +- e_p[m + a]  (0 beyond the end), one 32-bit word (8 nibbles / 4 bytes) per thread step
    constexpr int EPW = BYTES ? 4 : 8;                     // elements per table word
    constexpr int WPA = 512 / EPW;                         // words per (polynomial, sign, alignment) row
    for (int idx = threadIdx.x; idx < NP * 2 * 8 * WPA; idx += blockDim.x) {
        const int m = (idx % WPA) * EPW, al = (idx / WPA) & 7, sg = (idx / (WPA * 8)) & 1, p = idx / (WPA * 16);
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < EPW; q++) {
            const int j = m + q + al;
            int e = 0;
            if (j < 256) e = -(int)key_small[p * N + j];
            else if (j < 512) e = (int)key_small[p * N + j - 256];
            word |= (uint32_t)(ETA + (sg ? -e : e)) << ((32 / EPW) * q);
        }
        reinterpret_cast<uint32_t*>(tabs)[idx] = word;
    }
    __syncthreads();
    uint32_t* scr = scr_all + warp * SCRATCH_WORDS;
    uint32_t* hm = hm_all + warp * K * 8;
    uint16_t* terms = terms_all + warp * 64;

    // x[0..7] = (c * small_p)[8*lane .. 8*lane+7] from the term list of the current slot
    auto sparse_mul = [&](int p, int32_t (&x)[8]) {
        if constexpr (!BYTES) {
            const uint8_t* tp = tabs + (size_t)p * TABB + 4 * lane;
            uint32_t b0 = 0, b1 = 0, nacc = 0, tw = 0;
#pragma unroll
            for (int t = 0; t < TAU; t++) {
                uint32_t u;
                if ((t & 1) == 0) {   // term offsets are read two at a time
                    tw = (t + 1 < TAU) ? reinterpret_cast<const uint32_t*>(terms)[t >> 1] : (uint32_t)terms[t];
                    u = tw & 0xFFFFu;
                } else {
                    u = tw >> 16;
                }
                nacc += *reinterpret_cast<const uint32_t*>(tp + u);
                if (t % GROUP == GROUP - 1 || t == TAU - 1) {
                    b0 += nacc & 0x0F0F0F0Fu;          // coefficients 0, 2, 4, 6
                    b1 += (nacc >> 4) & 0x0F0F0F0Fu;   // coefficients 1, 3, 5, 7
                    nacc = 0;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (int32_t)((((i & 1) ? b1 : b0) >> (8 * (i >> 1))) & 0xFFu) - TAU * ETA;
        } else {
            const uint8_t* tp = tabs + (size_t)p * TABB + 8 * lane;
            uint32_t b0 = 0, b1 = 0, tw = 0;                 // byte sums of coefficients 0..3 / 4..7
            uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;         // 16-bit sums: (c0, c2), (c1, c3), (c4, c6), (c5, c7)
#pragma unroll
            for (int t = 0; t < TAU; t++) {
                uint32_t u;
                if ((t & 1) == 0) {
                    tw = (t + 1 < TAU) ? reinterpret_cast<const uint32_t*>(terms)[t >> 1] : (uint32_t)terms[t];
                    u = tw & 0xFFFFu;
                } else {
                    u = tw >> 16;
                }
                const uint2 v = *reinterpret_cast<const uint2*>(tp + u);
                b0 += v.x;
                b1 += v.y;
                if (t % GROUP == GROUP - 1 || t == TAU - 1) {
                    h0 += b0 & 0x00FF00FFu; h1 += (b0 >> 8) & 0x00FF00FFu;
                    h2 += b1 & 0x00FF00FFu; h3 += (b1 >> 8) & 0x00FF00FFu;
                    b0 = b1 = 0;
                }
            }
            x[0] = (int32_t)(h0 & 0xFFFFu) - TAU * ETA; x[1] = (int32_t)(h1 & 0xFFFFu) - TAU * ETA;
            x[2] = (int32_t)(h0 >> 16) - TAU * ETA;     x[3] = (int32_t)(h1 >> 16) - TAU * ETA;
            x[4] = (int32_t)(h2 & 0xFFFFu) - TAU * ETA; x[5] = (int32_t)(h3 & 0xFFFFu) - TAU * ETA;
            x[6] = (int32_t)(h2 >> 16) - TAU * ETA;     x[7] = (int32_t)(h3 >> 16) - TAU * ETA;
        }
    };

    // (Claiming two slots ahead and pulling the next slot's c / w into L2 with prefetches was measured 4 % SLOWER here: per-slot work
    // varies tenfold and the late rounds have only 5-9 slots per warp, so every slot a warp holds back costs balance.)
    uint32_t claim = 0;
    if (lane == 0) claim = atomicAdd(work_ctr, 1u);
    for (uint32_t a = __shfl_sync(0xffffffffu, claim, 0); a < n_slots; a = __shfl_sync(0xffffffffu, claim, 0)) {
        if (lane == 0) claim = atomicAdd(work_ctr, 1u);
        // term list: offset of every non-zero coefficient's table row (sign, alignment, shift)
        {
            const uint2 cb = *reinterpret_cast<const uint2*>(c + (size_t)a * N + 8 * lane);
            uint32_t cnt = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t byte = ((i < 4 ? cb.x : cb.y) >> (8 * (i & 3))) & 0xFFu;
                const uint32_t m = __ballot_sync(0xffffffffu, byte != 0);
                if (byte != 0) {
                    const int pos = 8 * lane + i, al = (-pos) & 7;
                    terms[cnt + __popc(m & ((1u << lane) - 1u))] =
                        BYTES ? (uint16_t)((byte == 0xFFu ? 4096 : 0) + al * 512 + (256 - pos - al))
                              : (uint16_t)((byte == 0xFFu ? 2048 : 0) + al * 256 + (256 - pos - al) / 2);
                }
                cnt += __popc(m);
            }
            __syncwarp();
        }
        // most selective check first (see sign_tail_kernel): r0 = LowBits(w) - c*s2, parked over w with the
        // HighBits(w) != 0 flag for the survivors
        bool bad = false;
        int32_t* wa = w + (size_t)a * K * N;
        {   // rows 1 .. K-1 of w and the first polynomial of y are on their way into L2 while row 0 is checked (no registers needed)
            constexpr int LINES = (K - 1) * 8 + 8;
#pragma unroll
            for (int q = 0; q < (LINES + 31) / 32; q++) {
                const int ln = 32 * q + lane;
                if (ln < (K - 1) * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(wa + N) + 128 * ln));
                else if (ln < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(y + (size_t)a * L * N) + 128 * (ln - (K - 1) * 8)));
            }
        }
#pragma unroll 1
        for (int i = 0; i < K && !bad; i++) {
            int4* wp = reinterpret_cast<int4*>(wa + i * N) + 2 * lane;
            const int4 w0 = wp[0], w1v = wp[1];
            int32_t wv[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
            int32_t x[8];
            sparse_mul(L + i, x);                   // c*s2_i
#pragma unroll
            for (int q = 0; q < 8; q++) {
                int32_t a0, w1;
                decompose<GAMMA2>(wv[q], w1, a0);
                const int32_t r0 = a0 - x[q];
                bad |= (r0 >= GAMMA2 - BETA) || (r0 <= -(GAMMA2 - BETA));
                wv[q] = r0 * 2 + (w1 != 0 ? 1 : 0);
            }
            bad = __any_sync(0xffffffffu, bad);
            if (!bad) {
                wp[0] = make_int4(wv[0], wv[1], wv[2], wv[3]);
                wp[1] = make_int4(wv[4], wv[5], wv[6], wv[7]);
            }
        }
        // z = y + c*s1, written over y
        int32_t* ya = y + (size_t)a * L * N;
#pragma unroll 1
        for (int j = 0; j < L && !bad; j++) {
            int4* yp = reinterpret_cast<int4*>(ya + j * N) + 2 * lane;
            const int4 y0 = yp[0], y1 = yp[1];
            int32_t x[8];
            sparse_mul(j, x);                       // c*s1_j
            const int32_t z[8] = {y0.x + x[0], y0.y + x[1], y0.z + x[2], y0.w + x[3], y1.x + x[4], y1.y + x[5], y1.z + x[6], y1.w + x[7]};
#pragma unroll
            for (int q = 0; q < 8; q++) bad |= (z[q] >= GAMMA1 - BETA) || (z[q] <= -(GAMMA1 - BETA));
            yp[0] = make_int4(z[0], z[1], z[2], z[3]);
            yp[1] = make_int4(z[4], z[5], z[6], z[7]);
            bad = __any_sync(0xffffffffu, bad);
        }
        uint32_t nh = 0;
        if (!bad) {
            // survivors: c*t0 through the transform path; coefficient lane + 32 r per register from here on
            __syncwarp();
            uint32_t ch[8];
            {
                const int8_t* cp = c + (size_t)a * N + lane;
#pragma unroll
                for (int r = 0; r < 8; r++) ch[r] = (uint32_t)(int32_t)cp[32 * r];
                FwdTw ftw;
                load_fwd_tw(ftw, &TW_FWD, lane);
                ntt_fwd_warp(ch, scr, ftw, lane);
                __syncwarp();
            }
            InvTw itw;   // loaded here (survivors only) so the twiddle registers are free during the sparse phases
            {
                const TwTable* tab = &TW_INV;
                asm volatile("" : "+l"(tab));
                load_inv_tw(itw, tab, lane);
            }
            bad = tail_ct0_hints<K, GAMMA2>(ch, t0_sm, wa + lane, scr, itw, hm, lane, nh) || nh > OMEGA;
        }
        tail_finish<K, L, G1BITS, OMEGA, true>(bad, a, y, h_slot, accepted, hm, ra, nullptr, lane);
    }
}

// ---------------------------------------------------------------------------------------
// S6: resolve, one warp per active item: the first accepted slot (smallest kappa) wins; its z is
// packed (encoder.v:96-133: gamma1 - z, 18 or 20 bits) straight into the signature, h and c~ are
// copied; items without an accepted slot advance kappa by `spec` and join the next round.
// ---------------------------------------------------------------------------------------
template <int L, int GAMMA1_BITS, int HB>
__global__ void __launch_bounds__(256) resolve_kernel(const ResolveArgs rargs, const int32_t* __restrict__ zslot,
                                                      const uint8_t* __restrict__ h_slot, const uint8_t* __restrict__ accepted) {
    const uint32_t spec = rargs.ctl->spec, n_items = rargs.ctl->n_active;
    if (spec == 1) return;   // the tail resolved this round's items itself
    const int lane = threadIdx.x & 31;
    const ResolveCtx ra = resolve_ctx(rargs);
    constexpr int ZB = L * 32 * (GAMMA1_BITS + 1);
    __shared__ __align__(16) uint8_t zstage[8][ZB];
    for (uint32_t idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < n_items; idx += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t item = ra.active[idx];
        const uint32_t a0 = idx * spec;
        const uint32_t acc = (lane < (int)spec) ? accepted[a0 + lane] : 0;
        const uint32_t mask = __ballot_sync(0xffffffffu, acc != 0);
        if (mask == 0) {
            if (lane == 0) {
                ra.kappa[item] += (uint16_t)spec;
                ra.next_active[atomicAdd(ra.next_count, 1u)] = item;
            }
            continue;
        }
        const uint32_t s = __ffs(mask) - 1;
        resolve_finish<L, GAMMA1_BITS, HB>(ra.zp, ra.h_out, ra.ct_out, ra.attempts, ra.kappa, ra.done_ctr, zslot, h_slot, ra.ct_slot,
                                           ra.done_list, item, a0 + s, s, zstage[threadIdx.x >> 5], lane);
    }
}

// Copies the signatures finished by one round from the device staging arrays into the caller's
// pinned host buffers (mapped into the device address space): posted PCIe writes, 512 contiguous bytes
// per warp store.  Runs on the copy stream next to the following rejection rounds (launch_drain).
__global__ void __launch_bounds__(512, 4) drain_kernel(uint8_t* __restrict__ hz, uint8_t* __restrict__ hh, uint8_t* __restrict__ hct,
                                                        uint32_t* __restrict__ hatt, const uint8_t* __restrict__ zp,
                                                        const uint8_t* __restrict__ h, const uint8_t* __restrict__ ct,
                                                        const uint32_t* __restrict__ att, const uint32_t* __restrict__ done_list,
                                                        const RoundCtl* __restrict__ ctl, uint32_t round, uint32_t zb, uint32_t hb) {
    // round r finished the done-list entries [done_snap[r-1], done_snap[r]); both were written by plan_kernel launches
    // that precede this kernel in stream order (the copy stream waits on an event recorded after round r)
    const uint32_t lo = round == 0 ? 0u : ctl->done_snap[(round - 1) & 63], n = ctl->done_snap[round & 63] - lo;
    const uint32_t* __restrict__ list = done_list + lo;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nv = zb >> 4;
    for (uint32_t i = warp; i < n; i += nwarps) {
        const uint32_t item = list[i];
        const uint4* s = reinterpret_cast<const uint4*>(zp + (size_t)item * zb);
        uint4* d = reinterpret_cast<uint4*>(hz + (size_t)item * zb);
        uint32_t t = lane;
        for (; t + 96 < nv; t += 128) {
            uint4 a = __ldcs(s + t), b = __ldcs(s + t + 32), c = __ldcs(s + t + 64), e = __ldcs(s + t + 96);
            d[t] = a; d[t + 32] = b; d[t + 64] = c; d[t + 96] = e;
        }
        for (; t < nv; t += 32) d[t] = __ldcs(s + t);
        if (lane < 2) reinterpret_cast<uint4*>(hct + (size_t)item * 32)[lane] = reinterpret_cast<const uint4*>(ct + (size_t)item * 32)[lane];
        for (uint32_t b = lane; b < hb; b += 32) hh[(size_t)item * hb + b] = h[(size_t)item * hb + b];
        if (hatt && lane == 0) hatt[item] = att[item];
    }
}

// ---------------------------------------------------------------------------------------
// Round control (device side of the rejection loop)
// ---------------------------------------------------------------------------------------
// slots per item for a round of n items: 1 while the GPU is saturated by distinct items, then fill up to the target
__device__ __forceinline__ uint32_t spec_policy(const RoundCtl* ctl, uint32_t n) {
    if (n == 0 || n >= ctl->spec_target) return 1;
    const uint32_t room = (ctl->slot_cap < ctl->spec_target ? ctl->slot_cap : ctl->spec_target) / n;
    return room < 1 ? 1 : (room > ctl->spec_max ? ctl->spec_max : room);
}

// start of a batch: active list 0 = all items, round 0 planned
__global__ void sign_begin_kernel(RoundCtl* __restrict__ ctl, uint32_t* __restrict__ active0, uint32_t n, uint32_t slot_cap,
                                  uint32_t spec_target, uint32_t spec_max) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) active0[t] = t;
    if (t == 0) {
        ctl->slot_cap = slot_cap; ctl->spec_target = spec_target; ctl->spec_max = spec_max;
        ctl->n_active = n; ctl->cur = 0;
        const uint32_t sp = spec_policy(ctl, n);
        ctl->spec = sp; ctl->n_slots = n * sp;
        ctl->next_count = 0; ctl->ctr_core = 0; ctl->ctr_tail = 0; ctl->done = 0; ctl->rounds = 0; ctl->total_slots = 0;
    }
}

// end of round `round`: the re-queued items become the next round
__global__ void plan_kernel(RoundCtl* __restrict__ ctl, uint32_t round) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctl->n_active != 0) { ctl->rounds += 1; ctl->total_slots += ctl->n_slots; }
    ctl->done_snap[round & 63] = ctl->done;
    const uint32_t n = ctl->next_count;
    const uint32_t sp = spec_policy(ctl, n);
    ctl->n_active = n; ctl->spec = sp; ctl->n_slots = n * sp; ctl->cur ^= 1u;
    ctl->next_count = 0; ctl->ctr_core = 0; ctl->ctr_tail = 0;
}

// the first 16 words of the round state into mapped pinned host memory: posted PCIe writes, so the host never
// queues a D2H copy behind bulk signature transfers
__global__ void publish_ctl_kernel(volatile uint32_t* host_dst, const RoundCtl* ctl) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ctl);
    if (threadIdx.x < 16) host_dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
cudaError_t launch_sign_begin(const SignBufs& b, uint32_t n, uint32_t slot_cap, uint32_t spec_target, uint32_t spec_max, cudaStream_t st) {
    sign_begin_kernel<<<(n + 255) / 256, 256, 0, st>>>(b.ctl, b.active[0], n, slot_cap, spec_target, spec_max);
    return cudaGetLastError();
}
cudaError_t launch_plan(RoundCtl* ctl, uint32_t round, cudaStream_t st) {
    plan_kernel<<<1, 32, 0, st>>>(ctl, round);
    return cudaGetLastError();
}
cudaError_t launch_publish_ctl(uint32_t* host_dst_dev, const RoundCtl* ctl, cudaStream_t st) {
    publish_ctl_kernel<<<1, 32, 0, st>>>(host_dst_dev, ctl);
    return cudaGetLastError();
}

cudaError_t launch_sign_init(uint64_t* mu, uint64_t* rhop, uint16_t* kappa, const uint8_t* tr, const uint8_t* key, size_t key_stride,
                             const uint8_t* msgs, const uint64_t* offsets, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sign_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(mu, rhop, kappa, tr, key, key_stride, msgs, offsets, n);
    return cudaGetLastError();
}

template <int L, int G1B>
static cudaError_t launch_expand_mask_t(const SignBufs& b, uint32_t cap_slots, cudaStream_t st) {
    constexpr int ROW = 32 * (G1B + 1) + 16;
    constexpr size_t smem = (size_t)4 * 32 * ROW;
    auto kern = expand_mask_kernel<L, G1B>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    const uint32_t n_polys = cap_slots * L;
    kern<<<(n_polys + 127) / 128, 128, smem, st>>>(b.y, b.rhop, b.kappa, b.active[0], b.active[1], b.ctl);
    return cudaGetLastError();
}

cudaError_t launch_expand_mask(int level, const SignBufs& b, uint32_t cap_slots, cudaStream_t st) {
    if (cap_slots == 0) return cudaSuccess;
    switch (level) {
        case 2: return launch_expand_mask_t<4, 17>(b, cap_slots, st);
        case 3: return launch_expand_mask_t<5, 19>(b, cap_slots, st);
        case 5: return launch_expand_mask_t<7, 19>(b, cap_slots, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_drain(uint8_t* hz, uint8_t* hh, uint8_t* hct, uint32_t* hatt, const SignBufs& b, uint32_t round, uint32_t zb,
                         uint32_t hb, cudaStream_t st) {
    // A handful of 16-warp CTAs, each asking for enough (unused) dynamic shared memory that the shared-memory
    // using signing kernels cannot be co-resident: the drain owns its few SMs instead of slowing every SM's
    // memory pipeline with PCIe-paced stores (thin many-CTA drains cost the signing kernels 20 %).
    constexpr int CTAS = 4, SMEM_KB = 200;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(drain_kernel, (size_t)SMEM_KB * 1024, configured); e != cudaSuccess) return e;
    drain_kernel<<<CTAS, 512, (size_t)SMEM_KB * 1024, st>>>(hz, hh, hct, hatt, b.zp, b.h_out, reinterpret_cast<const uint8_t*>(b.ct_out),
                                                            b.attempts, b.done_list, b.ctl, round, zb, hb);
    return cudaGetLastError();
}

cudaError_t launch_challenge(int level, const SignBufs& b, uint32_t cap_slots, cudaStream_t st) {
    if (cap_slots == 0) return cudaSuccess;
    // 64-thread CTAs: 1024 CTAs for a 65536-slot round spread evenly over 148 SMs
    const unsigned grid = (cap_slots + CH_THREADS - 1) / CH_THREADS;
    switch (level) {
        case 2: challenge_kernel<4, 192, 39><<<grid, CH_THREADS, 0, st>>>(b.c, b.ct_slot, b.mu, b.w1p, b.active[0], b.active[1], b.ctl); break;
        case 3: challenge_kernel<6, 128, 49><<<grid, CH_THREADS, 0, st>>>(b.c, b.ct_slot, b.mu, b.w1p, b.active[0], b.active[1], b.ctl); break;
        case 5: challenge_kernel<8, 128, 60><<<grid, CH_THREADS, 0, st>>>(b.c, b.ct_slot, b.mu, b.w1p, b.active[0], b.active[1], b.ctl); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

static ResolveArgs resolve_args(const SignBufs& b) {
    ResolveArgs ra;
    ra.zp = b.zp; ra.h_out = b.h_out; ra.ct_out = b.ct_out; ra.attempts = b.attempts; ra.kappa = b.kappa;
    ra.active0 = b.active[0]; ra.active1 = b.active[1]; ra.ctl = b.ctl; ra.ct_slot = b.ct_slot;
    ra.done_list = b.track_done ? b.done_list : nullptr;
    return ra;
}

// transform-only tail (any eta): one 24-warp CTA per SM
template <int K, int L, int32_t G1, int32_t G2, int BETA, int OMEGA, bool PER_ITEM = false>
static cudaError_t launch_sign_tail_t(const SignBufs& b, const int32_t* key_hat, uint32_t cap_slots, int sm_count, cudaStream_t st) {
    constexpr int WARPS = 24, CTAS = 1;
    constexpr int ZB = L * 32 * ((G1 == (1 << 17) ? 17 : 19) + 1);
    constexpr size_t smem = (size_t)((PER_ITEM ? 0 : (L + 2 * K) * N) + WARPS * SCRATCH_WORDS + WARPS * K * 8) * 4 + (size_t)WARPS * ZB;
    auto kern = sign_tail_kernel<K, L, G1, G2, BETA, OMEGA, WARPS, CTAS, PER_ITEM>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    unsigned want = (cap_slots + WARPS - 1) / WARPS;
    unsigned cap = (unsigned)sm_count * CTAS;
    kern<<<want < cap ? want : cap, WARPS * 32, smem, st>>>(b.y, b.h_slot, b.accepted, key_hat, b.w, b.c, resolve_args(b));
    return cudaGetLastError();
}

template <int K, int L, int32_t G1, int32_t G2, int BETA, int OMEGA, int TAU, int ETA>
static cudaError_t launch_sign_tail_sparse(const SignBufs& b, const int32_t* key_hat, const int8_t* key_small, uint32_t cap_slots,
                                           int sm_count, cudaStream_t st) {
    constexpr int WARPS = 24;
    constexpr size_t smem = (size_t)(K * N + WARPS * SCRATCH_WORDS + WARPS * K * 8) * 4 + (size_t)WARPS * 64 * 2 +
                            (size_t)(L + K) * 2 * 8 * (ETA > 3 ? 512 : 256);
    static_assert(smem <= 227 * 1024, "sparse tail tables do not fit in shared memory");
    auto kern = sign_tail_sparse_kernel<K, L, G1, G2, BETA, OMEGA, TAU, ETA, WARPS>;
    static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per device
    if (cudaError_t e = ensure_dyn_smem(kern, (size_t)(smem), configured); e != cudaSuccess) return e;
    unsigned want = (cap_slots + WARPS - 1) / WARPS;
    unsigned cap = (unsigned)sm_count;
    kern<<<want < cap ? want : cap, WARPS * 32, smem, st>>>(b.y, b.h_slot, b.accepted, key_hat, key_small, b.w, b.c, resolve_args(b));
    return cudaGetLastError();
}

// key_small != nullptr selects the sparse-product tail where it exists (eta = 2); a null key_small the transform-only tail
cudaError_t launch_sign_tail(int level, const SignBufs& b, const int32_t* key_hat, const int8_t* key_small, uint32_t cap_slots,
                             int sm_count, cudaStream_t st) {
    if (cap_slots == 0) return cudaSuccess;
    if (key_small != nullptr) {
        switch (level) {
            case 2: return launch_sign_tail_sparse<4, 4, 1 << 17, (Q_I - 1) / 88, 78, 80, 39, 2>(b, key_hat, key_small, cap_slots, sm_count, st);
            case 3: return launch_sign_tail_sparse<6, 5, 1 << 19, (Q_I - 1) / 32, 196, 55, 49, 4>(b, key_hat, key_small, cap_slots, sm_count, st);
            case 5: return launch_sign_tail_sparse<8, 7, 1 << 19, (Q_I - 1) / 32, 120, 75, 60, 2>(b, key_hat, key_small, cap_slots, sm_count, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (level) {
        case 2: return launch_sign_tail_t<4, 4, 1 << 17, (Q_I - 1) / 88, 78, 80>(b, key_hat, cap_slots, sm_count, st);
        case 3: return launch_sign_tail_t<6, 5, 1 << 19, (Q_I - 1) / 32, 196, 55>(b, key_hat, cap_slots, sm_count, st);
        case 5: return launch_sign_tail_t<8, 7, 1 << 19, (Q_I - 1) / 32, 120, 75>(b, key_hat, cap_slots, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// one key per signature: key_items = n records of (s1_hat | s2_hat | t0_hat), canonical, pre-multiplied by 256^-1
cudaError_t launch_sign_tail_multi(int level, const SignBufs& b, const int32_t* key_items, uint32_t cap_slots, int sm_count, cudaStream_t st) {
    if (cap_slots == 0) return cudaSuccess;
    switch (level) {
        case 2: return launch_sign_tail_t<4, 4, 1 << 17, (Q_I - 1) / 88, 78, 80, true>(b, key_items, cap_slots, sm_count, st);
        case 3: return launch_sign_tail_t<6, 5, 1 << 19, (Q_I - 1) / 32, 196, 55, true>(b, key_items, cap_slots, sm_count, st);
        case 5: return launch_sign_tail_t<8, 7, 1 << 19, (Q_I - 1) / 32, 120, 75, true>(b, key_items, cap_slots, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_resolve(int level, const SignBufs& b, uint32_t cap_items, cudaStream_t st) {
    if (cap_items == 0) return cudaSuccess;
    const unsigned grid = (cap_items + 7) / 8;
    const ResolveArgs ra = resolve_args(b);
    switch (level) {
        case 2: resolve_kernel<4, 17, 84><<<grid, 256, 0, st>>>(ra, b.y, b.h_slot, b.accepted); break;
        case 3: resolve_kernel<5, 19, 61><<<grid, 256, 0, st>>>(ra, b.y, b.h_slot, b.accepted); break;
        case 5: resolve_kernel<7, 19, 83><<<grid, 256, 0, st>>>(ra, b.y, b.h_slot, b.accepted); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace dil

// =======================================================================================
// Signing with ONE KEY PER SIGNATURE (dil_sign_multi_*): the reference's sign driver streams rho, tr, K, s1, s2, t0 in
// front of every message (rtl_tb/tb_sign_top.v:171-284), i.e. every signature may use a different key.  Key material is
// unpacked and transformed per item on the device; A_hat is expanded once per item into HBM and the mat-vec reads the
// item's own matrix (the per-round kernels below); ExpandMask, challenge and resolve are the shared-key kernels.
// =======================================================================================
namespace dil {

// out[item][first + p][8 g .. 8 g + 7] = bias - field(p, g) for the bit-packed polynomials of every item's key field
// (decoder.v:89-143: s as eta - x in 3 / 4 bits, t0 as 2^12 - x in 13 bits); one thread per 8 coefficients
__global__ void __launch_bounds__(256) unpack_key_field_kernel(int32_t* __restrict__ out, const uint8_t* __restrict__ in, size_t n_groups,
                                                               int width, int bias, int polys, int first, int nkey) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const size_t item = g / ((size_t)polys * 32);
    const uint32_t r = (uint32_t)(g % ((size_t)polys * 32)), p = r >> 5, q = r & 31;
    const uint8_t* src = in + g * (size_t)width;       // fields are contiguous per item: item * polys * 32 * width + ...
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < width; i++) {
        if (i < 8) lo |= (uint64_t)src[i] << (8 * i);
        else hi |= (uint64_t)src[i] << (8 * (i - 8));
    }
    int32_t o[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int pos = width * c;
        uint64_t v = pos < 64 ? (lo >> pos) | (pos && pos + width > 64 ? hi << (64 - pos) : 0) : hi >> (pos - 64);
        o[c] = bias - (int32_t)(v & ((1u << width) - 1));
    }
    int4* dst = reinterpret_cast<int4*>(out + ((item * nkey + first + p) * N)) + 2 * q;
    dst[0] = make_int4(o[0], o[1], o[2], o[3]);
    dst[1] = make_int4(o[4], o[5], o[6], o[7]);
}

// x <- x * 256^-1 mod Q (canonical in, canonical out); the tail's inverse transforms skip their scaling
__global__ void __launch_bounds__(256) scale_inv256_kernel(int32_t* __restrict__ x, size_t n_vec) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vec) return;
    int4 v = reinterpret_cast<int4*>(x)[t];
    reinterpret_cast<int4*>(x)[t] = make_int4((int)mul_full((uint32_t)v.x, INV256), (int)mul_full((uint32_t)v.y, INV256),
                                              (int)mul_full((uint32_t)v.z, INV256), (int)mul_full((uint32_t)v.w, INV256));
}

// w_hat[a][i] = sum_j A_hat[item(a)][i*l + j] o y_hat[a][j]: the MULT_A_Y loop nest (combined_top.v:1875-1913) with the
// slot's own matrix; one thread per 4 coefficients of one output polynomial
template <int K, int L>
__global__ void __launch_bounds__(256) matvec_multi_kernel(int32_t* __restrict__ wh, const int32_t* __restrict__ a_items,
                                                           const int32_t* __restrict__ yh, const uint32_t* __restrict__ active0,
                                                           const uint32_t* __restrict__ active1, const RoundCtl* __restrict__ ctl) {
    const uint32_t n_slots = ctl->n_slots, spec = ctl->spec;
    const uint32_t* __restrict__ active = ctl->cur ? active1 : active0;
    const size_t total = (size_t)n_slots * K * 64;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(t & 63), i = (uint32_t)((t >> 6) % K);
        const size_t a = t / ((size_t)K * 64);
        const int4* A = reinterpret_cast<const int4*>(a_items + ((size_t)active[a / spec] * K * L + (size_t)i * L) * N) + c;
        const int4* Y = reinterpret_cast<const int4*>(yh + a * L * N) + c;
        uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int4 av = __ldg(A + j * 64), yv = Y[j * 64];
            acc[0] += (uint64_t)(uint32_t)av.x * (uint32_t)yv.x; acc[1] += (uint64_t)(uint32_t)av.y * (uint32_t)yv.y;
            acc[2] += (uint64_t)(uint32_t)av.z * (uint32_t)yv.z; acc[3] += (uint64_t)(uint32_t)av.w * (uint32_t)yv.w;
        }
        reinterpret_cast<int4*>(wh + (a * K + i) * N)[c] =
            make_int4((int)reduce49(acc[0]), (int)reduce49(acc[1]), (int)reduce49(acc[2]), (int)reduce49(acc[3]));
    }
}

// w1 = HighBits(w), bit-packed (encoder.v:96-133: 6 bits for gamma2 = (Q-1)/88, else 4); one thread per 16 coefficients
template <int32_t GAMMA2>
__global__ void __launch_bounds__(256) pack_w1_kernel(uint32_t* __restrict__ w1p, const int32_t* __restrict__ w, size_t n_groups) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const int4* src = reinterpret_cast<const int4*>(w) + t * 4;
    uint32_t h[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int4 v = src[q];
        h[4 * q + 0] = highbits<GAMMA2>((uint32_t)v.x); h[4 * q + 1] = highbits<GAMMA2>((uint32_t)v.y);
        h[4 * q + 2] = highbits<GAMMA2>((uint32_t)v.z); h[4 * q + 3] = highbits<GAMMA2>((uint32_t)v.w);
    }
    if constexpr (GAMMA2 == (Q_I - 1) / 32) {
        uint32_t o0 = 0, o1 = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            o0 |= h[c] << (4 * c);
            o1 |= h[8 + c] << (4 * c);
        }
        w1p[t * 2] = o0;
        w1p[t * 2 + 1] = o1;
    } else {
        uint64_t lo = 0;
#pragma unroll
        for (int c = 0; c < 10; c++) lo |= (uint64_t)h[c] << (6 * c);           // bits 0..59
        lo |= (uint64_t)h[10] << 60;                                              // bits 60..65
        const uint32_t hi = (h[10] >> 4) | (h[11] << 2) | (h[12] << 8) | (h[13] << 14) | (h[14] << 20) | (h[15] << 26);
        w1p[t * 3] = (uint32_t)lo;
        w1p[t * 3 + 1] = (uint32_t)(lo >> 32);
        w1p[t * 3 + 2] = hi;
    }
}

cudaError_t launch_unpack_keys(int level, int32_t* key_items, const uint8_t* s1p, const uint8_t* s2p, const uint8_t* t0p, size_t n,
                               cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    const int nkey = P.l + 2 * P.k, sw = P.eta == 2 ? 3 : 4;
    auto go = [&](const uint8_t* in, int width, int bias, int polys, int first) {
        const size_t groups = n * (size_t)polys * 32;
        unpack_key_field_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(key_items, in, groups, width, bias, polys, first, nkey);
    };
    go(s1p, sw, P.eta, P.l, 0);
    go(s2p, sw, P.eta, P.k, P.l);
    go(t0p, 13, 1 << 12, P.k, P.l + P.k);
    return cudaGetLastError();
}
cudaError_t launch_scale_inv256(int32_t* x, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    const size_t n_vec = n_polys * 64;
    scale_inv256_kernel<<<(unsigned)((n_vec + 255) / 256), 256, 0, st>>>(x, n_vec);
    return cudaGetLastError();
}
cudaError_t launch_matvec_multi(int level, int32_t* wh, const int32_t* a_items, const int32_t* yh, const SignBufs& b, uint32_t cap_slots,
                                cudaStream_t st) {
    if (cap_slots == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    const size_t total = (size_t)cap_slots * P.k * 64;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148u * 64u ? (total + 255) / 256 : 148u * 64u);
    switch (level) {
        case 2: matvec_multi_kernel<4, 4><<<grid, 256, 0, st>>>(wh, a_items, yh, b.active[0], b.active[1], b.ctl); break;
        case 3: matvec_multi_kernel<6, 5><<<grid, 256, 0, st>>>(wh, a_items, yh, b.active[0], b.active[1], b.ctl); break;
        case 5: matvec_multi_kernel<8, 7><<<grid, 256, 0, st>>>(wh, a_items, yh, b.active[0], b.active[1], b.ctl); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_pack_w1(int level, uint32_t* w1p, const int32_t* w, size_t n_slots, cudaStream_t st) {
    if (n_slots == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    const size_t n_groups = n_slots * P.k * (N / 16);
    const unsigned grid = (unsigned)((n_groups + 255) / 256);
    if (level == 2) pack_w1_kernel<(Q_I - 1) / 88><<<grid, 256, 0, st>>>(w1p, w, n_groups);
    else pack_w1_kernel<(Q_I - 1) / 32><<<grid, 256, 0, st>>>(w1p, w, n_groups);
    return cudaGetLastError();
}

}  // namespace dil

// =======================================================================================
// Verification (rtl_src/combined_top.v:1080-1534, I/O order rtl_tb/tb_verify_top.v:144-249):
//   w' = INTT( A_hat * NTT(z) - NTT(c) o NTT(t1 * 2^13) ),  w1' = UseHint(h, w'),
//   accept iff SHAKE256(mu || pack(w1')) == c~   (and ||z|| < gamma1 - beta, h well formed).
// The core is the same fused kernel as signing with one extra column: inputs [z_0..z_{l-1}, c],
// matrix [A_hat | -t1_hat * 2^13] of shape k x (l+1)  (matvec_kernels.cu, launch_verify_core).
// =======================================================================================
namespace dil {

// mu = SHAKE256(tr || M); one thread per item
__global__ void __launch_bounds__(128) verify_mu_kernel(uint64_t* __restrict__ mu, const uint8_t* __restrict__ tr_base,
                                                        const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint32_t n,
                                                        uint32_t tr_stride) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint8_t* tr = tr_base + (size_t)t * tr_stride;   // stride 0: one key for the batch, 32: one key per item
    const uint8_t* m = msgs + offsets[t];
    // the host entry points validate the offsets; a non-monotonic pair in device memory reads as an empty message
    const size_t mlen = offsets[t + 1] >= offsets[t] ? (size_t)(offsets[t + 1] - offsets[t]) : 0;
    uint64_t A[25];
    shake256_absorb_lanes(A, 32 + mlen, [&](size_t idx) -> uint64_t {
        if (idx < 4) return load_lane_bytes(tr, idx * 8, 32);
        return load_lane_bytes(m, (idx - 4) * 8, mlen);
    });
#pragma unroll
    for (int i = 0; i < 8; i++) mu[(size_t)t * 8 + i] = A[i];
}

// tr = SHAKE256(rho || t1_packed)[0:32]   (combined_top.v:980); single thread, once per key
__global__ void tr_kernel(uint64_t* __restrict__ tr, const uint8_t* __restrict__ rho, const uint8_t* __restrict__ t1p, uint32_t t1_bytes) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t A[25];
    shake256_absorb_lanes(A, 32 + t1_bytes, [&](size_t idx) -> uint64_t {
        if (idx < 4) return load_lane_bytes(rho, idx * 8, 32);
        return load_lane_bytes(t1p, (idx - 4) * 8, t1_bytes);
    });
#pragma unroll
    for (int i = 0; i < 4; i++) tr[i] = A[i];
}

// z unpack (decoder.v:89-143: gamma1 - x, 18/20 bits) into v[item][j][256] of an (L+1)-poly item
// record, with the ||z||_inf < gamma1 - beta check; one thread per 8 coefficients
template <int L, int GAMMA1_BITS, int BETA>
__global__ void __launch_bounds__(256) unpack_z_kernel(int32_t* __restrict__ v, uint32_t* __restrict__ bad, const uint8_t* __restrict__ zp,
                                                       size_t n_groups) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    constexpr int BITS = GAMMA1_BITS + 1;
    constexpr int32_t G1 = 1 << GAMMA1_BITS;
    const size_t item = t / (L * 32);
    const uint32_t g = (uint32_t)(t % (L * 32));     // group inside the item's l polynomials
    const uint8_t* src = zp + item * (size_t)(L * 32 * BITS) + (size_t)g * BITS;
    uint64_t lo, mid;
    uint32_t hi;
    if constexpr (BITS == 18) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(src);
        lo = (uint64_t)h[0] | ((uint64_t)h[1] << 16) | ((uint64_t)h[2] << 32) | ((uint64_t)h[3] << 48);
        mid = (uint64_t)h[4] | ((uint64_t)h[5] << 16) | ((uint64_t)h[6] << 32) | ((uint64_t)h[7] << 48);
        hi = h[8];
    } else {
        const uint32_t* h = reinterpret_cast<const uint32_t*>(src);
        lo = (uint64_t)h[0] | ((uint64_t)h[1] << 32);
        mid = (uint64_t)h[2] | ((uint64_t)h[3] << 32);
        hi = h[4];
    }
    int32_t o[8];
    bool b = false;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int pos = c * BITS;
        uint64_t bits;
        if (pos + BITS <= 64) bits = lo >> pos;
        else if (pos < 64) bits = (lo >> pos) | (mid << (64 - pos));
        else if (pos + BITS <= 128) bits = mid >> (pos - 64);
        else if (pos < 128) bits = (mid >> (pos - 64)) | ((uint64_t)hi << (128 - pos));
        else bits = hi >> (pos - 128);
        o[c] = G1 - (int32_t)((uint32_t)bits & ((1u << BITS) - 1));
        b |= (o[c] >= G1 - BETA) || (o[c] <= -(G1 - BETA));
    }
    int4* dst = reinterpret_cast<int4*>(v + item * (size_t)((L + 1) * N) + (size_t)g * 8);
    dst[0] = make_int4(o[0], o[1], o[2], o[3]);
    dst[1] = make_int4(o[4], o[5], o[6], o[7]);
    if (b) bad[item] = 1;
}

// per item: hint decode (omega position bytes + k running counts) with the standard well-formedness
// checks -> 256-bit masks; c = SampleInBall(c~) written as the (L+1)-th polynomial of the item record.
// One thread per item, but everything that is indexed with data-dependent positions (hint bytes, hint masks,
// the squeezed bytes and the challenge polynomial) lives in a private shared-memory row, and the warp moves
// its 32 items' inputs and outputs cooperatively (coalesced) - as per-thread local arrays with per-thread
// global rows this kernel cost more than the transform core of the whole verification.
constexpr int VP_THREADS = 64;
template <int K, int L, int OMEGA, int TAU>
__global__ void __launch_bounds__(VP_THREADS) verify_prep_kernel(int32_t* __restrict__ v, uint32_t* __restrict__ hmask, uint32_t* __restrict__ bad,
                                                                 const uint8_t* __restrict__ h, const uint64_t* __restrict__ ctilde, uint32_t n) {
    constexpr int HB = OMEGA + K;
    constexpr int HB_STRIDE = (HB + 3) / 4 * 4 + 4;      // bytes; breaks the power-of-two bank pattern
    constexpr int M_STRIDE = K * 8 + 1;                  // words
    __shared__ __align__(16) uint8_t c_sm[VP_THREADS * CH_C_STRIDE];
    __shared__ __align__(16) uint8_t buf_sm[VP_THREADS * CH_BUF_STRIDE];
    __shared__ __align__(4) uint8_t hp_sm[VP_THREADS * HB_STRIDE];
    __shared__ uint32_t m_sm[VP_THREADS * M_STRIDE];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * VP_THREADS + warp * 32;   // first item of this warp
    if (base >= n) return;
    const uint32_t rows = n - base < 32 ? n - base : 32;
    const uint32_t t = base + lane;
    // stage the warp's hint strings (contiguous in global memory)
    {
        const uint8_t* src = h + (size_t)base * HB;
        uint8_t* dstw = hp_sm + (size_t)warp * 32 * HB_STRIDE;
        for (uint32_t i = lane; i < rows * HB; i += 32) dstw[(i / HB) * HB_STRIDE + (i % HB)] = src[i];
    }
    uint8_t* c = c_sm + threadIdx.x * CH_C_STRIDE;
    uint8_t* buf = buf_sm + threadIdx.x * CH_BUF_STRIDE;
    const uint8_t* hp = hp_sm + threadIdx.x * HB_STRIDE;
    uint32_t* m = m_sm + threadIdx.x * M_STRIDE;
    __syncwarp();
    if (t < n) {
#pragma unroll
        for (int i = 0; i < K * 8; i++) m[i] = 0;
        bool b = false;
        int idx = 0;
        for (int i = 0; i < K; i++) {
            int end = hp[OMEGA + i];
            if (end < idx || end > OMEGA) { b = true; break; }
            for (int j = idx; j < end; j++) {
                int pos = hp[j];
                if (j > idx && pos <= hp[j - 1]) b = true;
                m[i * 8 + (pos >> 5)] |= 1u << (pos & 31);
            }
            idx = end;
        }
        for (int j = idx; j < OMEGA && !b; j++)
            if (hp[j]) b = true;
        if (b) bad[t] = 1;
        // SampleInBall
        uint64_t A[25];
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) A[i] = ctilde[(size_t)t * 4 + i];
        A[4] = 0x1F;
        A[16] = 0x80ULL << 56;
        keccak_f1600(A);
        uint64_t signs = A[0];
#pragma unroll
        for (int i = 0; i < 17; i++) reinterpret_cast<uint64_t*>(buf)[i] = A[i];
#pragma unroll
        for (int i = 0; i < N / 16; i++) reinterpret_cast<uint4*>(c)[i] = make_uint4(0, 0, 0, 0);
        int pos = 8;
        for (int i = N - TAU; i < N; i++) {
            int bb;
            do {
                if (pos == 136) {
                    keccak_f1600(A);
#pragma unroll
                    for (int q = 0; q < 17; q++) reinterpret_cast<uint64_t*>(buf)[q] = A[q];
                    pos = 0;
                }
                bb = buf[pos++];
            } while (bb > i);
            c[i] = c[bb];
            c[bb] = (signs & 1) ? (uint8_t)0xFF : (uint8_t)1;
            signs >>= 1;
        }
    }
    __syncwarp();
    // cooperative write-out: hint masks (contiguous for the warp's items) and c as int32 polynomials
    {
        uint32_t* dstm = hmask + (size_t)base * (K * 8);
        const uint32_t* srcm = m_sm + (size_t)warp * 32 * M_STRIDE;
        for (uint32_t i = lane; i < rows * (K * 8); i += 32) dstm[i] = srcm[(i / (K * 8)) * M_STRIDE + (i % (K * 8))];
        const uint8_t* srcc = c_sm + (size_t)warp * 32 * CH_C_STRIDE;
        for (uint32_t r = 0; r < rows; r++) {
            int4* dst = reinterpret_cast<int4*>(v + (size_t)(base + r) * ((L + 1) * N) + L * N);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t wd = reinterpret_cast<const uint32_t*>(srcc + r * CH_C_STRIDE)[lane + 32 * q];
                dst[lane + 32 * q] = make_int4((int8_t)(wd & 0xFF), (int8_t)((wd >> 8) & 0xFF), (int8_t)((wd >> 16) & 0xFF), (int8_t)(wd >> 24));
            }
        }
    }
}

// w1' = UseHint(h, w') (usehint.v:134-155), packed; one thread per 16 coefficients
template <int K, int32_t GAMMA2>
__global__ void __launch_bounds__(256) usehint_pack_kernel(uint32_t* __restrict__ w1p, const int32_t* __restrict__ w,
                                                           const uint32_t* __restrict__ hmask, size_t n_groups) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const int4* src = reinterpret_cast<const int4*>(w) + t * 4;
    // group t covers coefficients 16*(t%16) .. +15 of polynomial t/16; hint word (t%16)>>1, half (t&1)
    const uint32_t hw = hmask[(t >> 4) * 8 + ((t & 15) >> 1)] >> ((t & 1) * 16);
    constexpr int32_t M = (Q_I - 1) / (2 * GAMMA2);   // 44 or 16
    int32_t hh[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int4 x = src[q];
        int32_t in[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int32_t a1, a0;
            decompose<GAMMA2>(in[e], a1, a0);
            if ((hw >> (4 * q + e)) & 1u) a1 = a0 > 0 ? (a1 + 1 == M ? 0 : a1 + 1) : (a1 == 0 ? M - 1 : a1 - 1);
            hh[4 * q + e] = a1;
        }
    }
    if constexpr (GAMMA2 == (Q_I - 1) / 32) {
        uint32_t o0 = 0, o1 = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            o0 |= (uint32_t)hh[c] << (4 * c);
            o1 |= (uint32_t)hh[8 + c] << (4 * c);
        }
        w1p[t * 2] = o0;
        w1p[t * 2 + 1] = o1;
    } else {
        uint64_t lo = 0;
#pragma unroll
        for (int c = 0; c < 10; c++) lo |= (uint64_t)hh[c] << (6 * c);
        lo |= (uint64_t)hh[10] << 60;
        uint32_t hi = ((uint32_t)hh[10] >> 4) | ((uint32_t)hh[11] << 2) | ((uint32_t)hh[12] << 8) | ((uint32_t)hh[13] << 14) |
                      ((uint32_t)hh[14] << 20) | ((uint32_t)hh[15] << 26);
        w1p[t * 3] = (uint32_t)lo;
        w1p[t * 3 + 1] = (uint32_t)(lo >> 32);
        w1p[t * 3 + 2] = hi;
    }
}

// c~' = SHAKE256(mu || w1'_packed)[0:32]; ok = (c~' == c~) && !bad; one thread per item
template <int K, int W1_BYTES>
__global__ void __launch_bounds__(128) verify_hash_kernel(uint8_t* __restrict__ ok, const uint64_t* __restrict__ mu,
                                                          const uint64_t* __restrict__ w1p, const uint64_t* __restrict__ ctilde,
                                                          const uint32_t* __restrict__ bad, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    constexpr int W1_LANES = K * W1_BYTES / 8;
    const uint64_t* m = mu + (size_t)t * 8;
    const uint64_t* w = w1p + (size_t)t * W1_LANES;
    uint64_t A[25];
    shake256_absorb_lanes(A, 64 + K * W1_BYTES, [&](size_t idx) -> uint64_t { return idx < 8 ? m[idx] : w[idx - 8]; });
    bool same = true;
#pragma unroll
    for (int i = 0; i < 4; i++) same &= (A[i] == ctilde[(size_t)t * 4 + i]);
    ok[t] = (same && bad[t] == 0) ? 1 : 0;
}

// ---- launchers ----
cudaError_t launch_verify_mu(uint64_t* mu, const uint8_t* tr, const uint8_t* msgs, const uint64_t* offsets, uint32_t n,
                             uint32_t tr_stride, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    verify_mu_kernel<<<(n + 127) / 128, 128, 0, st>>>(mu, tr, msgs, offsets, n, tr_stride);
    return cudaGetLastError();
}

// -t1 * 2^13 mod Q from the 10-bit packed t1 (decoder.v:96-100); one thread per 8 coefficients (10 bytes)
__global__ void __launch_bounds__(256) unpack_t1neg_kernel(int32_t* __restrict__ out, const uint8_t* __restrict__ t1p, size_t n_groups) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint8_t* src = t1p + g * 10;
    uint64_t lo = 0;
    uint32_t hi;
#pragma unroll
    for (int i = 0; i < 8; i++) lo |= (uint64_t)src[i] << (8 * i);
    hi = (uint32_t)src[8] | ((uint32_t)src[9] << 8);
    int32_t o[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int pos = 10 * c;
        uint32_t v = pos + 10 <= 64 ? (uint32_t)(lo >> pos)
                     : (pos < 64 ? (uint32_t)((lo >> pos) | ((uint64_t)hi << (64 - pos))) : (hi >> (pos - 64)));
        v &= 1023u;
        uint32_t m = (v << D_BITS) % Q;        // t1 * 2^13 < 2^23: at most one subtraction
        o[c] = (int32_t)((Q - m) % Q);
    }
    int4* dst = reinterpret_cast<int4*>(out) + g * 2;
    dst[0] = make_int4(o[0], o[1], o[2], o[3]);
    dst[1] = make_int4(o[4], o[5], o[6], o[7]);
}
cudaError_t launch_unpack_t1neg(int32_t* out, const uint8_t* t1p, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    size_t n_groups = n_polys * 32;
    unpack_t1neg_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(out, t1p, n_groups);
    return cudaGetLastError();
}
cudaError_t launch_tr(uint64_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, cudaStream_t st) {
    tr_kernel<<<1, 32, 0, st>>>(tr, rho, t1p, t1_bytes);
    return cudaGetLastError();
}
cudaError_t launch_unpack_z(int level, int32_t* v, uint32_t* bad, const uint8_t* zp, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    size_t n_groups = (size_t)n * P.l * 32;
    unsigned grid = (unsigned)((n_groups + 255) / 256);
    switch (level) {
        case 2: unpack_z_kernel<4, 17, 78><<<grid, 256, 0, st>>>(v, bad, zp, n_groups); break;
        case 3: unpack_z_kernel<5, 19, 196><<<grid, 256, 0, st>>>(v, bad, zp, n_groups); break;
        case 5: unpack_z_kernel<7, 19, 120><<<grid, 256, 0, st>>>(v, bad, zp, n_groups); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_verify_prep(int level, int32_t* v, uint32_t* hmask, uint32_t* bad, const uint8_t* h, const uint64_t* ctilde,
                               uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned grid = (n + VP_THREADS - 1) / VP_THREADS;
    switch (level) {
        case 2: verify_prep_kernel<4, 4, 80, 39><<<grid, VP_THREADS, 0, st>>>(v, hmask, bad, h, ctilde, n); break;
        case 3: verify_prep_kernel<6, 5, 55, 49><<<grid, VP_THREADS, 0, st>>>(v, hmask, bad, h, ctilde, n); break;
        case 5: verify_prep_kernel<8, 7, 75, 60><<<grid, VP_THREADS, 0, st>>>(v, hmask, bad, h, ctilde, n); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_usehint_pack(int level, uint32_t* w1p, const int32_t* w, const uint32_t* hmask, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    size_t n_groups = (size_t)n * P.k * 16;
    unsigned grid = (unsigned)((n_groups + 255) / 256);
    switch (level) {
        case 2: usehint_pack_kernel<4, (Q_I - 1) / 88><<<grid, 256, 0, st>>>(w1p, w, hmask, n_groups); break;
        case 3: usehint_pack_kernel<6, (Q_I - 1) / 32><<<grid, 256, 0, st>>>(w1p, w, hmask, n_groups); break;
        case 5: usehint_pack_kernel<8, (Q_I - 1) / 32><<<grid, 256, 0, st>>>(w1p, w, hmask, n_groups); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_verify_hash(int level, uint8_t* ok, const uint64_t* mu, const uint64_t* w1p, const uint64_t* ctilde,
                               const uint32_t* bad, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned grid = (n + 127) / 128;
    switch (level) {
        case 2: verify_hash_kernel<4, 192><<<grid, 128, 0, st>>>(ok, mu, w1p, ctilde, bad, n); break;
        case 3: verify_hash_kernel<6, 128><<<grid, 128, 0, st>>>(ok, mu, w1p, ctilde, bad, n); break;
        case 5: verify_hash_kernel<8, 128><<<grid, 128, 0, st>>>(ok, mu, w1p, ctilde, bad, n); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace dil

// =======================================================================================
// Key generation (rtl_src/combined_top.v:754-1079; outputs in the order of
// rtl_tb/tb_keygen_top.v:180-275: rho, K, s1, s2, t1, t0, tr)
// =======================================================================================
namespace dil {

// (rho, rho', K) = SHAKE256(xi)[0:128]  (combined_top.v:776, :805-827); thread per item
__global__ void __launch_bounds__(128) keygen_seed_kernel(uint8_t* __restrict__ rho, uint64_t* __restrict__ rhop, uint8_t* __restrict__ key,
                                                          const uint8_t* __restrict__ xi, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) A[i] = load_lane_bytes(xi + (size_t)t * 32, i * 8, 32);
    A[4] = 0x1F;
    A[16] = 0x80ULL << 56;
    keccak_f1600(A);
    uint64_t* r = reinterpret_cast<uint64_t*>(rho) + (size_t)t * 4;
    uint64_t* kk = reinterpret_cast<uint64_t*>(key) + (size_t)t * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = A[i];
#pragma unroll
    for (int i = 0; i < 8; i++) rhop[(size_t)t * 8 + i] = A[4 + i];
#pragma unroll
    for (int i = 0; i < 4; i++) kk[i] = A[12 + i];
}

// s1_j = RejEta(SHAKE256(rho' || u16le(j))), s2_i uses nonce l+i (gen_s.v:115, sampler_s.v:117-135,
// rejection_s.v:47-51,:85-138); one Keccak state per polynomial.  The accepted coefficients land in a private
// shared-memory row per thread (bytes, data-dependent index), then the warp writes its 32 polynomials out
// cooperatively as centred int32 coefficients: HBM sees coalesced 1 KiB polynomial writes.  (Per-thread global
// stores - 32 lanes writing to 32 different polynomials, one coefficient at a time - made this kernel five times
// slower than its Keccak work: 315 us per 16 384 level-2 keys, a third of key generation.)
constexpr int ETA_ROW = N + 16;   // bytes per staged polynomial: rows stay 16-byte aligned, banks are skewed
template <int K, int L, int ETA>
__global__ void __launch_bounds__(128) eta_sample_kernel(int32_t* __restrict__ s1, int32_t* __restrict__ s2,
                                                         const uint64_t* __restrict__ rhop, uint32_t n) {
    __shared__ __align__(16) int8_t stage_sm[128 * ETA_ROW];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;    // first polynomial of this warp
    const uint32_t total = n * (K + L);
    if (wbase >= total) return;
    const uint32_t t = wbase + lane;
    int8_t* row = stage_sm + threadIdx.x * ETA_ROW;
    if (t < total) {
        const uint32_t item = t / (K + L), p = t % (K + L);
        uint64_t A[25];
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) A[i] = rhop[(size_t)item * 8 + i];
        A[8] = (uint64_t)p | (0x1FULL << 16);
        A[16] = 0x80ULL << 56;
        int cnt = 0;
        while (cnt < N) {
            keccak_f1600(A);
#pragma unroll
            for (int ln = 0; ln < 17; ln++) {
                if (cnt >= N) break;          // the second block usually contributes only a handful of coefficients
                const uint64_t v = A[ln];
#pragma unroll
                for (int nib = 0; nib < 16; nib++) {
                    const uint32_t x = (uint32_t)(v >> (4 * nib)) & 15u;
                    if (ETA == 2) {
                        if (x < 15 && cnt < N) row[cnt++] = (int8_t)(2 - (int32_t)(x - (205 * x >> 10) * 5));
                    } else {
                        if (x < 9 && cnt < N) row[cnt++] = (int8_t)(4 - (int32_t)x);
                    }
                }
            }
        }
    }
    __syncwarp();
    // polynomial q of the warp: lane writes coefficients 8*lane .. 8*lane+7
    const uint32_t rows = total - wbase < 32 ? total - wbase : 32;
    for (uint32_t q = 0; q < rows; q++) {
        const uint32_t tq = wbase + q, item = tq / (K + L), p = tq % (K + L);
        int32_t* out = p < L ? s1 + ((size_t)item * L + p) * N : s2 + ((size_t)item * K + (p - L)) * N;
        const uint2 b = *reinterpret_cast<const uint2*>(stage_sm + (size_t)(warp * 32 + q) * ETA_ROW + 8 * lane);
        int4* dst = reinterpret_cast<int4*>(out) + 2 * lane;
        dst[0] = make_int4((int8_t)(b.x), (int8_t)(b.x >> 8), (int8_t)(b.x >> 16), (int8_t)(b.x >> 24));
        dst[1] = make_int4((int8_t)(b.y), (int8_t)(b.y >> 8), (int8_t)(b.y >> 16), (int8_t)(b.y >> 24));
    }
    __syncwarp();
    // s1 / s2 are secret: nothing of them stays behind in shared memory
#pragma unroll
    for (int i = 0; i < ETA_ROW / 16; i++) reinterpret_cast<uint4*>(row)[i] = make_uint4(0, 0, 0, 0);
}

// t = t + s2 (canonical); Power2Round (uncenter_coeff.v:54-55); pack t1 (10 bit) and 2^12 - t0 (13 bit).
// One thread per 8 coefficients -> 10 + 13 bytes.
__global__ void __launch_bounds__(256) t_pack_kernel(uint8_t* __restrict__ t1p, uint8_t* __restrict__ t0p, const int32_t* __restrict__ t,
                                                     const int32_t* __restrict__ s2, size_t n_groups) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int4* a = reinterpret_cast<const int4*>(t) + g * 2;
    const int4* b = reinterpret_cast<const int4*>(s2) + g * 2;
    int4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    int32_t tv[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    int32_t sv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint64_t hi10 = 0, lo10 = 0;   // 80 bits of t1
    uint64_t p0 = 0, p1 = 0;       // 104 bits of t0
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint32_t x = csub((uint32_t)tv[c] + canon_signed(sv[c]));
        uint32_t r1 = (x + (1u << (D_BITS - 1)) - 1) >> D_BITS;
        uint32_t r0 = (uint32_t)((1 << (D_BITS - 1)) - ((int32_t)x - (int32_t)(r1 << D_BITS)));   // 2^12 - t0, 13 bits
        const int q1 = 10 * c, q0 = 13 * c;
        if (q1 < 64) { lo10 |= (uint64_t)r1 << q1; if (q1 + 10 > 64) hi10 |= (uint64_t)r1 >> (64 - q1); } else hi10 |= (uint64_t)r1 << (q1 - 64);
        if (q0 < 64) { p0 |= (uint64_t)r0 << q0; if (q0 + 13 > 64) p1 |= (uint64_t)r0 >> (64 - q0); } else p1 |= (uint64_t)r0 << (q0 - 64);
    }
    // a warp holds one polynomial (32 groups): 320 + 416 packed bytes.  With 16-byte aligned outputs they are staged in shared
    // memory and leave as 20 + 26 16-byte vectors; otherwise every thread stores its 10 + 13 bytes itself.
    __shared__ __align__(16) uint8_t stage[8 * 736];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = ((reinterpret_cast<uintptr_t>(t1p) | reinterpret_cast<uintptr_t>(t0p)) & 15u) == 0;   // uniform
    uint8_t* d1 = vec ? stage + warp * 736 + lane * 10 : t1p + g * 10;
    uint8_t* d0 = vec ? stage + warp * 736 + 320 + lane * 13 : t0p + g * 13;
#pragma unroll
    for (int i = 0; i < 8; i++) d1[i] = (uint8_t)(lo10 >> (8 * i));
    d1[8] = (uint8_t)hi10; d1[9] = (uint8_t)(hi10 >> 8);
#pragma unroll
    for (int i = 0; i < 8; i++) d0[i] = (uint8_t)(p0 >> (8 * i));
#pragma unroll
    for (int i = 0; i < 5; i++) d0[8 + i] = (uint8_t)(p1 >> (8 * i));
    if (vec) {   // n_groups is a multiple of 32: the whole warp is here
        __syncwarp();
        const size_t poly = g >> 5;
        const uint4* src = reinterpret_cast<const uint4*>(stage + warp * 736);
        if (lane < 20) reinterpret_cast<uint4*>(t1p + poly * 320)[lane] = src[lane];
        if (lane < 26) reinterpret_cast<uint4*>(t0p + poly * 416)[lane] = src[20 + lane];
    }
}

// pack eta - s (3 or 4 bits); one thread per 8 coefficients -> 3 or 4 bytes
template <int ETA>
__global__ void __launch_bounds__(256) s_pack_kernel(uint8_t* __restrict__ sp, const int32_t* __restrict__ s, size_t n_groups) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int4* a = reinterpret_cast<const int4*>(s) + g * 2;
    int4 a0 = a[0], a1 = a[1];
    int32_t v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    constexpr int W = ETA == 2 ? 3 : 4;
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) o |= (uint32_t)(ETA - v[c]) << (W * c);
    uint8_t* d = sp + g * W;
#pragma unroll
    for (int i = 0; i < W; i++) d[i] = (uint8_t)(o >> (8 * i));
}

// tr = SHAKE256(rho || t1_packed)[0:32] per item: one Keccak state per thread, but the 1312 / 1952 / 2592 input bytes of the
// warp's 32 items are fetched COOPERATIVELY, one 136-byte rate block of every item at a time (contiguous 136-byte runs,
// 8-byte loads) into a shared-memory tile from which each thread absorbs its own row.  (Every thread walking its own
// record with private loads ran at 44 % of the hash rate: 1.42 ms per 131 072 level-5 keys.)  Needs 8-byte aligned
// inputs and t1_bytes % 8 == 0 (true for the packed t1 of every level); otherwise the per-thread path below is used.
constexpr int TR_ROW = 18;   // 64-bit words per staged row (17 rate lanes + 1: skews the banks)
__global__ void __launch_bounds__(128) tr_batch_kernel(uint8_t* __restrict__ tr, const uint8_t* __restrict__ rho,
                                                       const uint8_t* __restrict__ t1p, uint32_t t1_bytes, uint32_t n) {
    __shared__ uint64_t tile[4 * 32 * TR_ROW];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    if (wbase >= n) return;
    const uint32_t t = wbase + lane;
    const bool coop = (((reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(t1p)) & 7u) == 0) && (t1_bytes % 8u == 0);   // uniform
    uint64_t A[25];
    if (coop) {
        uint64_t* wt = tile + warp * 32 * TR_ROW;
        const uint32_t total = 32 + t1_bytes, nfull = total / 136, rem_lanes = (total - nfull * 136) / 8;
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] = 0;
        for (uint32_t blk = 0; blk <= nfull; blk++) {
            const uint32_t lanes_here = blk < nfull ? 17u : rem_lanes;
            // stream word s of an item: words 0..3 are rho, the rest t1
            for (uint32_t idx = lane; idx < 32 * lanes_here; idx += 32) {
                const uint32_t row = idx / lanes_here, w = idx % lanes_here, sw = blk * 17 + w;
                const uint32_t item = wbase + row < n ? wbase + row : n - 1;
                wt[row * TR_ROW + w] = sw < 4 ? reinterpret_cast<const uint64_t*>(rho)[(size_t)item * 4 + sw]
                                              : reinterpret_cast<const uint64_t*>(t1p + (size_t)item * t1_bytes)[sw - 4];
            }
            __syncwarp();
            const uint64_t* my = wt + lane * TR_ROW;
#pragma unroll
            for (int i = 0; i < 17; i++)
                if ((uint32_t)i < lanes_here) A[i] ^= my[i];
            if (blk == nfull) {
#pragma unroll
                for (int i = 0; i < 17; i++)
                    if ((uint32_t)i == rem_lanes) A[i] ^= 0x1FULL;
                A[16] ^= 0x80ULL << 56;
            }
            keccak_f1600(A);
            __syncwarp();
        }
    } else if (t < n) {
        const uint8_t* r = rho + (size_t)t * 32;
        const uint8_t* p = t1p + (size_t)t * t1_bytes;
        shake256_absorb_lanes(A, 32 + t1_bytes, [&](size_t idx) -> uint64_t {
            if (idx < 4) return load_lane_bytes(r, idx * 8, 32);
            return load_lane_bytes(p, (idx - 4) * 8, t1_bytes);
        });
    }
    if (t >= n) return;
    uint64_t* o = reinterpret_cast<uint64_t*>(tr) + (size_t)t * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = A[i];
}

cudaError_t launch_keygen_seed(uint8_t* rho, uint64_t* rhop, uint8_t* key, const uint8_t* xi, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    keygen_seed_kernel<<<(n + 127) / 128, 128, 0, st>>>(rho, rhop, key, xi, n);
    return cudaGetLastError();
}
cudaError_t launch_eta_sample(int level, int32_t* s1, int32_t* s2, const uint64_t* rhop, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const LevelParams P = level_params(level);
    unsigned grid = (n * (P.k + P.l) + 127) / 128;
    switch (level) {
        case 2: eta_sample_kernel<4, 4, 2><<<grid, 128, 0, st>>>(s1, s2, rhop, n); break;
        case 3: eta_sample_kernel<6, 5, 4><<<grid, 128, 0, st>>>(s1, s2, rhop, n); break;
        case 5: eta_sample_kernel<8, 7, 2><<<grid, 128, 0, st>>>(s1, s2, rhop, n); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_t_pack(uint8_t* t1p, uint8_t* t0p, const int32_t* t, const int32_t* s2, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    size_t n_groups = n_polys * 32;
    t_pack_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(t1p, t0p, t, s2, n_groups);
    return cudaGetLastError();
}
cudaError_t launch_s_pack(int eta, uint8_t* sp, const int32_t* s, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    size_t n_groups = n_polys * 32;
    unsigned grid = (unsigned)((n_groups + 255) / 256);
    if (eta == 2) s_pack_kernel<2><<<grid, 256, 0, st>>>(sp, s, n_groups);
    else s_pack_kernel<4><<<grid, 256, 0, st>>>(sp, s, n_groups);
    return cudaGetLastError();
}
cudaError_t launch_tr_batch(uint8_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    tr_batch_kernel<<<(n + 127) / 128, 128, 0, st>>>(tr, rho, t1p, t1_bytes, n);
    return cudaGetLastError();
}

}  // namespace dil

// =======================================================================================
// Diagnostics: the pure Keccak-f[1600] rate of this GPU (the ALU-pipe speed of light that bench.py measures in the
// same run as the signing step and reports the Keccak-bound kernels against).  Same permutation code as every
// hash kernel above (keccak.cuh), one state per thread, nothing but permutations.
// =======================================================================================
namespace dil {
__global__ void __launch_bounds__(128) keccak_rate_kernel(uint64_t* __restrict__ out, uint32_t perms) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = (uint64_t)(gid + 1) * 0x9E3779B97F4A7C15ULL + (uint64_t)i;
    for (uint32_t p = 0; p < perms; p++) keccak_f1600(A);
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) x ^= A[i];
    out[gid] = x;
}
cudaError_t launch_keccak_rate(uint64_t* out, unsigned ctas, uint32_t perms, cudaStream_t st) {
    if (ctas == 0) return cudaSuccess;
    keccak_rate_kernel<<<ctas, 128, 0, st>>>(out, perms);
    return cudaGetLastError();
}
}  // namespace dil
