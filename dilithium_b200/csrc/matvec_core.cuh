// matvec_core.cuh - the per-item core shared by the fused kernels: w = [INTT] ( A_hat * v_hat ) for one item, executed
// by one warp with A_hat in shared memory (matvec_kernels.cu: sign / verify / keygen cores; mask_core.cu: the signing
// core with ExpandMask fused in front).  MULT_A_Y -> NTTI_W of rtl_src/combined_top.v:1875-1933.
#pragma once
#include "ntt_core.cuh"
#include "rounding.cuh"

namespace dil {

constexpr int A_STRIDE = 260;  // words per A polynomial in shared memory (pad 4: spreads the
                               // sampler's same-index stores over 8 bank groups, keeps 16-B alignment)

#ifdef __CUDACC__

// Rows [i_begin, i_end) of the product for one item whose NTT-domain inputs yh are already in registers (layout C).
// a_sm: k*l polynomials in shared memory (stride A_STRIDE), pre-multiplied by 256^-1 when INTT_OUT.
// EXTRA: the last input is multiplied by a per-item column extra_item[i] from global memory instead of a matrix column.
// W1: also emit w1 = HighBits(w), bit-packed as encoder.v:96-133 (6 bits for gamma2 = (Q-1)/88, i.e. K = 4, else 4 bits).
// PRESCALED: a_sm (and nothing else) already carries the 256^-1 factor of the inverse transform - worth it where one
// matrix serves a whole batch (scaled once per CTA); the per-item-rho kernels leave their matrices unscaled (scaling 256
// coefficients of each of the k*l generated polynomials would cost more than the 12 multiplications per inverse transform).
template <int K, int LA, bool INTT_OUT, bool EXTRA, bool W1, bool PRESCALED = true>
__device__ __forceinline__ void item_rows(int32_t* __restrict__ w_item, const uint32_t (&yh)[LA + (EXTRA ? 1 : 0)][8],
                                          const uint32_t* __restrict__ a_sm, uint32_t* __restrict__ scr, int lane,
                                          const int32_t* __restrict__ extra_item, uint8_t* __restrict__ w1_item, int i_begin,
                                          int i_end, int a_first_row = 0, const uint32_t* __restrict__ hint_item = nullptr) {
    // a_first_row: the matrix row held first in a_sm (kernels that stream A through shared memory a few rows at a time)
    // hint_item (W1 only; verification): k x 8 words of hint bits, bit b of word r of row i <-> coefficient 32 r + b.  The packed
    // output is then w1' = UseHint(h, w) (usehint.v:134-155) instead of HighBits(w), and w itself is not stored at all.
    constexpr int L = LA + (EXTRA ? 1 : 0);
#pragma unroll 1
    for (int i = i_begin; i < i_end; i++) {
        uint64_t acc[8];
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = 0;
#pragma unroll
        for (int j = 0; j < L; j++) {
            uint4 lo, hi;
            if (EXTRA && j == LA) {
                const int4* ep = reinterpret_cast<const int4*>(extra_item + i * N) + lane;
                int4 elo = __ldg(ep), ehi = __ldg(ep + 32);
                auto sc = [](int32_t x) -> uint32_t { return INTT_OUT && PRESCALED ? mul_full(canon_signed(x), INV256) : canon_signed(x); };
                lo = make_uint4(sc(elo.x), sc(elo.y), sc(elo.z), sc(elo.w));
                hi = make_uint4(sc(ehi.x), sc(ehi.y), sc(ehi.z), sc(ehi.w));
            } else {
                const uint4* ap = reinterpret_cast<const uint4*>(a_sm + ((i - a_first_row) * LA + j) * A_STRIDE) + lane;
                lo = ap[0];
                hi = ap[32];
            }
            acc[0] += (uint64_t)lo.x * yh[j][0]; acc[1] += (uint64_t)lo.y * yh[j][1];
            acc[2] += (uint64_t)lo.z * yh[j][2]; acc[3] += (uint64_t)lo.w * yh[j][3];
            acc[4] += (uint64_t)hi.x * yh[j][4]; acc[5] += (uint64_t)hi.y * yh[j][5];
            acc[6] += (uint64_t)hi.z * yh[j][6]; acc[7] += (uint64_t)hi.w * yh[j][7];
        }
        uint32_t x[8];
#pragma unroll
        for (int r = 0; r < 8; r++) x[r] = reduce49(acc[r]);
        if constexpr (INTT_OUT) {
            InvTw itw;
            {
                const TwTable* tab = &TW_INV;
                asm volatile("" : "+l"(tab));
                load_inv_tw(itw, tab, lane);
            }
            ntt_inv_warp<PRESCALED>(x, scr, itw, lane);
            __syncwarp();
            if (!(W1 && hint_item != nullptr)) {
                int32_t* o = w_item + i * N + lane;
#pragma unroll
                for (int r = 0; r < 8; r++) o[32 * r] = (int32_t)x[r];
            }
            if constexpr (W1) {
                static_assert(INTT_OUT, "w1 is defined on the time-domain w");
                constexpr int32_t G2 = K == 4 ? (Q_I - 1) / 88 : (Q_I - 1) / 32;
                // this lane holds coefficients lane + 32 r: stage HighBits as bytes, re-read 8 consecutive ones
                uint8_t* sb = reinterpret_cast<uint8_t*>(scr);
                if (hint_item != nullptr) {   // uniform
                    constexpr int32_t M = (Q_I - 1) / (2 * G2);   // 44 or 16
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        int32_t a1, a0;
                        decompose<G2>((int32_t)x[r], a1, a0);
                        if ((__ldg(hint_item + i * 8 + r) >> lane) & 1u) a1 = a0 > 0 ? (a1 + 1 == M ? 0 : a1 + 1) : (a1 == 0 ? M - 1 : a1 - 1);
                        sb[32 * r + lane] = (uint8_t)a1;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; r++) sb[32 * r + lane] = (uint8_t)highbits<G2>(x[r]);
                }
                __syncwarp();
                const uint2 q = reinterpret_cast<const uint2*>(sb)[lane];
                if constexpr (G2 == (Q_I - 1) / 32) {
                    auto p4 = [](uint32_t v) { return (v & 0xFu) | ((v >> 4) & 0xF0u) | ((v >> 8) & 0xF00u) | ((v >> 12) & 0xF000u); };
                    reinterpret_cast<uint32_t*>(w1_item + i * 128)[lane] = p4(q.x) | (p4(q.y) << 16);
                } else {
                    auto p6 = [](uint32_t v) { return (v & 0x3Fu) | ((v >> 2) & 0xFC0u) | ((v >> 4) & 0x3F000u) | ((v >> 6) & 0xFC0000u); };
                    const uint32_t a = p6(q.x), b = p6(q.y);   // 24 bits each
                    uint16_t* d = reinterpret_cast<uint16_t*>(w1_item + i * 192 + 6 * lane);
                    d[0] = (uint16_t)a;
                    d[1] = (uint16_t)((a >> 16) | (b << 8));
                    d[2] = (uint16_t)(b >> 8);
                }
                __syncwarp();   // the scratch is reused by the next transform
            }
        } else {
            int4* o = reinterpret_cast<int4*>(w_item + i * N) + lane;
            o[0] = make_int4((int)x[0], (int)x[1], (int)x[2], (int)x[3]);
            o[32] = make_int4((int)x[4], (int)x[5], (int)x[6], (int)x[7]);
        }
    }
}

// The item's L input polynomials into registers, NTT domain, layout C (transformed here when NTT_IN).
template <int L, bool NTT_IN>
__device__ __forceinline__ void item_inputs(uint32_t (&yh)[L][8], const int32_t* __restrict__ v_item, uint32_t* __restrict__ scr, int lane) {
    if constexpr (NTT_IN) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int32_t* p = v_item + j * N + lane;
#pragma unroll
            for (int r = 0; r < 8; r++) yh[j][r] = (uint32_t)p[32 * r];  // layout A: 128-B line per access
        }
        FwdTw ftw;
        {   // 2 KiB table, L1 resident; reloaded per item (opaque pointer defeats hoisting) to keep registers low
            const TwTable* tab = &TW_FWD;
            asm volatile("" : "+l"(tab));
            load_fwd_tw(ftw, tab, lane);
        }
#pragma unroll
        for (int j = 0; j < L; j++) {
            ntt_fwd_warp(yh[j], scr, ftw, lane);
            __syncwarp();
        }
    } else {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int4* p = reinterpret_cast<const int4*>(v_item + j * N) + lane;
            int4 lo = p[0], hi = p[32];
            yh[j][0] = canon_signed(lo.x); yh[j][1] = canon_signed(lo.y); yh[j][2] = canon_signed(lo.z); yh[j][3] = canon_signed(lo.w);
            yh[j][4] = canon_signed(hi.x); yh[j][5] = canon_signed(hi.y); yh[j][6] = canon_signed(hi.z); yh[j][7] = canon_signed(hi.w);
        }
    }
}

// Verification inputs straight from the signature: the item's LZ polynomials of z are read from their packed form
// (gamma1 - z in ZBITS = 18 / 20 bits per coefficient, encoder.v:96-133 / decoder.v:89-143) - no unpack pass, no 4 KiB per
// polynomial of int32 in HBM - and the challenge polynomial c (int32, time domain, from verify_prep) is the last input.
// Every coefficient is a bit field inside two aligned 32-bit words (one funnel shift); the ||z|| < gamma1 - beta check of
// the unpack pass happens here.  Returns (warp-uniform) whether the item violates the bound.
template <int LZ, int ZBITS, int BETA>
__device__ __forceinline__ bool item_inputs_zpacked(uint32_t (&yh)[LZ + 1][8], const uint8_t* __restrict__ z_item,
                                                    const int32_t* __restrict__ c_item, uint32_t* __restrict__ scr, int lane) {
    constexpr int ROWB = 32 * ZBITS;              // packed bytes per polynomial: 576 / 640
    constexpr int32_t G1 = 1 << (ZBITS - 1);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < LZ; j++) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(z_item + j * ROWB);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int bit = (lane + 32 * r) * ZBITS, byte = bit >> 3, aw = byte >> 2, sh = (byte & 3) * 8 + (bit & 7);
            const uint32_t w0 = __ldg(p + aw), w1 = (4 * aw + 4 < ROWB) ? __ldg(p + aw + 1) : 0u;
            const int32_t zc = G1 - (int32_t)(__funnelshift_r(w0, w1, sh) & ((1u << ZBITS) - 1));
            bad |= (zc >= G1 - BETA) || (zc <= -(G1 - BETA));
            yh[j][r] = (uint32_t)zc;              // signed representative: lifted inside the first butterfly
        }
    }
    {
        const int32_t* p = c_item + lane;
#pragma unroll
        for (int r = 0; r < 8; r++) yh[LZ][r] = (uint32_t)p[32 * r];
    }
    FwdTw ftw;
    {
        const TwTable* tab = &TW_FWD;
        asm volatile("" : "+l"(tab));
        load_fwd_tw(ftw, tab, lane);
    }
#pragma unroll
    for (int j = 0; j < LZ + 1; j++) {
        ntt_fwd_warp(yh[j], scr, ftw, lane);
        __syncwarp();
    }
    return __any_sync(0xffffffffu, bad);
}

// ---- per-item core, executed by one warp ----
// v_item: l polys (global), w_item: k polys (global), a_sm: k*l polys in shared memory (stride A_STRIDE)
// EXTRA = true (verification with per-item public keys): v_item holds L+1 polynomials, and the last one is
// multiplied by a per-item column extra_item[i] read from global memory (-NTT(t1_i * 2^13)) instead of a
// shared-memory matrix column.
// W1 = true (signing): besides w the core also emits w1 = HighBits(w), bit-packed as encoder.v:96-133 (6 bits for
// gamma2 = (Q-1)/88, i.e. K = 4, else 4 bits), into w1_item - the input of the challenge hash - so that no
// separate pass has to read w again.
// SPLIT = true (per-item kernel of level 5, two warps per item): both warps of the CTA call this together; warp
// `part` transforms every second input and publishes it through yh_sm (layout C), and computes rows
// [part*K/2, (part+1)*K/2) of the result - the single-warp latency of the per-item core is halved.
template <int K, int LA, bool NTT_IN, bool INTT_OUT, bool EXTRA = false, bool W1 = false, bool SPLIT = false, bool PRESCALED = true>
__device__ __forceinline__ void item_core(int32_t* __restrict__ w_item, const int32_t* __restrict__ v_item,
                                          const uint32_t* __restrict__ a_sm, uint32_t* __restrict__ scr, int lane,
                                          const int32_t* __restrict__ extra_item = nullptr,
                                          uint8_t* __restrict__ w1_item = nullptr, uint32_t* __restrict__ yh_sm = nullptr,
                                          int part = 0, const uint32_t* __restrict__ hint_item = nullptr) {
    constexpr int L = LA + (EXTRA ? 1 : 0);   // number of input polynomials
    uint32_t yh[L][8];  // NTT-domain inputs in layout C
    if constexpr (NTT_IN && SPLIT) {
        FwdTw ftw;
        {
            const TwTable* tab = &TW_FWD;
            asm volatile("" : "+l"(tab));
            load_fwd_tw(ftw, tab, lane);
        }
#pragma unroll 1
        for (int j = part; j < L; j += 2) {
            uint32_t x[8];
            const int32_t* p = v_item + j * N + lane;
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = (uint32_t)p[32 * r];
            ntt_fwd_warp(x, scr, ftw, lane);
            __syncwarp();
            uint4* o = reinterpret_cast<uint4*>(yh_sm + j * N) + lane;
            o[0] = make_uint4(x[0], x[1], x[2], x[3]);
            o[32] = make_uint4(x[4], x[5], x[6], x[7]);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < L; j++) {
            const uint4* q = reinterpret_cast<const uint4*>(yh_sm + j * N) + lane;
            const uint4 lo = q[0], hi = q[32];
            yh[j][0] = lo.x; yh[j][1] = lo.y; yh[j][2] = lo.z; yh[j][3] = lo.w;
            yh[j][4] = hi.x; yh[j][5] = hi.y; yh[j][6] = hi.z; yh[j][7] = hi.w;
        }
    } else {
        item_inputs<L, NTT_IN>(yh, v_item, scr, lane);
    }
    const int i_begin = SPLIT ? part * (K / 2) : 0, i_end = SPLIT ? (part + 1) * (K / 2) : K;
    item_rows<K, LA, INTT_OUT, EXTRA, W1, PRESCALED>(w_item, yh, a_sm, scr, lane, extra_item, w1_item, i_begin, i_end, 0, hint_item);
}

#endif  // __CUDACC__
}  // namespace dil
