// ntt_core.cuh — one-polynomial-per-warp 256-point negacyclic NTT / INTT over Z_8380417.
//
// Same maps as the reference's ntt()/invntt() (ref_ntt.cpp:28-47, :59-87) and its radix-2x2
// restatement (ref_ntt2x2.cpp:37-145), re-scheduled for a 32-lane warp:
//
//   * each lane owns 8 coefficients; three register-resident "phases" of 3+3+2 butterfly
//     layers (forward) / 2+3+3 (inverse) cover the 8 layers, so a polynomial is exchanged
//     between lanes only twice, through a padded 288-word shared-memory scratch whose
//     addressing (word 36*(i>>5) + (i&31)) makes every access bank-conflict-free with
//     compile-time offsets.  (The FPGA does the same job with its 4x4 FIFO transposer,
//     ntt_fifo.v:51-62 / address_resolver.v:38-52.)
//   * register layouts, i = coefficient index:
//       layout A : x[r] <-> i = 32*r + lane                    (r = i bits 7..5)
//       layout B : x[r] <-> i = 32*(lane>>2) + 4*r + (lane&3)  (r = i bits 4..2)
//       layout C : x[r] <-> i = 128*(r>>2) + 4*lane + (r&3)    (r = i bit 7, bits 1..0)
//     Layout C is the "four-wide packed" layout: a lane's x[0..3] / x[4..7] are 16-byte
//     vectors and a warp covers 512 contiguous bytes per vector access.
//     forward: A -> B -> C,   inverse: C -> B -> A.
//   * twiddles: layers whose twiddle index depends only on register bits use compile-time
//     immediates; the 13 lane-dependent (w, w') pairs live in registers for the lifetime of
//     a persistent warp (twiddle_resolver.v's index walk becomes a function of the lane id).
//   * all butterflies use Shoup multiplication and lazy unsigned accumulation
//     (dil_field.cuh); outputs are canonical [0,Q).
#pragma once
#include "dil_field.cuh"

namespace dil {

struct TwTable {
    Twiddle t[256];
};
constexpr TwTable make_fwd_table() {
    TwTable r{};
    for (unsigned k = 1; k < 256; k++) r.t[k] = Twiddle{zeta_fwd(k), shoup(zeta_fwd(k))};
    return r;
}
constexpr TwTable make_inv_table() {
    TwTable r{};
    for (unsigned k = 1; k < 256; k++) r.t[k] = Twiddle{zeta_inv(k), shoup(zeta_inv(k))};
    return r;
}

// compile-time twiddle constants usable as immediates in device code
template <unsigned K>
struct ZF {
    static constexpr uint32_t w = zeta_fwd(K), wp = shoup(zeta_fwd(K));
};
template <unsigned K>
struct ZI {
    static constexpr uint32_t w = zeta_inv(K), wp = shoup(zeta_inv(K));
};
template <unsigned K>
struct ZIF {   // zeta_inv(K) * 256^-1: inverse twiddles that also carry the scaling (see ntt_inv_warp)
    static constexpr uint32_t w = cmulq(zeta_inv(K), INV256), wp = shoup(cmulq(zeta_inv(K), INV256));
};
struct ZLast {  // 256^-1 and zeta_inv(1)*256^-1 for the merged last inverse layer
    static constexpr uint32_t f = INV256, fp = shoup(INV256);
    static constexpr uint32_t wf = cmulq(zeta_inv(1), INV256), wfp = shoup(cmulq(zeta_inv(1), INV256));
};

constexpr int SCRATCH_WORDS = 288;  // 8 rows of 32 words, row stride 36

struct FwdTw {  // lane-dependent forward twiddles (layers 3..7)
    Twiddle l3, l4[2], l5[4], l6[2], l7[4];
};
struct InvTw {  // lane-dependent inverse twiddles (spans 1,2,4,8,16)
    Twiddle s1[4], s2[2], s4[4], s8[2], s16;
};

#ifdef __CUDACC__

// per-translation-unit copies of the 2 KiB twiddle tables (no relocatable device code needed)
static __device__ const TwTable TW_FWD = make_fwd_table();
static __device__ const TwTable TW_INV = make_inv_table();

__device__ __forceinline__ Twiddle ld_tw(const TwTable* __restrict__ tab, int k) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(&tab->t[k]));
    return Twiddle{v.x, v.y};
}

__device__ __forceinline__ void load_fwd_tw(FwdTw& tw, const TwTable* __restrict__ tab, int lane) {
    const int g = lane >> 2;
    tw.l3 = ld_tw(tab, 8 + g);
#pragma unroll
    for (int h = 0; h < 2; h++) tw.l4[h] = ld_tw(tab, 16 + 2 * g + h);
#pragma unroll
    for (int m = 0; m < 4; m++) tw.l5[m] = ld_tw(tab, 32 + 4 * g + m);
#pragma unroll
    for (int h = 0; h < 2; h++) tw.l6[h] = ld_tw(tab, 64 + 32 * h + lane);
    tw.l7[0] = ld_tw(tab, 128 + 2 * lane);
    tw.l7[1] = ld_tw(tab, 129 + 2 * lane);
    tw.l7[2] = ld_tw(tab, 192 + 2 * lane);
    tw.l7[3] = ld_tw(tab, 193 + 2 * lane);
}

__device__ __forceinline__ void load_inv_tw(InvTw& tw, const TwTable* __restrict__ tab, int lane) {
    const int g = lane >> 2;
    // span 1: block m = i>>1, index 255 - m
    tw.s1[0] = ld_tw(tab, 255 - 2 * lane);
    tw.s1[1] = ld_tw(tab, 254 - 2 * lane);
    tw.s1[2] = ld_tw(tab, 191 - 2 * lane);
    tw.s1[3] = ld_tw(tab, 190 - 2 * lane);
    // span 2: m = i>>2, index 127 - m
    tw.s2[0] = ld_tw(tab, 127 - lane);
    tw.s2[1] = ld_tw(tab, 95 - lane);
    // span 4: m = i>>3 = 4g + (r>>1), index 63 - m
#pragma unroll
    for (int m = 0; m < 4; m++) tw.s4[m] = ld_tw(tab, 63 - 4 * g - m);
    // span 8: m = i>>4 = 2g + (r>>2), index 31 - m
#pragma unroll
    for (int h = 0; h < 2; h++) tw.s8[h] = ld_tw(tab, 31 - 2 * g - h);
    // span 16: m = g, index 15 - m
    tw.s16 = ld_tw(tab, 15 - g);
}

// ---- butterflies ----
// Cooley-Tukey: (a, b) -> (a + w*b, a - w*b), lazy: outputs grow by < 2Q per layer.
__device__ __forceinline__ void ct(uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp) {
    uint32_t t = mul_shoup(b, w, wp);
    b = a + 2 * Q - t;
    a = a + t;
}
__device__ __forceinline__ void ct(uint32_t& a, uint32_t& b, const Twiddle& z) { ct(a, b, z.w, z.wp); }
template <unsigned K>
__device__ __forceinline__ void ct_k(uint32_t& a, uint32_t& b) {
    ct(a, b, ZF<K>::w, ZF<K>::wp);
}
// first layer: inputs are signed representatives in (-Q, Q); lift by +Q inside the butterfly
template <unsigned K>
__device__ __forceinline__ void ct_first(uint32_t& a, uint32_t& b) {
    uint32_t t = mul_shoup(b + Q, ZF<K>::w, ZF<K>::wp);
    b = a + 3 * Q - t;
    a = a + Q + t;
}
// Gentleman-Sande: (a, b) -> (a + b, w*(a - b)); inputs < C (a multiple of Q), a-output < 2C.
template <uint32_t C>
__device__ __forceinline__ void gs(uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp) {
    uint32_t d = a + C - b;
    a = a + b;
    b = mul_shoup(d, w, wp);
}
template <uint32_t C>
__device__ __forceinline__ void gs(uint32_t& a, uint32_t& b, const Twiddle& z) { gs<C>(a, b, z.w, z.wp); }
template <uint32_t C, unsigned K>
__device__ __forceinline__ void gs_k(uint32_t& a, uint32_t& b) {
    gs<C>(a, b, ZI<K>::w, ZI<K>::wp);
}
// first inverse layer: signed inputs in (-Q, Q)
__device__ __forceinline__ void gs_first(uint32_t& a, uint32_t& b, const Twiddle& z) {
    uint32_t d = a + 2 * Q - b;
    a = a + b + 2 * Q;
    b = mul_shoup(d, z.w, z.wp);
}

// scratch addressing (32-bit word offsets inside a warp's 288-word scratch)
__device__ __forceinline__ int scr_off_a(int lane) { return lane; }                               // + 36*r
__device__ __forceinline__ int scr_off_b(int lane) { return 36 * (lane >> 2) + (lane & 3); }       // + 4*r
__device__ __forceinline__ int scr_off_c(int lane) { return 36 * (lane >> 3) + 4 * (lane & 7); }   // + 144*(r>>2) + (r&3)

// Forward NTT.  in: x in layout A, signed/unsigned representatives in (-Q, Q).
//               out: x in layout C, canonical [0, Q), bit-reversed NTT order as ref ntt().
__device__ __forceinline__ void ntt_fwd_warp(uint32_t (&x)[8], uint32_t* __restrict__ scr, const FwdTw& tw, int lane) {
    // phase A: layers 0..2 (spans 128, 64, 32); register bits = i[7:5]; twiddles 1..7
    ct_first<1>(x[0], x[4]); ct_first<1>(x[1], x[5]); ct_first<1>(x[2], x[6]); ct_first<1>(x[3], x[7]);
    ct_k<2>(x[0], x[2]); ct_k<2>(x[1], x[3]); ct_k<3>(x[4], x[6]); ct_k<3>(x[5], x[7]);
    ct_k<4>(x[0], x[1]); ct_k<5>(x[2], x[3]); ct_k<6>(x[4], x[5]); ct_k<7>(x[6], x[7]);
    {
        uint32_t* p = scr + scr_off_a(lane);
#pragma unroll
        for (int r = 0; r < 8; r++) p[36 * r] = x[r];
    }
    __syncwarp();
    // phase B: layers 3..5 (spans 16, 8, 4); register bits = i[4:2]
    uint32_t* pb = scr + scr_off_b(lane);
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = pb[4 * r];
    ct(x[0], x[4], tw.l3); ct(x[1], x[5], tw.l3); ct(x[2], x[6], tw.l3); ct(x[3], x[7], tw.l3);
    ct(x[0], x[2], tw.l4[0]); ct(x[1], x[3], tw.l4[0]); ct(x[4], x[6], tw.l4[1]); ct(x[5], x[7], tw.l4[1]);
    ct(x[0], x[1], tw.l5[0]); ct(x[2], x[3], tw.l5[1]); ct(x[4], x[5], tw.l5[2]); ct(x[6], x[7], tw.l5[3]);
#pragma unroll
    for (int r = 0; r < 8; r++) pb[4 * r] = x[r];
    __syncwarp();
    // phase C: layers 6, 7 (spans 2, 1); register bits = i[7], i[1:0]
    {
        const uint4* pc = reinterpret_cast<const uint4*>(scr + scr_off_c(lane));
        uint4 lo = pc[0], hi = pc[36];  // +144 words
        x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
        x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
    }
    ct(x[0], x[2], tw.l6[0]); ct(x[1], x[3], tw.l6[0]); ct(x[4], x[6], tw.l6[1]); ct(x[5], x[7], tw.l6[1]);
    ct(x[0], x[1], tw.l7[0]); ct(x[2], x[3], tw.l7[1]); ct(x[4], x[5], tw.l7[2]); ct(x[6], x[7], tw.l7[3]);
    // values < 18Q -> canonical
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = canon_small(x[r]);
}

// Inverse NTT incl. 256^-1.  in: x in layout C, representatives in (-Q, Q) (bit-reversed order).
//                            out: x in layout A, canonical [0, Q), natural order.
// PRESCALED = true: the caller has already folded 256^-1 into the NTT-domain input (the fused kernels
// scale their shared-memory matrix / key polynomials once per CTA), so the sum path of the last layer
// needs only a canonicalisation instead of a Shoup multiplication (12 of ~180 multiply slots per transform).
// The caller must __syncwarp() before the scratch is reused by another transform.
template <bool PRESCALED = false>
__device__ __forceinline__ void ntt_inv_warp(uint32_t (&x)[8], uint32_t* __restrict__ scr, const InvTw& tw, int lane) {
    // phase C: spans 1, 2
    gs_first(x[0], x[1], tw.s1[0]); gs_first(x[2], x[3], tw.s1[1]); gs_first(x[4], x[5], tw.s1[2]); gs_first(x[6], x[7], tw.s1[3]);
    gs<4 * Q>(x[0], x[2], tw.s2[0]); gs<4 * Q>(x[1], x[3], tw.s2[0]); gs<4 * Q>(x[4], x[6], tw.s2[1]); gs<4 * Q>(x[5], x[7], tw.s2[1]);
    {
        uint4* pc = reinterpret_cast<uint4*>(scr + scr_off_c(lane));
        pc[0] = make_uint4(x[0], x[1], x[2], x[3]);
        pc[36] = make_uint4(x[4], x[5], x[6], x[7]);
    }
    __syncwarp();
    // phase B: spans 4, 8, 16
    uint32_t* pb = scr + scr_off_b(lane);
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = pb[4 * r];
    gs<8 * Q>(x[0], x[1], tw.s4[0]); gs<8 * Q>(x[2], x[3], tw.s4[1]); gs<8 * Q>(x[4], x[5], tw.s4[2]); gs<8 * Q>(x[6], x[7], tw.s4[3]);
    gs<16 * Q>(x[0], x[2], tw.s8[0]); gs<16 * Q>(x[1], x[3], tw.s8[0]); gs<16 * Q>(x[4], x[6], tw.s8[1]); gs<16 * Q>(x[5], x[7], tw.s8[1]);
    gs<32 * Q>(x[0], x[4], tw.s16); gs<32 * Q>(x[1], x[5], tw.s16); gs<32 * Q>(x[2], x[6], tw.s16); gs<32 * Q>(x[3], x[7], tw.s16);
#pragma unroll
    for (int r = 0; r < 8; r++) pb[4 * r] = x[r];
    __syncwarp();
    // phase A: spans 32, 64, 128
    {
        const uint32_t* p = scr + scr_off_a(lane);
#pragma unroll
        for (int r = 0; r < 8; r++) x[r] = p[36 * r];
    }
    if constexpr (PRESCALED) {
        gs_k<64 * Q, 7>(x[0], x[1]); gs_k<64 * Q, 6>(x[2], x[3]); gs_k<64 * Q, 5>(x[4], x[5]); gs_k<64 * Q, 4>(x[6], x[7]);
        gs_k<128 * Q, 3>(x[0], x[2]); gs_k<128 * Q, 3>(x[1], x[3]); gs_k<128 * Q, 2>(x[4], x[6]); gs_k<128 * Q, 2>(x[5], x[7]);
        // last layer (span 128, twiddle index 1): the input already carries 256^-1
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t d = x[r] + 256 * Q - x[r + 4];
            uint32_t s = x[r] + x[r + 4];
            x[r] = canon_small(s);
            x[r + 4] = csub(mul_shoup(d, ZI<1>::w, ZI<1>::wp));
        }
    } else {
        // The 256^-1 scaling rides on the twiddles: a difference output is multiplied by its twiddle anyway, so the
        // twiddle of the FIRST difference a register takes in this phase is zeta * 256^-1 (compile-time constants, free),
        // and everything downstream of it is scaled.  Only x[0], which is a sum in all three layers, needs one explicit
        // multiplication - 1 instead of 4 (both inputs of every butterfly below have the same scaled / unscaled status).
        auto gsf = [](uint32_t C, uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp) {
            uint32_t d = a + C - b;
            a = a + b;
            b = mul_shoup(d, w, wp);
        };
        gsf(64 * Q, x[0], x[1], ZIF<7>::w, ZIF<7>::wp); gsf(64 * Q, x[2], x[3], ZIF<6>::w, ZIF<6>::wp);
        gsf(64 * Q, x[4], x[5], ZIF<5>::w, ZIF<5>::wp); gsf(64 * Q, x[6], x[7], ZIF<4>::w, ZIF<4>::wp);
        // unscaled pairs (0,2), (4,6): scale on the difference; pairs (1,3), (5,7) are scaled already
        gsf(128 * Q, x[0], x[2], ZIF<3>::w, ZIF<3>::wp); gs_k<4 * Q, 3>(x[1], x[3]);
        gsf(128 * Q, x[4], x[6], ZIF<2>::w, ZIF<2>::wp); gs_k<4 * Q, 2>(x[5], x[7]);
        {   // pair (0,4): both unscaled
            uint32_t d = x[0] + 256 * Q - x[4];
            uint32_t s = x[0] + x[4];
            x[0] = csub(mul_shoup(s, ZLast::f, ZLast::fp));
            x[4] = csub(mul_shoup(d, ZLast::wf, ZLast::wfp));
        }
#pragma unroll
        for (int r = 1; r < 4; r++) {   // pairs (r, r+4): scaled, values < 8Q
            uint32_t d = x[r] + 8 * Q - x[r + 4];
            uint32_t s = x[r] + x[r + 4];
            x[r] = canon_small(s);
            x[r + 4] = csub(mul_shoup(d, ZI<1>::w, ZI<1>::wp));
        }
    }
}

#endif  // __CUDACC__
}  // namespace dil
