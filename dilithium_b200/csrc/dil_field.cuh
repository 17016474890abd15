// dil_field.cuh — Z_Q arithmetic for the B200 engine, Q = 8380417 (params.h:33).
//
// Reduction strategy (replaces the reference's signed 64-bit `%`, ref_ntt.cpp:41-43,
// and the RTL's shift-add Barrett, Barrett_8380417.v:189-219):
//   * multiplication by a KNOWN constant w (every NTT twiddle) uses Shoup's method:
//       w' = floor(w * 2^32 / Q)            (precomputed)
//       t  = b*w - umulhi(b, w') * Q        (3 integer multiplies, no division)
//     valid for ANY 32-bit b, result in [0, 2Q).
//   * additions/subtractions are lazy: values stay unreduced unsigned 32-bit with
//     statically tracked bounds (forward < 18Q, inverse < 512Q < 2^32) and are brought
//     to canonical [0, Q) once, at the end.
// All twiddles are plain-domain zeta^brv8(k) as in consts.cpp:64-97 (no Montgomery
// factor), generated here at compile time from zeta = 1753.
#pragma once
#include <cstdint>

namespace dil {

constexpr uint32_t Q = 8380417u;
constexpr int N = 256;
constexpr uint32_t INV256 = 8347681u;  // 256^-1 mod Q (ref_ntt.cpp:64)

constexpr uint32_t cmulq(uint64_t a, uint64_t b) { return (uint32_t)((a * b) % Q); }
constexpr uint32_t cpowq(uint32_t base, unsigned e) {
    uint64_t r = 1, b = base;
    while (e) {
        if (e & 1) r = (r * b) % Q;
        b = (b * b) % Q;
        e >>= 1;
    }
    return (uint32_t)r;
}
constexpr unsigned cbrv8(unsigned k) {
    unsigned r = 0;
    for (int b = 0; b < 8; b++) r |= ((k >> b) & 1u) << (7 - b);
    return r;
}
// canonical forward twiddle for table index k (k = 1..255)
constexpr uint32_t zeta_fwd(unsigned k) { return cpowq(1753u, cbrv8(k)); }
// canonical inverse twiddle: -zetas[k]  (ref_ntt.cpp:72)
constexpr uint32_t zeta_inv(unsigned k) { return (Q - zeta_fwd(k)) % Q; }
constexpr uint32_t shoup(uint32_t w) { return (uint32_t)(((uint64_t)w << 32) / Q); }

struct Twiddle {  // (w, floor(w*2^32/Q))
    uint32_t w, wp;
};

#ifdef __CUDACC__
// t = b*w mod Q, lazily reduced to [0, 2Q).  b: any uint32.
__device__ __forceinline__ uint32_t mul_shoup(uint32_t b, uint32_t w, uint32_t wp) {
    uint32_t qh = __umulhi(b, wp);
    return b * w - qh * Q;
}
// x < 2^32 arbitrary -> canonical [0,Q).  Uses 2^23 = 2^13 - 1 (mod Q):
// x = h*2^23 + lo -> h*(2^13-1) + lo < 2Q for h <= 511, then one conditional subtract.
__device__ __forceinline__ uint32_t canon_small(uint32_t x) {
    uint32_t h = x >> 23;
    uint32_t r = x - h * Q;              // == h*8191 + (x & 0x7FFFFF), < 2Q
    return min(r, r - Q);                // unsigned min: r-Q wraps high when r < Q
}
// r in [0, 2Q) -> [0, Q)
__device__ __forceinline__ uint32_t csub(uint32_t r) { return min(r, r - Q); }
// signed representative in (-Q, Q) -> [0, Q)
__device__ __forceinline__ uint32_t canon_signed(int32_t a) { return (uint32_t)(a + ((a >> 31) & (int32_t)Q)); }
// Barrett for wide accumulators: x < 2^49 (up to 8 products of canonical residues) -> [0,Q).
//   M = floor(2^54/Q); qh = floor(floor(x/2^22) * M / 2^32) underestimates floor(x/Q) by at
//   most 1 for x < 2^49 (error < x/2^54 + 2^22/Q + 1 < 1.54), so one conditional subtract.
// Same construction as the RTL reducer (Barrett_8380417.v:189-219: m = floor(2^46/Q),
// (x>>22)*m>>24, one correction), widened from 46-bit to 49-bit inputs.
constexpr uint32_t BARRETT_M54 = (uint32_t)((1ull << 54) / Q);
__device__ __forceinline__ uint32_t reduce49(uint64_t x) {
    uint32_t xs = (uint32_t)(x >> 22);
    uint32_t qh = __umulhi(xs, BARRETT_M54);
    uint32_t r = (uint32_t)x - qh * Q;
    return csub(r);
}
// exact a*b mod Q for canonical a, b
__device__ __forceinline__ uint32_t mul_full(uint32_t a, uint32_t b) { return reduce49((uint64_t)a * b); }
#endif

}  // namespace dil
