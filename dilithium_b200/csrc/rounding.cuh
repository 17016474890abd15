// rounding.cuh - Decompose / HighBits of the scheme (shared by the sign core's fused w1 packing and the
// sign / verify kernels).
#pragma once
#include <cstdint>

#include "dil_field.cuh"
#include "dil_params.h"

namespace dil {
#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------
// Decompose (decomp_map1.v / coeff_decomposer.v:80-88): a = a1*2*gamma2 + a0,
// -gamma2 < a0 <= gamma2, with the wrap-around row mapped to a1 = 0.
// ---------------------------------------------------------------------------------------
template <int32_t GAMMA2>
__device__ __forceinline__ void decompose(int32_t a, int32_t& a1, int32_t& a0) {
    int32_t t = (a + 127) >> 7;
    if constexpr (GAMMA2 == (Q_I - 1) / 32) {
        t = (t * 1025 + (1 << 21)) >> 22;
        t &= 15;
    } else {
        t = (t * 11275 + (1 << 23)) >> 24;
        t ^= ((43 - t) >> 31) & t;
    }
    a1 = t;
    a0 = a - t * 2 * GAMMA2;
    a0 -= (((Q_I - 1) / 2 - a0) >> 31) & Q_I;
}

// centred representative in (-Q/2, Q/2] of a canonical coefficient
__device__ __forceinline__ int32_t centre(uint32_t a) { return (int32_t)a - (int32_t)((a > (Q - 1) / 2) ? Q : 0); }

// HighBits only (w1), for a canonical coefficient
template <int32_t GAMMA2>
__device__ __forceinline__ uint32_t highbits(uint32_t a) {
    int32_t t = ((int32_t)a + 127) >> 7;
    if constexpr (GAMMA2 == (Q_I - 1) / 32) {
        t = (t * 1025 + (1 << 21)) >> 22;
        t &= 15;
    } else {
        t = (t * 11275 + (1 << 23)) >> 24;
        t ^= ((43 - t) >> 31) & t;
    }
    return (uint32_t)t;
}

#endif  // __CUDACC__
}  // namespace dil
