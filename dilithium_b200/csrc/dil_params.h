// dil_params.h — Dilithium round-3.1 parameter sets as used by the reference RTL
// (rtl_src/combined_top.v:520-551, norm_check.v:43-51, gen_c.v:107-124, makehint.v:48-55,
// rejection_y.v:45-54, rejection_s.v:45-51).  Shared by host and device code.
#pragma once
#include <cstdint>

namespace dil {

struct LevelParams {
    int level, k, l, eta, tau, gamma1_bits, omega, beta;
    int32_t gamma1, gamma2;
    int z_bytes;    // packed z bytes per polynomial (18 or 20 bits/coeff)
    int w1_bytes;   // packed w1 bytes per polynomial (6 or 4 bits/coeff)
    int s_bytes;    // packed s1/s2 bytes per polynomial (3 or 4 bits/coeff)
};

constexpr int32_t Q_I = 8380417;
constexpr int D_BITS = 13;

#ifdef __CUDACC__
__host__ __device__
#endif
    constexpr LevelParams level_params(int level) {
    return level == 2   ? LevelParams{2, 4, 4, 2, 39, 17, 80, 78, 1 << 17, (Q_I - 1) / 88, 576, 192, 96}
           : level == 3 ? LevelParams{3, 6, 5, 4, 49, 19, 55, 196, 1 << 19, (Q_I - 1) / 32, 640, 128, 128}
                        : LevelParams{5, 8, 7, 2, 60, 19, 75, 120, 1 << 19, (Q_I - 1) / 32, 640, 128, 96};
}

}  // namespace dil
