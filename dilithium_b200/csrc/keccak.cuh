// keccak.cuh — register-resident Keccak-f[1600] and the SHAKE-128 ExpandA sampler.
//
// The reference realises Keccak as a VHDL core doing one round per clock
// (rtl_src/keccak_round.vhd, keccak_datapath.vhd:190-203, round constants
// keccak_cons.vhd:25-33).  Here one THREAD owns one 1600-bit state in 25 64-bit registers
// and runs the 24 rounds fully unrolled (LOP3 for theta/chi, funnel shifts for rho).
//
// ExpandA restates sampler_a_ext.v:107-133 (SHAKE-128 over rho || byte(j) || byte(i),
// padding 0x1F .. 0x80 as keccak_bytepad.vhd:37-44, rate 168 B) and rejection_a.v:67-92
// (3-byte little-endian chunks, low 23 bits, accept < Q, first 256 in stream order).
#pragma once
#include <cstdint>

#include "dil_field.cuh"

namespace dil {

#ifdef __CUDACC__

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) {
    // n is a compile-time constant after unrolling
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n < 32) {
        uint32_t nhi = __funnelshift_l(lo, hi, n);
        uint32_t nlo = __funnelshift_l(hi, lo, n);
        return ((uint64_t)nhi << 32) | nlo;
    }
    uint32_t nhi = __funnelshift_l(hi, lo, n - 32);
    uint32_t nlo = __funnelshift_l(lo, hi, n - 32);
    return ((uint64_t)nhi << 32) | nlo;
}

// round constants (keccak_cons.vhd:25-33); uniform index -> constant-cache broadcast
static __constant__ uint64_t KECCAK_RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

// 24 rounds; the round body is fully unrolled (all lane indices static), the round loop is
// kept rolled so one permutation is ~190 instructions of code instead of ~4500.
__device__ __forceinline__ void keccak_f1600(uint64_t (&A)[25]) {
    // rho rotation of lane x+5y
    constexpr int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t C[5], R1[5], B[25];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) R1[x] = rotl64(C[x], 1);
        // theta folded into the rho/pi input: A ^ C[x-1] ^ rot(C[x+1],1) is one three-input LOP3 per half,
        // so the five D words are never materialised (10 LOP3 fewer per round)
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++)
                B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(A[x + 5 * y] ^ C[(x + 4) % 5] ^ R1[(x + 1) % 5], RHO[x + 5 * y]);
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= KECCAK_RC[r];
    }
}

// SHAKE-128 state after absorbing rho[32] || j || i and padding (34 bytes < rate 168)
__device__ __forceinline__ void expand_a_absorb(uint64_t (&A)[25], const uint8_t* __restrict__ rho, int i, int j) {
#pragma unroll
    for (int t = 0; t < 25; t++) A[t] = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        uint64_t v = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) v |= (uint64_t)rho[8 * t + b] << (8 * b);
        A[t] = v;
    }
    A[4] = (uint64_t)(uint8_t)j | ((uint64_t)(uint8_t)i << 8) | (0x1FULL << 16);
    A[20] = 0x80ULL << 56;  // last byte of the 168-byte rate
}

// Generate the 256 coefficients of A_hat[i][j] and hand each accepted value, in stream
// order, to emit(index, value).
template <class Emit>
__device__ __forceinline__ void expand_a_poly(const uint8_t* __restrict__ rho, int i, int j, Emit emit) {
    uint64_t A[25];
    expand_a_absorb(A, rho, i, j);
    int cnt = 0;
    while (cnt < N) {
        keccak_f1600(A);
        // 168 bytes = 21 lanes = 56 three-byte candidates
#pragma unroll
        for (int c = 0; c < 56; c++) {
            const int byte = 3 * c, ln = byte >> 3, sh = (byte & 7) * 8;
            uint64_t v = A[ln] >> sh;
            if (sh > 40) v |= A[ln + 1] << (64 - sh);
            uint32_t t = (uint32_t)v & 0x7FFFFFu;
            if (t < Q && cnt < N) {
                emit(cnt, t);
                cnt++;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// SHAKE-256 helpers (rate 136 B = 17 lanes) for the sign/verify pipeline.  All message
// schedules used by the scheme are expressed as a "lane source" lane_at(idx) -> uint64 so the
// state registers are only ever indexed with compile-time constants.
// ---------------------------------------------------------------------------------------
constexpr int SHAKE256_LANES = 17;

// Absorb `total_bytes` bytes delivered as 64-bit little-endian lanes by lane_at(idx) (bytes
// beyond total_bytes inside the last lane must read as zero), pad (0x1F .. 0x80) and permute
// once more, leaving the first squeezable block in A.
template <class LaneAt>
__device__ __forceinline__ void shake256_absorb_lanes(uint64_t (&A)[25], size_t total_bytes, LaneAt lane_at) {
#pragma unroll
    for (int t = 0; t < 25; t++) A[t] = 0;
    const size_t nfull = total_bytes / 136;
    for (size_t b = 0; b < nfull; b++) {
#pragma unroll
        for (int i = 0; i < SHAKE256_LANES; i++) A[i] ^= lane_at(b * SHAKE256_LANES + i);
        keccak_f1600(A);
    }
    const unsigned rem = (unsigned)(total_bytes - nfull * 136);
    const unsigned rem_lanes = (rem + 7) >> 3;
#pragma unroll
    for (int i = 0; i < SHAKE256_LANES; i++) {
        if ((unsigned)i < rem_lanes) A[i] ^= lane_at(nfull * SHAKE256_LANES + i);
        if ((unsigned)i == (rem >> 3)) A[i] ^= 0x1FULL << (8 * (rem & 7));
    }
    A[16] ^= 0x80ULL << 56;
    keccak_f1600(A);
}

// little-endian 64-bit load from an arbitrarily aligned byte string, zero beyond `len`; one 8-byte load when the
// lane lies inside the string and happens to be 8-byte aligned (packed keys, fixed-length messages), else byte loads
__device__ __forceinline__ uint64_t load_lane_bytes(const uint8_t* __restrict__ p, size_t off, size_t len) {
    uint64_t v = 0;
    if (off + 8 <= len) {
        if ((reinterpret_cast<uintptr_t>(p + off) & 7u) == 0) return *reinterpret_cast<const uint64_t*>(p + off);
#pragma unroll
        for (int b = 0; b < 8; b++) v |= (uint64_t)p[off + b] << (8 * b);
    } else {
        for (int b = 0; b < 8 && off + b < len; b++) v |= (uint64_t)p[off + b] << (8 * b);
    }
    return v;
}

#endif  // __CUDACC__
}  // namespace dil
