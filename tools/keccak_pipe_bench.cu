// keccak_pipe_bench.cu — micro-benchmark: Keccak-f[1600] with some of the 64-bit rotations moved from the
// ALU pipe (SHF funnel shifts) to the FMA pipe (IMAD.WIDE / IMAD / IMAD.HI with a power-of-two multiplier
// read from the constant bank so ptxas cannot strength-reduce it back into shifts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o keccak_pipe_bench keccak_pipe_bench.cu
//   ./keccak_pipe_bench            (prints G permutations/s for each variant, checks they agree)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static __constant__ uint32_t POW2[32] = {1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,
                                         1u << 8,  1u << 9,  1u << 10, 1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15,
                                         1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23,
                                         1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};
static __constant__ uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

__device__ __forceinline__ uint64_t rotl64_shf(uint64_t x, int n) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n < 32) return ((uint64_t)__funnelshift_l(lo, hi, n) << 32) | __funnelshift_l(hi, lo, n);
    return ((uint64_t)__funnelshift_l(hi, lo, n - 32) << 32) | __funnelshift_l(lo, hi, n - 32);
}

// Rotation variants that move work from the ALU pipe to the FMA pipe (m = 2^r from the constant bank):
//  B: P = lo*m, Q = hi*m (two IMAD.WIDE), out_lo = P.lo + Q.hi, out_hi = Q.lo + P.hi (two IMADs by an opaque 1)
//     -> 0 ALU, 4 FMA instructions
//  D: S = {hi >> (32-r), hi*m} (one SHF.R + one IMAD), U = lo*m + S (one IMAD.WIDE with 64-bit addend)
//     -> 1 ALU, 2 FMA instructions
__device__ __forceinline__ uint64_t rotl64_v(const int MODE, uint64_t x, int n) {
    if (MODE == 0) return rotl64_shf(x, n);
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n > 32) { uint32_t t = lo; lo = hi; hi = t; n -= 32; }
    const uint32_t m = POW2[n];
    if (MODE == 1) {
        const uint32_t one = POW2[0];
        uint32_t plo, phi, qlo, qhi, ohi, olo;
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(plo), "=r"(phi) : "r"(lo), "r"(m));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(qlo), "=r"(qhi) : "r"(hi), "r"(m));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(olo) : "r"(qhi), "r"(one), "r"(plo));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(phi), "r"(one), "r"(qlo));
        return ((uint64_t)ohi << 32) | olo;
    } else {
        uint32_t shi, slo = hi >> (32 - n), ohi, olo;
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(shi) : "r"(hi), "r"(m));
        asm("{\n\t.reg .u64 s, u;\n\tmov.b64 s, {%2, %3};\n\tmad.wide.u32 u, %4, %5, s;\n\tmov.b64 {%0, %1}, u;\n\t}"
            : "=r"(olo), "=r"(ohi) : "r"(slo), "r"(shi), "r"(lo), "r"(m));
        return ((uint64_t)ohi << 32) | olo;
    }
}

// bit i of MB / MD set -> the i-th rotation site of the round (0..4: theta's rot-by-1, 5..29: rho of lane i-5)
// uses variant B / D
template <uint32_t MB, uint32_t MD>
__device__ __forceinline__ void keccak_f1600(uint64_t (&A)[25]) {
    constexpr int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t C[5], R1[5], B[25];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) R1[x] = rotl64_v(((MB >> x) & 1) ? 1 : ((MD >> x) & 1) ? 2 : 0, C[x], 1);
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) {
                const uint64_t v = A[x + 5 * y] ^ C[(x + 4) % 5] ^ R1[(x + 1) % 5];
                B[y + 5 * ((2 * x + 3 * y) % 5)] =
                    rotl64_v(((MB >> (5 + x + 5 * y)) & 1) ? 1 : ((MD >> (5 + x + 5 * y)) & 1) ? 2 : 0, v, RHO[x + 5 * y]);
            }
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= KECCAK_RC[r];
    }
}

template <uint32_t MB, uint32_t MD>
__global__ void __launch_bounds__(128) bench_kernel(uint64_t* out, int perms) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0x9E3779B97F4A7C15ULL * (t + 1) + i;
    for (int p = 0; p < perms; p++) keccak_f1600<MB, MD>(A);
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) acc ^= A[i] * (2 * i + 1);
    out[t] = acc;
}

template <uint32_t MB, uint32_t MD>
static double run(const char* name, uint64_t* d_out, std::vector<uint64_t>& host, int blocks, int perms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench_kernel<MB, MD><<<blocks, 128>>>(d_out, 2);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        bench_kernel<MB, MD><<<blocks, 128>>>(d_out, perms);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaMemcpy(host.data(), d_out, host.size() * 8, cudaMemcpyDeviceToHost);
    const double gps = (double)blocks * 128 * perms / (best * 1e-3) / 1e9;
    printf("%-34s B=%08x D=%08x  %8.3f ms  %7.3f G perms/s\n", name, MB, MD, best, gps);
    return gps;
}

int main(int argc, char** argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int per_sm = argc > 1 ? atoi(argv[1]) : 6;   // CTAs of 128 threads per SM
    const int blocks = sms * per_sm, perms = 200;
    uint64_t* d_out;
    cudaMalloc(&d_out, (size_t)blocks * 128 * 8);
    std::vector<uint64_t> ref((size_t)blocks * 128), got(ref.size());
    printf("%d SMs, %d CTAs x 128 threads (%d per SM), %d permutations per thread\n", sms, blocks, per_sm, perms);
    const double base = run<0u, 0u>("all SHF (baseline)", d_out, ref, blocks, perms);
    auto check = [&](double g) {
        bool ok = got == ref;
        printf("    -> %s, %.3fx baseline\n", ok ? "bit-identical" : "MISMATCH", g / base);
        if (!ok) exit(1);
    };
    check(run<0x0000001Fu, 0u>("B: theta x5", d_out, got, blocks, perms));
    check(run<0x00001FE0u, 0u>("B: 8 rho", d_out, got, blocks, perms));
    check(run<0x0001FFFFu, 0u>("B: 12 rho + 5 theta (17)", d_out, got, blocks, perms));
    check(run<0x003FFFFFu, 0u>("B: 22", d_out, got, blocks, perms));
    check(run<0x03FFFFFFu, 0u>("B: 26", d_out, got, blocks, perms));
    check(run<0x3FFFFFFFu, 0u>("B: all 29", d_out, got, blocks, perms));
    check(run<0u, 0x0000001Fu>("D: theta x5", d_out, got, blocks, perms));
    check(run<0u, 0x0001FFFFu>("D: 17", d_out, got, blocks, perms));
    check(run<0u, 0x3FFFFFFFu>("D: all 29", d_out, got, blocks, perms));
    check(run<0x0001FFFFu, 0x3FFE0000u>("B: 17, D: 12", d_out, got, blocks, perms));
    check(run<0x000FFFFFu, 0x3FF00000u>("B: 20, D: 9", d_out, got, blocks, perms));
    cudaFree(d_out);
    return 0;
}
