// pipe_overlap_bench.cu — do the integer ALU pipe (LOP3 / SHF / IADD3) and the FMA pipe (IMAD, IMAD.WIDE, IMAD.HI, FFMA)
// of an sm_100a sub-partition issue at the same time?  Every signing kernel is bound by one of the two (Keccak: ALU;
// transforms: the multiplier half of the FMA pipe), and every "fuse the hash into the transform kernel" idea assumes
// the answer is yes.  The kernel runs NA ALU instructions and NF FMA-pipe instructions per loop iteration on independent
// register chains (inline PTX, so ptxas keeps the instruction selection) and reports warp instructions per cycle per SM
// sub-partition for each mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o pipe_overlap_bench pipe_overlap_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

enum FmaKind { IMAD = 0, IMAD_WIDE = 1, IMAD_HI = 2, FFMA = 3 };
enum AluKind { LOP3 = 0, SHF = 1, IADD3 = 2 };

template <int NA, int AK, int NF, int FK>
__global__ void __launch_bounds__(256) mix_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], f[8];
    uint64_t w[8];
    float ff[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (threadIdx.x + 1) + i;
        f[i] = seed + 3 * i + threadIdx.x;
        w[i] = ((uint64_t)(seed + threadIdx.x) << 20) ^ (i * 77u + threadIdx.x);
        ff[i] = (float)(threadIdx.x + i) * 1e-3f;
    }
    const uint32_t m = seed | 1u, c = seed >> 3;
    const float fm = 1.0000001f, fc = 1e-7f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < (NA > NF ? NA : NF); u++) {
            if (u < NA) {
                if (AK == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[u & 7]) : "r"(m), "r"(c));
                else if (AK == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[u & 7]) : "r"(m));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[u & 7]) : "r"(m));
            }
            if (u < NF) {
                if (FK == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(f[u & 7]) : "r"(m), "r"(c));
                else if (FK == IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(f[u & 7]) : "r"(m), "r"(c));
                else if (FK == IMAD_WIDE)
                    asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmul.wide.u32 %0, lo, hi;\n\t}" : "+l"(w[u & 7]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(ff[u & 7]) : "f"(fm), "f"(fc));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] ^ f[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ __float_as_uint(ff[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static int g_sms = 148, g_clock_khz = 1965000;

template <int NA, int AK, int NF, int FK>
static void run(const char* name, uint32_t* d_out, int ctas_per_sm) {
    const int iters = 4096, blocks = g_sms * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    mix_kernel<NA, AK, NF, FK><<<blocks, 256>>>(d_out, 16, 12345u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        mix_kernel<NA, AK, NF, FK><<<blocks, 256>>>(d_out, iters, 12345u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    // warp instructions of each kind per sub-partition per cycle
    const double warps_per_smsp = ctas_per_sm * 8 / 4.0;
    const double cycles = best * 1e-3 * g_clock_khz * 1e3;
    const double alu_rate = warps_per_smsp * iters * NA / cycles, fma_rate = warps_per_smsp * iters * NF / cycles;
    printf("%-44s %d warps/SMSP  %8.3f ms   ALU %.3f  FMA-pipe %.3f  total %.3f warp-instr/clk/SMSP\n", name, (int)warps_per_smsp, best,
           alu_rate, fma_rate, alu_rate + fma_rate);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
}

int main(int argc, char** argv) {
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&g_clock_khz, cudaDevAttrClockRate, 0);
    const int cps = argc > 1 ? atoi(argv[1]) : 2;   // CTAs of 256 threads per SM
    uint32_t* d_out;
    cudaMalloc(&d_out, (size_t)g_sms * cps * 256 * 4);
    printf("%d SMs, max clock %d kHz (rates assume it), %d CTAs x 256 threads per SM\n", g_sms, g_clock_khz, cps);
    run<16, LOP3, 0, IMAD>("16 LOP3", d_out, cps);
    run<16, SHF, 0, IMAD>("16 SHF", d_out, cps);
    run<16, IADD3, 0, IMAD>("16 IADD", d_out, cps);
    run<0, LOP3, 16, IMAD>("16 IMAD", d_out, cps);
    run<0, LOP3, 16, IMAD_HI>("16 IMAD.HI", d_out, cps);
    run<0, LOP3, 16, IMAD_WIDE>("16 IMAD.WIDE", d_out, cps);
    run<0, LOP3, 16, FFMA>("16 FFMA", d_out, cps);
    run<16, LOP3, 16, IMAD>("16 LOP3 + 16 IMAD", d_out, cps);
    run<16, LOP3, 8, IMAD>("16 LOP3 + 8 IMAD", d_out, cps);
    run<16, LOP3, 4, IMAD>("16 LOP3 + 4 IMAD", d_out, cps);
    run<16, SHF, 16, IMAD>("16 SHF + 16 IMAD", d_out, cps);
    run<16, LOP3, 16, FFMA>("16 LOP3 + 16 FFMA", d_out, cps);
    run<16, LOP3, 8, IMAD_HI>("16 LOP3 + 8 IMAD.HI", d_out, cps);
    run<16, LOP3, 8, IMAD_WIDE>("16 LOP3 + 8 IMAD.WIDE", d_out, cps);
    run<16, LOP3, 4, IMAD_WIDE>("16 LOP3 + 4 IMAD.WIDE", d_out, cps);
    run<8, LOP3, 16, IMAD>("8 LOP3 + 16 IMAD", d_out, cps);
    run<16, IADD3, 16, FFMA>("16 IADD + 16 FFMA", d_out, cps);
    cudaFree(d_out);
    return 0;
}
