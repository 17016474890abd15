// overlap_probe.cu — experiment: how much do the Keccak-bound ExpandMask kernel (ALU pipe) and the
// multiply-bound fused sign core (FMA-heavy pipe) gain from running concurrently on one GPU?
// Times each alone and both on two streams; kernel footprints are selected with the library's
// development knobs (DIL_EM_CTAS, DIL_SC_HALF, DIL_SC_WARPS).  Links against the engine's objects:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Idilithium_b200/csrc -Iinclude \
//        -o build/overlap_probe tools/overlap_probe.cu dilithium_b200/csrc/build/{matvec,sign,ntt,poly}_kernels.o
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 65536;
    const int K = 4, L = 4;
    uint64_t* rhop; uint16_t* kappa; uint32_t* active; int32_t *y, *y2, *w, *a_hat; uint8_t* w1p; uint32_t* ctr;
    CK(cudaMalloc(&rhop, (size_t)n * 64));
    CK(cudaMalloc(&kappa, (size_t)n * 2));
    CK(cudaMalloc(&active, (size_t)n * 4));
    CK(cudaMalloc(&y, (size_t)n * L * 1024));
    CK(cudaMalloc(&y2, (size_t)n * L * 1024));
    CK(cudaMalloc(&w, (size_t)n * K * 1024));
    CK(cudaMalloc(&a_hat, (size_t)K * L * 1024));
    CK(cudaMalloc(&w1p, (size_t)n * K * 192));
    CK(cudaMalloc(&ctr, 64));
    std::vector<uint64_t> hr((size_t)n * 8);
    for (size_t i = 0; i < hr.size(); i++) hr[i] = 0x9E3779B97F4A7C15ULL * (i + 1);
    CK(cudaMemcpy(rhop, hr.data(), hr.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(kappa, 0, (size_t)n * 2));
    CK(dil::launch_iota(active, n, 0));
    std::vector<int32_t> ha((size_t)K * L * 256);
    for (size_t i = 0; i < ha.size(); i++) ha[i] = (int32_t)((i * 2654435761u) % 8380417u);
    CK(cudaMemcpy(a_hat, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f));
    auto em = [&](cudaStream_t st) { CK(dil::launch_expand_mask(2, y2, rhop, kappa, active, n, 1, st)); };
    auto sc = [&](cudaStream_t st) {
        CK(cudaMemsetAsync(ctr, 0, 4, st));
        CK(dil::launch_signcore(w, a_hat, y, K, L, n, 148, st, ctr, w1p));
    };
    CK(dil::launch_expand_mask(2, y, rhop, kappa, active, n, 1, s1));   // valid y for the core
    CK(cudaDeviceSynchronize());
    auto timeit = [&](const char* name, bool do_em, bool do_sc) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s1));
            CK(cudaStreamWaitEvent(s2, e0, 0));
            if (do_em) em(s1);
            if (do_sc) sc(s2);
            CK(cudaEventRecord(f, s2));
            CK(cudaStreamWaitEvent(s1, f, 0));
            CK(cudaEventRecord(e1, s1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("%-28s %8.1f us\n", name, best * 1e3);
        return best;
    };
    printf("n=%u  DIL_EM_CTAS=%s DIL_SC_HALF=%s DIL_SC_WARPS=%s\n", n, getenv("DIL_EM_CTAS") ? getenv("DIL_EM_CTAS") : "-",
           getenv("DIL_SC_HALF") ? getenv("DIL_SC_HALF") : "-", getenv("DIL_SC_WARPS") ? getenv("DIL_SC_WARPS") : "-");
    float a = timeit("ExpandMask alone", true, false);
    float b = timeit("sign core alone", false, true);
    float c = timeit("both, two streams", true, true);
    printf("sum %.1f us, concurrent %.1f us -> %.2fx\n", (a + b) * 1e3, c * 1e3, (a + b) / c);
    return 0;
}
