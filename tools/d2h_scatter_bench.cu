// d2h_scatter_bench.cu — how fast can scattered signature rows (2304 B each) leave the device for pinned host
// memory?  Compares the copy engine (contiguous memcpy), SM stores into mapped host memory, TMA bulk stores
// (cp.async.bulk shared -> host) and cudaMemcpyBatchAsync.  Development aid for dil_sign_batch_host's drain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o d2h_scatter_bench d2h_scatter_bench.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../dilithium_b200/csrc/tma.cuh"
using namespace dil;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 16) drain_st(uint8_t* __restrict__ hz, const uint8_t* __restrict__ zp, const uint32_t* __restrict__ list,
                                                    uint32_t n, uint32_t zb) {
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
    }
}

// one thread per in-flight row: bulk load into the thread's shared-memory slot, bulk store to the host
template <int SLOT_BYTES, int THREADS>
__global__ void __launch_bounds__(THREADS) drain_tma(uint8_t* __restrict__ hz, const uint8_t* __restrict__ zp, const uint32_t* __restrict__ list,
                                                     uint32_t n, uint32_t zb) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[THREADS];
    const uint32_t t = threadIdx.x;
    const uint32_t slot = smem_u32(smem) + t * SLOT_BYTES, bar = smem_u32(&bars[t]);
    mbar_init(bar, 1);
    fence_mbar_init();
    uint32_t phase = 0;
    for (uint32_t i = blockIdx.x * THREADS + t; i < n; i += gridDim.x * THREADS) {
        const uint32_t item = list[i];
        bulk_wait_read<0>();   // the previous store has finished reading the slot
        mbar_expect_tx(bar, zb);
        bulk_g2s(slot, zp + (size_t)item * zb, zb, bar);
        mbar_wait(bar, phase);
        phase ^= 1;
        bulk_s2g(hz + (size_t)item * zb, slot, zb);
        bulk_commit();
    }
    bulk_wait_all<0>();
}

int main(int argc, char** argv) {
    const uint32_t n = 65536, zb = argc > 1 ? atoi(argv[1]) : 2304;
    const size_t bytes = (size_t)n * zb;
    uint8_t *d_z, *h_z, *h_alias;
    uint32_t* d_list;
    CK(cudaMalloc(&d_z, bytes));
    CK(cudaMalloc(&d_list, n * 4));
    CK(cudaHostAlloc(&h_z, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_alias), h_z, 0));
    std::vector<uint8_t> src(bytes);
    std::mt19937_64 rng(1);
    for (size_t i = 0; i < bytes; i += 8) { uint64_t v = rng(); memcpy(&src[i], &v, 8); }
    CK(cudaMemcpy(d_z, src.data(), bytes, cudaMemcpyHostToDevice));
    std::vector<uint32_t> list(n);
    std::iota(list.begin(), list.end(), 0u);
    std::shuffle(list.begin(), list.end(), rng);
    CK(cudaMemcpy(d_list, list.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto report = [&](const char* name, float ms, size_t moved) {
        bool ok = memcmp(h_z, src.data(), moved == bytes ? bytes : 0) == 0;
        printf("%-44s %8.3f ms  %6.1f GB/s %s\n", name, ms, moved / (ms * 1e-3) / 1e9, moved == bytes ? (ok ? "ok" : "MISMATCH") : "");
    };
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        memset(h_z, 0, bytes);
        CK(cudaEventRecord(a));
        CK(cudaMemcpyAsync(h_z, d_z, bytes, cudaMemcpyDeviceToHost));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) report("copy engine, contiguous", ms, bytes);
    }
    for (int ctas : {4, 8, 16, 32, 64, 148, 296}) {
        memset(h_z, 0, bytes);
        CK(cudaEventRecord(a));
        drain_st<<<ctas, 128>>>(h_alias, d_z, d_list, n, zb);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        char nm[64];
        snprintf(nm, 64, "SM stores (st.global v4), %d CTAs x 4 warps", ctas);
        report(nm, ms, bytes);
    }
    {
        constexpr int SLOT = 4608, TH = 32;
        CK(cudaFuncSetAttribute(drain_tma<SLOT, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOT * TH));
        for (int ctas : {1, 2, 4, 8, 16, 32, 64}) {
            memset(h_z, 0, bytes);
            CK(cudaEventRecord(a));
            drain_tma<SLOT, TH><<<ctas, TH, SLOT * TH>>>(h_alias, d_z, d_list, n, zb);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&ms, a, b));
            char nm[64];
            snprintf(nm, 64, "TMA bulk stores, %d CTAs x 32 rows in flight", ctas);
            report(nm, ms, bytes);
        }
    }
    {
        // cudaMemcpyBatchAsync: one entry per row
        const size_t cnt = 16384;
        std::vector<void*> dsts(cnt), srcs(cnt);
        std::vector<size_t> sizes(cnt, zb);
        for (size_t i = 0; i < cnt; i++) { dsts[i] = h_z + (size_t)list[i] * zb; srcs[i] = d_z + (size_t)list[i] * zb; }
        cudaMemcpyAttributes attr{};
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
        size_t idx = 0, fail = 0;
        cudaStream_t st;
        CK(cudaStreamCreate(&st));
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a, st));
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), cnt, &attr, &idx, 1, &fail, st);
            if (e != cudaSuccess) { printf("cudaMemcpyBatchAsync: %s\n", cudaGetErrorString(e)); break; }
            CK(cudaEventRecord(b, st));
            CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep) report("cudaMemcpyBatchAsync, 16384 rows", ms, cnt * zb);
        }
    }
    return 0;
}
