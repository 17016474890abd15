// d2h_ceiling.cu — what can this box's host memory absorb when N GPUs stream signatures to it AT THE SAME TIME?
//
// Every GPU holds `rows` signature rows (2424 B each: Dilithium-2 z | h | c~, 159 MB for 65 536 signatures) and writes
// them `reps` times into its own pinned host buffer.  All GPUs start on a common barrier; the figure of merit is
// bytes of ALL GPUs / wall time from the barrier to the last GPU's completion.  Variants:
//   ce        copy engine, one contiguous cudaMemcpyAsync per repetition
//   ce16      copy engine, 16 contiguous chunks per repetition
//   sm<G>     SM stores (st.global.v4 into mapped pinned memory), G CTAs x 512 threads, contiguous
//   smrow<G>  SM stores, one warp per 2304-B row in a random row order (the shape of the engine's drain kernel)
// Modes: threads (one process, one host thread per GPU - the shape of dil_pool) or procs (one process per GPU,
// fork before CUDA is touched - the shape of torchrun).  Subsets N = 1, 2, 4, ... up to --gpus.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/d2h_ceiling tools/d2h_ceiling.cu -lpthread
//   tools/bin/d2h_ceiling [--gpus 8] [--rows 65536] [--reps 8] [--mode threads|procs|both] [--wc 0|1] [--min-n 1]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr uint32_t ROW = 2424, ZROW = 2304;

__global__ void __launch_bounds__(512) st_contig(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < nvec; i += stride) dst[i] = __ldcs(src + i);
}

__global__ void __launch_bounds__(512) st_rows(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, const uint32_t* __restrict__ list,
                                               uint32_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n; i += nwarps) {
        const uint32_t item = list[i];
        const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)item * ZROW);
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)item * ZROW);
        uint32_t t = lane;
        for (; t + 96 < ZROW / 16; t += 128) {
            uint4 a = __ldcs(s + t), b = __ldcs(s + t + 32), c = __ldcs(s + t + 64), e = __ldcs(s + t + 96);
            d[t] = a; d[t + 32] = b; d[t + 64] = c; d[t + 96] = e;
        }
        for (; t < ZROW / 16; t += 32) d[t] = __ldcs(s + t);
    }
}

struct Shared {   // lives in MAP_SHARED memory so that forked processes can use it too
    std::atomic<int> arrive[64];
    std::atomic<int> go[64];
    double ms[64][16];
};

struct Variant { const char* name; int kind; int ctas; };
static const Variant VARIANTS[] = {{"ce", 0, 0}, {"ce16", 1, 0}, {"sm4", 2, 4}, {"sm16", 2, 16}, {"sm64", 2, 64}, {"smrow4", 3, 4}, {"smrow16", 3, 16}};
constexpr int NV = sizeof(VARIANTS) / sizeof(VARIANTS[0]);

static void barrier(Shared* sh, int phase, int n) {
    sh->arrive[phase].fetch_add(1);
    while (sh->arrive[phase].load() < n) { }
}

// one GPU's work: for every (subset, variant) phase this GPU takes part in, wait on the barrier, run, record its time
static int g_wc = 0, g_min_n = 1;   // --wc 1: write-combined pinned memory (not snooped by the CPU caches); --min-n: skip smaller subsets

static void worker(int dev, int max_gpus, uint32_t rows, int reps, Shared* sh) {
    CK(cudaSetDevice(dev));
    const size_t bytes = (size_t)rows * ROW, zbytes = (size_t)rows * ZROW;
    uint8_t *dsrc, *hdst, *hdst_dev;
    uint32_t* dlist;
    CK(cudaMalloc(&dsrc, bytes));
    CK(cudaMemset(dsrc, 0x5a, bytes));
    CK(cudaHostAlloc(&hdst, bytes, cudaHostAllocMapped | cudaHostAllocPortable | (g_wc ? cudaHostAllocWriteCombined : 0)));
    memset(hdst, 0, bytes);
    CK(cudaHostGetDevicePointer(&hdst_dev, hdst, 0));
    std::vector<uint32_t> list(rows);
    for (uint32_t i = 0; i < rows; i++) list[i] = i;
    std::shuffle(list.begin(), list.end(), std::mt19937(dev + 1));
    CK(cudaMalloc(&dlist, rows * 4));
    CK(cudaMemcpy(dlist, list.data(), rows * 4, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int phase = 0;
    for (int n = 1; n <= max_gpus; n *= 2) {
        for (int v = 0; v < NV; v++, phase++) {
            if (dev >= n || n < g_min_n) continue;
            const Variant& V = VARIANTS[v];
            auto run = [&](int r) {
                for (int i = 0; i < r; i++) {
                    if (V.kind == 0) CK(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, st));
                    else if (V.kind == 1) {
                        const size_t ch = ((bytes / 16) + 15) & ~(size_t)15;
                        for (size_t o = 0; o < bytes; o += ch) CK(cudaMemcpyAsync(hdst + o, dsrc + o, std::min(ch, bytes - o), cudaMemcpyDeviceToHost, st));
                    } else if (V.kind == 2) st_contig<<<V.ctas, 512, 0, st>>>((uint4*)hdst_dev, (const uint4*)dsrc, bytes / 16);
                    else st_rows<<<V.ctas, 512, 0, st>>>(hdst_dev, dsrc, dlist, rows);
                }
            };
            run(1);   // warm-up
            CK(cudaStreamSynchronize(st));
            barrier(sh, phase, n);
            auto t0 = std::chrono::steady_clock::now();
            run(reps);
            CK(cudaStreamSynchronize(st));
            sh->ms[dev][phase % 16] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            sh->go[phase].fetch_add(1);
            while (sh->go[phase].load() < n) { }
            if (dev == 0) {
                double worst = 0, sum_bw = 0;
                const size_t moved = (V.kind == 3 ? zbytes : bytes) * reps;
                for (int g = 0; g < n; g++) { worst = std::max(worst, sh->ms[g][phase % 16]); sum_bw += moved / (sh->ms[g][phase % 16] * 1e6); }
                printf("  N=%d %-8s aggregate %7.1f GB/s (all GPUs / slowest GPU's time)   per-GPU min %5.1f  mean %5.1f GB/s\n", n, V.name,
                       n * moved / (worst * 1e6), moved / (worst * 1e6), sum_bw / n);
                fflush(stdout);
            }
            sh->go[phase].fetch_add(100);   // second stage: results printed
            while (sh->go[phase].load() < n + 100) { }
        }
    }
    cudaFreeHost(hdst);
    cudaFree(dsrc);
    cudaFree(dlist);
}

int main(int argc, char** argv) {
    int gpus = 8, reps = 8;
    uint32_t rows = 65536;
    std::string mode = "both";
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--gpus")) gpus = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--rows")) rows = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--mode")) mode = argv[i + 1];
        else if (!strcmp(argv[i], "--wc")) g_wc = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--min-n")) g_min_n = atoi(argv[i + 1]);
    }
    auto fresh = []() {
        void* p = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        Shared* sh = new (p) Shared();
        for (auto& a : sh->arrive) a.store(0);
        for (auto& a : sh->go) a.store(0);
        return sh;
    };
    printf("d2h_ceiling: %u rows x %u B = %.1f MB per GPU per repetition, %d repetitions, up to %d GPUs, %ld host CPUs\n", rows, ROW,
           rows * (double)ROW / 1e6, reps, gpus, sysconf(_SC_NPROCESSORS_ONLN));
    printf("host buffers: %s pinned memory\n", g_wc ? "WRITE-COMBINED" : "ordinary (cacheable)");
    if (mode == "procs" || mode == "both") {
        // fork BEFORE any CUDA call: each child creates its own context, as torchrun ranks do
        printf("mode procs (one process per GPU):\n");
        fflush(stdout);
        Shared* sh = fresh();
        std::vector<pid_t> kids;
        for (int d = 0; d < gpus; d++) {
            pid_t p = fork();
            if (p == 0) { worker(d, gpus, rows, reps, sh); fflush(stdout); _exit(0); }
            kids.push_back(p);
        }
        for (pid_t p : kids) { int s; waitpid(p, &s, 0); }
    }
    if (mode == "threads" || mode == "both") {
        int count = 0;
        CK(cudaGetDeviceCount(&count));
        if (count < gpus) { fprintf(stderr, "only %d GPUs visible\n", count); gpus = count; }
        printf("mode threads (one process, one host thread per GPU):\n");
        fflush(stdout);
        Shared* sh = fresh();
        std::vector<std::thread> th;
        for (int d = 0; d < gpus; d++) th.emplace_back(worker, d, gpus, rows, reps, sh);
        for (auto& t : th) t.join();
    }
    return 0;
}
