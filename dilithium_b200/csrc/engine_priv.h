// engine_priv.h - private definition of the engine handle (shared by capi.cu and sign_api.cu)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>

struct dil_engine {
    int device = -1;
    int sm_count = 0;
    std::atomic<uint64_t> launches{0};
    std::atomic<int> sign_in_flight{0};   // sign batches inside their round loop right now (load-adaptive speculation, sign_api.cu)
    std::mutex mu;           // guards staging (host-pointer calls serialise on it)
    std::mutex err_mu;       // guards last_error (set from any entry point, whichever lock it holds)
    void* staging[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t staging_bytes[4] = {0, 0, 0, 0};
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // D2H of finished chunks overlaps compute of the next chunk
    cudaEvent_t arena_done = nullptr;     // recorded after the last kernel that uses staging[3] (multi-key verification arena)
    std::string last_error;
    void* multi_sign[3] = {nullptr, nullptr, nullptr};   // per-level workspace of dil_sign_multi_* (sign_api.cu owns the type)
};

void dil_internal_free_multi_sign(dil_engine* e);   // sign_api.cu


namespace dil {
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
}  // namespace dil
