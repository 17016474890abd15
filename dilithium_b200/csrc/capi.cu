// capi.cu — the C ABI of libdilithium_b200.so (include/dilithium_b200.h).
// Thin host layer: argument checks, device selection, stream hand-off, launch counting and
// the host-pointer variants (staging copies).  There is no CPU fallback anywhere: without a
// CUDA device dil_engine_create fails and nothing else is callable.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "dilithium_b200.h"
#include "engine_priv.h"
#include "kernels.h"

namespace {

int fail_cuda(dil_engine* e, cudaError_t err, const char* what) {
    if (e) {
        std::lock_guard<std::mutex> g(e->err_mu);
        e->last_error = std::string(what) + ": " + cudaGetErrorString(err);
    }
    return DIL_ERR_CUDA;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

using dil::DeviceGuard;

bool dims_ok(int k, int l) { return k >= 1 && k <= 8 && l >= 1 && l <= 8; }
bool level_dims(int k, int l) { return (k == 4 && l == 4) || (k == 6 && l == 5) || (k == 8 && l == 7); }

#define DIL_CHECK_ENGINE(e) \
    if (!(e)) return DIL_ERR_ARG
#define DIL_LAUNCH(e, expr, n_launches)                      \
    do {                                                     \
        DeviceGuard guard_((e)->device);                     \
        if (!guard_.ok) return DIL_ERR_CUDA;                 \
        cudaError_t err_ = (expr);                           \
        if (err_ != cudaSuccess) return fail_cuda((e), err_, #expr); \
        (e)->launches += (n_launches);                       \
    } while (0)

// grow-only device staging buffer `slot`
int stage(dil_engine* e, int slot, size_t bytes, void** out) {
    if (bytes > e->staging_bytes[slot]) {
        if (e->staging[slot]) cudaFree(e->staging[slot]);
        e->staging[slot] = nullptr;
        e->staging_bytes[slot] = 0;
        cudaError_t err = cudaMalloc(&e->staging[slot], bytes);
        if (err != cudaSuccess) {
            std::lock_guard<std::mutex> g(e->err_mu);
            e->last_error = std::string("cudaMalloc staging: ") + cudaGetErrorString(err);
            return DIL_ERR_ALLOC;
        }
        e->staging_bytes[slot] = bytes;
    }
    *out = e->staging[slot];
    return DIL_OK;
}

}  // namespace

extern "C" {

int dil_engine_create(dil_engine_t** out, int device) {
    if (!out) return DIL_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return DIL_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return DIL_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DIL_ERR_CUDA;
    if (prop.major < 10) return DIL_ERR_UNSUPPORTED;  // sm_100a SASS only
    dil_engine* e = new (std::nothrow) dil_engine();
    if (!e) return DIL_ERR_ALLOC;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    DeviceGuard g(device);
    if (!g.ok || cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return DIL_ERR_CUDA;
    }
    *out = e;
    return DIL_OK;
}

int dil_engine_destroy(dil_engine_t* e) {
    if (!e) return DIL_OK;
    {
        DeviceGuard g(e->device);
        dil_internal_free_multi_sign(e);
        for (auto& p : e->staging)
            if (p) cudaFree(p);
        if (e->arena_done) cudaEventDestroy(e->arena_done);
        if (e->host_stream) cudaStreamDestroy(e->host_stream);
        if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    }
    delete e;
    return DIL_OK;
}

const char* dil_status_string(int status) {
    switch (status) {
        case DIL_OK: return "ok";
        case DIL_ERR_NO_DEVICE: return "no CUDA device";
        case DIL_ERR_CUDA: return "CUDA error";
        case DIL_ERR_ARG: return "invalid argument";
        case DIL_ERR_ALLOC: return "allocation failed";
        case DIL_ERR_UNSUPPORTED: return "unsupported device or configuration";
        default: return "unknown status";
    }
}
const char* dil_last_error(const dil_engine_t* e) {
    // a per-thread copy: the engine's string may be rewritten by a failing call on another thread
    static thread_local std::string copy;
    if (!e) return "";
    std::lock_guard<std::mutex> g(const_cast<dil_engine*>(e)->err_mu);
    copy = e->last_error;
    return copy.c_str();
}
int dil_engine_device(const dil_engine_t* e) { return e ? e->device : -1; }
int dil_engine_sm_count(const dil_engine_t* e) { return e ? e->sm_count : 0; }
uint64_t dil_engine_launch_count(const dil_engine_t* e) { return e ? e->launches.load() : 0; }

int dil_level_dims(int level, int* k, int* l) {
    int kk, ll;
    switch (level) {
        case 2: kk = 4; ll = 4; break;
        case 3: kk = 6; ll = 5; break;
        case 5: kk = 8; ll = 7; break;
        default: return DIL_ERR_ARG;
    }
    if (k) *k = kk;
    if (l) *l = ll;
    return DIL_OK;
}

// ---- device-pointer API ----
int dil_ntt_dev(dil_engine_t* e, int32_t* dst, const int32_t* src, size_t n, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (n == 0) return DIL_OK;
    if (!dst || !src || !aligned16(dst) || !aligned16(src) || n > 0x7FFFFFFFu) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_ntt_fwd(dst, src, n, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
int dil_invntt_dev(dil_engine_t* e, int32_t* dst, const int32_t* src, size_t n, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (n == 0) return DIL_OK;
    if (!dst || !src || !aligned16(dst) || !aligned16(src) || n > 0x7FFFFFFFu) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_ntt_inv(dst, src, n, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
static int elementwise(dil_engine_t* e, dil::EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (n == 0) return DIL_OK;
    if (!c || !a || !b || !aligned16(c) || !aligned16(a) || !aligned16(b)) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_elementwise(op, c, a, b, n, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
int dil_pointwise_dev(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* s) {
    return elementwise(e, dil::EwOp::MUL, c, a, b, n, s);
}
int dil_pointwise_acc_dev(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* s) {
    return elementwise(e, dil::EwOp::MULACC, c, a, b, n, s);
}
int dil_add_dev(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* s) {
    return elementwise(e, dil::EwOp::ADD, c, a, b, n, s);
}
int dil_sub_dev(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* s) {
    return elementwise(e, dil::EwOp::SUB, c, a, b, n, s);
}
int dil_matvec_dev(dil_engine_t* e, int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (!dims_ok(k, l)) return DIL_ERR_ARG;
    if (batch == 0) return DIL_OK;
    if (!w || !a_hat || !v || !aligned16(w) || !aligned16(a_hat) || !aligned16(v)) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_matvec(w, a_hat, v, k, l, batch, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
int dil_expand_a_dev(dil_engine_t* e, int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (!dims_ok(k, l)) return DIL_ERR_ARG;
    if (n_rho == 0) return DIL_OK;
    if (!a_hat || !rho || !aligned16(a_hat)) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_expand_a(a_hat, rho, n_rho, k, l, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
int dil_matvec_expand_dev(dil_engine_t* e, int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch,
                          unsigned flags, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (!level_dims(k, l) || (flags & ~7u)) return DIL_ERR_ARG;
    if (batch == 0) return DIL_OK;
    if (!w || !rho || !v || !aligned16(w) || !aligned16(v) || batch > 0x7FFFFFFFu) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_matvec_expand(w, rho, v, k, l, batch, flags, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}
int dil_signcore_dev(dil_engine_t* e, int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (!level_dims(k, l)) return DIL_ERR_ARG;
    if (batch == 0) return DIL_OK;
    if (!w || !a_hat || !y || !aligned16(w) || !aligned16(a_hat) || !aligned16(y) || batch > 0x7FFFFFFFu) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_signcore(w, a_hat, y, k, l, batch, e->sm_count, (cudaStream_t)stream), 1);
    return DIL_OK;
}

int dil_diag_item_rows_threshold(size_t min_batch) {
    dil::set_item_rows_threshold(min_batch);
    return DIL_OK;
}

int dil_diag_keccak_dev(dil_engine_t* e, uint64_t* d_out, unsigned ctas_per_sm, unsigned perms_per_thread, void* stream) {
    DIL_CHECK_ENGINE(e);
    if (!d_out || ctas_per_sm == 0 || ctas_per_sm > 16 || perms_per_thread == 0) return DIL_ERR_ARG;
    DIL_LAUNCH(e, dil::launch_keccak_rate(d_out, (unsigned)e->sm_count * ctas_per_sm, perms_per_thread, (cudaStream_t)stream), 1);
    return DIL_OK;
}

int dil_invntt_tomont_dev(dil_engine_t* e, int32_t* d, const int32_t* s, size_t n, void* st) { return dil_invntt_dev(e, d, s, n, st); }
int dil_poly_pointwise_dev(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n, void* st) {
    return dil_pointwise_dev(e, c, a, b, n, st);
}
int dil_polyvec_matrix_pointwise_dev(dil_engine_t* e, int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l,
                                     size_t batch, void* st) {
    return dil_matvec_dev(e, w, a_hat, v, k, l, batch, st);
}

// ---- host-pointer API ----
// One lock per engine: host calls serialise on the engine's staging buffers.
#define H2D(dst, src, bytes) \
    if ((err = cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda_locked(e, err, "H2D")
#define D2H(dst, src, bytes) \
    if ((err = cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda_locked(e, err, "D2H")

static int fail_cuda_locked(dil_engine* e, cudaError_t err, const char* what) {   // caller holds e->mu (not err_mu)
    std::lock_guard<std::mutex> g(e->err_mu);
    e->last_error = std::string(what) + ": " + cudaGetErrorString(err);
    return DIL_ERR_CUDA;
}

static int host_unary(dil_engine_t* e, bool inverse, int32_t* polys, size_t n) {
    DIL_CHECK_ENGINE(e);
    if (n == 0) return DIL_OK;
    if (!polys) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    cudaStream_t st = e->host_stream;
    cudaError_t err;
    size_t bytes = n * DIL_N * sizeof(int32_t);
    void* d;
    int rc = stage(e, 0, bytes, &d);
    if (rc) return rc;
    H2D(d, polys, bytes);
    err = inverse ? dil::launch_ntt_inv((int32_t*)d, (const int32_t*)d, n, e->sm_count, st)
                  : dil::launch_ntt_fwd((int32_t*)d, (const int32_t*)d, n, e->sm_count, st);
    if (err != cudaSuccess) return fail_cuda_locked(e, err, "launch ntt");
    e->launches += 1;
    D2H(polys, d, bytes);
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda_locked(e, err, "sync");
    return DIL_OK;
}
int dil_ntt_host(dil_engine_t* e, int32_t* polys, size_t n) { return host_unary(e, false, polys, n); }
int dil_invntt_host(dil_engine_t* e, int32_t* polys, size_t n) { return host_unary(e, true, polys, n); }

static int host_binary(dil_engine_t* e, dil::EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n) {
    DIL_CHECK_ENGINE(e);
    if (n == 0) return DIL_OK;
    if (!c || !a || !b) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    cudaStream_t st = e->host_stream;
    cudaError_t err;
    size_t bytes = n * DIL_N * sizeof(int32_t);
    void *da, *db, *dc;
    int rc;
    if ((rc = stage(e, 0, bytes, &da)) || (rc = stage(e, 1, bytes, &db)) || (rc = stage(e, 2, bytes, &dc))) return rc;
    H2D(da, a, bytes);
    H2D(db, b, bytes);
    if (op == dil::EwOp::MULACC) H2D(dc, c, bytes);
    err = dil::launch_elementwise(op, (int32_t*)dc, (const int32_t*)da, (const int32_t*)db, n, e->sm_count, st);
    if (err != cudaSuccess) return fail_cuda_locked(e, err, "launch elementwise");
    e->launches += 1;
    D2H(c, dc, bytes);
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda_locked(e, err, "sync");
    return DIL_OK;
}
int dil_pointwise_host(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n) { return host_binary(e, dil::EwOp::MUL, c, a, b, n); }
int dil_pointwise_acc_host(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n) { return host_binary(e, dil::EwOp::MULACC, c, a, b, n); }
int dil_add_host(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n) { return host_binary(e, dil::EwOp::ADD, c, a, b, n); }
int dil_sub_host(dil_engine_t* e, int32_t* c, const int32_t* a, const int32_t* b, size_t n) { return host_binary(e, dil::EwOp::SUB, c, a, b, n); }

// mode: 0 = matvec (a_hat), 1 = signcore (a_hat), 2 = matvec_expand (rho, flags)
static int host_matvec(dil_engine_t* e, int mode, int32_t* w, const int32_t* a_hat, const uint8_t* rho, const int32_t* v,
                       int k, int l, size_t batch, unsigned flags) {
    DIL_CHECK_ENGINE(e);
    if (mode == 0 ? !dims_ok(k, l) : !level_dims(k, l)) return DIL_ERR_ARG;
    if (mode == 2 && (flags & ~7u)) return DIL_ERR_ARG;
    if (batch == 0) return DIL_OK;
    if (!w || !v || (mode == 2 ? !rho : !a_hat) || batch > 0x7FFFFFFFu) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    cudaStream_t st = e->host_stream;
    cudaError_t err;
    const size_t poly = DIL_N * sizeof(int32_t);
    size_t vbytes = batch * l * poly, wbytes = batch * k * poly;
    size_t abytes = mode == 2 ? ((flags & DIL_RHO_PER_ITEM) ? batch * 32 : 32) : (size_t)k * l * poly;
    void *dv, *dw, *da;
    int rc;
    if ((rc = stage(e, 0, vbytes, &dv)) || (rc = stage(e, 1, wbytes, &dw)) || (rc = stage(e, 2, (abytes + 15) & ~(size_t)15, &da))) return rc;
    H2D(dv, v, vbytes);
    H2D(da, mode == 2 ? (const void*)rho : (const void*)a_hat, abytes);
    if (mode == 0) err = dil::launch_matvec((int32_t*)dw, (const int32_t*)da, (const int32_t*)dv, k, l, batch, e->sm_count, st);
    else if (mode == 1) err = dil::launch_signcore((int32_t*)dw, (const int32_t*)da, (const int32_t*)dv, k, l, batch, e->sm_count, st);
    else err = dil::launch_matvec_expand((int32_t*)dw, (const uint8_t*)da, (const int32_t*)dv, k, l, batch, flags, e->sm_count, st);
    if (err != cudaSuccess) return fail_cuda_locked(e, err, "launch matvec");
    e->launches += 1;
    D2H(w, dw, wbytes);
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda_locked(e, err, "sync");
    return DIL_OK;
}
int dil_matvec_host(dil_engine_t* e, int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch) {
    return host_matvec(e, 0, w, a_hat, nullptr, v, k, l, batch, 0);
}
int dil_signcore_host(dil_engine_t* e, int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch) {
    return host_matvec(e, 1, w, a_hat, nullptr, y, k, l, batch, 0);
}
int dil_matvec_expand_host(dil_engine_t* e, int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch, unsigned flags) {
    return host_matvec(e, 2, w, nullptr, rho, v, k, l, batch, flags);
}
int dil_expand_a_host(dil_engine_t* e, int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l) {
    DIL_CHECK_ENGINE(e);
    if (!dims_ok(k, l)) return DIL_ERR_ARG;
    if (n_rho == 0) return DIL_OK;
    if (!a_hat || !rho) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    cudaStream_t st = e->host_stream;
    cudaError_t err;
    size_t abytes = n_rho * (size_t)(k * l) * DIL_N * sizeof(int32_t), rbytes = n_rho * 32;
    void *da, *dr;
    int rc;
    if ((rc = stage(e, 0, abytes, &da)) || (rc = stage(e, 1, (rbytes + 15) & ~(size_t)15, &dr))) return rc;
    H2D(dr, rho, rbytes);
    err = dil::launch_expand_a((int32_t*)da, (const uint8_t*)dr, n_rho, k, l, e->sm_count, st);
    if (err != cudaSuccess) return fail_cuda_locked(e, err, "launch expand_a");
    e->launches += 1;
    D2H(a_hat, da, abytes);
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda_locked(e, err, "sync");
    return DIL_OK;
}

}  // extern "C"
