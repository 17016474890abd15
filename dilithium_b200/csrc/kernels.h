// kernels.h — host-side launchers of the engine's CUDA kernels (internal to the library).
#pragma once
#include <atomic>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace dil {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute: `done` remembers, per kernel,
// on which devices it has been set (one bit per device ordinal), so engines on several GPUs of one process work.
template <class Kern>
inline cudaError_t ensure_dyn_smem(Kern kern, size_t smem, std::atomic<uint64_t>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}


enum class EwOp : int { MUL = 0, MULACC = 1, ADD = 2, SUB = 3 };

cudaError_t launch_ntt_fwd(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_ntt_inv(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_elementwise(EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n_polys, int sm_count,
                               cudaStream_t st);
cudaError_t launch_matvec(int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch, int sm_count,
                          cudaStream_t st);
cudaError_t launch_expand_a(int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l, int sm_count, cudaStream_t st);
cudaError_t launch_matvec_expand(int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch,
                                 unsigned flags, int sm_count, cudaStream_t st);
cudaError_t launch_signcore(int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch, int sm_count,
                            cudaStream_t st, uint32_t* work_ctr = nullptr, uint8_t* w1p = nullptr);

// ---- sign pipeline (sign_kernels.cu) ----
cudaError_t launch_sign_init(uint64_t* mu, uint64_t* rhop, uint16_t* kappa, const uint8_t* tr, const uint8_t* key,
                             const uint8_t* msgs, const uint64_t* offsets, uint32_t n, cudaStream_t st);
cudaError_t launch_expand_mask(int level, int32_t* y, const uint64_t* rhop, const uint16_t* kappa, const uint32_t* active,
                               uint32_t n_slots, uint32_t spec, cudaStream_t st);
cudaError_t launch_pack_w1(int level, uint32_t* w1p, const int32_t* w, uint32_t n_slots, cudaStream_t st);
cudaError_t launch_challenge(int level, int8_t* c, uint64_t* ct_slot, const uint64_t* mu, const uint64_t* w1p,
                             const uint32_t* active, uint32_t n_slots, uint32_t spec, cudaStream_t st);
// Outputs and queues of the resolve step, handed to the tail kernel in rounds with one slot per item: the tail
// then finishes (packs the signature) or re-queues its item itself and launch_resolve is skipped.
struct TailResolve {
    uint8_t* zp;
    uint8_t* h_out;
    uint64_t* ct_out;
    uint32_t* attempts;
    uint16_t* kappa;
    uint32_t* next_active;
    uint32_t* next_count;
    const uint64_t* ct_slot;
    const uint32_t* active;
    uint32_t* done_list;
};
cudaError_t launch_sign_tail(int level, int32_t* y, uint8_t* h_slot, uint8_t* accepted, const int32_t* key_hat,
                             int32_t* w, const int8_t* c, uint32_t n_slots, int sm_count, cudaStream_t st,
                             uint32_t* work_ctr = nullptr, const TailResolve* fused = nullptr, const int8_t* key_small = nullptr);
cudaError_t launch_resolve(int level, uint8_t* zp, uint8_t* h_out, uint64_t* ct_out, uint32_t* attempts, uint16_t* kappa,
                           uint32_t* next_active, uint32_t* next_count, const int32_t* zslot, const uint8_t* h_slot,
                           const uint64_t* ct_slot, const uint8_t* accepted, const uint32_t* active, uint32_t n_items,
                           uint32_t spec, uint32_t* done_list, cudaStream_t st);
cudaError_t launch_drain(uint8_t* hz, uint8_t* hh, uint8_t* hct, uint32_t* hatt, const uint8_t* zp, const uint8_t* h,
                         const uint8_t* ct, const uint32_t* att, const uint32_t* list, uint32_t n, uint32_t zb, uint32_t hb,
                         cudaStream_t st);
cudaError_t launch_iota(uint32_t* dst, uint32_t n, cudaStream_t st);
cudaError_t launch_publish_count(uint32_t* host_dst_dev, const uint32_t* src, cudaStream_t st);
// ---- verify pipeline ----
cudaError_t launch_verify_core(int32_t* w, const int32_t* a_ext, const int32_t* v, int level, size_t batch, int sm_count,
                               cudaStream_t st);
cudaError_t launch_verify_mu(uint64_t* mu, const uint8_t* tr, const uint8_t* msgs, const uint64_t* offsets, uint32_t n,
                             uint32_t tr_stride, cudaStream_t st);
cudaError_t launch_unpack_t1neg(int32_t* out, const uint8_t* t1p, size_t n_polys, cudaStream_t st);
cudaError_t launch_verify_core_item(int32_t* w, const uint8_t* rho, const int32_t* v, const int32_t* extra, int level,
                                    size_t batch, cudaStream_t st);
cudaError_t launch_tr(uint64_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, cudaStream_t st);
cudaError_t launch_unpack_z(int level, int32_t* v, uint32_t* bad, const uint8_t* zp, uint32_t n, cudaStream_t st);
cudaError_t launch_verify_prep(int level, int32_t* v, uint32_t* hmask, uint32_t* bad, const uint8_t* h, const uint64_t* ctilde,
                               uint32_t n, cudaStream_t st);
cudaError_t launch_usehint_pack(int level, uint32_t* w1p, const int32_t* w, const uint32_t* hmask, uint32_t n, cudaStream_t st);
cudaError_t launch_verify_hash(int level, uint8_t* ok, const uint64_t* mu, const uint64_t* w1p, const uint64_t* ctilde,
                               const uint32_t* bad, uint32_t n, cudaStream_t st);

// ---- keygen pipeline ----
cudaError_t launch_keygen_seed(uint8_t* rho, uint64_t* rhop, uint8_t* key, const uint8_t* xi, uint32_t n, cudaStream_t st);
cudaError_t launch_eta_sample(int level, int32_t* s1, int32_t* s2, const uint64_t* rhop, uint32_t n, cudaStream_t st);
cudaError_t launch_t_pack(uint8_t* t1p, uint8_t* t0p, const int32_t* t, const int32_t* s2, size_t n_polys, cudaStream_t st);
cudaError_t launch_s_pack(int eta, uint8_t* sp, const int32_t* s, size_t n_polys, cudaStream_t st);
cudaError_t launch_tr_batch(uint8_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, uint32_t n, cudaStream_t st);

}  // namespace dil
