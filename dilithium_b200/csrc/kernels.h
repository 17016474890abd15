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

// Device-resident state of the rejection-round loop of one signing batch.  The host enqueues whole rounds ahead of
// time without knowing how many items are still active: every kernel of a round reads the round's size from here,
// plan_kernel (the last launch of a round) turns the re-queued items into the next round, and launches of a round
// that finds nothing left exit at once.  combined_top.v:2217-2228 restarts ONE signature with the next kappa; this
// is that loop for a whole batch, kept on the device.
struct RoundCtl {
    uint32_t n_active;     // items of the current round                       (written between rounds only)
    uint32_t spec;         // speculative attempt slots per item, >= 1
    uint32_t n_slots;      // n_active * spec
    uint32_t cur;          // which of the two active lists holds the current round's items
    uint32_t next_count;   // items re-queued for the next round so far        (atomic, during a round)
    uint32_t ctr_core;     // work counter of the sign core                    (atomic)
    uint32_t ctr_tail;     // work counter of the tail                         (atomic)
    uint32_t done;         // finished items of the batch = length of the completion-ordered done list (atomic)
    uint32_t rounds;       // rounds that had work
    uint32_t total_slots;  // attempt slots processed so far
    uint32_t slot_cap;     // slot capacity of the workspace for this batch    (policy constants)
    uint32_t spec_target;  // once fewer items remain, a round is filled up to this many slots ...
    uint32_t spec_max;     // ... with at most this many slots per item
    uint32_t pad[3];
    uint32_t done_snap[64];   // `done` at the end of round r (index r mod 64): what the host path's drain of round r may copy
};

cudaError_t launch_ntt_fwd(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_ntt_inv(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_elementwise(EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n_polys, int sm_count,
                               cudaStream_t st);
cudaError_t launch_matvec(int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch, int sm_count,
                          cudaStream_t st);
cudaError_t launch_expand_a(int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l, int sm_count, cudaStream_t st);
cudaError_t launch_matvec_expand(int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch,
                                 unsigned flags, int sm_count, cudaStream_t st);
// batch_dev != nullptr: the number of items is read from device memory (round loop); `batch` then only sizes the grid
cudaError_t launch_signcore(int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch, int sm_count,
                            cudaStream_t st, uint32_t* work_ctr = nullptr, uint8_t* w1p = nullptr,
                            const uint32_t* batch_dev = nullptr);

// ---- sign pipeline (sign_kernels.cu) ----
// Buffers of one signing batch that the round kernels share (all device pointers).
struct SignBufs {
    RoundCtl* ctl;
    uint32_t* active[2];       // item lists of the current / next round
    uint32_t* done_list;       // finished items in completion order
    uint64_t *mu, *rhop;       // per item: mu[8], rho'[8]
    uint16_t* kappa;           // per item: next attempt number
    int32_t *y, *w;            // per slot: l / k polynomials
    uint64_t* w1p;             // per slot: packed HighBits(w)
    int8_t* c;                 // per slot: challenge polynomial
    uint64_t* ct_slot;         // per slot: c~
    uint8_t *h_slot, *accepted;
    uint8_t *zp, *h_out;       // per item outputs: packed z, hint bytes
    uint64_t* ct_out;          //                   c~
    uint32_t* attempts;
    bool track_done;           // host path: record completion order for the per-round drain
};
cudaError_t launch_sign_begin(const SignBufs& b, uint32_t n, uint32_t slot_cap, uint32_t spec_target, uint32_t spec_max, cudaStream_t st);
cudaError_t launch_sign_init(uint64_t* mu, uint64_t* rhop, uint16_t* kappa, const uint8_t* tr, const uint8_t* key, size_t key_stride,
                             const uint8_t* msgs, const uint64_t* offsets, uint32_t n, cudaStream_t st);
// cap_slots / cap_items bound the round's size (the kernels read the real size from b.ctl) and size the grids
cudaError_t launch_expand_mask(int level, const SignBufs& b, uint32_t cap_slots, cudaStream_t st);
// ExpandMask + sign core in one kernel (mask_core.cu): y, w and the packed HighBits(w) of every slot of the round
cudaError_t launch_mask_core(int level, const SignBufs& b, const int32_t* a_hat, uint32_t cap_slots, int sm_count, cudaStream_t st,
                             int maxp = 0);
cudaError_t launch_challenge(int level, const SignBufs& b, uint32_t cap_slots, cudaStream_t st);
cudaError_t launch_sign_tail(int level, const SignBufs& b, const int32_t* key_hat, const int8_t* key_small, uint32_t cap_slots,
                             int sm_count, cudaStream_t st);
cudaError_t launch_resolve(int level, const SignBufs& b, uint32_t cap_items, cudaStream_t st);
// ---- one key per signature (dil_sign_multi_*) ----
cudaError_t launch_unpack_keys(int level, int32_t* key_items, const uint8_t* s1p, const uint8_t* s2p, const uint8_t* t0p, size_t n,
                               cudaStream_t st);
cudaError_t launch_scale_inv256(int32_t* x, size_t n_polys, cudaStream_t st);
cudaError_t launch_matvec_multi(int level, int32_t* wh, const int32_t* a_items, const int32_t* yh, const SignBufs& b, uint32_t cap_slots,
                                cudaStream_t st);
cudaError_t launch_pack_w1(int level, uint32_t* w1p, const int32_t* w, size_t n_slots, cudaStream_t st);
cudaError_t launch_sign_tail_multi(int level, const SignBufs& b, const int32_t* key_items, uint32_t cap_slots, int sm_count, cudaStream_t st);
cudaError_t launch_plan(RoundCtl* ctl, uint32_t round, cudaStream_t st);
// copies round `round`'s finished signatures (done list entries [done_snap[round-1], done_snap[round])) to the host
cudaError_t launch_drain(uint8_t* hz, uint8_t* hh, uint8_t* hct, uint32_t* hatt, const SignBufs& b, uint32_t round, uint32_t zb,
                         uint32_t hb, cudaStream_t st);
cudaError_t launch_publish_ctl(uint32_t* host_dst_dev, const RoundCtl* ctl, cudaStream_t st);
// ---- verify pipeline ----
// w1p / hmask non-null: the core applies the hints itself and emits the packed w1' = UseHint(h, w); w is then not written.
// zp / bad_flags non-null as well (levels 2 and 3): z is read from the packed signature (4-byte aligned) inside the core, with
// the ||z|| check; v then only has to hold the challenge polynomial (last of the l + 1 polynomials of an item record)
cudaError_t launch_verify_core(int32_t* w, const int32_t* a_ext, const int32_t* v, int level, size_t batch, int sm_count,
                               cudaStream_t st, uint8_t* w1p = nullptr, const uint32_t* hmask = nullptr, const uint8_t* zp = nullptr,
                               uint32_t* bad_flags = nullptr);
cudaError_t launch_verify_mu(uint64_t* mu, const uint8_t* tr, const uint8_t* msgs, const uint64_t* offsets, uint32_t n,
                             uint32_t tr_stride, cudaStream_t st);
cudaError_t launch_unpack_t1neg(int32_t* out, const uint8_t* t1p, size_t n_polys, cudaStream_t st);
cudaError_t launch_verify_core_item(int32_t* w, const uint8_t* rho, const int32_t* v, const int32_t* extra, int level,
                                    size_t batch, cudaStream_t st, uint8_t* w1p = nullptr, const uint32_t* hmask = nullptr,
                                    bool* fused = nullptr);
cudaError_t launch_tr(uint64_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, cudaStream_t st);
cudaError_t launch_unpack_z(int level, int32_t* v, uint32_t* bad, const uint8_t* zp, uint32_t n, cudaStream_t st);
cudaError_t launch_verify_prep(int level, int32_t* v, uint32_t* hmask, uint32_t* bad, const uint8_t* h, const uint64_t* ctilde,
                               uint32_t n, cudaStream_t st);
cudaError_t launch_usehint_pack(int level, uint32_t* w1p, const int32_t* w, const uint32_t* hmask, uint32_t n, cudaStream_t st);
cudaError_t launch_verify_hash(int level, uint8_t* ok, const uint64_t* mu, const uint64_t* w1p, const uint64_t* ctilde,
                               const uint32_t* bad, uint32_t n, cudaStream_t st);

// ---- diagnostics ----
void set_item_rows_threshold(size_t n);
cudaError_t launch_keccak_rate(uint64_t* out, unsigned ctas, uint32_t perms, cudaStream_t st);

// ---- keygen pipeline ----
cudaError_t launch_keygen_seed(uint8_t* rho, uint64_t* rhop, uint8_t* key, const uint8_t* xi, uint32_t n, cudaStream_t st);
cudaError_t launch_eta_sample(int level, int32_t* s1, int32_t* s2, const uint64_t* rhop, uint32_t n, cudaStream_t st);
cudaError_t launch_t_pack(uint8_t* t1p, uint8_t* t0p, const int32_t* t, const int32_t* s2, size_t n_polys, cudaStream_t st);
cudaError_t launch_s_pack(int eta, uint8_t* sp, const int32_t* s, size_t n_polys, cudaStream_t st);
cudaError_t launch_tr_batch(uint8_t* tr, const uint8_t* rho, const uint8_t* t1p, uint32_t t1_bytes, uint32_t n, cudaStream_t st);

}  // namespace dil
