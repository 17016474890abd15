/*
 * dilithium_b200.h — C ABI of the B200-native batched Dilithium polynomial-arithmetic
 * engine (libdilithium_b200.so).  Plain pointers and sizes only; no C++/torch types.
 *
 * This is the drop-in boundary for the hot path of GMUCERG/Dilithium (paths relative to
 * the reference root).  Each entry point names the reference interface it replaces:
 *
 *   dil_ntt*            void ntt(data_t a[256])                 dilithium-256/reference_code/ref_ntt.h:30
 *                       void ntt2x2_ref(data_t a[256])          dilithium-256/reference_code/ref_ntt2x2.h:31
 *                       operation_module mode 0 (FORWARD)       rtl_src/operation_module.v:28-55
 *   dil_invntt*         void invntt(data_t a[256])              dilithium-256/reference_code/ref_ntt.h:36
 *                       void invntt2x2_ref(data_t a[256])       dilithium-256/reference_code/ref_ntt2x2.h:33
 *                       operation_module mode 1 (INVERSE)       rtl_src/operation_module.v:28-55
 *   dil_pointwise*      void pointwise_barrett(c, a, b)         dilithium-256/reference_code/ref_ntt.h:32-34
 *   dil_pointwise_acc*  operation_module mode 2 (MULT = a*b+acc) rtl_src/butterfly.v:144-150
 *   dil_add* / dil_sub* operation_module modes 3 / 4            rtl_src/butterfly.v:151-164
 *   dil_matvec*         MULT_MODE loop nest ("polyvec_matrix_pointwise")
 *                                                               rtl_src/combined_top.v:921-958, :1347-1386, :1875-1913
 *   dil_expand_a*       gen_a_ext / sampler_a_ext / rejection_a rtl_src/gen_a_ext.v:29-409,
 *                                                               sampler_a_ext.v:100-146, rejection_a.v:63-113
 *   dil_matvec_expand*  ExpandA fused into the mat-vec (A never written to HBM); optional
 *                       fused NTT on the input vector and INTT on the output vector
 *                       (sign: NTT_Y -> MULT_A_Y -> NTTI_W, combined_top.v:1850-1933)
 *
 * Conventions:
 *   - a polynomial is 256 contiguous int32 (params.h:30-35); a batch is n contiguous polys.
 *   - inputs may be any representative in (-Q, Q); outputs are CANONICAL in [0, Q).  The
 *     reference emits signed non-canonical residues and its own tests compare mod Q
 *     (ref_test_ntt_ntt2x2.cpp:31-42); results here are equal to the reference's mod Q.
 *   - arithmetic is plain-domain (no Montgomery factor), forward NTT natural -> bit-reversed
 *     order, inverse includes the 256^-1 scaling, exactly as ref_ntt.cpp:28-87.
 *   - caller owns every buffer; no allocation happens inside *_dev calls; dst == src
 *     (in-place) is legal everywhere, as is c == a for pointwise (ntt2x2_test.cpp:102).
 *   - *_dev entry points take DEVICE pointers (16-byte aligned) and a cudaStream_t passed as
 *     void* (NULL = default stream); they enqueue work and return without synchronising.
 *   - *_host entry points take HOST pointers, copy in, run, copy out and synchronise.
 *   - every function returns 0 on success or a negative dil_status; there is no CPU fallback:
 *     without a usable CUDA device dil_engine_create fails with DIL_ERR_NO_DEVICE.
 *   - an engine handle is bound to one device; calls on one handle from several threads are
 *     safe as long as they use different streams for overlapping buffers.  Key handles own
 *     internal workspaces: calls on one key are serialised by the library (host side by a
 *     per-key lock, device side by an event the next user of the workspace waits on), so
 *     _dev calls on one key from several streams are legal and simply run back to back.
 */
#ifndef DILITHIUM_B200_H
#define DILITHIUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIL_Q 8380417
#define DIL_N 256

typedef struct dil_engine dil_engine_t;

typedef enum {
    DIL_OK = 0,
    DIL_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
    DIL_ERR_CUDA = -2,        /* a CUDA call failed; see dil_last_error() */
    DIL_ERR_ARG = -3,         /* null / misaligned pointer, bad level, bad dims */
    DIL_ERR_ALLOC = -4,       /* host or device allocation failed (host-variant staging) */
    DIL_ERR_UNSUPPORTED = -5
} dil_status;

/* flags for dil_matvec_expand_* */
#define DIL_RHO_SHARED   0u   /* one 32-byte rho for the whole batch               */
#define DIL_RHO_PER_ITEM 1u   /* rho[item] : batch x 32 bytes                       */
#define DIL_NTT_INPUT    2u   /* v is in the time domain: apply NTT before the product */
#define DIL_INTT_OUTPUT  4u   /* apply INTT to every output polynomial              */

/* ---- engine lifetime ---- */
int dil_engine_create(dil_engine_t **out, int device);
int dil_engine_destroy(dil_engine_t *e);
const char *dil_status_string(int status);
const char *dil_last_error(const dil_engine_t *e);      /* text of the last CUDA error on e */
int dil_engine_device(const dil_engine_t *e);
int dil_engine_sm_count(const dil_engine_t *e);
/* (k, l) for level in {2,3,5} (combined_top.v:520-551); returns DIL_ERR_ARG otherwise */
int dil_level_dims(int level, int *k, int *l);
/* number of engine kernels launched on this handle since creation (bench evidence) */
uint64_t dil_engine_launch_count(const dil_engine_t *e);

/* ---- device-pointer API (hot path) ---- */
int dil_ntt_dev(dil_engine_t *e, int32_t *dst, const int32_t *src, size_t n_polys, void *stream);
int dil_invntt_dev(dil_engine_t *e, int32_t *dst, const int32_t *src, size_t n_polys, void *stream);
int dil_pointwise_dev(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys, void *stream);
/* c = c + a o b */
int dil_pointwise_acc_dev(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys, void *stream);
int dil_add_dev(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys, void *stream);
int dil_sub_dev(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys, void *stream);
/* w[b][i] = sum_j a_hat[i*l+j] o v[b][j];  a_hat: k*l polys shared by the batch,
   v: batch*l polys, w: batch*k polys, all NTT domain */
int dil_matvec_dev(dil_engine_t *e, int32_t *w, const int32_t *a_hat, const int32_t *v,
                   int k, int l, size_t batch, void *stream);
/* a_hat[r][i*l+j] = ExpandA(rho[r])_{i,j}: n_rho * k*l polys (materialised; for tests/keys) */
int dil_expand_a_dev(dil_engine_t *e, int32_t *a_hat, const uint8_t *rho, size_t n_rho,
                     int k, int l, void *stream);
/* fused: w[b] = [INTT] ( ExpandA(rho) * [NTT] v[b] ); A is never written to HBM */
int dil_matvec_expand_dev(dil_engine_t *e, int32_t *w, const uint8_t *rho, const int32_t *v,
                          int k, int l, size_t batch, unsigned flags, void *stream);
/* cfg2 sign core with a pre-expanded shared A: w[b] = INTT(a_hat * NTT(y[b])) in ONE kernel */
int dil_signcore_dev(dil_engine_t *e, int32_t *w, const int32_t *a_hat, const int32_t *y,
                     int k, int l, size_t batch, void *stream);

/* ---- host-pointer API (copies inside; synchronous) ---- */
int dil_ntt_host(dil_engine_t *e, int32_t *polys, size_t n_polys);
int dil_invntt_host(dil_engine_t *e, int32_t *polys, size_t n_polys);
int dil_pointwise_host(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys);
int dil_pointwise_acc_host(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys);
int dil_add_host(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys);
int dil_sub_host(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys);
int dil_matvec_host(dil_engine_t *e, int32_t *w, const int32_t *a_hat, const int32_t *v,
                    int k, int l, size_t batch);
int dil_expand_a_host(dil_engine_t *e, int32_t *a_hat, const uint8_t *rho, size_t n_rho, int k, int l);
int dil_matvec_expand_host(dil_engine_t *e, int32_t *w, const uint8_t *rho, const int32_t *v,
                           int k, int l, size_t batch, unsigned flags);
int dil_signcore_host(dil_engine_t *e, int32_t *w, const int32_t *a_hat, const int32_t *y,
                      int k, int l, size_t batch);

/* ---- batched deterministic signing (the caller of the hot path) ----
 * Replaces, for a whole batch, the sign mode of the reference's top level
 * (rtl_src/combined_top.v:31-41 mode 2, FSMs :1535-2229) with the I/O of rtl_tb/tb_sign_top.v:171-335:
 * inputs rho, tr, K, s1, s2, t0 (bit-packed exactly as the KAT files / decoder.v:89-143) and
 * messages; outputs per signature z (l * 576|640 bytes, gamma1 - z packed), h (omega + k bytes),
 * c~ (32 bytes) and the number of attempts.  Round-3.1, deterministic (rho' = SHAKE256(K || mu)).
 * msgs = all messages concatenated; offsets[i]..offsets[i+1] delimit message i (n+1 entries). */
typedef struct dil_sign_key dil_sign_key_t;
int dil_sign_sizes(int level, size_t *z_bytes, size_t *h_bytes);
int dil_sign_key_create(dil_engine_t *e, dil_sign_key_t **out, int level, const uint8_t *rho, const uint8_t *key,
                        const uint8_t *tr, const uint8_t *s1_packed, const uint8_t *s2_packed, const uint8_t *t0_packed);
int dil_sign_key_destroy(dil_engine_t *e, dil_sign_key_t *k);
/* Streaming use: a key handle signs one batch at a time (calls on one handle serialise).  To keep several batches in
 * flight on one GPU - the next batch signs while the small last rejection rounds of the previous one leave SMs idle -
 * create two to four handles of the same key and either call from one host thread per handle (dil_sign_batch_host uses streams
 * owned by the handle; dil_sign_batch_dev the stream it is given) or drive them all from one thread with the asynchronous
 * *_begin / dil_sign_batch_finish pair below.  Measured on one B200, Dilithium-2, 65 536-message
 * batches: 12.1 M signs/s with one batch at a time, 13.9 M with two and 14.6 M with four in flight.
 * host pointers.  When z, h, ctilde (and attempts, if given) are all pinned host memory the device can address
 * (cudaHostAlloc / cudaHostRegister; z and ctilde 16-byte aligned), finished signatures are streamed into them
 * round by round while the batch is still signing; pageable buffers take chunked copy-engine transfers.  The
 * results are identical either way.  attempts may be NULL. */
int dil_sign_batch_host(dil_engine_t *e, dil_sign_key_t *k, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                        uint8_t *z, uint8_t *h, uint8_t *ctilde, uint32_t *attempts);
/* device pointers (d_z 16-byte, d_ctilde 8-byte aligned).  The rejection loop runs on the device: the call enqueues the
 * expected number of rounds at once and synchronises the stream once at the end (again only if items are still active). */
int dil_sign_batch_dev(dil_engine_t *e, dil_sign_key_t *k, const uint8_t *d_msgs, const uint64_t *d_offsets, size_t n,
                       uint8_t *d_z, uint8_t *d_h, uint8_t *d_ctilde, uint32_t *d_attempts, void *stream);
/* Asynchronous pair: several batches in flight from ONE host thread.  *_begin enqueues the batch (the rejection loop runs on the
 * device) and returns at once; dil_sign_batch_finish waits for it, enqueues further rounds if messages are still unsigned, and
 * returns the batch's status - only then are the outputs complete.  A key handle carries one batch at a time (a second *_begin,
 * a synchronous call or dil_sign_key_set_tuning on a busy handle returns DIL_ERR_ARG); use one handle of the same key per batch in
 * flight.  Limits: at most 1.25 x dev_chunk (2^20) messages for _dev_begin and 1.25 x host_chunk (2^18) for _host_begin; _host_begin
 * needs pinned, device-addressable output buffers (the streaming path); every buffer passed to *_begin (and the stream given to
 * _dev_begin) must stay valid until the finish.  dil_sign_key_destroy completes a batch that was begun and never finished.
 * dil_sign_batch_dev / _host are exactly begin + finish for batches within those limits. */
int dil_sign_batch_dev_begin(dil_engine_t *e, dil_sign_key_t *k, const uint8_t *d_msgs, const uint64_t *d_offsets, size_t n,
                             uint8_t *d_z, uint8_t *d_h, uint8_t *d_ctilde, uint32_t *d_attempts, void *stream);
int dil_sign_batch_host_begin(dil_engine_t *e, dil_sign_key_t *k, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                              uint8_t *z, uint8_t *h, uint8_t *ctilde, uint32_t *attempts);
int dil_sign_batch_finish(dil_engine_t *e, dil_sign_key_t *k);
uint32_t dil_sign_last_rounds(const dil_sign_key_t *k);   /* rejection rounds of the last batch */
uint64_t dil_sign_last_slots(const dil_sign_key_t *k);    /* signing attempts (slots) the last batch executed, speculative ones included */
/* Per-key tuning of the batch scheduler; zero fields keep the defaults.  Results never depend on these. */
typedef struct {
    uint32_t spec_target;    /* straggler speculation: rounds with fewer items are filled up to this many attempt slots         */
    uint32_t spec_max;       /* ... with at most this many consecutive attempts per item (at most 32).  Defaults follow the
                                engine's load: 32768 / 32 for a lone batch, 16384 / 16 with two and 8192 / 8 with three or
                                more sign batches in flight on the engine (other batches fill the idle SMs instead)            */
    size_t dev_chunk;        /* dil_sign_batch_dev signs larger requests in pieces of this many messages (2^20)               */
    size_t host_chunk;       /* the same for the streaming path of dil_sign_batch_host (2^18)                                 */
    int host_copy_path;      /* non-zero: dil_sign_batch_host always uses chunked copy-engine transfers                       */
    int fused_mask;          /* non-zero: ExpandMask runs inside the sign core (mask_core kernel) instead of as its own kernel;
                                measured slower on B200, kept selectable                                         */
    int mask_producers;      /* fused mask_core: warps per CTA that may squeeze masks at once (0 = one per ring buffer)        */
} dil_sign_tuning;
int dil_sign_key_set_tuning(dil_sign_key_t *k, const dil_sign_tuning *t);   /* t == NULL restores the defaults */
/* One key PER SIGNATURE, as the reference's sign driver streams it (rtl_tb/tb_sign_top.v:171-284: rho, tr, K, s1, s2, t0
 * precede every message): rho, key (K), tr are n x 32 bytes, s1 / s2 / t0 n bit-packed records as in dil_sign_key_create.
 * A_hat = ExpandA(rho[i]) and NTT(s1, s2, t0) are computed per item on the device.  Same outputs as dil_sign_batch_*. */
int dil_sign_multi_host(dil_engine_t *e, int level, const uint8_t *rho, const uint8_t *key, const uint8_t *tr,
                        const uint8_t *s1_packed, const uint8_t *s2_packed, const uint8_t *t0_packed, const uint8_t *msgs,
                        const uint64_t *offsets, size_t n, uint8_t *z, uint8_t *h, uint8_t *ctilde, uint32_t *attempts);
int dil_sign_multi_dev(dil_engine_t *e, int level, const uint8_t *d_rho, const uint8_t *d_key, const uint8_t *d_tr,
                       const uint8_t *d_s1_packed, const uint8_t *d_s2_packed, const uint8_t *d_t0_packed, const uint8_t *d_msgs,
                       const uint64_t *d_offsets, size_t n, uint8_t *d_z, uint8_t *d_h, uint8_t *d_ctilde, uint32_t *d_attempts,
                       void *stream);
/* optional device timing of the pipeline's kernel classes (CUDA events on the launching stream):
   index 0 init (mu, rho'), 1 ExpandMask, 2 fused sign core, 3 w1 pack, 4 challenge, 5 tail, 6 resolve;
   ms[8] = summed kernel time of the last batch, units[8] = slots (attempts) each class processed */
int dil_sign_set_profile(dil_sign_key_t *k, int on);
int dil_sign_get_profile(const dil_sign_key_t *k, double *ms, uint64_t *units);

/* ---- batched verification (combined_top.v mode 1, FSM :1080-1534) ----
 * I/O of rtl_tb/tb_verify_top.v:144-249: public key rho + t1 (10-bit packed), per signature c~, z, h and the
 * message; result ok[i] = 1 (accept) / 0 (reject: hash mismatch, ||z|| >= gamma1 - beta or malformed hint).
 * One public key per handle; A_hat and -NTT(t1*2^13) are expanded once, tr = SHAKE256(rho || t1) on device. */
typedef struct dil_verify_key dil_verify_key_t;
int dil_verify_key_create(dil_engine_t *e, dil_verify_key_t **out, int level, const uint8_t *rho, const uint8_t *t1_packed);
int dil_verify_key_destroy(dil_engine_t *e, dil_verify_key_t *k);
int dil_verify_batch_host(dil_engine_t *e, dil_verify_key_t *k, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                          const uint8_t *z, const uint8_t *h, const uint8_t *ctilde, uint8_t *ok);
int dil_verify_batch_dev(dil_engine_t *e, dil_verify_key_t *k, const uint8_t *d_msgs, const uint64_t *d_offsets, size_t n,
                         const uint8_t *d_z, const uint8_t *d_h, const uint8_t *d_ctilde, uint8_t *d_ok, void *stream);

/* one public key PER signature (cfg4 "per-item rho"): rho n x 32 B, t1 n x k*320 B; A is expanded on chip per item */
int dil_verify_multi_host(dil_engine_t *e, int level, const uint8_t *rho, const uint8_t *t1_packed, const uint8_t *msgs,
                          const uint64_t *offsets, size_t n, const uint8_t *z, const uint8_t *h, const uint8_t *ctilde,
                          uint8_t *ok);
int dil_verify_multi_dev(dil_engine_t *e, int level, const uint8_t *d_rho, const uint8_t *d_t1_packed, const uint8_t *d_msgs,
                         const uint64_t *d_offsets, size_t n, const uint8_t *d_z, const uint8_t *d_h, const uint8_t *d_ctilde,
                         uint8_t *d_ok, void *stream);

/* ---- batched key generation (combined_top.v mode 0, FSM :754-1079; outputs as tb_keygen_top.v:180-275) ----
 * xi: n x 32-byte seeds.  Outputs per key, bit-packed exactly as the KAT files: rho, K, tr (32 B each),
 * s1 (l polys), s2 (k polys) as eta - s, t1 (10 bit), t0 as 2^12 - t0 (13 bit). */
int dil_keygen_batch_host(dil_engine_t *e, int level, const uint8_t *xi, size_t n, uint8_t *rho, uint8_t *key, uint8_t *tr,
                          uint8_t *s1_packed, uint8_t *s2_packed, uint8_t *t1_packed, uint8_t *t0_packed);
/* device pointers (d_rho, d_key, d_tr 8-byte aligned); enqueues on `stream` and returns */
int dil_keygen_batch_dev(dil_engine_t *e, int level, const uint8_t *d_xi, size_t n, uint8_t *d_rho, uint8_t *d_key, uint8_t *d_tr,
                         uint8_t *d_s1_packed, uint8_t *d_s2_packed, uint8_t *d_t1_packed, uint8_t *d_t0_packed, void *stream);

/* ---- several GPUs in one process (items are independent, shards need no exchange) ----
 * A pool owns one engine per device (devices == NULL or n_devices == 0: every visible device; a device may be listed
 * twice).  dil_pool_sign_batch_host splits the batch into contiguous shards, one per engine, and signs them concurrently
 * (one host thread per engine, each through dil_sign_batch_host: pinned output buffers - cudaHostAllocPortable /
 * cudaHostRegisterPortable - are streamed to round by round).  Results are identical to signing the whole batch on one
 * engine.  The only thing every GPU receives is the packed key (the path's single broadcast). */
typedef struct dil_pool dil_pool_t;
typedef struct dil_pool_sign_key dil_pool_sign_key_t;
int dil_pool_create(dil_pool_t **out, const int *devices, int n_devices);
int dil_pool_destroy(dil_pool_t *p);
int dil_pool_size(const dil_pool_t *p);
dil_engine_t *dil_pool_engine(dil_pool_t *p, int i);
int dil_pool_sign_key_create(dil_pool_t *p, dil_pool_sign_key_t **out, int level, const uint8_t *rho, const uint8_t *key,
                             const uint8_t *tr, const uint8_t *s1_packed, const uint8_t *s2_packed, const uint8_t *t0_packed);
int dil_pool_sign_key_destroy(dil_pool_t *p, dil_pool_sign_key_t *k);
int dil_pool_sign_batch_host(dil_pool_t *p, dil_pool_sign_key_t *k, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                             uint8_t *z, uint8_t *h, uint8_t *ctilde, uint32_t *attempts);
/* asynchronous pair over the pool: every engine begins its shard, dil_pool_sign_batch_finish completes them all (one batch per
   pool key at a time; same limits per shard as dil_sign_batch_host_begin: pinned device-addressable outputs, <= 1.25 x 2^18 per GPU) */
int dil_pool_sign_batch_host_begin(dil_pool_t *p, dil_pool_sign_key_t *k, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                   uint8_t *z, uint8_t *h, uint8_t *ctilde, uint32_t *attempts);
int dil_pool_sign_batch_finish(dil_pool_t *p, dil_pool_sign_key_t *k);

/* ---- diagnostics ----
 * Pure Keccak-f[1600] rate of the device: sm_count * ctas_per_sm CTAs of 128 threads, every thread runs perms_per_thread
 * permutations on its own state (the permutation code of every hash kernel of the engine) and writes one word to
 * d_out[thread].  The caller times the launch; bench.py uses it as the measured ALU-pipe roofline of the run. */
int dil_diag_keccak_dev(dil_engine_t *e, uint64_t *d_out, unsigned ctas_per_sm, unsigned perms_per_thread, void *stream);
/* Per-item-rho batches of at least min_batch items use the row-streamed kernel (8 items per CTA, A generated a few rows
 * at a time), smaller ones one CTA per item.  Default: never (the one-CTA-per-item kernel measured 5-10 % faster);
 * process-wide, atomic; results are identical - this exists for A/B measurements. */
int dil_diag_item_rows_threshold(size_t min_batch);

/* ---- north_star aliases (plain domain, identical to the above) ---- */
int dil_invntt_tomont_dev(dil_engine_t *e, int32_t *dst, const int32_t *src, size_t n_polys, void *stream);
int dil_poly_pointwise_dev(dil_engine_t *e, int32_t *c, const int32_t *a, const int32_t *b, size_t n_polys, void *stream);
int dil_polyvec_matrix_pointwise_dev(dil_engine_t *e, int32_t *w, const int32_t *a_hat, const int32_t *v,
                                     int k, int l, size_t batch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DILITHIUM_B200_H */
