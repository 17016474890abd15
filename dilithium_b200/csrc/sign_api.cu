// sign_api.cu — host orchestration of batched signing behind the C ABI.
// Replaces, for a batch, what rtl_src/combined_top.v does for one signature in mode 2 with the
// I/O order of rtl_tb/tb_sign_top.v:171-335 (inputs rho, mlen, tr, M, K, s1, s2, t0; outputs z,
// h, c~).  Key material is expanded once per key (ExpandA + NTT of s1, s2, t0: the LOAD_RHO /
// NTT_S1 / NTT_S2 / NTT_T0 states, combined_top.v:1560-1767); each round then runs
// ExpandMask -> fused sign core (+ w1 pack) -> challenge -> tail (-> resolve) over the active items.
// The rejection loop itself lives on the device (RoundCtl, kernels.h): the host enqueues rounds ahead
// of time and looks at the state once per burst of rounds.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "dil_params.h"
#include "dilithium_b200.h"
#include "engine_priv.h"
#include "kernels.h"

using dil::DeviceGuard;
using dil::LevelParams;

struct dil_sign_key {
    LevelParams P{};
    int device = -1;
    int32_t* a_hat = nullptr;    // k*l polys
    int32_t* key_hat = nullptr;  // s1_hat (l) | s2_hat (k) | t0_hat (k)
    int8_t* key_small = nullptr; // s1 (l) | s2 (k) in the time domain, coefficients in [-eta, eta] (sparse c*s products)
    uint8_t* seeds = nullptr;    // tr[32] | K[32] | rho[32]
    // workspace (grow-only, sized for `cap` items)
    std::mutex mu;
    size_t cap = 0;
    uint64_t *mu_d = nullptr, *rhop = nullptr, *w1p = nullptr;
    uint16_t* kappa = nullptr;
    uint32_t* active[2] = {nullptr, nullptr};
    dil::RoundCtl* ctl = nullptr;    // device-resident round state
    uint32_t* done_list = nullptr;   // finished items in completion order (host path: per-round drain)
    int32_t *y = nullptr, *w = nullptr;
    int8_t* c = nullptr;
    uint8_t *h_slot = nullptr, *accepted = nullptr;
    uint64_t* ct_slot = nullptr;
    size_t slots = 0;            // slot capacity of y/w/c/w1p/h_slot/ct_slot/accepted
    // the host path's own streams: two key handles (even of the same key) can have a batch in flight each, so that one
    // batch's last signatures drain while the next batch already signs (bench.py's e2e loop, examples/pool_sign.cpp)
    cudaStream_t st_own = nullptr, cs_own = nullptr;
    uint32_t* ctl_host = nullptr;      // mapped pinned copy of the first 16 words of the round state
    uint32_t* ctl_host_dev = nullptr;  // its device alias
    cudaEvent_t round_ev = nullptr;    // orders the copy stream's drains after the rounds they copy
    cudaEvent_t block_ev = nullptr;    // blocking-sync event: the host thread sleeps instead of spinning when other batches are in flight
    // host-variant staging
    uint8_t *msgs_d = nullptr, *zp_d = nullptr, *h_d = nullptr, *ct_d = nullptr;
    uint64_t* off_d = nullptr;
    uint32_t* att_d = nullptr;
    size_t msgs_cap = 0, out_cap = 0;
    uint32_t last_rounds = 0;
    uint64_t last_slots = 0;
    void* pend = nullptr;        // Pending*: the batch between *_begin and dil_sign_batch_finish (one per handle)
    dil_sign_tuning tune{};      // zero = defaults
    // one key per signature (dil_sign_multi_*): per-item key material and the unfused core's intermediates
    int32_t *a_items = nullptr, *key_items = nullptr, *yh = nullptr, *wh = nullptr;
    size_t multi_cap = 0, multi_slots = 0;
    uint8_t* multi_in = nullptr;   // host variant: staged rho | K | tr | s1 | s2 | t0 records
    size_t multi_in_cap = 0;
    // optional per-kernel-class device timing (CUDA events on the launching stream)
    bool profile = false;
    cudaEvent_t ev[16] = {};
    double prof_ms[8] = {};
    uint64_t prof_slots[8] = {};
};

namespace {

int fail(dil_engine* e, cudaError_t err, const char* what) {
    std::lock_guard<std::mutex> g(e->err_mu);
    e->last_error = std::string(what) + ": " + cudaGetErrorString(err);
    return DIL_ERR_CUDA;
}
int fail_msg(dil_engine* e, int status, const std::string& msg) {
    std::lock_guard<std::mutex> g(e->err_mu);
    e->last_error = msg;
    return status;
}
#define CK(expr)                                                   \
    do {                                                           \
        cudaError_t err_ = (expr);                                 \
        if (err_ != cudaSuccess) return fail(e, err_, #expr);      \
    } while (0)

// little-endian fixed-width bit stream -> values (decoder.v:89-143)
void bits_get(uint32_t* v, const uint8_t* in, int n, int width) {
    size_t bit = 0;
    for (int i = 0; i < n; i++) {
        uint32_t x = 0;
        for (int b = 0; b < width; b++, bit++) x |= (uint32_t)((in[bit >> 3] >> (bit & 7)) & 1u) << b;
        v[i] = x;
    }
}

template <class T>
cudaError_t dmalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

// secrets (s1, s2, t0, K, rho', y) never outlive their buffers: wipe, then free
void wipe_free(void* p, size_t bytes) {
    if (!p) return;
    cudaMemset(p, 0, bytes);
    cudaFree(p);
}
void wipe_host(void* p, size_t bytes) {
    volatile unsigned char* q = static_cast<volatile unsigned char*>(p);
    for (size_t i = 0; i < bytes; i++) q[i] = 0;
}

// message offsets come from the caller: offsets[0] == 0 and non-decreasing, otherwise a kernel would read out of bounds
bool offsets_ok(const uint64_t* off, size_t n) {
    if (off[0] != 0) return false;
    for (size_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return false;
    return true;
}

// Straggler speculation: once fewer than spec_target items are still active the GPU is no longer saturated by
// distinct items, so each remaining item tries up to spec_max consecutive kappa values per round in parallel
// slots; the smallest accepted kappa wins, which is exactly the sequential result of combined_top.v:2217-2228
// (restart with the next kappa).  The policy itself runs on the device (spec_policy, sign_kernels.cu).
constexpr uint32_t SPEC_TARGET_DEFAULT = 32768, SPEC_MAX_DEFAULT = 32;
struct Spec {
    uint32_t target, max;
};
// Without explicit tuning the policy follows the load of the engine: when other sign batches are in flight on the same GPU
// (two or more key handles signing from different host threads, the streaming use of the API), their kernels fill the
// SMs that the small late rounds of this batch leave idle, so speculating costs more than it hides - fewer slots per
// item then, and later.  Measured on one B200, 65 536-message Dilithium-2 batches (profiles/r2e_concurrent_spec.txt):
// 1 in flight 12.1 M signs/s with 32768 / 32; 2 in flight 13.9 M with 16384 / 16; 4 in flight 14.6 M with 8192 / 8.
Spec spec_for(const dil_sign_key* k, int in_flight) {
    Spec s{SPEC_TARGET_DEFAULT, SPEC_MAX_DEFAULT};
    if (in_flight == 2) s = Spec{16384, 16};
    else if (in_flight >= 3) s = Spec{8192, 8};
    if (k->tune.spec_target) s.target = k->tune.spec_target;
    if (k->tune.spec_max && k->tune.spec_max <= 32) s.max = k->tune.spec_max;   // one warp lane per slot in resolve
    return s;
}
size_t slots_for(Spec s, size_t n) {
    const size_t T = s.target, M = s.max;
    return n > T ? n : (n * M < T ? n * M : T);
}
// the workspace is sized for the policy that needs the most slots (a lone batch)
size_t slots_alloc(const dil_sign_key* k, size_t n) { return slots_for(spec_for(k, 1), n); }

// Waiting for a stream: a lone batch spins (lowest latency); with other batches in flight on the engine the host thread
// sleeps on a blocking-sync event instead, so that T threads per GPU do not burn T cores while their batches sign.
cudaError_t wait_stream(dil_sign_key* k, cudaStream_t s, bool loaded) {
    if (loaded && k->block_ev) {
        cudaError_t er = cudaEventRecord(k->block_ev, s);
        return er != cudaSuccess ? er : cudaEventSynchronize(k->block_ev);
    }
    return cudaStreamSynchronize(s);
}

// sign batches currently inside sign_rounds on this engine (any key handle, any host thread)
struct InFlight {
    std::atomic<int>& c;
    int seen;
    explicit InFlight(std::atomic<int>& ctr) : c(ctr), seen(++ctr) {}
    ~InFlight() { --c; }
};

void free_ws(dil_sign_key* k) {
    const LevelParams& P = k->P;
    wipe_free(k->rhop, k->cap * 64);
    wipe_free(k->y, k->slots * (size_t)P.l * 1024);
    void* ptrs[] = {k->mu_d, k->w1p, k->kappa, k->active[0], k->active[1], k->ctl, k->w, k->c, k->h_slot, k->accepted, k->ct_slot, k->done_list};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    k->mu_d = k->rhop = k->w1p = k->ct_slot = nullptr;
    k->kappa = nullptr;
    k->active[0] = k->active[1] = k->done_list = nullptr;
    k->ctl = nullptr;
    k->y = k->w = nullptr;
    k->c = nullptr;
    k->h_slot = k->accepted = nullptr;
    k->cap = k->slots = 0;
}

int ensure_ws(dil_engine* e, dil_sign_key* k, size_t n) {
    const size_t slots = slots_alloc(k, n);
    if (n <= k->cap && slots <= k->slots) return DIL_OK;
    free_ws(k);
    const LevelParams& P = k->P;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    // slots: every item of a round owns `spec` speculative attempt slots (spec = 1 in the big rounds)
    A(dmalloc(&k->mu_d, n * 8));
    A(dmalloc(&k->rhop, n * 8));
    A(dmalloc(&k->kappa, n));
    A(dmalloc(&k->active[0], n));
    A(dmalloc(&k->active[1], n));
    A(dmalloc(&k->ctl, 1));
    A(dmalloc(&k->done_list, n));
    A(dmalloc(&k->w1p, slots * (size_t)(P.k * P.w1_bytes / 8)));
    A(dmalloc(&k->y, slots * (size_t)P.l * 256));
    A(dmalloc(&k->w, slots * (size_t)P.k * 256));
    A(dmalloc(&k->c, slots * 256));
    A(dmalloc(&k->h_slot, slots * (size_t)(P.omega + P.k)));
    A(dmalloc(&k->accepted, slots));
    A(dmalloc(&k->ct_slot, slots * 4));
    if (err == cudaSuccess && !k->ctl_host) {
        A(cudaHostAlloc(reinterpret_cast<void**>(&k->ctl_host), 64, cudaHostAllocMapped));
        if (err == cudaSuccess) A(cudaHostGetDevicePointer(reinterpret_cast<void**>(&k->ctl_host_dev), k->ctl_host, 0));
    }
    if (err == cudaSuccess && !k->round_ev) A(cudaEventCreateWithFlags(&k->round_ev, cudaEventDisableTiming));
    if (err == cudaSuccess && !k->block_ev) A(cudaEventCreateWithFlags(&k->block_ev, cudaEventDisableTiming | cudaEventBlockingSync));
    if (err != cudaSuccess) {
        free_ws(k);
        return fail_msg(e, DIL_ERR_ALLOC, std::string("sign workspace: ") + cudaGetErrorString(err));
    }
    k->cap = n;
    k->slots = slots;
    return DIL_OK;
}

// Host-path streaming: device aliases of the caller's pinned (mapped) output buffers for the batch being
// signed.  After every rejection round the signatures that round finished are copied out by drain_kernel
// on `stream`, concurrently with the following rounds, so that only the last round's few signatures are
// still on the device when signing ends.
struct DrainTarget {
    uint8_t *z = nullptr, *h = nullptr, *ct = nullptr;
    uint32_t* att = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t idle = nullptr;   // recorded on `stream` after the last drain of a batch
};

// Expected number of active items at the start of every round of a batch of n items (the device decides what really
// happens; this only sizes the first burst of launches and their grids): an attempt is rejected with probability
// 1 - 1/repetitions, the scheme's expected repetitions being 4.25 / 5.1 / 3.85 at levels 2 / 3 / 5.
std::vector<double> expected_trajectory(const dil_sign_key* k, Spec sp, size_t n) {
    const double q = k->P.level == 2 ? 0.765 : (k->P.level == 3 ? 0.804 : 0.74);
    const size_t T = sp.target, M = sp.max, cap = slots_for(sp, n);
    std::vector<double> tr;
    double rem = (double)n;
    while (rem >= 0.3 && tr.size() < 200) {
        tr.push_back(rem);
        size_t spec = 1;
        if (rem < (double)T) {
            const double room = (double)(cap < T ? cap : T) / (rem < 1 ? 1 : rem);
            spec = room < 1 ? 1 : (room > (double)M ? M : (size_t)room);
        }
        rem *= std::pow(q, (double)spec);
    }
    return tr;
}

// one key per signature: per-item tr / K (32-byte records) instead of the key handle's seeds
struct MultiKeys {
    const uint8_t *tr, *key;
};

// ---- the round loop of a batch under one key, split in two so that a caller can keep several batches in flight from ONE
// host thread: rounds_begin enqueues the expected number of rounds (the loop itself runs on the device) and returns;
// rounds_finish waits, looks at the round state and enqueues more rounds until nothing is left.
struct Pending {
    dil::SignBufs b{};
    Spec sp{};
    size_t n = 0;
    uint32_t cap_slots = 0, enq = 0, seen = 0, seen_at = 0;
    std::vector<double> traj;
    cudaStream_t st = nullptr;
    bool has_drain = false, host = false, loaded = false;
    DrainTarget drain{};
    uint64_t launches = 0;
};

dil::SignBufs sign_bufs(const dil_sign_key* k, uint8_t* d_zp, uint8_t* d_h, uint8_t* d_ct, uint32_t* d_att, bool track_done) {
    dil::SignBufs b{};
    b.ctl = k->ctl; b.active[0] = k->active[0]; b.active[1] = k->active[1]; b.done_list = k->done_list;
    b.mu = k->mu_d; b.rhop = k->rhop; b.kappa = k->kappa; b.y = k->y; b.w = k->w; b.w1p = k->w1p; b.c = k->c;
    b.ct_slot = k->ct_slot; b.h_slot = k->h_slot; b.accepted = k->accepted;
    b.zp = d_zp; b.h_out = d_h; b.ct_out = reinterpret_cast<uint64_t*>(d_ct); b.attempts = d_att;
    b.track_done = track_done;
    return b;
}

// enqueue `burst` more rounds of the pending batch (grid sizes are hints derived from the expected trajectory), then
// the publication of the round state into mapped host memory
int rounds_enqueue(dil_engine* e, dil_sign_key* k, Pending& p, int burst) {
    const LevelParams& P = k->P;
    const uint32_t T = p.sp.target, M = p.sp.max;
    cudaStream_t st = p.st;
    for (int i = 0; i < burst; i++, p.enq++) {
        uint32_t ci = p.seen;   // hint: items of this round
        if (p.enq > p.seen_at) {
            const size_t ti = p.enq < p.traj.size() ? p.enq : p.traj.size() - 1, ts = p.seen_at < p.traj.size() ? p.seen_at : p.traj.size() - 1;
            const double ratio = p.traj[ts] > 0 ? p.traj[ti] / p.traj[ts] : 1.0;
            const double guess = (double)p.seen * ratio * 1.25 + 512.0;
            if (guess < (double)p.seen) ci = (uint32_t)guess;
        }
        const uint64_t cs64 = ci >= T ? ci : ((uint64_t)ci * M < p.cap_slots ? (uint64_t)ci * M : p.cap_slots);
        const uint32_t cs = (uint32_t)(cs64 < ci ? ci : cs64);   // hint: slots of this round
        if (k->tune.fused_mask) {
            CK(dil::launch_mask_core(P.level, p.b, k->a_hat, cs, e->sm_count, st, k->tune.mask_producers));
            p.launches += 1;
        } else {
            CK(dil::launch_expand_mask(P.level, p.b, cs, st));
            CK(dil::launch_signcore(k->w, k->a_hat, k->y, P.k, P.l, cs, e->sm_count, st, &k->ctl->ctr_core,
                                    reinterpret_cast<uint8_t*>(k->w1p), &k->ctl->n_slots));
            p.launches += 2;
        }
        CK(dil::launch_challenge(P.level, p.b, cs, st));
        CK(dil::launch_sign_tail(P.level, p.b, k->key_hat, k->key_small, cs, e->sm_count, st));
        // only a first round of >= spec_target items is known to run without speculation; later rounds decide on the device
        if (!(p.enq == 0 && p.n >= T)) { CK(dil::launch_resolve(P.level, p.b, ci < T ? ci : T, st)); p.launches++; }
        CK(dil::launch_plan(k->ctl, p.enq, st));
        p.launches += 3;
        if (p.has_drain) {
            CK(cudaEventRecord(k->round_ev, st));
            CK(cudaStreamWaitEvent(p.drain.stream, k->round_ev, 0));
            CK(dil::launch_drain(p.drain.z, p.drain.h, p.drain.ct, p.drain.att, p.b, p.enq, (uint32_t)(P.l * P.z_bytes),
                                 (uint32_t)(P.omega + P.k), p.drain.stream));
            p.launches++;
        }
    }
    CK(dil::launch_publish_ctl(k->ctl_host_dev, k->ctl, st));
    return DIL_OK;
}

void pending_drop(dil_engine* e, dil_sign_key* k) {
    if (!k->pend) return;
    --e->sign_in_flight;
    delete static_cast<Pending*>(k->pend);
    k->pend = nullptr;
}

int rounds_begin(dil_engine* e, dil_sign_key* k, const uint8_t* d_msgs, const uint64_t* d_off, size_t n, uint8_t* d_zp, uint8_t* d_h,
                 uint8_t* d_ct, uint32_t* d_att, cudaStream_t st, const DrainTarget* drain) {
    if (k->pend) return fail_msg(e, DIL_ERR_ARG, "sign: this key handle already has a batch in flight (dil_sign_batch_finish it first)");
    int rc = ensure_ws(e, k, n);
    if (rc) return rc;
    // the previous batch's last drains still read the done list and the round state: let them finish first
    if (drain) CK(cudaStreamWaitEvent(st, drain->idle, 0));
    Pending* p = new (std::nothrow) Pending();
    if (!p) return DIL_ERR_ALLOC;
    k->pend = p;
    const int load = ++e->sign_in_flight;
    p->loaded = load > 1;
    p->sp = spec_for(k, load);
    p->n = n;
    p->st = st;
    p->cap_slots = (uint32_t)slots_for(p->sp, n);
    p->b = sign_bufs(k, d_zp, d_h, d_ct, d_att, drain != nullptr);
    p->has_drain = drain != nullptr;
    if (drain) p->drain = *drain;
    p->seen = (uint32_t)n;
    p->traj = expected_trajectory(k, p->sp, n);
    auto bail = [&](int code) { pending_drop(e, k); return code; };
    cudaError_t er = dil::launch_sign_begin(p->b, (uint32_t)n, p->cap_slots, p->sp.target, p->sp.max, st);
    if (er == cudaSuccess) er = dil::launch_sign_init(k->mu_d, k->rhop, k->kappa, k->seeds, k->seeds + 32, 0, d_msgs, d_off, (uint32_t)n, st);
    if (er != cudaSuccess) return bail(fail(e, er, "sign: first launches"));
    p->launches = 2;
    rc = rounds_enqueue(e, k, *p, (int)p->traj.size() + 1);
    if (rc) return bail(rc);
    return DIL_OK;
}

// may_sleep: the synchronous calls (one host thread per batch in flight) sleep on a blocking event under load; the asynchronous
// pair is driven by one thread for many batches, which should see completions as early as possible, so it spins
int rounds_finish(dil_engine* e, dil_sign_key* k, bool may_sleep) {
    Pending* p = static_cast<Pending*>(k->pend);
    if (!p) return fail_msg(e, DIL_ERR_ARG, "sign: no batch in flight on this key handle");
    volatile uint32_t* hc = k->ctl_host;
    int rc = DIL_OK;
    for (;;) {
        cudaError_t er = wait_stream(k, p->st, may_sleep && (p->loaded || e->sign_in_flight.load() > 1));
        if (er != cudaSuccess) { rc = fail(e, er, "sign: waiting for the rounds"); break; }
        p->seen = hc[0];
        p->seen_at = p->enq;
        if (p->seen == 0) break;
        if (p->enq > 4000) { rc = fail_msg(e, DIL_ERR_CUDA, "sign: rejection loop did not terminate"); break; }
        rc = rounds_enqueue(e, k, *p, 2);
        if (rc) break;
    }
    if (rc == DIL_OK && p->has_drain) {
        cudaError_t er = cudaEventRecord(p->drain.idle, p->drain.stream);
        if (er != cudaSuccess) rc = fail(e, er, "sign drain event");
    }
    e->launches += p->launches;
    if (rc == DIL_OK) {
        k->last_rounds = hc[8];
        k->last_slots = hc[9];
    }
    pending_drop(e, k);
    return rc;
}

// the round loop, one call (profiling and one-key-per-signature batches observe every round); all pointers are device pointers
int sign_rounds(dil_engine* e, dil_sign_key* k, const uint8_t* d_msgs, const uint64_t* d_off, size_t n, uint8_t* d_zp,
                uint8_t* d_h, uint8_t* d_ct, uint32_t* d_att, cudaStream_t st, const DrainTarget* drain = nullptr,
                const MultiKeys* mk = nullptr) {
    const LevelParams& P = k->P;
    if (!k->profile && mk == nullptr) {   // the ordinary batch: begin + finish
        int rcb = rounds_begin(e, k, d_msgs, d_off, n, d_zp, d_h, d_ct, d_att, st, drain);
        return rcb ? rcb : rounds_finish(e, k, true);
    }
    if (k->pend) return fail_msg(e, DIL_ERR_ARG, "sign: this key handle already has a batch in flight (dil_sign_batch_finish it first)");
    int rc = ensure_ws(e, k, n);
    if (rc) return rc;
    // the previous batch's last drains still read the done list and the round state: let them finish first
    if (drain) CK(cudaStreamWaitEvent(st, drain->idle, 0));
    uint64_t launches = 0;
    const bool prof = k->profile;
    if (prof) {
        for (int i = 0; i < 8; i++) k->prof_ms[i] = 0, k->prof_slots[i] = 0;
        for (int i = 0; i < 16; i++)
            if (!k->ev[i]) CK(cudaEventCreate(&k->ev[i]));
    }
#define PROF_BEGIN(cls) do { if (prof) CK(cudaEventRecord(k->ev[2 * (cls)], st)); } while (0)
#define PROF_END(cls) do { if (prof) CK(cudaEventRecord(k->ev[2 * (cls) + 1], st)); } while (0)
    const dil::SignBufs b = sign_bufs(k, d_zp, d_h, d_ct, d_att, drain != nullptr);
    InFlight load(e->sign_in_flight);
    const Spec sp = spec_for(k, load.seen);
    const uint32_t cap_slots = (uint32_t)slots_for(sp, n);
    CK(dil::launch_sign_begin(b, (uint32_t)n, cap_slots, sp.target, sp.max, st));
    PROF_BEGIN(0);
    if (mk) CK(dil::launch_sign_init(k->mu_d, k->rhop, k->kappa, mk->tr, mk->key, 32, d_msgs, d_off, (uint32_t)n, st));
    else CK(dil::launch_sign_init(k->mu_d, k->rhop, k->kappa, k->seeds, k->seeds + 32, 0, d_msgs, d_off, (uint32_t)n, st));
    PROF_END(0);
    launches += 2;
    if (prof) {
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, k->ev[0], k->ev[1]));
        k->prof_ms[0] += ms;
        k->prof_slots[0] += n;
    }
    // This variant observes EVERY round (profiling reads its events, the one-key-per-signature core sizes its stand-alone
    // kernels on the host); the ordinary batch (rounds_begin / rounds_finish above) enqueues its rounds in bursts without
    // waiting for their outcome.  Grid sizes are hints (every kernel strides or claims work dynamically, so any grid is
    // correct): the expected number of active items plus a margin, never more than the last count the host has seen.
    const std::vector<double> traj = expected_trajectory(k, sp, n);
    const uint32_t T = sp.target, M = sp.max;
    uint32_t enq = 0, seen = (uint32_t)n, seen_at = 0;   // last observed item count and the round it belongs to
    volatile uint32_t* hc = k->ctl_host;
    // per-item keys: the unfused core's stand-alone kernels take their sizes from the host, so every round is observed
    const bool step = prof || mk != nullptr;
    int burst = step ? 1 : (int)traj.size() + 1;
    for (;;) {
        for (int i = 0; i < burst; i++, enq++) {
            uint32_t ci = seen;   // hint: items of this round
            if (enq > seen_at) {
                const size_t ti = enq < traj.size() ? enq : traj.size() - 1, ts = seen_at < traj.size() ? seen_at : traj.size() - 1;
                const double ratio = traj[ts] > 0 ? traj[ti] / traj[ts] : 1.0;
                const double guess = (double)seen * ratio * 1.25 + 512.0;
                if (guess < (double)seen) ci = (uint32_t)guess;
            }
            const uint64_t cs64 = ci >= T ? ci : ((uint64_t)ci * M < cap_slots ? (uint64_t)ci * M : cap_slots);
            const uint32_t cs = (uint32_t)(cs64 < ci ? ci : cs64);   // hint: slots of this round
            if (mk) {
                // exact size of this round (the host saw the state after the previous one)
                const uint32_t sp = seen >= T ? 1u : (uint32_t)std::min<uint64_t>(std::max<uint64_t>(std::min<uint64_t>(cap_slots, T) / seen, 1), M);
                const size_t ns = (size_t)seen * sp;
                PROF_BEGIN(1);
                CK(dil::launch_expand_mask(P.level, b, (uint32_t)ns, st));
                PROF_END(1);
                PROF_BEGIN(2);
                CK(dil::launch_ntt_fwd(k->yh, k->y, ns * P.l, e->sm_count, st));
                CK(dil::launch_matvec_multi(P.level, k->wh, k->a_items, k->yh, b, (uint32_t)ns, st));
                CK(dil::launch_ntt_inv(k->w, k->wh, ns * P.k, e->sm_count, st));
                CK(dil::launch_pack_w1(P.level, reinterpret_cast<uint32_t*>(k->w1p), k->w, ns, st));
                PROF_END(2);
                launches += 3;
            } else if (k->tune.fused_mask) {
                // ExpandMask, the transforms and the mat-vec in one kernel (mask_core.cu); timed as class 2
                PROF_BEGIN(1);
                PROF_END(1);
                PROF_BEGIN(2);
                CK(dil::launch_mask_core(P.level, b, k->a_hat, cs, e->sm_count, st, k->tune.mask_producers));
                PROF_END(2);
                launches--;
            } else {
                PROF_BEGIN(1);
                CK(dil::launch_expand_mask(P.level, b, cs, st));
                PROF_END(1);
                PROF_BEGIN(2);
                // the core also emits the packed w1 = HighBits(w)
                CK(dil::launch_signcore(k->w, k->a_hat, k->y, P.k, P.l, cs, e->sm_count, st, &k->ctl->ctr_core,
                                        reinterpret_cast<uint8_t*>(k->w1p), &k->ctl->n_slots));
                PROF_END(2);
            }
            PROF_BEGIN(4);
            CK(dil::launch_challenge(P.level, b, cs, st));
            PROF_END(4);
            // rounds with one slot per item: the tail finishes or re-queues its item itself and resolve returns at once
            PROF_BEGIN(5);
            if (mk) CK(dil::launch_sign_tail_multi(P.level, b, k->key_items, cs, e->sm_count, st));
            else CK(dil::launch_sign_tail(P.level, b, k->key_hat, k->key_small, cs, e->sm_count, st));
            PROF_END(5);
            PROF_BEGIN(6);
            // only a first round of >= spec_target items is known to run without speculation; later rounds decide on the device
            if (!(enq == 0 && n >= T)) { CK(dil::launch_resolve(P.level, b, ci < T ? ci : T, st)); launches++; }
            PROF_END(6);
            CK(dil::launch_plan(k->ctl, enq, st));
            launches += 5;
            if (drain) {
                CK(cudaEventRecord(k->round_ev, st));
                CK(cudaStreamWaitEvent(drain->stream, k->round_ev, 0));
                CK(dil::launch_drain(drain->z, drain->h, drain->ct, drain->att, b, enq, (uint32_t)(P.l * P.z_bytes),
                                     (uint32_t)(P.omega + P.k), drain->stream));
                launches++;
            }
            if (step && !prof) {
                CK(dil::launch_publish_ctl(k->ctl_host_dev, k->ctl, st));
                CK(cudaStreamSynchronize(st));
                seen = hc[0];
                seen_at = enq + 1;
            }
            if (prof) {
                // per-class device time of this round; the slots it processed = growth of total_slots
                CK(dil::launch_publish_ctl(k->ctl_host_dev, k->ctl, st));
                CK(cudaStreamSynchronize(st));
                const uint64_t total = hc[9];
                for (int cls = 1; cls <= 6; cls++) {
                    if (cls == 3) continue;   // w1 packing is part of the sign core (class 2)
                    float ms = 0;
                    CK(cudaEventElapsedTime(&ms, k->ev[2 * cls], k->ev[2 * cls + 1]));
                    k->prof_ms[cls] += ms;
                    k->prof_slots[cls] = total;
                }
                seen = hc[0];
                seen_at = enq + 1;
            }
        }
        if (!step) {
            CK(dil::launch_publish_ctl(k->ctl_host_dev, k->ctl, st));
            CK(wait_stream(k, st, load.seen > 1));
            seen = hc[0];
            seen_at = enq;
        }
        if (seen == 0) break;
        if (enq > 4000) return fail_msg(e, DIL_ERR_CUDA, "sign: rejection loop did not terminate");
        burst = step ? 1 : 2;
    }
    if (drain) CK(cudaEventRecord(drain->idle, drain->stream));
    e->launches += launches;
    k->last_rounds = hc[8];
    k->last_slots = hc[9];
    return DIL_OK;
}

}  // namespace

extern "C" {

int dil_sign_sizes(int level, size_t* z_bytes, size_t* h_bytes) {
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    LevelParams P = dil::level_params(level);
    if (z_bytes) *z_bytes = (size_t)P.l * P.z_bytes;
    if (h_bytes) *h_bytes = (size_t)P.omega + P.k;
    return DIL_OK;
}

int dil_sign_key_create(dil_engine_t* e, dil_sign_key_t** out, int level, const uint8_t* rho, const uint8_t* key,
                        const uint8_t* tr, const uint8_t* s1p, const uint8_t* s2p, const uint8_t* t0p) {
    if (!e || !out || !rho || !key || !tr || !s1p || !s2p || !t0p) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    dil_sign_key* k = new (std::nothrow) dil_sign_key();
    if (!k) return DIL_ERR_ALLOC;
    k->P = dil::level_params(level);
    k->device = e->device;
    const LevelParams& P = k->P;
    const int nkey = P.l + 2 * P.k;
    // unpack s1, s2 (eta - x, 3 or 4 bits) and t0 (2^12 - x, 13 bits): decoder.v:89-143
    std::vector<int32_t> polys((size_t)nkey * 256);
    std::vector<uint32_t> tmp(256);
    const int sw = P.eta == 2 ? 3 : 4;
    for (int p = 0; p < P.l + P.k; p++) {
        const uint8_t* src = p < P.l ? s1p + (size_t)p * P.s_bytes : s2p + (size_t)(p - P.l) * P.s_bytes;
        bits_get(tmp.data(), src, 256, sw);
        for (int i = 0; i < 256; i++) polys[(size_t)p * 256 + i] = P.eta - (int32_t)tmp[i];
    }
    for (int p = 0; p < P.k; p++) {
        bits_get(tmp.data(), t0p + (size_t)p * 416, 256, 13);
        for (int i = 0; i < 256; i++) polys[(size_t)(P.l + P.k + p) * 256 + i] = (1 << 12) - (int32_t)tmp[i];
    }
    cudaStream_t st = e->host_stream;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    uint8_t seeds[96];
    std::memcpy(seeds, tr, 32);
    std::memcpy(seeds + 32, key, 32);
    std::memcpy(seeds + 64, rho, 32);
    A(dmalloc(&k->a_hat, (size_t)P.k * P.l * 256));
    A(dmalloc(&k->key_hat, (size_t)nkey * 256));
    A(dmalloc(&k->key_small, (size_t)(P.l + P.k) * 256));
    A(dmalloc(&k->seeds, 96));
    std::vector<int8_t> small((size_t)(P.l + P.k) * 256);
    for (size_t i = 0; i < small.size(); i++) small[i] = (int8_t)polys[i];
    if (err == cudaSuccess) {
        A(cudaMemcpyAsync(k->seeds, seeds, 96, cudaMemcpyHostToDevice, st));
        A(cudaMemcpyAsync(k->key_hat, polys.data(), polys.size() * 4, cudaMemcpyHostToDevice, st));
        A(cudaMemcpyAsync(k->key_small, small.data(), small.size(), cudaMemcpyHostToDevice, st));
        A(dil::launch_expand_a(k->a_hat, k->seeds + 64, 1, P.k, P.l, e->sm_count, st));
        A(dil::launch_ntt_fwd(k->key_hat, k->key_hat, nkey, e->sm_count, st));
        A(cudaStreamSynchronize(st));
    }
    // host copies of the secret key leave no residue
    wipe_host(seeds, sizeof seeds);
    wipe_host(polys.data(), polys.size() * 4);
    wipe_host(small.data(), small.size());
    wipe_host(tmp.data(), tmp.size() * 4);
    if (err != cudaSuccess) {
        if (k->a_hat) cudaFree(k->a_hat);
        wipe_free(k->key_hat, (size_t)nkey * 1024);
        wipe_free(k->key_small, (size_t)(P.l + P.k) * 256);
        wipe_free(k->seeds, 96);
        delete k;
        return fail(e, err, "dil_sign_key_create");
    }
    e->launches += 2;
    *out = k;
    return DIL_OK;
}

int dil_sign_key_destroy(dil_engine_t* e, dil_sign_key_t* k) {
    if (!k) return DIL_OK;
    DeviceGuard dg(k->device);
    if (k->pend && e) dil_sign_batch_finish(e, k);   // a batch begun and never finished: let it run out before its buffers go
    const LevelParams& P = k->P;
    free_ws(k);
    wipe_free(k->key_hat, (size_t)(P.l + 2 * P.k) * 1024);
    wipe_free(k->key_small, (size_t)(P.l + P.k) * 256);
    wipe_free(k->seeds, 96);
    void* ptrs[] = {k->a_hat, k->msgs_d, k->zp_d, k->h_d, k->ct_d, k->off_d, k->att_d};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& ev : k->ev)
        if (ev) cudaEventDestroy(ev);
    if (k->round_ev) cudaEventDestroy(k->round_ev);
    if (k->block_ev) cudaEventDestroy(k->block_ev);
    if (k->st_own) cudaStreamDestroy(k->st_own);
    if (k->cs_own) cudaStreamDestroy(k->cs_own);
    if (k->ctl_host) cudaFreeHost(k->ctl_host);
    (void)e;
    delete k;
    return DIL_OK;
}

uint32_t dil_sign_last_rounds(const dil_sign_key_t* k) { return k ? k->last_rounds : 0; }
uint64_t dil_sign_last_slots(const dil_sign_key_t* k) { return k ? k->last_slots : 0; }

int dil_sign_key_set_tuning(dil_sign_key_t* k, const dil_sign_tuning* t) {
    if (!k) return DIL_ERR_ARG;
    if (t && t->spec_max > 32) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    if (k->pend) return DIL_ERR_ARG;   // not between *_begin and dil_sign_batch_finish
    k->tune = t ? *t : dil_sign_tuning{};
    return DIL_OK;
}

int dil_sign_set_profile(dil_sign_key_t* k, int on) {
    if (!k) return DIL_ERR_ARG;
    k->profile = on != 0;
    return DIL_OK;
}
int dil_sign_get_profile(const dil_sign_key_t* k, double* ms, uint64_t* units) {
    if (!k || !ms || !units) return DIL_ERR_ARG;
    for (int i = 0; i < 8; i++) ms[i] = k->prof_ms[i], units[i] = k->prof_slots[i];
    return DIL_OK;
}

int dil_sign_batch_dev(dil_engine_t* e, dil_sign_key_t* k, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                       uint8_t* d_z, uint8_t* d_h, uint8_t* d_ctilde, uint32_t* d_attempts, void* stream) {
    if (!e || !k) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!d_msgs || !d_offsets || !d_z || !d_h || !d_ctilde || !d_attempts || n > 0x07FFFFFFu) return DIL_ERR_ARG;
    // z leaves the device as 16-byte vectors, c~ as 64-bit words
    if ((reinterpret_cast<uintptr_t>(d_ctilde) & 7u) || (reinterpret_cast<uintptr_t>(d_z) & 15u)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    if (k->pend) return fail_msg(e, DIL_ERR_ARG, "sign: this key handle already has a batch in flight (dil_sign_batch_finish it first)");
    // very large batches are signed in 2^20-message pieces so that the per-attempt workspace (about
    // 9.5 / 13 / 17 KB per message at levels 2 / 3 / 5) stays bounded
    const size_t chunk = k->tune.dev_chunk ? k->tune.dev_chunk : ((size_t)1 << 20);
    const size_t zb = (size_t)k->P.l * k->P.z_bytes, hb = (size_t)k->P.omega + k->P.k;
    for (size_t lo = 0; lo < n;) {
        const size_t m = n - lo <= chunk + chunk / 4 ? n - lo : chunk;
        int rc = sign_rounds(e, k, d_msgs, d_offsets + lo, m, d_z + lo * zb, d_h + lo * hb, d_ctilde + lo * 32, d_attempts + lo,
                             (cudaStream_t)stream);
        if (rc) return rc;
        lo += m;
    }
    return DIL_OK;
}

}  // extern "C"

namespace {

// host path, common part: the handle's own streams, device staging for the messages and the packed outputs, H2D of the inputs
int host_stage(dil_engine* e, dil_sign_key* k, const uint8_t* msgs, const uint64_t* offsets, size_t n) {
    const LevelParams& P = k->P;
    if (!k->st_own) {
        CK(cudaStreamCreateWithFlags(&k->st_own, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&k->cs_own, cudaStreamNonBlocking));
    }
    const size_t mbytes = offsets[n] > 0 ? offsets[n] : 1;
    const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k;
    if (mbytes > k->msgs_cap) {
        if (k->msgs_d) cudaFree(k->msgs_d);
        k->msgs_d = nullptr;
        k->msgs_cap = 0;
        CK(dmalloc(&k->msgs_d, mbytes));
        k->msgs_cap = mbytes;
    }
    if (n > k->out_cap) {
        void* ptrs[] = {k->zp_d, k->h_d, k->ct_d, k->off_d, k->att_d};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        k->zp_d = k->h_d = k->ct_d = nullptr;
        k->off_d = nullptr;
        k->att_d = nullptr;
        k->out_cap = 0;
        CK(dmalloc(&k->zp_d, n * zb));
        CK(dmalloc(&k->h_d, n * hb));
        CK(dmalloc(&k->ct_d, n * 32));
        CK(dmalloc(&k->off_d, n + 1));
        CK(dmalloc(&k->att_d, n));
        k->out_cap = n;
    }
    CK(cudaMemcpyAsync(k->msgs_d, msgs, offsets[n], cudaMemcpyHostToDevice, k->st_own));
    CK(cudaMemcpyAsync(k->off_d, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, k->st_own));
    return DIL_OK;
}

// Streaming path: when every output buffer is pinned host memory the device can address (cudaHostAlloc /
// cudaHostRegister; torch's pinned tensors are), finished signatures leave round by round (DrainTarget)
// and the transfer hides behind the remaining rounds.  Returns false for pageable (or misaligned) buffers.
bool host_drain_target(const dil_sign_key* k, uint8_t* z, uint8_t* h, uint8_t* ctilde, uint32_t* attempts, DrainTarget& dt) {
    if (k->tune.host_copy_path) return false;
    auto dev_alias = [&](const void* p, size_t align) -> void* {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
        if (reinterpret_cast<uintptr_t>(a.devicePointer) & (align - 1)) return nullptr;
        return a.devicePointer;
    };
    dt.z = static_cast<uint8_t*>(dev_alias(z, 16));
    dt.h = static_cast<uint8_t*>(dev_alias(h, 1));
    dt.ct = static_cast<uint8_t*>(dev_alias(ctilde, 16));
    dt.att = attempts ? static_cast<uint32_t*>(dev_alias(attempts, 4)) : nullptr;
    dt.stream = k->cs_own;
    return dt.z && dt.h && dt.ct && (!attempts || dt.att);
}

}  // namespace

extern "C" {

// ---- asynchronous pair: several batches in flight from ONE host thread (one batch per key handle) ----
int dil_sign_batch_dev_begin(dil_engine_t* e, dil_sign_key_t* k, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                             uint8_t* d_z, uint8_t* d_h, uint8_t* d_ctilde, uint32_t* d_attempts, void* stream) {
    if (!e || !k) return DIL_ERR_ARG;
    if (!d_msgs || !d_offsets || !d_z || !d_h || !d_ctilde || !d_attempts || n == 0 || n > 0x07FFFFFFu) return DIL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_ctilde) & 7u) || (reinterpret_cast<uintptr_t>(d_z) & 15u)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const size_t chunk = k->tune.dev_chunk ? k->tune.dev_chunk : ((size_t)1 << 20);
    if (n > chunk + chunk / 4 || k->profile)
        return fail_msg(e, DIL_ERR_ARG, "dil_sign_batch_dev_begin: at most 1.25 x dev_chunk messages per asynchronous batch, profiling off");
    return rounds_begin(e, k, d_msgs, d_offsets, n, d_z, d_h, d_ctilde, d_attempts, (cudaStream_t)stream, nullptr);
}

int dil_sign_batch_host_begin(dil_engine_t* e, dil_sign_key_t* k, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                              uint8_t* z, uint8_t* h, uint8_t* ctilde, uint32_t* attempts) {
    if (!e || !k) return DIL_ERR_ARG;
    if (!msgs || !offsets || !z || !h || !ctilde || n == 0 || n > 0x07FFFFFFu) return DIL_ERR_ARG;
    if (!offsets_ok(offsets, n)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    if (k->pend) return fail_msg(e, DIL_ERR_ARG, "sign: this key handle already has a batch in flight (dil_sign_batch_finish it first)");
    const size_t CHUNK = k->tune.host_chunk ? k->tune.host_chunk : 262144;
    if (n > CHUNK + CHUNK / 4 || k->profile)
        return fail_msg(e, DIL_ERR_ARG, "dil_sign_batch_host_begin: at most 1.25 x host_chunk messages per asynchronous batch, profiling off");
    int rc = host_stage(e, k, msgs, offsets, n);
    if (rc) return rc;
    DrainTarget dt;
    if (!host_drain_target(k, z, h, ctilde, attempts, dt))
        return fail_msg(e, DIL_ERR_ARG, "dil_sign_batch_host_begin: z, h, ctilde (and attempts) must be pinned host memory the device can "
                                        "address (cudaHostAlloc / cudaHostRegister), z and ctilde 16-byte aligned");
    CK(cudaEventCreateWithFlags(&dt.idle, cudaEventDisableTiming));
    cudaError_t er = cudaEventRecord(dt.idle, dt.stream);
    rc = er == cudaSuccess ? DIL_OK : fail(e, er, "sign drain event");
    if (rc == DIL_OK) rc = rounds_begin(e, k, k->msgs_d, k->off_d, n, k->zp_d, k->h_d, k->ct_d, k->att_d, k->st_own, &dt);
    if (rc != DIL_OK) {
        cudaEventDestroy(dt.idle);
        return rc;
    }
    static_cast<Pending*>(k->pend)->host = true;
    return DIL_OK;
}

int dil_sign_batch_finish(dil_engine_t* e, dil_sign_key_t* k) {
    if (!e || !k) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    Pending* p = static_cast<Pending*>(k->pend);
    if (!p) return fail_msg(e, DIL_ERR_ARG, "sign: no batch in flight on this key handle");
    const bool host = p->host, loaded = p->loaded;
    const DrainTarget dt = p->drain;
    int rc = rounds_finish(e, k, false);
    if (host) {   // the last drains: every signature is in the caller's buffers when this returns
        cudaError_t e1 = wait_stream(k, dt.stream, false);
        (void)loaded;
        cudaEventDestroy(dt.idle);
        if (rc == DIL_OK && e1 != cudaSuccess) rc = fail(e, e1, "sign drain sync");
    }
    return rc;
}

int dil_sign_batch_host(dil_engine_t* e, dil_sign_key_t* k, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                        uint8_t* z, uint8_t* h, uint8_t* ctilde, uint32_t* attempts) {
    if (!e || !k) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!msgs || !offsets || !z || !h || !ctilde || n > 0x07FFFFFFu) return DIL_ERR_ARG;
    if (!offsets_ok(offsets, n)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    if (k->pend) return fail_msg(e, DIL_ERR_ARG, "sign: this key handle already has a batch in flight (dil_sign_batch_finish it first)");
    const LevelParams& P = k->P;
    int rcs = host_stage(e, k, msgs, offsets, n);
    if (rcs) return rcs;
    cudaStream_t st = k->st_own, cs = k->cs_own;
    const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k;
    {
        const size_t CHUNK = k->tune.host_chunk ? k->tune.host_chunk : 262144;
        DrainTarget dt;
        if (host_drain_target(k, z, h, ctilde, attempts, dt)) {
            CK(cudaEventCreateWithFlags(&dt.idle, cudaEventDisableTiming));
            cudaError_t er = cudaEventRecord(dt.idle, cs);
            int rc = er == cudaSuccess ? DIL_OK : fail(e, er, "sign drain event");
            // large batches amortise the rejection tail better; 2^18 items keep the workspace at a few GiB
            for (size_t lo = 0; lo < n && rc == DIL_OK;) {
                const size_t m = n - lo <= CHUNK + CHUNK / 4 ? n - lo : CHUNK;
                DrainTarget t = dt;
                t.z += lo * zb; t.h += lo * hb; t.ct += lo * 32;
                if (t.att) t.att += lo;
                rc = sign_rounds(e, k, k->msgs_d, k->off_d + lo, m, k->zp_d + lo * zb, k->h_d + lo * hb, k->ct_d + lo * 32,
                                 k->att_d + lo, st, &t);
                lo += m;
            }
            const bool loaded = e->sign_in_flight.load() > 0;
            cudaError_t e2 = wait_stream(k, st, loaded);
            cudaError_t e1 = wait_stream(k, cs, loaded);
            cudaEventDestroy(dt.idle);
            if (rc) return rc;
            if (e1 != cudaSuccess) return fail(e, e1, "sign drain sync");
            if (e2 != cudaSuccess) return fail(e, e2, "sign sync");
            return DIL_OK;
        }
    }
    // Chunked so that the D2H of a finished chunk's signatures (2.4-4.6 KB each, ~47 ns per signature over
    // PCIe) overlaps the signing of the next chunk.  Signing has a fixed per-batch latency (rejection
    // tail), so chunks are unequal: 64 K-message chunks while a lot remains, then a 3:1 split so that only
    // the small last chunk's transfer is exposed.
    std::vector<size_t> chunks;
    {
        size_t rem = n;
        while (rem > 98304) { chunks.push_back(65536); rem -= 65536; }
        if (rem > 16384) { size_t a = (rem * 3 / 4 + 255) & ~(size_t)255; chunks.push_back(a); chunks.push_back(rem - a); }
        else chunks.push_back(rem);
    }
    cudaEvent_t done = nullptr;
    CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    int rc = DIL_OK;
    size_t lo = 0;
    for (size_t ci = 0; ci < chunks.size() && rc == DIL_OK; lo += chunks[ci], ci++) {
        const size_t m = chunks[ci];
        rc = sign_rounds(e, k, k->msgs_d, k->off_d + lo, m, k->zp_d + lo * zb, k->h_d + lo * hb, k->ct_d + lo * 32, k->att_d + lo, st);
        if (rc) break;
        cudaError_t er = cudaEventRecord(done, st);
        if (er == cudaSuccess) er = cudaStreamWaitEvent(cs, done, 0);
        if (er == cudaSuccess) er = cudaMemcpyAsync(z + lo * zb, k->zp_d + lo * zb, m * zb, cudaMemcpyDeviceToHost, cs);
        if (er == cudaSuccess) er = cudaMemcpyAsync(h + lo * hb, k->h_d + lo * hb, m * hb, cudaMemcpyDeviceToHost, cs);
        if (er == cudaSuccess) er = cudaMemcpyAsync(ctilde + lo * 32, k->ct_d + lo * 32, m * 32, cudaMemcpyDeviceToHost, cs);
        if (er == cudaSuccess && attempts) er = cudaMemcpyAsync(attempts + lo, k->att_d + lo, m * 4, cudaMemcpyDeviceToHost, cs);
        if (er != cudaSuccess) rc = fail(e, er, "sign D2H");
    }
    cudaError_t er2 = cudaStreamSynchronize(cs);
    cudaError_t er3 = cudaStreamSynchronize(st);
    cudaEventDestroy(done);
    if (rc) return rc;
    if (er2 != cudaSuccess) return fail(e, er2, "sign D2H sync");
    if (er3 != cudaSuccess) return fail(e, er3, "sign sync");
    return DIL_OK;
}

// ---- one key per signature -------------------------------------------------------------------------------------
// The reference's sign driver feeds rho, tr, K, s1, s2, t0 in front of EVERY message (rtl_tb/tb_sign_top.v:171-284), so
// a batch may carry a different key per signature (e.g. all 100 KAT vectors of a level in one call).  The engine keeps
// one workspace per level for this mode; key material is unpacked, transformed and ExpandA'd per item on the device.
}  // extern "C"

namespace {

dil_sign_key* multi_ctx(dil_engine* e, int level) {
    const int idx = level == 2 ? 0 : (level == 3 ? 1 : 2);
    if (!e->multi_sign[idx]) {
        dil_sign_key* k = new (std::nothrow) dil_sign_key();
        if (!k) return nullptr;
        k->P = dil::level_params(level);
        k->device = e->device;
        e->multi_sign[idx] = k;
    }
    return static_cast<dil_sign_key*>(e->multi_sign[idx]);
}

void free_multi(dil_sign_key* k) {
    const LevelParams& P = k->P;
    wipe_free(k->key_items, k->multi_cap * (size_t)(P.l + 2 * P.k) * 1024);
    wipe_free(k->multi_in, k->multi_in_cap);
    if (k->a_items) cudaFree(k->a_items);
    if (k->yh) cudaFree(k->yh);
    if (k->wh) cudaFree(k->wh);
    k->a_items = k->key_items = k->yh = k->wh = nullptr;
    k->multi_in = nullptr;
    k->multi_cap = k->multi_slots = k->multi_in_cap = 0;
}

int ensure_multi(dil_engine* e, dil_sign_key* k, size_t n) {
    const LevelParams& P = k->P;
    const size_t slots = slots_alloc(k, n);
    if (n <= k->multi_cap && slots <= k->multi_slots) return DIL_OK;
    const size_t in_cap = k->multi_in_cap;
    uint8_t* in = k->multi_in;
    k->multi_in = nullptr;           // the input staging is managed separately
    k->multi_in_cap = 0;
    free_multi(k);
    k->multi_in = in;
    k->multi_in_cap = in_cap;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    A(dmalloc(&k->a_items, n * (size_t)P.k * P.l * 256));
    A(dmalloc(&k->key_items, n * (size_t)(P.l + 2 * P.k) * 256));
    A(dmalloc(&k->yh, slots * (size_t)P.l * 256));
    A(dmalloc(&k->wh, slots * (size_t)P.k * 256));
    if (err != cudaSuccess) {
        free_multi(k);
        return fail_msg(e, DIL_ERR_ALLOC, std::string("multi-key sign workspace: ") + cudaGetErrorString(err));
    }
    k->multi_cap = n;
    k->multi_slots = slots;
    return DIL_OK;
}

constexpr size_t MULTI_CHUNK = 65536;   // per-item A_hat is 16 / 30 / 56 KiB: 1 - 3.5 GiB per chunk

// all pointers device pointers; key arrays hold n records
int sign_multi_run(dil_engine* e, dil_sign_key* k, const uint8_t* d_rho, const uint8_t* d_key, const uint8_t* d_tr, const uint8_t* d_s1p,
                   const uint8_t* d_s2p, const uint8_t* d_t0p, const uint8_t* d_msgs, const uint64_t* d_off, size_t n, uint8_t* d_z,
                   uint8_t* d_h, uint8_t* d_ct, uint32_t* d_att, cudaStream_t st) {
    const LevelParams& P = k->P;
    const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k, nkey = (size_t)P.l + 2 * P.k;
    for (size_t lo = 0; lo < n;) {
        const size_t m = n - lo <= MULTI_CHUNK + MULTI_CHUNK / 4 ? n - lo : MULTI_CHUNK;
        int rc = ensure_multi(e, k, m);
        if (rc) return rc;
        // per item: A_hat = ExpandA(rho); (s1 | s2 | t0) unpacked, transformed, scaled by 256^-1 for the tail
        CK(dil::launch_expand_a(k->a_items, d_rho + lo * 32, m, P.k, P.l, e->sm_count, st));
        CK(dil::launch_unpack_keys(P.level, k->key_items, d_s1p + lo * P.l * P.s_bytes, d_s2p + lo * P.k * P.s_bytes,
                                   d_t0p + lo * P.k * 416, m, st));
        CK(dil::launch_ntt_fwd(k->key_items, k->key_items, m * nkey, e->sm_count, st));
        CK(dil::launch_scale_inv256(k->key_items, m * nkey, st));
        e->launches += 6;
        MultiKeys mk{d_tr + lo * 32, d_key + lo * 32};
        rc = sign_rounds(e, k, d_msgs, d_off + lo, m, d_z + lo * zb, d_h + lo * hb, d_ct + lo * 32, d_att + lo, st, nullptr, &mk);
        if (rc) return rc;
        lo += m;
    }
    return DIL_OK;
}

}  // namespace

void dil_internal_free_multi_sign(dil_engine* e) {
    for (auto& p : e->multi_sign) {
        if (!p) continue;
        dil_sign_key* k = static_cast<dil_sign_key*>(p);
        free_multi(k);
        dil_sign_key_destroy(e, k);
        p = nullptr;
    }
}

extern "C" {

int dil_sign_multi_dev(dil_engine_t* e, int level, const uint8_t* d_rho, const uint8_t* d_key, const uint8_t* d_tr, const uint8_t* d_s1p,
                       const uint8_t* d_s2p, const uint8_t* d_t0p, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                       uint8_t* d_z, uint8_t* d_h, uint8_t* d_ctilde, uint32_t* d_attempts, void* stream) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!d_rho || !d_key || !d_tr || !d_s1p || !d_s2p || !d_t0p || !d_msgs || !d_offsets || !d_z || !d_h || !d_ctilde || !d_attempts ||
        n > 0x07FFFFFFu)
        return DIL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_ctilde) & 7u) || (reinterpret_cast<uintptr_t>(d_z) & 15u)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    dil_sign_key* k = multi_ctx(e, level);
    if (!k) return DIL_ERR_ALLOC;
    return sign_multi_run(e, k, d_rho, d_key, d_tr, d_s1p, d_s2p, d_t0p, d_msgs, d_offsets, n, d_z, d_h, d_ctilde, d_attempts,
                          (cudaStream_t)stream);
}

int dil_sign_multi_host(dil_engine_t* e, int level, const uint8_t* rho, const uint8_t* key, const uint8_t* tr, const uint8_t* s1p,
                        const uint8_t* s2p, const uint8_t* t0p, const uint8_t* msgs, const uint64_t* offsets, size_t n, uint8_t* z,
                        uint8_t* h, uint8_t* ctilde, uint32_t* attempts) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!rho || !key || !tr || !s1p || !s2p || !t0p || !msgs || !offsets || !z || !h || !ctilde || n > 0x07FFFFFFu) return DIL_ERR_ARG;
    if (!offsets_ok(offsets, n)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    dil_sign_key* k = multi_ctx(e, level);
    if (!k) return DIL_ERR_ALLOC;
    const LevelParams& P = k->P;
    cudaStream_t st = e->host_stream;
    const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k;
    const size_t s1b = (size_t)P.l * P.s_bytes, s2b = (size_t)P.k * P.s_bytes, t0b = (size_t)P.k * 416;
    const size_t mbytes = offsets[n] > 0 ? offsets[n] : 1;
    // staging: inputs | outputs in one grow-only allocation
    size_t total = 0;
    auto seg = [&](size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_rho = seg(n * 32), o_key = seg(n * 32), o_tr = seg(n * 32), o_s1 = seg(n * s1b), o_s2 = seg(n * s2b), o_t0 = seg(n * t0b),
                 o_msg = seg(mbytes), o_off = seg((n + 1) * 8), o_z = seg(n * zb), o_h = seg(n * hb), o_ct = seg(n * 32), o_att = seg(n * 4);
    if (total > k->multi_in_cap) {
        wipe_free(k->multi_in, k->multi_in_cap);
        k->multi_in = nullptr;
        k->multi_in_cap = 0;
        CK(dmalloc(&k->multi_in, total));
        k->multi_in_cap = total;
    }
    uint8_t* base = k->multi_in;
    CK(cudaMemcpyAsync(base + o_rho, rho, n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_key, key, n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_tr, tr, n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_s1, s1p, n * s1b, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_s2, s2p, n * s2b, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_t0, t0p, n * t0b, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_msg, msgs, offsets[n], cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = sign_multi_run(e, k, base + o_rho, base + o_key, base + o_tr, base + o_s1, base + o_s2, base + o_t0, base + o_msg,
                            reinterpret_cast<const uint64_t*>(base + o_off), n, base + o_z, base + o_h, base + o_ct,
                            reinterpret_cast<uint32_t*>(base + o_att), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(z, base + o_z, n * zb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h, base + o_h, n * hb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctilde, base + o_ct, n * 32, cudaMemcpyDeviceToHost, st));
    if (attempts) CK(cudaMemcpyAsync(attempts, base + o_att, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DIL_OK;
}

}  // extern "C"

// =======================================================================================
// Verification API (combined_top.v mode 1; I/O order of rtl_tb/tb_verify_top.v:144-249:
// rho, c~, z, t1, mlen, M, h -> one accept/reject word)
// =======================================================================================
struct dil_verify_key {
    LevelParams P{};
    int device = -1;
    int32_t* a_ext = nullptr;   // k x (l+1) polys: [A_hat | -NTT(t1 * 2^13)]
    uint8_t* seeds = nullptr;   // tr[32] | rho[32] | t1_packed
    std::mutex mu;
    size_t cap = 0;
    uint64_t *mu_d = nullptr, *w1p = nullptr;
    int32_t *v = nullptr, *w = nullptr;
    uint32_t *hmask = nullptr, *bad = nullptr;
    // The workspace above is shared by every call on this key.  _dev calls only enqueue, so a later call (possibly on
    // another stream) first waits for the previous user's last kernel: ws_done is recorded after it.
    cudaEvent_t ws_done = nullptr;
    // host-variant staging
    uint8_t *msgs_d = nullptr, *z_d = nullptr, *h_d = nullptr, *ct_d = nullptr, *ok_d = nullptr;
    uint64_t* off_d = nullptr;
    size_t msgs_cap = 0, io_cap = 0;
};

namespace {
void vfree(void* p) { if (p) cudaFree(p); }
void free_vws(dil_verify_key* k) {
    vfree(k->mu_d); vfree(k->w1p); vfree(k->v); vfree(k->w); vfree(k->hmask); vfree(k->bad);
    k->mu_d = k->w1p = nullptr; k->v = k->w = nullptr; k->hmask = k->bad = nullptr; k->cap = 0;
}
int ensure_vws(dil_engine* e, dil_verify_key* k, size_t n) {
    if (n <= k->cap) return DIL_OK;
    free_vws(k);
    const LevelParams& P = k->P;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    A(dmalloc(&k->mu_d, n * 8));
    A(dmalloc(&k->w1p, n * (size_t)(P.k * P.w1_bytes / 8)));
    A(dmalloc(&k->v, n * (size_t)(P.l + 1) * 256));
    A(dmalloc(&k->w, n * (size_t)P.k * 256));
    A(dmalloc(&k->hmask, n * (size_t)P.k * 8));
    A(dmalloc(&k->bad, n));
    if (err != cudaSuccess) {
        free_vws(k);
        return fail_msg(e, DIL_ERR_ALLOC, std::string("verify workspace: ") + cudaGetErrorString(err));
    }
    k->cap = n;
    return DIL_OK;
}
int verify_run(dil_engine* e, dil_verify_key* k, const uint8_t* d_msgs, const uint64_t* d_off, size_t n, const uint8_t* d_z,
               const uint8_t* d_h, const uint8_t* d_ct, uint8_t* d_ok, cudaStream_t st) {
    const LevelParams& P = k->P;
    // very large batches are verified in 2^20-signature pieces so that the unpacked intermediates (about 9 / 12 / 17 KB per
    // signature at levels 2 / 3 / 5) stay bounded; the pieces run back to back on the caller's stream
    constexpr size_t CHUNK = (size_t)1 << 20;
    if (n > CHUNK + CHUNK / 4) {
        const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k;
        for (size_t lo = 0; lo < n;) {
            const size_t m = n - lo <= CHUNK + CHUNK / 4 ? n - lo : CHUNK;
            int rc = verify_run(e, k, d_msgs, d_off + lo, m, d_z + lo * zb, d_h + lo * hb, d_ct + lo * 32, d_ok + lo, st);
            if (rc) return rc;
            lo += m;
        }
        return DIL_OK;
    }
    int rc = ensure_vws(e, k, n);
    if (rc) return rc;
    const uint32_t nn = (uint32_t)n;
    if (!k->ws_done) CK(cudaEventCreateWithFlags(&k->ws_done, cudaEventDisableTiming));
    else CK(cudaStreamWaitEvent(st, k->ws_done, 0));   // the previous batch may still be running on another stream
    CK(cudaMemsetAsync(k->bad, 0, n * 4, st));
    CK(dil::launch_verify_mu(k->mu_d, k->seeds, d_msgs, d_off, nn, 0, st));
    // Levels 2 and 3: the core reads z straight from the packed signature (with the ||z|| check) and applies the hints and packs
    // w1' = UseHint(h, w') itself - no unpack pass, no UseHint pass, neither z as int32 nor w' ever reaches HBM.  Level 5: the
    // 8 x 8 core (one CTA per SM) gains nothing from it: the extra work per row costs it as much as the passes save
    // (fused UseHint: 55.6 vs 65.8 M/s with 8 warps per SM; everything fused with 12 warps: 70.4 vs 72.7 M/s at 2^20), so it keeps the
    // separate unpack_z / usehint_pack passes.
    if (P.level != 5) {
        CK(dil::launch_verify_prep(P.level, k->v, k->hmask, k->bad, d_h, reinterpret_cast<const uint64_t*>(d_ct), nn, st));
        CK(dil::launch_verify_core(k->w, k->a_ext, k->v, P.level, n, e->sm_count, st, reinterpret_cast<uint8_t*>(k->w1p), k->hmask, d_z, k->bad));
    } else {
        CK(dil::launch_unpack_z(P.level, k->v, k->bad, d_z, nn, st));
        CK(dil::launch_verify_prep(P.level, k->v, k->hmask, k->bad, d_h, reinterpret_cast<const uint64_t*>(d_ct), nn, st));
        CK(dil::launch_verify_core(k->w, k->a_ext, k->v, P.level, n, e->sm_count, st));
        CK(dil::launch_usehint_pack(P.level, reinterpret_cast<uint32_t*>(k->w1p), k->w, k->hmask, nn, st));
        e->launches += 2;
    }
    CK(dil::launch_verify_hash(P.level, d_ok, k->mu_d, k->w1p, reinterpret_cast<const uint64_t*>(d_ct), k->bad, nn, st));
    CK(cudaEventRecord(k->ws_done, st));
    e->launches += 4;
    return DIL_OK;
}
}  // namespace

extern "C" {

int dil_verify_key_create(dil_engine_t* e, dil_verify_key_t** out, int level, const uint8_t* rho, const uint8_t* t1p) {
    if (!e || !out || !rho || !t1p) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    dil_verify_key* k = new (std::nothrow) dil_verify_key();
    if (!k) return DIL_ERR_ALLOC;
    k->P = dil::level_params(level);
    k->device = e->device;
    const LevelParams& P = k->P;
    const int K = P.k, L = P.l;
    // -t1 * 2^13 mod Q (decoder.v:96-100 delivers t1 * 2^13), unpacked on the host
    std::vector<int32_t> t1neg((size_t)K * 256);
    std::vector<uint32_t> tmp(256);
    for (int p = 0; p < K; p++) {
        bits_get(tmp.data(), t1p + (size_t)p * 320, 256, 10);
        for (int i = 0; i < 256; i++) {
            int64_t v = ((int64_t)tmp[i] << dil::D_BITS) % dil::Q_I;
            t1neg[(size_t)p * 256 + i] = (int32_t)((dil::Q_I - v) % dil::Q_I);
        }
    }
    cudaStream_t st = e->host_stream;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    int32_t* a_tmp = nullptr;
    const size_t t1_bytes = (size_t)K * 320;
    std::vector<uint8_t> seeds(64 + t1_bytes);
    std::memcpy(seeds.data() + 32, rho, 32);
    std::memcpy(seeds.data() + 64, t1p, t1_bytes);
    A(dmalloc(&k->a_ext, (size_t)K * (L + 1) * 256));
    A(dmalloc(&a_tmp, (size_t)K * L * 256));
    A(dmalloc(&k->seeds, seeds.size()));
    if (err == cudaSuccess) {
        A(cudaMemcpyAsync(k->seeds, seeds.data(), seeds.size(), cudaMemcpyHostToDevice, st));
        A(dil::launch_tr(reinterpret_cast<uint64_t*>(k->seeds), k->seeds + 32, k->seeds + 64, (uint32_t)t1_bytes, st));
        A(dil::launch_expand_a(a_tmp, k->seeds + 32, 1, K, L, e->sm_count, st));
        // rows of A_hat into the first l columns of a_ext, -t1*2^13 into column l, then NTT that column
        A(cudaMemcpy2DAsync(k->a_ext, (size_t)(L + 1) * 1024, a_tmp, (size_t)L * 1024, (size_t)L * 1024, K, cudaMemcpyDeviceToDevice, st));
        A(cudaMemcpy2DAsync(k->a_ext + (size_t)L * 256, (size_t)(L + 1) * 1024, t1neg.data(), 1024, 1024, K, cudaMemcpyHostToDevice, st));
        for (int i = 0; i < K; i++) {
            int32_t* col = k->a_ext + ((size_t)i * (L + 1) + L) * 256;
            A(dil::launch_ntt_fwd(col, col, 1, e->sm_count, st));
        }
        A(cudaStreamSynchronize(st));
    }
    if (a_tmp) cudaFree(a_tmp);
    if (err != cudaSuccess) {
        vfree(k->a_ext); vfree(k->seeds);
        delete k;
        return fail(e, err, "dil_verify_key_create");
    }
    e->launches += 2 + K;
    *out = k;
    return DIL_OK;
}

int dil_verify_key_destroy(dil_engine_t* e, dil_verify_key_t* k) {
    (void)e;
    if (!k) return DIL_OK;
    DeviceGuard dg(k->device);
    free_vws(k);
    vfree(k->a_ext); vfree(k->seeds); vfree(k->msgs_d); vfree(k->z_d); vfree(k->h_d); vfree(k->ct_d); vfree(k->ok_d); vfree(k->off_d);
    if (k->ws_done) cudaEventDestroy(k->ws_done);
    delete k;
    return DIL_OK;
}

int dil_verify_batch_dev(dil_engine_t* e, dil_verify_key_t* k, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                         const uint8_t* d_z, const uint8_t* d_h, const uint8_t* d_ctilde, uint8_t* d_ok, void* stream) {
    if (!e || !k) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!d_msgs || !d_offsets || !d_z || !d_h || !d_ctilde || !d_ok || n > 0x7FFFFFFFu) return DIL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_ctilde) & 7u) || (reinterpret_cast<uintptr_t>(d_z) & 3u)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    return verify_run(e, k, d_msgs, d_offsets, n, d_z, d_h, d_ctilde, d_ok, (cudaStream_t)stream);
}

int dil_verify_batch_host(dil_engine_t* e, dil_verify_key_t* k, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                          const uint8_t* z, const uint8_t* h, const uint8_t* ctilde, uint8_t* ok) {
    if (!e || !k) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!msgs || !offsets || !z || !h || !ctilde || !ok || n > 0x7FFFFFFFu) return DIL_ERR_ARG;
    if (!offsets_ok(offsets, n)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(k->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const LevelParams& P = k->P;
    cudaStream_t st = e->host_stream;
    const size_t mbytes = offsets[n] > 0 ? offsets[n] : 1;
    const size_t zb = (size_t)P.l * P.z_bytes, hb = (size_t)P.omega + P.k;
    if (mbytes > k->msgs_cap) {
        vfree(k->msgs_d);
        k->msgs_d = nullptr;
        k->msgs_cap = 0;
        CK(dmalloc(&k->msgs_d, mbytes));
        k->msgs_cap = mbytes;
    }
    if (n > k->io_cap) {
        vfree(k->z_d); vfree(k->h_d); vfree(k->ct_d); vfree(k->ok_d); vfree(k->off_d);
        k->z_d = k->h_d = k->ct_d = k->ok_d = nullptr;
        k->off_d = nullptr;
        k->io_cap = 0;
        CK(dmalloc(&k->z_d, n * zb));
        CK(dmalloc(&k->h_d, n * hb));
        CK(dmalloc(&k->ct_d, n * 32));
        CK(dmalloc(&k->ok_d, n));
        CK(dmalloc(&k->off_d, n + 1));
        k->io_cap = n;
    }
    // Streaming: signatures (2.4-4.6 KB each) cross PCIe in 64 K-signature pieces on the copy stream while the previous
    // piece is being verified on the host stream; messages and offsets (small) go first.
    cudaStream_t cs = e->copy_stream;
    cudaEvent_t ready = nullptr;
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    A(cudaMemcpyAsync(k->msgs_d, msgs, offsets[n], cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(k->off_d, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    // the staging buffers may still be read by the previous call's kernels on the host stream
    A(cudaEventRecord(ready, st));
    A(cudaStreamWaitEvent(cs, ready, 0));
    constexpr size_t PIECE = 65536;
    int rc = DIL_OK;
    for (size_t lo = 0; lo < n && err == cudaSuccess && rc == DIL_OK;) {
        const size_t m = n - lo <= PIECE + PIECE / 4 ? n - lo : PIECE;
        A(cudaMemcpyAsync(k->z_d + lo * zb, z + lo * zb, m * zb, cudaMemcpyHostToDevice, cs));
        A(cudaMemcpyAsync(k->h_d + lo * hb, h + lo * hb, m * hb, cudaMemcpyHostToDevice, cs));
        A(cudaMemcpyAsync(k->ct_d + lo * 32, ctilde + lo * 32, m * 32, cudaMemcpyHostToDevice, cs));
        A(cudaEventRecord(ready, cs));
        A(cudaStreamWaitEvent(st, ready, 0));
        if (err == cudaSuccess) rc = verify_run(e, k, k->msgs_d, k->off_d + lo, m, k->z_d + lo * zb, k->h_d + lo * hb, k->ct_d + lo * 32, k->ok_d + lo, st);
        lo += m;
    }
    if (err == cudaSuccess && rc == DIL_OK) A(cudaMemcpyAsync(ok, k->ok_d, n, cudaMemcpyDeviceToHost, st));
    cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
    cudaEventDestroy(ready);
    if (rc) return rc;
    if (err != cudaSuccess) return fail(e, err, "dil_verify_batch_host");
    if (e1 != cudaSuccess) return fail(e, e1, "verify H2D sync");
    if (e2 != cudaSuccess) return fail(e, e2, "verify sync");
    return DIL_OK;
}

}  // extern "C"

// =======================================================================================
// Batched key generation (combined_top.v mode 0; outputs as rtl_tb/tb_keygen_top.v:180-275)
// =======================================================================================
namespace {
struct Arena {   // carve 256-byte aligned segments out of one grow-only device allocation (engine staging slot 3)
    size_t total = 0;
    size_t seg(size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; }
};

int arena_reserve(dil_engine* e, size_t bytes, uint8_t** base) {
    if (bytes > e->staging_bytes[3]) {
        if (e->staging[3]) cudaFree(e->staging[3]);
        e->staging[3] = nullptr;
        e->staging_bytes[3] = 0;
        cudaError_t err = cudaMalloc(&e->staging[3], bytes);
        if (err != cudaSuccess) {
            return fail_msg(e, DIL_ERR_ALLOC, std::string("verify arena: ") + cudaGetErrorString(err));
        }
        e->staging_bytes[3] = bytes;
    }
    *base = static_cast<uint8_t*>(e->staging[3]);
    return DIL_OK;
}

// one piece of a key-generation batch; all pointers device pointers; `work` holds the intermediates (keygen_work_bytes)
size_t keygen_work_bytes(const LevelParams& P, size_t n) {
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return r(n * 64) + r(n * P.l * 1024) + 2 * r(n * P.k * 1024);
}
cudaError_t keygen_piece(dil_engine* e, const LevelParams& P, uint8_t* work, const uint8_t* d_xi, size_t n, uint8_t* d_rho, uint8_t* d_key,
                         uint8_t* d_tr, uint8_t* d_s1p, uint8_t* d_s2p, uint8_t* d_t1p, uint8_t* d_t0p, cudaStream_t st) {
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t K = P.k, L = P.l;
    uint64_t* rhop = reinterpret_cast<uint64_t*>(work);
    int32_t* s1 = reinterpret_cast<int32_t*>(work + r(n * 64));
    int32_t* s2 = reinterpret_cast<int32_t*>(work + r(n * 64) + r(n * L * 1024));
    int32_t* t = reinterpret_cast<int32_t*>(work + r(n * 64) + r(n * L * 1024) + r(n * K * 1024));
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t x) { if (err == cudaSuccess) err = x; };
    A(dil::launch_keygen_seed(d_rho, rhop, d_key, d_xi, (uint32_t)n, st));
    A(dil::launch_eta_sample(P.level, s1, s2, rhop, (uint32_t)n, st));
    // t = INTT(ExpandA(rho) * NTT(s1)): per-item rho, A generated on chip
    A(dil::launch_matvec_expand(t, d_rho, s1, P.k, P.l, n, DIL_RHO_PER_ITEM | DIL_NTT_INPUT | DIL_INTT_OUTPUT, e->sm_count, st));
    A(dil::launch_t_pack(d_t1p, d_t0p, t, s2, n * K, st));
    A(dil::launch_s_pack(P.eta, d_s1p, s1, n * L, st));
    A(dil::launch_s_pack(P.eta, d_s2p, s2, n * K, st));
    A(dil::launch_tr_batch(d_tr, d_rho, d_t1p, (uint32_t)(K * 320), (uint32_t)n, st));
    if (err == cudaSuccess) e->launches += 7;
    return err;
}
constexpr size_t KEYGEN_PIECE = 16384;   // host path, keys per piece: 0.2-0.4 GB of intermediates (s1, s2, t as int32 polynomials)
// device path: larger pieces fill the GPU better (the one-thread-per-key hashes run 2048 instead of 512 warps); 0.8-1.5 GB of intermediates
constexpr size_t KEYGEN_PIECE_DEV = 65536;
}  // namespace

// device pointers in, device pointers out; enqueues on `stream` (the secret intermediates live in an engine-owned
// workspace that the next user waits for)
extern "C" int dil_keygen_batch_dev(dil_engine_t* e, int level, const uint8_t* d_xi, size_t n, uint8_t* d_rho, uint8_t* d_key,
                                    uint8_t* d_tr, uint8_t* d_s1p, uint8_t* d_s2p, uint8_t* d_t1p, uint8_t* d_t0p, void* stream) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!d_xi || !d_rho || !d_key || !d_tr || !d_s1p || !d_s2p || !d_t1p || !d_t0p || n > 0x00FFFFFFu) return DIL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_rho) | reinterpret_cast<uintptr_t>(d_key) | reinterpret_cast<uintptr_t>(d_tr)) & 7u) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const LevelParams P = dil::level_params(level);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t K = P.k, L = P.l, sb = P.s_bytes;
    const size_t piece = n < KEYGEN_PIECE_DEV ? n : KEYGEN_PIECE_DEV;
    uint8_t* work = nullptr;
    int rc = arena_reserve(e, keygen_work_bytes(P, piece), &work);
    if (rc) return rc;
    cudaError_t err = cudaSuccess;
    if (!e->arena_done) err = cudaEventCreateWithFlags(&e->arena_done, cudaEventDisableTiming);
    else err = cudaStreamWaitEvent(st, e->arena_done, 0);
    for (size_t lo = 0; lo < n && err == cudaSuccess; lo += piece) {
        const size_t m = n - lo < piece ? n - lo : piece;
        err = keygen_piece(e, P, work, d_xi + lo * 32, m, d_rho + lo * 32, d_key + lo * 32, d_tr + lo * 32, d_s1p + lo * L * sb,
                           d_s2p + lo * K * sb, d_t1p + lo * K * 320, d_t0p + lo * K * 416, st);
    }
    if (err == cudaSuccess) err = cudaEventRecord(e->arena_done, st);
    if (err != cudaSuccess) return fail(e, err, "dil_keygen_batch_dev");
    return DIL_OK;
}

// host pointers; pieces of 16 K keys: the packed keys of piece i cross PCIe on the copy stream while piece i + 1 is generated
extern "C" int dil_keygen_batch_host(dil_engine_t* e, int level, const uint8_t* xi, size_t n, uint8_t* rho, uint8_t* key,
                                     uint8_t* tr, uint8_t* s1p, uint8_t* s2p, uint8_t* t1p, uint8_t* t0p) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!xi || !rho || !key || !tr || !s1p || !s2p || !t1p || !t0p || n > 0x00FFFFFFu) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const LevelParams P = dil::level_params(level);
    cudaStream_t st = e->host_stream, cs = e->copy_stream;
    const size_t K = P.k, L = P.l, sb = P.s_bytes;
    const size_t piece = n < KEYGEN_PIECE ? n : KEYGEN_PIECE;
    // one transient allocation: intermediates | two sets of packed outputs (double buffer) | seeds
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t out_bytes = r(piece * 32) * 3 + r(piece * L * sb) + r(piece * K * sb) + r(piece * K * 320) + r(piece * K * 416);
    const size_t wbytes = keygen_work_bytes(P, piece);
    const size_t total = wbytes + 2 * out_bytes + r(n * 32);
    uint8_t* base = nullptr;
    cudaError_t aerr = cudaMalloc(reinterpret_cast<void**>(&base), total);
    if (aerr != cudaSuccess) return fail_msg(e, DIL_ERR_ALLOC, std::string("keygen arena: ") + cudaGetErrorString(aerr));
    uint8_t* d_xi = base + wbytes + 2 * out_bytes;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t x) { if (err == cudaSuccess) err = x; };
    for (int i = 0; i < 2; i++) {
        A(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    }
    A(cudaMemcpyAsync(d_xi, xi, n * 32, cudaMemcpyHostToDevice, st));
    int idx = 0;
    for (size_t lo = 0; lo < n && err == cudaSuccess; lo += piece, idx ^= 1) {
        const size_t m = n - lo < piece ? n - lo : piece;
        uint8_t* o = base + wbytes + (size_t)idx * out_bytes;
        uint8_t* o_rho = o; uint8_t* o_key = o_rho + r(piece * 32); uint8_t* o_tr = o_key + r(piece * 32);
        uint8_t* o_s1 = o_tr + r(piece * 32); uint8_t* o_s2 = o_s1 + r(piece * L * sb); uint8_t* o_t1 = o_s2 + r(piece * K * sb);
        uint8_t* o_t0 = o_t1 + r(piece * K * 320);
        if (lo >= 2 * piece) A(cudaStreamWaitEvent(st, copied[idx], 0));   // this output set has left for the host
        A(keygen_piece(e, P, base, d_xi + lo * 32, m, o_rho, o_key, o_tr, o_s1, o_s2, o_t1, o_t0, st));
        A(cudaEventRecord(done[idx], st));
        A(cudaStreamWaitEvent(cs, done[idx], 0));
        A(cudaMemcpyAsync(rho + lo * 32, o_rho, m * 32, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(key + lo * 32, o_key, m * 32, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(tr + lo * 32, o_tr, m * 32, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(s1p + lo * L * sb, o_s1, m * L * sb, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(s2p + lo * K * sb, o_s2, m * K * sb, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(t1p + lo * K * 320, o_t1, m * K * 320, cudaMemcpyDeviceToHost, cs));
        A(cudaMemcpyAsync(t0p + lo * K * 416, o_t0, m * K * 416, cudaMemcpyDeviceToHost, cs));
        A(cudaEventRecord(copied[idx], cs));
    }
    cudaError_t e1 = cudaStreamSynchronize(st), e2 = cudaStreamSynchronize(cs);
    for (int i = 0; i < 2; i++) {
        if (done[i]) cudaEventDestroy(done[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
    }
    wipe_free(base, total);   // s1, s2, K, rho' were here
    if (err != cudaSuccess) return fail(e, err, "dil_keygen_batch_host");
    if (e1 != cudaSuccess) return fail(e, e1, "keygen sync");
    if (e2 != cudaSuccess) return fail(e, e2, "keygen D2H sync");
    return DIL_OK;
}

// =======================================================================================
// Verification with one public key PER signature (SURVEY.md §8d cfg4 "per-item rho"): A is expanded
// on chip from rho[i] inside the fused kernel; t1[i] is unpacked, negated, scaled and NTT'd per item.
// =======================================================================================
namespace {
// all pointers device; `work` holds the intermediates laid out by the caller with multi_work_bytes()
struct MultiWork { size_t tr, mu, bad, hm, v, w, t1n, w1p, total; };
MultiWork multi_work_layout(const LevelParams& P, size_t n) {
    Arena ar;
    MultiWork m{};
    const size_t K = P.k, L = P.l;
    m.tr = ar.seg(n * 32); m.mu = ar.seg(n * 64); m.bad = ar.seg(n * 4); m.hm = ar.seg(n * K * 32);
    m.v = ar.seg(n * (L + 1) * 1024); m.w = ar.seg(n * K * 1024); m.t1n = ar.seg(n * K * 1024); m.w1p = ar.seg(n * K * P.w1_bytes);
    m.total = ar.total;
    return m;
}
cudaError_t verify_multi_run(dil_engine* e, const LevelParams& P, uint8_t* work, const MultiWork& m, const uint8_t* d_rho,
                             const uint8_t* d_t1p, const uint8_t* d_msgs, const uint64_t* d_off, size_t n, const uint8_t* d_z,
                             const uint8_t* d_h, const uint8_t* d_ct, uint8_t* d_ok, cudaStream_t st) {
    const uint32_t nn = (uint32_t)n;
    const size_t K = P.k;
    auto P32 = [&](size_t o) { return reinterpret_cast<int32_t*>(work + o); };
    auto PU32 = [&](size_t o) { return reinterpret_cast<uint32_t*>(work + o); };
    auto PU64 = [&](size_t o) { return reinterpret_cast<uint64_t*>(work + o); };
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    // the arena is shared by every multi-key call on this engine: wait for its previous user (possibly another stream)
    if (!e->arena_done) A(cudaEventCreateWithFlags(&e->arena_done, cudaEventDisableTiming));
    else A(cudaStreamWaitEvent(st, e->arena_done, 0));
    A(cudaMemsetAsync(work + m.bad, 0, n * 4, st));
    A(dil::launch_tr_batch(work + m.tr, d_rho, d_t1p, (uint32_t)(K * 320), nn, st));
    A(dil::launch_verify_mu(PU64(m.mu), work + m.tr, d_msgs, d_off, nn, 32, st));
    A(dil::launch_unpack_z(P.level, P32(m.v), PU32(m.bad), d_z, nn, st));
    A(dil::launch_verify_prep(P.level, P32(m.v), PU32(m.hm), PU32(m.bad), d_h, reinterpret_cast<const uint64_t*>(d_ct), nn, st));
    A(dil::launch_unpack_t1neg(P32(m.t1n), d_t1p, n * K, st));
    A(dil::launch_ntt_fwd(P32(m.t1n), P32(m.t1n), n * K, e->sm_count, st));
    // per-key verification keeps the separate UseHint pass: its core is Keccak-bound, one warp (two at level 5) per item, and
    // the fused variant (matvec_item_kernel<..., W1>, kept for A/B) measured 2 % slower at level 5
    A(dil::launch_verify_core_item(P32(m.w), d_rho, P32(m.v), P32(m.t1n), P.level, n, st));
    A(dil::launch_usehint_pack(P.level, PU32(m.w1p), P32(m.w), PU32(m.hm), nn, st));
    A(dil::launch_verify_hash(P.level, d_ok, PU64(m.mu), PU64(m.w1p), reinterpret_cast<const uint64_t*>(d_ct), PU32(m.bad), nn, st));
    if (err == cudaSuccess) A(cudaEventRecord(e->arena_done, st));
    return err;
}
}  // namespace

extern "C" int dil_verify_multi_dev(dil_engine_t* e, int level, const uint8_t* d_rho, const uint8_t* d_t1p, const uint8_t* d_msgs,
                                    const uint64_t* d_offsets, size_t n, const uint8_t* d_z, const uint8_t* d_h,
                                    const uint8_t* d_ctilde, uint8_t* d_ok, void* stream) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!d_rho || !d_t1p || !d_msgs || !d_offsets || !d_z || !d_h || !d_ctilde || !d_ok || n > 0x00FFFFFFu) return DIL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_ctilde) & 7u) || (reinterpret_cast<uintptr_t>(d_z) & 3u)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const LevelParams P = dil::level_params(level);
    const MultiWork m = multi_work_layout(P, n);
    uint8_t* work = nullptr;
    int rc = arena_reserve(e, m.total, &work);
    if (rc) return rc;
    cudaError_t err = verify_multi_run(e, P, work, m, d_rho, d_t1p, d_msgs, d_offsets, n, d_z, d_h, d_ctilde, d_ok, (cudaStream_t)stream);
    if (err != cudaSuccess) return fail(e, err, "dil_verify_multi_dev");
    e->launches += 9;
    return DIL_OK;
}

extern "C" int dil_verify_multi_host(dil_engine_t* e, int level, const uint8_t* rho, const uint8_t* t1p, const uint8_t* msgs,
                                     const uint64_t* offsets, size_t n, const uint8_t* z, const uint8_t* h,
                                     const uint8_t* ctilde, uint8_t* ok) {
    if (!e) return DIL_ERR_ARG;
    if (level != 2 && level != 3 && level != 5) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!rho || !t1p || !msgs || !offsets || !z || !h || !ctilde || !ok || n > 0x00FFFFFFu) return DIL_ERR_ARG;
    if (!offsets_ok(offsets, n)) return DIL_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    DeviceGuard dg(e->device);
    if (!dg.ok) return DIL_ERR_CUDA;
    const LevelParams P = dil::level_params(level);
    cudaStream_t st = e->host_stream;
    const size_t K = P.k, L = P.l, zb = L * P.z_bytes, hb = (size_t)P.omega + P.k, t1b = K * 320;
    const size_t mbytes = offsets[n] > 0 ? offsets[n] : 1;
    const MultiWork m = multi_work_layout(P, n);
    Arena io;
    io.total = m.total;
    const size_t o_rho = io.seg(n * 32), o_t1p = io.seg(n * t1b), o_msgs = io.seg(mbytes), o_off = io.seg((n + 1) * 8), o_z = io.seg(n * zb),
                 o_h = io.seg(n * hb), o_ct = io.seg(n * 32), o_ok = io.seg(n);
    uint8_t* base = nullptr;
    int rc = arena_reserve(e, io.total, &base);
    if (rc) return rc;
    cudaError_t err = cudaSuccess;
    auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
    if (e->arena_done) A(cudaStreamWaitEvent(st, e->arena_done, 0));   // a _dev call on another stream may still use the arena
    A(cudaMemcpyAsync(base + o_rho, rho, n * 32, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_t1p, t1p, n * t1b, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_msgs, msgs, offsets[n], cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_z, z, n * zb, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_h, h, n * hb, cudaMemcpyHostToDevice, st));
    A(cudaMemcpyAsync(base + o_ct, ctilde, n * 32, cudaMemcpyHostToDevice, st));
    if (err == cudaSuccess)
        err = verify_multi_run(e, P, base, m, base + o_rho, base + o_t1p, base + o_msgs, reinterpret_cast<const uint64_t*>(base + o_off), n,
                               base + o_z, base + o_h, base + o_ct, base + o_ok, st);
    A(cudaMemcpyAsync(ok, base + o_ok, n, cudaMemcpyDeviceToHost, st));
    A(cudaStreamSynchronize(st));
    if (err != cudaSuccess) return fail(e, err, "dil_verify_multi_host");
    e->launches += 9;
    return DIL_OK;
}
