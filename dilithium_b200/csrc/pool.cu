// pool.cu — single-process multi-GPU entry points of the C ABI (dil_pool_*).
//
// The reference's sign driver is one program that streams messages into one device and collects signatures
// (rtl_tb/tb_sign_top.v:171-335).  With several B200s in a box the same shape is: one process, one engine per GPU,
// the batch split into contiguous shards (items are independent), every shard signed by its GPU
// through the ordinary host-pointer path on its own host thread, nothing exchanged between GPUs - the "broadcast"
// of the key is the host handing the same packed key to every engine.  A C++ caller therefore needs neither
// torchrun nor NCCL; bench.py's multi-process arm and this pool produce identical signatures.
#include <cuda_runtime.h>

#include <new>
#include <thread>
#include <vector>

#include "dilithium_b200.h"

struct dil_pool {
    std::vector<dil_engine_t*> engines;
    std::vector<int> devices;
};
struct dil_pool_sign_key {
    int level = 0;
    std::vector<dil_sign_key_t*> keys;   // one expanded key per engine
    // the batch between dil_pool_sign_batch_host_begin and dil_pool_sign_batch_finish: which engines carry a shard, and the
    // shards' rebased offset arrays (they must outlive the asynchronous H2D copies)
    std::vector<char> busy;
    std::vector<std::vector<uint64_t>> offs;
};

namespace {
// contiguous shard [lo, hi) of `total` items for member g of G (sizes differ by at most one)
void shard(size_t total, size_t G, size_t g, size_t* lo, size_t* hi) {
    const size_t base = total / G, rem = total % G;
    *lo = g * base + (g < rem ? g : rem);
    *hi = *lo + base + (g < rem ? 1 : 0);
}
}  // namespace

extern "C" {

int dil_pool_create(dil_pool_t** out, const int* devices, int n_devices) {
    if (!out || n_devices < 0) return DIL_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return DIL_ERR_NO_DEVICE;
    dil_pool* p = new (std::nothrow) dil_pool();
    if (!p) return DIL_ERR_ALLOC;
    if (!devices || n_devices == 0) {
        for (int d = 0; d < count; d++) p->devices.push_back(d);
    } else {
        for (int i = 0; i < n_devices; i++) p->devices.push_back(devices[i]);
    }
    for (int d : p->devices) {
        dil_engine_t* e = nullptr;
        int rc = dil_engine_create(&e, d);
        if (rc != DIL_OK) {
            dil_pool_destroy(p);
            return rc;
        }
        p->engines.push_back(e);
    }
    *out = p;
    return DIL_OK;
}

int dil_pool_destroy(dil_pool_t* p) {
    if (!p) return DIL_OK;
    for (dil_engine_t* e : p->engines) dil_engine_destroy(e);
    delete p;
    return DIL_OK;
}

int dil_pool_size(const dil_pool_t* p) { return p ? (int)p->engines.size() : 0; }
dil_engine_t* dil_pool_engine(dil_pool_t* p, int i) { return p && i >= 0 && i < (int)p->engines.size() ? p->engines[i] : nullptr; }

int dil_pool_sign_key_create(dil_pool_t* p, dil_pool_sign_key_t** out, int level, const uint8_t* rho, const uint8_t* key,
                             const uint8_t* tr, const uint8_t* s1p, const uint8_t* s2p, const uint8_t* t0p) {
    if (!p || !out) return DIL_ERR_ARG;
    *out = nullptr;
    dil_pool_sign_key* k = new (std::nothrow) dil_pool_sign_key();
    if (!k) return DIL_ERR_ALLOC;
    k->level = level;
    for (dil_engine_t* e : p->engines) {
        dil_sign_key_t* sk = nullptr;
        int rc = dil_sign_key_create(e, &sk, level, rho, key, tr, s1p, s2p, t0p);
        if (rc != DIL_OK) {
            dil_pool_sign_key_destroy(p, k);
            return rc;
        }
        k->keys.push_back(sk);
    }
    *out = k;
    return DIL_OK;
}

int dil_pool_sign_key_destroy(dil_pool_t* p, dil_pool_sign_key_t* k) {
    if (!k) return DIL_OK;
    for (size_t i = 0; i < k->keys.size(); i++) dil_sign_key_destroy(p && i < p->engines.size() ? p->engines[i] : nullptr, k->keys[i]);
    delete k;
    return DIL_OK;
}

// Asynchronous pair over the pool: every engine begins its shard (dil_sign_batch_host_begin; the enqueueing runs on one short-lived
// host thread per engine so that the GPUs start together), the call returns, and dil_pool_sign_batch_finish completes all shards.
// One batch per pool key at a time; use one pool key per batch in flight.  Same limits as dil_sign_batch_host_begin per shard
// (pinned, device-addressable output buffers; at most 1.25 x host_chunk messages per GPU).
int dil_pool_sign_batch_host_begin(dil_pool_t* p, dil_pool_sign_key_t* k, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                   uint8_t* z, uint8_t* h, uint8_t* ctilde, uint32_t* attempts) {
    if (!p || !k || k->keys.size() != p->engines.size()) return DIL_ERR_ARG;
    if (!msgs || !offsets || !z || !h || !ctilde || n == 0) return DIL_ERR_ARG;
    for (char b : k->busy)
        if (b) return DIL_ERR_ARG;   // a batch is already in flight on this pool key
    size_t zb = 0, hb = 0;
    if (dil_sign_sizes(k->level, &zb, &hb) != DIL_OK) return DIL_ERR_ARG;
    const size_t G = p->engines.size();
    if (offsets[0] != 0) return DIL_ERR_ARG;
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return DIL_ERR_ARG;
    k->busy.assign(G, 0);
    k->offs.assign(G, {});
    for (size_t g = 0; g < G; g++) {
        size_t lo, hi;
        shard(n, G, g, &lo, &hi);
        if (hi == lo) continue;
        k->offs[g].resize(hi - lo + 1);
        for (size_t i = lo; i <= hi; i++) k->offs[g][i - lo] = offsets[i] - offsets[lo];
    }
    std::vector<int> rcs(G, DIL_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        size_t lo, hi;
        shard(n, G, g, &lo, &hi);
        if (hi == lo) continue;
        th.emplace_back([=, &rcs] {
            rcs[g] = dil_sign_batch_host_begin(p->engines[g], k->keys[g], msgs + offsets[lo], k->offs[g].data(), hi - lo, z + lo * zb,
                                               h + lo * hb, ctilde + lo * 32, attempts ? attempts + lo : nullptr);
            k->busy[g] = rcs[g] == DIL_OK;
        });
    }
    for (auto& t : th) t.join();
    int rc = DIL_OK;
    for (int r : rcs)
        if (r != DIL_OK) rc = r;
    if (rc != DIL_OK) dil_pool_sign_batch_finish(p, k);   // let the shards that did start run out
    return rc;
}

int dil_pool_sign_batch_finish(dil_pool_t* p, dil_pool_sign_key_t* k) {
    if (!p || !k || k->keys.size() != p->engines.size()) return DIL_ERR_ARG;
    int rc = DIL_OK;
    for (size_t g = 0; g < k->busy.size(); g++) {
        if (!k->busy[g]) continue;
        const int r = dil_sign_batch_finish(p->engines[g], k->keys[g]);
        if (r != DIL_OK) rc = r;
        k->busy[g] = 0;
    }
    return rc;
}

int dil_pool_sign_batch_host(dil_pool_t* p, dil_pool_sign_key_t* k, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                             uint8_t* z, uint8_t* h, uint8_t* ctilde, uint32_t* attempts) {
    if (!p || !k || k->keys.size() != p->engines.size()) return DIL_ERR_ARG;
    if (n == 0) return DIL_OK;
    if (!msgs || !offsets || !z || !h || !ctilde) return DIL_ERR_ARG;
    size_t zb = 0, hb = 0;
    if (dil_sign_sizes(k->level, &zb, &hb) != DIL_OK) return DIL_ERR_ARG;
    const size_t G = p->engines.size();
    std::vector<int> rcs(G, DIL_OK);
    std::vector<std::vector<uint64_t>> offs(G);
    // every shard sees its own message blob starting at offset 0 (validated before any thread starts)
    if (offsets[0] != 0) return DIL_ERR_ARG;
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return DIL_ERR_ARG;
    for (size_t g = 0; g < G; g++) {
        size_t lo, hi;
        shard(n, G, g, &lo, &hi);
        if (hi == lo) continue;
        offs[g].resize(hi - lo + 1);
        for (size_t i = lo; i <= hi; i++) offs[g][i - lo] = offsets[i] - offsets[lo];
    }
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        size_t lo, hi;
        shard(n, G, g, &lo, &hi);
        if (hi == lo) continue;
        th.emplace_back([=, &rcs, &offs] {
            rcs[g] = dil_sign_batch_host(p->engines[g], k->keys[g], msgs + offsets[lo], offs[g].data(), hi - lo, z + lo * zb, h + lo * hb,
                                         ctilde + lo * 32, attempts ? attempts + lo : nullptr);
        });
    }
    for (auto& t : th) t.join();
    for (int rc : rcs)
        if (rc != DIL_OK) return rc;
    return DIL_OK;
}

}  // extern "C"
