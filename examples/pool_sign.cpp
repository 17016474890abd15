// pool_sign.cpp — one process, every GPU of the box: the shape of the reference's sign driver (rtl_tb/tb_sign_top.v:171-335:
// one program feeds messages and collects signatures) scaled out with dil_pool_* - no torchrun, no NCCL.
// Signs `steps` batches of n_per_gpu * G messages (32-byte synthetic messages, KAT key `key_index` read from the
// reference's KAT files) from pinned host memory into pinned host memory, checks that the pool's signatures equal the
// ones a single engine produces for a sample shard, and prints one JSON line with the end-to-end rate.
//
// `in_flight` batches are signed at the same time (one pool key and one set of output buffers per batch in flight, all driven by
// this one thread through dil_pool_sign_batch_host_begin / dil_pool_sign_batch_finish): the next batch signs while the last
// rejection rounds of the previous one leave SMs idle and its last signatures drain.
//
//   usage: pool_sign <KAT dir> [level=2] [n_per_gpu=65536] [steps=5] [gpus=all] [in_flight=2]
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "dilithium_b200.h"

using Bytes = std::vector<uint8_t>;

static Bytes read_hex_line(const std::string& path, size_t index) {
    std::ifstream f(path);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    std::string line;
    auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
    for (size_t i = 0; std::getline(f, line); ) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        if (i++ < index) continue;
        Bytes b(line.size() / 2);
        for (size_t j = 0; j < b.size(); j++) b[j] = (uint8_t)(nib(line[2 * j]) * 16 + nib(line[2 * j + 1]));
        return b;
    }
    std::fprintf(stderr, "%s: no line %zu\n", path.c_str(), index);
    std::exit(2);
}

#define CK(call) do { int rc_ = (call); if (rc_ != DIL_OK) { std::fprintf(stderr, "%s failed: %s\n", #call, dil_status_string(rc_)); return 3; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 3; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: pool_sign <KAT dir> [level] [n_per_gpu] [steps] [gpus] [in_flight]\n"); return 2; }
    const std::string dir = argv[1];
    const int level = argc > 2 ? std::atoi(argv[2]) : 2;
    const size_t per_gpu = argc > 3 ? (size_t)std::atol(argv[3]) : 65536;
    const int steps = argc > 4 ? std::atoi(argv[4]) : 5;
    int gpus = argc > 5 ? std::atoi(argv[5]) : 0;
    const int T = argc > 6 && std::atoi(argv[6]) > 0 ? std::atoi(argv[6]) : 2;
    const std::string sfx = "_" + std::to_string(level) + ".txt";
    const size_t key_index = 0;
    Bytes rho = read_hex_line(dir + "/rho" + sfx, key_index), key = read_hex_line(dir + "/k" + sfx, key_index),
          tr = read_hex_line(dir + "/tr" + sfx, key_index), s1 = read_hex_line(dir + "/s1" + sfx, key_index),
          s2 = read_hex_line(dir + "/s2" + sfx, key_index), t0 = read_hex_line(dir + "/t0" + sfx, key_index);
    std::vector<int> devs;
    for (int d = 0; d < gpus; d++) devs.push_back(d);
    dil_pool_t* pool = nullptr;
    CK(dil_pool_create(&pool, devs.empty() ? nullptr : devs.data(), (int)devs.size()));
    const size_t G = (size_t)dil_pool_size(pool), n = per_gpu * G, mlen = 32;
    std::vector<dil_pool_sign_key_t*> pks(T, nullptr);
    for (int t = 0; t < T; t++)
        CK(dil_pool_sign_key_create(pool, &pks[t], level, rho.data(), key.data(), tr.data(), s1.data(), s2.data(), t0.data()));
    size_t zb = 0, hb = 0;
    CK(dil_sign_sizes(level, &zb, &hb));
    // pinned, portable buffers: every engine streams its shard into its slice; one output set per batch in flight
    uint8_t* msgs;
    uint64_t* off;
    std::vector<uint8_t*> zs(T), hs(T), cts(T);
    std::vector<uint32_t*> atts(T);
    CU(cudaHostAlloc((void**)&msgs, n * mlen, cudaHostAllocPortable | cudaHostAllocMapped));
    CU(cudaHostAlloc((void**)&off, (n + 1) * 8, cudaHostAllocPortable | cudaHostAllocMapped));
    for (int t = 0; t < T; t++) {
        CU(cudaHostAlloc((void**)&zs[t], n * zb, cudaHostAllocPortable | cudaHostAllocMapped));
        CU(cudaHostAlloc((void**)&hs[t], n * hb, cudaHostAllocPortable | cudaHostAllocMapped));
        CU(cudaHostAlloc((void**)&cts[t], n * 32, cudaHostAllocPortable | cudaHostAllocMapped));
        CU(cudaHostAlloc((void**)&atts[t], n * 4, cudaHostAllocPortable | cudaHostAllocMapped));
    }
    uint8_t *z = zs[0], *h = hs[0], *ct = cts[0];
    uint32_t* att = atts[0];
    uint64_t x = 0x44494C32ull;
    for (size_t i = 0; i < n * mlen; i++) { x = x * 6364136223846793005ull + 1442695040888963407ull; msgs[i] = (uint8_t)(x >> 56); }
    for (size_t i = 0; i <= n; i++) off[i] = i * mlen;
    for (int t = 0; t < T; t++)   // warm-up (workspaces, kernel attributes)
        CK(dil_pool_sign_batch_host(pool, pks[t], msgs, off, n, zs[t], hs[t], cts[t], atts[t]));
    // ONE host thread keeps T batches in flight: step s is begun on pool key s mod T as soon as that key's previous batch is finished
    auto t0c = std::chrono::steady_clock::now();
    if (T == 1) {
        for (int s = 0; s < steps; s++) CK(dil_pool_sign_batch_host(pool, pks[0], msgs, off, n, zs[0], hs[0], cts[0], atts[0]));
    } else {
        for (int s = 0; s < steps; s++) {
            const int t = s % T;
            if (s >= T) CK(dil_pool_sign_batch_finish(pool, pks[t]));
            CK(dil_pool_sign_batch_host_begin(pool, pks[t], msgs, off, n, zs[t], hs[t], cts[t], atts[t]));
        }
        for (int s = steps > T ? steps - T : 0; s < steps; s++) CK(dil_pool_sign_batch_finish(pool, pks[s % T]));
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0c).count();
    bool sets_equal = true;
    for (int t = 1; t < T && t < steps; t++)
        sets_equal = sets_equal && !std::memcmp(zs[t], z, n * zb) && !std::memcmp(hs[t], h, n * hb) && !std::memcmp(cts[t], ct, n * 32);
    // cross-check: the last shard, signed alone on engine 0 (pageable outputs -> copy path), must be identical
    const size_t lo = (G - 1) * per_gpu, m = per_gpu < 4096 ? per_gpu : 4096;
    dil_sign_key_t* k0 = nullptr;
    dil_engine_t* e0 = dil_pool_engine(pool, 0);
    CK(dil_sign_key_create(e0, &k0, level, rho.data(), key.data(), tr.data(), s1.data(), s2.data(), t0.data()));
    std::vector<uint8_t> z1(m * zb), h1(m * hb), c1(m * 32);
    std::vector<uint32_t> a1(m);
    std::vector<uint64_t> o1(m + 1);
    for (size_t i = 0; i <= m; i++) o1[i] = i * mlen;
    CK(dil_sign_batch_host(e0, k0, msgs + lo * mlen, o1.data(), m, z1.data(), h1.data(), c1.data(), a1.data()));
    const bool same = sets_equal && !std::memcmp(z1.data(), z + lo * zb, m * zb) && !std::memcmp(h1.data(), h + lo * hb, m * hb) &&
                      !std::memcmp(c1.data(), ct + lo * 32, m * 32) && !std::memcmp(a1.data(), att + lo, m * 4);
    double att_mean = 0;
    for (size_t i = 0; i < n; i++) att_mean += att[i];
    std::printf("{\"tool\": \"pool_sign\", \"level\": %d, \"n_gpus\": %zu, \"batch\": %zu, \"steps\": %d, \"in_flight\": %d, \"signs_per_s\": %.1f, "
                "\"ms_per_batch\": %.3f, \"gb_per_s_to_host\": %.2f, \"mean_attempts\": %.4f, \"matches_single_engine\": %s}\n",
                level, G, n, steps, T, n * steps / sec, sec / steps * 1e3, n * steps * (double)(zb + hb + 36) / sec / 1e9, att_mean / n,
                same ? "true" : "false");
    dil_sign_key_destroy(e0, k0);
    for (auto* pk : pks) dil_pool_sign_key_destroy(pool, pk);
    dil_pool_destroy(pool);
    return same ? 0 : 1;
}
