// kat_driver.cpp — host-side C++ driver over the C ABI that replays the reference's three
// testbenches on the B200 engine:
//     rtl_tb/tb_keygen_top.v  (seed -> rho, K, s1, s2, t1, t0, tr;  :145-275)
//     rtl_tb/tb_sign_top.v    (rho, mlen, tr, M, K, s1, s2, t0 -> z, h, c~;  :171-335)
//     rtl_tb/tb_verify_top.v  (rho, c~, z, t1, mlen, M, h -> accept/reject;  :144-249)
// It reads the KAT files in the reference's own format (KAT/<field>_<level>.txt, one upper-case hex
// vector per line, $readmemh-style as tb_sign_top.v:110-140 loads them), runs all vectors of a level
// as ONE batch per operation where the scheme allows it, compares every output field and prints
// one line per testbench in the style of the reference ("KG2 ... completed").
//
//   usage: kat_driver <KAT dir> [level=2] [num_vectors=100]
//
// Exit code 0 iff every vector matches.  No CPU arithmetic happens here: all math is in
// libdilithium_b200.so.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "dilithium_b200.h"

using Bytes = std::vector<uint8_t>;

static std::vector<Bytes> read_hex(const std::string& path, size_t limit) {
    std::ifstream f(path);
    if (!f) {
        std::fprintf(stderr, "cannot open %s\n", path.c_str());
        std::exit(2);
    }
    std::vector<Bytes> rows;
    std::string line;
    auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
    while (rows.size() < limit && std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        Bytes b(line.size() / 2);
        for (size_t i = 0; i < b.size(); i++) b[i] = (uint8_t)(nib(line[2 * i]) * 16 + nib(line[2 * i + 1]));
        rows.push_back(std::move(b));
    }
    return rows;
}

static Bytes flat(const std::vector<Bytes>& rows) {
    Bytes out;
    for (auto& r : rows) out.insert(out.end(), r.begin(), r.end());
    return out;
}

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != DIL_OK) {                                                                \
            std::fprintf(stderr, "%s failed: %s (%s)\n", #call, dil_status_string(rc_), dil_last_error(eng)); \
            return 3;                                                                       \
        }                                                                                   \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <KAT dir> [level] [num_vectors]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    const int level = argc > 2 ? std::atoi(argv[2]) : 2;
    const size_t nv = argc > 3 ? (size_t)std::atoi(argv[3]) : 100;
    auto file = [&](const char* stem) { return dir + "/" + stem + "_" + std::to_string(level) + ".txt"; };

    dil_engine_t* eng = nullptr;
    if (dil_engine_create(&eng, 0) != DIL_OK) {
        std::fprintf(stderr, "no usable B200: the engine has no CPU fallback\n");
        return 3;
    }
    int k = 0, l = 0;
    CHECK(dil_level_dims(level, &k, &l));
    size_t zb = 0, hb = 0;
    CHECK(dil_sign_sizes(level, &zb, &hb));

    auto seeds = read_hex(file("z"), nv), rho = read_hex(file("rho"), nv), key = read_hex(file("k"), nv), tr = read_hex(file("tr"), nv);
    auto s1 = read_hex(file("s1"), nv), s2 = read_hex(file("s2"), nv), t1 = read_hex(file("t1"), nv), t0 = read_hex(file("t0"), nv);
    auto zs = read_hex(file("zs"), nv), hh = read_hex(file("h"), nv), cc = read_hex(file("c"), nv);
    auto msgs = read_hex(file("m"), nv), mlen = read_hex(file("mlen"), nv);
    const size_t n = seeds.size();
    int failures = 0;

    // ---- keygen: all seeds as one batch (tb_keygen_top.v) ----
    {
        Bytes xi = flat(seeds);
        Bytes o_rho(n * 32), o_key(n * 32), o_tr(n * 32), o_s1(n * s1[0].size()), o_s2(n * s2[0].size()), o_t1(n * t1[0].size()),
            o_t0(n * t0[0].size());
        CHECK(dil_keygen_batch_host(eng, level, xi.data(), n, o_rho.data(), o_key.data(), o_tr.data(), o_s1.data(), o_s2.data(),
                                    o_t1.data(), o_t0.data()));
        int bad = 0;
        bad += o_rho != flat(rho); bad += o_key != flat(key); bad += o_tr != flat(tr); bad += o_s1 != flat(s1);
        bad += o_s2 != flat(s2); bad += o_t1 != flat(t1); bad += o_t0 != flat(t0);
        std::printf("KG%d[0..%zu] %s (rho, K, s1, s2, t1, t0, tr)\n", level, n - 1, bad ? "MISMATCH" : "completed, all fields match");
        failures += bad;
    }
    // ---- sign + verify: one key per vector (the KATs use a fresh key per message) ----
    int sign_bad = 0, verify_bad = 0;
    unsigned long attempts_total = 0;
    for (size_t i = 0; i < n; i++) {
        const size_t ml = ((size_t)mlen[i][0] << 8) | mlen[i][1];         // mlen_*.txt: 16-bit big-endian hex
        uint64_t off[2] = {0, ml};
        dil_sign_key_t* sk = nullptr;
        CHECK(dil_sign_key_create(eng, &sk, level, rho[i].data(), key[i].data(), tr[i].data(), s1[i].data(), s2[i].data(), t0[i].data()));
        Bytes z(zb), h(hb), c(32);
        uint32_t att = 0;
        CHECK(dil_sign_batch_host(eng, sk, msgs[i].data(), off, 1, z.data(), h.data(), c.data(), &att));
        attempts_total += att;
        if (z != zs[i] || h != hh[i] || c != cc[i]) sign_bad++;
        dil_sign_key_destroy(eng, sk);

        dil_verify_key_t* vk = nullptr;
        CHECK(dil_verify_key_create(eng, &vk, level, rho[i].data(), t1[i].data()));
        uint8_t ok = 0;
        CHECK(dil_verify_batch_host(eng, vk, msgs[i].data(), off, 1, zs[i].data(), hh[i].data(), cc[i].data(), &ok));
        if (ok != 1) verify_bad++;
        Bytes zt = zs[i];
        zt[zt.size() / 2] ^= 0x08;                                          // a tampered signature must be rejected
        CHECK(dil_verify_batch_host(eng, vk, msgs[i].data(), off, 1, zt.data(), hh[i].data(), cc[i].data(), &ok));
        if (ok != 0) verify_bad++;
        dil_verify_key_destroy(eng, vk);
    }
    std::printf("SIGN%d[0..%zu] %s (z, h, c~), mean attempts %.2f\n", level, n - 1, sign_bad ? "MISMATCH" : "completed, all fields match",
                (double)attempts_total / (double)n);
    std::printf("VERIFY%d[0..%zu] %s\n", level, n - 1, verify_bad ? "MISMATCH" : "completed, all accepted, all tampered rejected");
    failures += sign_bad + verify_bad;
    std::printf("engine kernels launched: %llu\n", (unsigned long long)dil_engine_launch_count(eng));
    dil_engine_destroy(eng);
    return failures ? 1 : 0;
}
