// ref_shim.cpp — drop-in definitions of the reference's C++ API
//   void ntt(data_t a[256]); void invntt(data_t a[256]);
//   void pointwise_barrett(data_t c[256], const data_t a[256], const data_t b[256]);
//                                         (dilithium-256/reference_code/ref_ntt.h:30-36)
//   void ntt2x2_ref(data_t a[256]); void invntt2x2_ref(data_t a[256]);
//                                         (dilithium-256/reference_code/ref_ntt2x2.h:31-33)
//   extern const data_t zetas_barrett[256];                (dilithium-256/consts.h:30)
// with the reference's C++ linkage (mangled _Z3nttPi, _Z6invnttPi, _Z17pointwise_barrettPiPKiS1_,
// _Z10ntt2x2_refPi, _Z13invntt2x2_refPi), each a batch-of-one call into the B200 engine through
// the C ABI.  Linking the reference's own test mains (ref_test_ntt_ntt2x2.cpp,
// hardware_code/ntt2x2_test.cpp) against libdilithium_b200_shim.so instead of ref_ntt.cpp /
// ref_ntt2x2.cpp / consts.cpp runs them on the GPU unchanged.
//
// Results are canonical in [0,Q) where the reference returns signed representatives; the
// reference's callers compare mod Q (ref_test_ntt_ntt2x2.cpp:31-42, util.cpp:97-114).
// The functions are `void` like the reference's: an engine failure aborts with a message.
// Batch-of-one is a compatibility path, not a fast one (one H2D/D2H round trip per call);
// batched callers should use dil_*_host / dil_*_dev directly.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "dil_field.cuh"
#include "dilithium_b200.h"

namespace {
dil_engine_t* g_engine = nullptr;
std::once_flag g_once;

void die(const char* what, int rc) {
    std::fprintf(stderr, "dilithium_b200 shim: %s failed: %s (%s)\n", what, dil_status_string(rc),
                 g_engine ? dil_last_error(g_engine) : "");
    std::abort();
}
dil_engine_t* engine() {
    std::call_once(g_once, [] {
        const char* dev = std::getenv("DIL_DEVICE");
        int rc = dil_engine_create(&g_engine, dev ? std::atoi(dev) : 0);
        if (rc != DIL_OK) die("dil_engine_create", rc);
    });
    return g_engine;
}

struct ZetaTable {
    int32_t v[256];
};
constexpr ZetaTable make_zetas() {
    ZetaTable t{};
    for (unsigned k = 1; k < 256; k++) {
        uint32_t z = dil::zeta_fwd(k);
        t.v[k] = z > dil::Q / 2 ? (int32_t)z - (int32_t)dil::Q : (int32_t)z;  // centred, as consts.cpp:64-97
    }
    return t;
}
constexpr ZetaTable kZetas = make_zetas();
}  // namespace

// data symbol of consts.h:30 (unmangled at namespace scope)
extern const int32_t zetas_barrett[256];
const int32_t zetas_barrett[256] = {
#define Z4(i) kZetas.v[i], kZetas.v[i + 1], kZetas.v[i + 2], kZetas.v[i + 3]
#define Z16(i) Z4(i), Z4(i + 4), Z4(i + 8), Z4(i + 12)
#define Z64(i) Z16(i), Z16(i + 16), Z16(i + 32), Z16(i + 48)
    Z64(0), Z64(64), Z64(128), Z64(192)};

// The reference reduces with a signed 64-bit % (ref_ntt.cpp:41-43) and therefore accepts ANY int32 coefficient, e.g.
// unreduced sums of residues; the engine's lazy arithmetic expects representatives in (-Q, Q).  The shim brings
// arbitrary inputs into that range on the host so that a reference caller can never get a silently wrong result.
static void reduce_in(int32_t* dst, const int32_t* src) {
    for (int i = 0; i < 256; i++) dst[i] = src[i] % (int32_t)dil::Q;   // C++ %: result in (-Q, Q)
}
void ntt(int32_t a[256]) {
    reduce_in(a, a);
    int rc = dil_ntt_host(engine(), a, 1);
    if (rc) die("ntt", rc);
}
void invntt(int32_t a[256]) {
    reduce_in(a, a);
    int rc = dil_invntt_host(engine(), a, 1);
    if (rc) die("invntt", rc);
}
void pointwise_barrett(int32_t c[256], const int32_t a[256], const int32_t b[256]) {
    int32_t ra[256], rb[256];
    reduce_in(ra, a);
    reduce_in(rb, b);
    int rc = dil_pointwise_host(engine(), c, ra, rb, 1);
    if (rc) die("pointwise_barrett", rc);
}
// The radix-2x2 schedule IS how the engine computes every transform (ntt_core.cuh); results
// equal ntt()/invntt() mod Q exactly as the reference's own differential test asserts.
void ntt2x2_ref(int32_t a[256]) { ntt(a); }
void invntt2x2_ref(int32_t a[256]) { invntt(a); }

// north_star aliases with the reference's calling style (plain domain)
void invntt_tomont(int32_t a[256]) { invntt(a); }
void poly_pointwise(int32_t c[256], const int32_t a[256], const int32_t b[256]) { pointwise_barrett(c, a, b); }
// w[k][256] = A_hat[k*l][256] * v[l][256]  (MULT_MODE loop nest, combined_top.v:921-958)
void polyvec_matrix_pointwise(int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l) {
    int rc = dil_matvec_host(engine(), w, a_hat, v, k, l, 1);
    if (rc) die("polyvec_matrix_pointwise", rc);
}
