/*
 * ref_bridge.cpp - routes the oracle's ring arithmetic (orc_ntt / orc_invntt / orc_pointwise ...)
 * to the UNMODIFIED reference functions ntt() / invntt() / pointwise_barrett()
 * (dilithium-256/reference_code/ref_ntt.h:30-36), canonicalising their signed outputs mod Q.
 * TEST/BENCH INFRASTRUCTURE ONLY.  Linked, instead of dil_arith.c, into oracle/_ref/libdilref.so
 * together with the oracle's scheme glue (dil_scheme.c, dil_expand.c, dil_keccak.c) so that the CPU
 * baseline for full signing runs the reference's own C++ on every NTT, inverse NTT and pointwise
 * product - "the reference_code C++ path timed on the host cores" (BASELINE.json north_star).
 * The reference has no C/C++ keygen/sign/verify; only that glue is a port.
 */
#include <cstddef>
#include <cstdint>

void ntt(int32_t a[256]);
void invntt(int32_t a[256]);
void pointwise_barrett(int32_t c[256], const int32_t a[256], const int32_t b[256]);
void ntt2x2_ref(int32_t a[256]);
void invntt2x2_ref(int32_t a[256]);
extern const int32_t zetas_barrett[256];

namespace {
constexpr int32_t Q = 8380417;
inline int32_t canon(int64_t x) { x %= Q; return (int32_t)(x < 0 ? x + Q : x); }
inline void canon_poly(int32_t a[256]) { for (int i = 0; i < 256; i++) a[i] = canon(a[i]); }
}

extern "C" {
const int32_t *orc_zetas(void) { return zetas_barrett; }
void orc_ntt(int32_t a[256]) { ntt(a); canon_poly(a); }
void orc_invntt(int32_t a[256]) { invntt(a); canon_poly(a); }
void orc_ntt2x2(int32_t a[256]) { ntt2x2_ref(a); canon_poly(a); }
void orc_invntt2x2(int32_t a[256]) { invntt2x2_ref(a); canon_poly(a); }
void orc_pointwise(int32_t c[256], const int32_t a[256], const int32_t b[256]) { pointwise_barrett(c, a, b); canon_poly(c); }
void orc_pointwise_acc(int32_t c[256], const int32_t a[256], const int32_t b[256]) {
    int32_t t[256];
    pointwise_barrett(t, a, b);
    for (int i = 0; i < 256; i++) c[i] = canon((int64_t)c[i] + t[i]);
}
void orc_add(int32_t c[256], const int32_t a[256], const int32_t b[256]) { for (int i = 0; i < 256; i++) c[i] = canon((int64_t)a[i] + b[i]); }
void orc_sub(int32_t c[256], const int32_t a[256], const int32_t b[256]) { for (int i = 0; i < 256; i++) c[i] = canon((int64_t)a[i] - b[i]); }
void orc_ntt_batch(int32_t *a, size_t n) { for (size_t p = 0; p < n; p++) orc_ntt(a + p * 256); }
void orc_invntt_batch(int32_t *a, size_t n) { for (size_t p = 0; p < n; p++) orc_invntt(a + p * 256); }
void orc_pointwise_batch(int32_t *c, const int32_t *a, const int32_t *b, size_t n) {
    for (size_t p = 0; p < n; p++) orc_pointwise(c + p * 256, a + p * 256, b + p * 256);
}
}
