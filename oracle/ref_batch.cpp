/*
 * ref_batch.cpp — threaded batch drivers around the UNMODIFIED reference
 * functions (dilithium-256/reference_code/ref_ntt.h:30-36, ref_ntt2x2.h:31-33).
 * TEST/BENCH INFRASTRUCTURE ONLY.  Compiled together with the reference's own
 * sources (taken in place from /root/reference by oracle/Makefile) into
 * oracle/_ref/libdilref.so; it adds nothing to the arithmetic, it only splits a
 * batch of independent polynomials across std::threads so bench.py can time the
 * reference on all host cores (SURVEY.md §8d "CPU baseline").
 */
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

/* the reference's own API (C++ linkage, as compiled by reference_code/Makefile:24,36) */
void ntt(int32_t a[256]);
void invntt(int32_t a[256]);
void pointwise_barrett(int32_t c[256], const int32_t a[256], const int32_t b[256]);
void ntt2x2_ref(int32_t a[256]);
void invntt2x2_ref(int32_t a[256]);
extern const int32_t zetas_barrett[256];

namespace {
template <class F>
void split(size_t n, int threads, F f) {
    if (threads <= 1 || n < 2) { f(0, n); return; }
    std::vector<std::thread> th;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto &t : th) t.join();
}
}  // namespace

extern "C" {
const int32_t *ref_zetas(void) { return zetas_barrett; }
void ref_ntt_batch(int32_t *a, size_t n, int threads) {
    split(n, threads, [=](size_t lo, size_t hi) { for (size_t p = lo; p < hi; p++) ntt(a + p * 256); });
}
void ref_invntt_batch(int32_t *a, size_t n, int threads) {
    split(n, threads, [=](size_t lo, size_t hi) { for (size_t p = lo; p < hi; p++) invntt(a + p * 256); });
}
void ref_ntt2x2_batch(int32_t *a, size_t n, int threads) {
    split(n, threads, [=](size_t lo, size_t hi) { for (size_t p = lo; p < hi; p++) ntt2x2_ref(a + p * 256); });
}
void ref_invntt2x2_batch(int32_t *a, size_t n, int threads) {
    split(n, threads, [=](size_t lo, size_t hi) { for (size_t p = lo; p < hi; p++) invntt2x2_ref(a + p * 256); });
}
void ref_pointwise_batch(int32_t *c, const int32_t *a, const int32_t *b, size_t n, int threads) {
    split(n, threads, [=](size_t lo, size_t hi) {
        for (size_t p = lo; p < hi; p++) pointwise_barrett(c + p * 256, a + p * 256, b + p * 256);
    });
}
/* cfg2 sign-core per item with the reference's own functions:
   NTT(y_j) j<l ; w_i = sum_j A_ij o y_j (pointwise_barrett + add, mod Q) ; INTT(w_i) i<k.
   The reference has no C++ mat-vec; the accumulation below is the
   only non-reference arithmetic and mirrors butterfly.v:144-150. */
void ref_signcore_batch(int32_t *w, int32_t *y, const int32_t *a_hat, int k, int l, size_t batch, int threads) {
    split(batch, threads, [=](size_t lo, size_t hi) {
        int32_t prod[256];
        for (size_t b = lo; b < hi; b++) {
            int32_t *yb = y + b * l * 256, *wb = w + b * k * 256;
            for (int j = 0; j < l; j++) ntt(yb + j * 256);
            for (int i = 0; i < k; i++) {
                int32_t *wi = wb + i * 256;
                for (int j = 0; j < l; j++) {
                    pointwise_barrett(j ? prod : wi, a_hat + (size_t)(i * l + j) * 256, yb + j * 256);
                    if (j) for (int c = 0; c < 256; c++) wi[c] = (wi[c] + prod[c]) % 8380417;
                }
                invntt(wi);
            }
        }
    });
}
}
