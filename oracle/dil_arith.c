/*
 * dil_arith.c — oracle ring arithmetic over Z_Q[X]/(X^256+1), Q = 8380417.
 * TEST INFRASTRUCTURE ONLY (see dil_oracle.h).
 *
 * Restates, with canonical [0,Q) outputs:
 *   ntt               dilithium-256/reference_code/ref_ntt.cpp:28-47
 *   pointwise         dilithium-256/reference_code/ref_ntt.cpp:49-57
 *   invntt            dilithium-256/reference_code/ref_ntt.cpp:59-87
 *   ntt2x2/invntt2x2  dilithium-256/reference_code/ref_ntt2x2.cpp:37-82, :100-145
 *   mul-acc/add/sub   rtl_src/butterfly.v:144-164, :224-237
 * The twiddle table is regenerated from zeta = 1753 (consts.cpp:64-97 holds
 * zeta^brv8(k) centred); tests compare it with the reference's table.
 */
#include "dil_oracle.h"

#define Q ORC_Q
#define N ORC_N

static int32_t zetas_tab[N];
static int zetas_ready;

static inline int32_t canon(int64_t x) {
    x %= Q;
    return (int32_t)(x < 0 ? x + Q : x);
}
static inline int32_t mulq(int64_t a, int64_t b) { return canon(a * b); }

static unsigned brv8(unsigned k) {
    unsigned r = 0;
    for (int b = 0; b < 8; b++) r |= ((k >> b) & 1u) << (7 - b);
    return r;
}

const int32_t *orc_zetas(void) {
    if (!zetas_ready) {
        /* powers of the 512-th root of unity 1753 */
        int32_t pw[N];
        pw[0] = 1;
        for (int e = 1; e < N; e++) pw[e] = mulq(pw[e - 1], 1753);
        zetas_tab[0] = 0; /* slot 0 unused, reference stores 0 there */
        for (unsigned k = 1; k < N; k++) {
            int32_t z = pw[brv8(k)];
            zetas_tab[k] = z > Q / 2 ? z - Q : z; /* centred like consts.cpp */
        }
        zetas_ready = 1;
    }
    return zetas_tab;
}

/* Forward: layer s (0..7) has span = 128>>s and 2^s blocks; block m uses
   twiddle index 2^s + m  (ref_ntt.cpp:34-38, pre-incremented k). */
void orc_ntt(int32_t a[N]) {
    const int32_t *z = orc_zetas();
    for (int i = 0; i < N; i++) a[i] = canon(a[i]);
    for (int s = 0; s < 8; s++) {
        int span = 128 >> s;
        for (int m = 0; m < (1 << s); m++) {
            int32_t w = z[(1 << s) + m];
            int base = m * 2 * span;
            for (int j = base; j < base + span; j++) {
                int32_t t = mulq(w, a[j + span]);
                int32_t u = a[j];
                a[j] = canon((int64_t)u + t);
                a[j + span] = canon((int64_t)u - t);
            }
        }
    }
}

/* Inverse: spans 1,2,..,128; for span the blocks count down the table from
   index 2*nblocks-1 and use the negated twiddle (ref_ntt.cpp:66-72); final
   scaling by 256^-1 = 8347681 (ref_ntt.cpp:64, :83-86). */
void orc_invntt(int32_t a[N]) {
    const int32_t *z = orc_zetas();
    for (int i = 0; i < N; i++) a[i] = canon(a[i]);
    for (int span = 1; span < N; span <<= 1) {
        int nblocks = N / (2 * span);
        for (int m = 0; m < nblocks; m++) {
            int32_t w = -z[2 * nblocks - 1 - m];
            int base = m * 2 * span;
            for (int j = base; j < base + span; j++) {
                int32_t u = a[j], v = a[j + span];
                a[j] = canon((int64_t)u + v);
                a[j + span] = mulq(w, (int64_t)u - v);
            }
        }
    }
    for (int i = 0; i < N; i++) a[i] = mulq(a[i], 8347681);
}

/* Radix-2x2 forward: four passes, each fusing two layers on quadruples
   (j, j+q, j+2q, j+3q), q = quarter of the block; twiddle indices
   k1=(256+i)>>l, k2=(256+i)>>(l-1), k2+1 (ref_ntt2x2.cpp:46-79). */
void orc_ntt2x2(int32_t a[N]) {
    const int32_t *z = orc_zetas();
    for (int i = 0; i < N; i++) a[i] = canon(a[i]);
    for (int l = 8; l > 0; l -= 2) {
        int quarter = 1 << (l - 2);
        for (int blk = 0; blk < N; blk += 1 << l) {
            int32_t w1 = z[(N + blk) >> l];
            int32_t w2lo = z[(N + blk) >> (l - 1)];
            int32_t w2hi = z[((N + blk) >> (l - 1)) + 1];
            for (int j = blk; j < blk + quarter; j++) {
                int64_t p0 = a[j], p1 = a[j + quarter], p2 = a[j + 2 * quarter], p3 = a[j + 3 * quarter];
                /* layer A: (p0,p2) and (p1,p3) with w1 */
                int32_t t0 = mulq(w1, p2), t1 = mulq(w1, p3);
                int64_t e0 = canon(p0 + t0), e2 = canon(p0 - t0);
                int64_t e1 = canon(p1 + t1), e3 = canon(p1 - t1);
                /* layer B: (e0,e1) with w2lo, (e2,e3) with w2hi */
                int32_t u0 = mulq(w2lo, e1), u1 = mulq(w2hi, e3);
                a[j] = canon(e0 + u0);
                a[j + quarter] = canon(e0 - u0);
                a[j + 2 * quarter] = canon(e2 + u1);
                a[j + 3 * quarter] = canon(e2 - u1);
            }
        }
    }
}

/* exact halving mod Q: (t odd) ? (t>>1) + (Q+1)/2 : t>>1   (ref_ntt2x2.cpp:91) */
static inline int32_t half(int32_t t) { return (t & 1) ? (t >> 1) + (Q + 1) / 2 : (t >> 1); }

/* Radix-2x2 inverse with per-butterfly halving instead of the final 256^-1
   (ref_ntt2x2.cpp:100-145; RTL butterfly.v:214-222). */
void orc_invntt2x2(int32_t a[N]) {
    const int32_t *z = orc_zetas();
    for (int i = 0; i < N; i++) a[i] = canon(a[i]);
    for (int l = 0; l < 8; l += 2) {
        int span = 1 << l;
        for (int blk = 0; blk < N; blk += 4 * span) {
            int ka = ((N - blk / 2) >> l) - 1;
            int kb = ((N - blk / 2) >> (l + 1)) - 1;
            int32_t wa0 = -z[ka], wa1 = -z[ka - 1], wb = -z[kb];
            for (int j = blk; j < blk + span; j++) {
                int32_t p0 = a[j], p1 = a[j + span], p2 = a[j + 2 * span], p3 = a[j + 3 * span];
                /* layer A: (p0,p1) with wa0, (p2,p3) with wa1 */
                int32_t s0 = half(canon((int64_t)p0 + p1)), d0 = mulq(wa0, half(canon((int64_t)p0 - p1)));
                int32_t s1 = half(canon((int64_t)p2 + p3)), d1 = mulq(wa1, half(canon((int64_t)p2 - p3)));
                /* layer B: (s0,s1) and (d0,d1) with wb */
                a[j] = half(canon((int64_t)s0 + s1));
                a[j + 2 * span] = mulq(wb, half(canon((int64_t)s0 - s1)));
                a[j + span] = half(canon((int64_t)d0 + d1));
                a[j + 3 * span] = mulq(wb, half(canon((int64_t)d0 - d1)));
            }
        }
    }
}

void orc_pointwise(int32_t c[N], const int32_t a[N], const int32_t b[N]) {
    for (int i = 0; i < N; i++) c[i] = mulq(a[i], b[i]);
}
/* butterfly.v MULT mode: out = acc + a*b  (butterfly.v:144-150, :224-230) */
void orc_pointwise_acc(int32_t c[N], const int32_t a[N], const int32_t b[N]) {
    for (int i = 0; i < N; i++) c[i] = canon((int64_t)canon(c[i]) + mulq(a[i], b[i]));
}
void orc_add(int32_t c[N], const int32_t a[N], const int32_t b[N]) {
    for (int i = 0; i < N; i++) c[i] = canon((int64_t)a[i] + b[i]);
}
void orc_sub(int32_t c[N], const int32_t a[N], const int32_t b[N]) {
    for (int i = 0; i < N; i++) c[i] = canon((int64_t)a[i] - b[i]);
}

void orc_ntt_batch(int32_t *a, size_t n) {
    for (size_t p = 0; p < n; p++) orc_ntt(a + p * N);
}
void orc_invntt_batch(int32_t *a, size_t n) {
    for (size_t p = 0; p < n; p++) orc_invntt(a + p * N);
}
void orc_pointwise_batch(int32_t *c, const int32_t *a, const int32_t *b, size_t n) {
    for (size_t p = 0; p < n; p++) orc_pointwise(c + p * N, a + p * N, b + p * N);
}
