/*
 * dil_scheme.c — oracle restatement of the Dilithium (round-3.1, deterministic)
 * keygen / sign / verify data flow, used ONLY to pin the polynomial-arithmetic
 * hot path (ExpandA, NTT, mat-vec, INTT) end-to-end against the reference's
 * KAT/ vectors.  TEST INFRASTRUCTURE ONLY (see dil_oracle.h).
 *
 * Reference pointers (GMUCERG/Dilithium):
 *   parameter table      rtl_src/combined_top.v:520-551, norm_check.v:43-51,
 *                        gen_c.v:107-124, makehint.v:48-55
 *   bit packing          rtl_src/decoder.v:89-143, encoder.v:96-133,
 *                        uncenter_coeff.v:51-64  (s: eta-s, t0: 2^12-t0, z: gamma1-z)
 *   Power2Round          rtl_src/uncenter_coeff.v:54-55
 *   Decompose            rtl_src/decomp_map1.v, coeff_decomposer.v:80-88
 *   MakeHint / UseHint   rtl_src/makehint.v:99-102, usehint.v:134-155
 *   SampleInBall         rtl_src/gen_c.v:192-222, :317-343
 *   ExpandMask           rtl_src/expandmask_ext.v:98,:131-185,:284-294; rejection_y.v:76-99
 *   eta sampler          rtl_src/gen_s.v:115; sampler_s.v:117-135; rejection_s.v:47-51,:85-138
 *   keygen flow          rtl_src/combined_top.v:754-1079
 *   verify flow          rtl_src/combined_top.v:1080-1534
 *   sign flow            rtl_src/combined_top.v:1535-2229 (checks :2098-2101,
 *                        :2144-2147, :2176-2179; restart :2217-2228)
 * All small-norm polynomials are kept as centred signed values here; NTT-domain
 * values are canonical.
 */
#include <stdlib.h>
#include <string.h>
#include "dil_oracle.h"

#define Q ORC_Q
#define N ORC_N
#define D 13

int orc_params(orc_params_t *p, int level) {
    memset(p, 0, sizeof *p);
    p->level = level;
    switch (level) {
    case 2: p->k = 4; p->l = 4; p->eta = 2; p->tau = 39; p->gamma1_bits = 17; p->gamma2 = (Q - 1) / 88; p->beta = 78; p->omega = 80; break;
    case 3: p->k = 6; p->l = 5; p->eta = 4; p->tau = 49; p->gamma1_bits = 19; p->gamma2 = (Q - 1) / 32; p->beta = 196; p->omega = 55; break;
    case 5: p->k = 8; p->l = 7; p->eta = 2; p->tau = 60; p->gamma1_bits = 19; p->gamma2 = (Q - 1) / 32; p->beta = 120; p->omega = 75; break;
    default: return -1;
    }
    p->gamma1 = 1 << p->gamma1_bits;
    p->z_bytes = N * (p->gamma1_bits + 1) / 8;
    p->w1_bytes = p->gamma2 == (Q - 1) / 88 ? 192 : 128;
    p->s_bytes = p->eta == 2 ? 96 : 128;
    return 0;
}

/* ---- generic little-endian fixed-width bit streams ---- */
static void bits_put(uint8_t *out, const uint32_t *v, int n, int width) {
    memset(out, 0, (size_t)(n * width + 7) / 8);
    size_t bit = 0;
    for (int i = 0; i < n; i++)
        for (int b = 0; b < width; b++, bit++)
            if ((v[i] >> b) & 1u) out[bit >> 3] |= (uint8_t)(1u << (bit & 7));
}
static void bits_get(uint32_t *v, const uint8_t *in, int n, int width) {
    size_t bit = 0;
    for (int i = 0; i < n; i++) {
        uint32_t x = 0;
        for (int b = 0; b < width; b++, bit++) x |= (uint32_t)((in[bit >> 3] >> (bit & 7)) & 1u) << b;
        v[i] = x;
    }
}
static inline int32_t centre(int32_t a) { /* any int32 -> (-Q/2, Q/2] */
    int32_t r = (int32_t)(((int64_t)a % Q + Q) % Q);
    return r > Q / 2 ? r - Q : r;
}

void orc_unpack_s(int32_t *s, const uint8_t *in, int npoly, int eta) {
    int w = eta == 2 ? 3 : 4;
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        bits_get(tmp, in + (size_t)p * N * w / 8, N, w);
        for (int i = 0; i < N; i++) s[p * N + i] = eta - (int32_t)tmp[i];
    }
}
void orc_pack_s(uint8_t *out, const int32_t *s, int npoly, int eta) {
    int w = eta == 2 ? 3 : 4;
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        for (int i = 0; i < N; i++) tmp[i] = (uint32_t)(eta - centre(s[p * N + i]));
        bits_put(out + (size_t)p * N * w / 8, tmp, N, w);
    }
}
void orc_unpack_t1(int32_t *t1, const uint8_t *in, int npoly) {
    for (int p = 0; p < npoly; p++) bits_get((uint32_t *)t1 + p * N, in + (size_t)p * 320, N, 10);
}
void orc_pack_t1(uint8_t *out, const int32_t *t1, int npoly) {
    for (int p = 0; p < npoly; p++) bits_put(out + (size_t)p * 320, (const uint32_t *)t1 + p * N, N, 10);
}
void orc_unpack_t0(int32_t *t0, const uint8_t *in, int npoly) {
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        bits_get(tmp, in + (size_t)p * 416, N, 13);
        for (int i = 0; i < N; i++) t0[p * N + i] = (1 << (D - 1)) - (int32_t)tmp[i];
    }
}
void orc_pack_t0(uint8_t *out, const int32_t *t0, int npoly) {
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        for (int i = 0; i < N; i++) tmp[i] = (uint32_t)((1 << (D - 1)) - centre(t0[p * N + i]));
        bits_put(out + (size_t)p * 416, tmp, N, 13);
    }
}
void orc_unpack_z(int32_t *z, const uint8_t *in, int npoly, int gamma1_bits) {
    int w = gamma1_bits + 1;
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        bits_get(tmp, in + (size_t)p * N * w / 8, N, w);
        for (int i = 0; i < N; i++) z[p * N + i] = (1 << gamma1_bits) - (int32_t)tmp[i];
    }
}
void orc_pack_z(uint8_t *out, const int32_t *z, int npoly, int gamma1_bits) {
    int w = gamma1_bits + 1;
    uint32_t tmp[N];
    for (int p = 0; p < npoly; p++) {
        for (int i = 0; i < N; i++) tmp[i] = (uint32_t)((1 << gamma1_bits) - centre(z[p * N + i]));
        bits_put(out + (size_t)p * N * w / 8, tmp, N, w);
    }
}
void orc_pack_w1(uint8_t *out, const int32_t *w1, int npoly, int32_t gamma2) {
    int w = gamma2 == (Q - 1) / 88 ? 6 : 4;
    for (int p = 0; p < npoly; p++) bits_put(out + (size_t)p * N * w / 8, (const uint32_t *)w1 + p * N, N, w);
}

/* t canonical in [0,Q): t1 = (t + 2^12 - 1) >> 13, t0 = t - t1*2^13 */
void orc_power2round(int32_t *t1, int32_t *t0, const int32_t *t, int n) {
    for (int i = 0; i < n; i++) {
        int32_t a = (int32_t)(((int64_t)t[i] % Q + Q) % Q);
        t1[i] = (a + (1 << (D - 1)) - 1) >> D;
        t0[i] = a - (t1[i] << D);
    }
}

/* a = a1*2*gamma2 + a0 with -gamma2 < a0 <= gamma2, except the wrap case
   a1 == (Q-1)/(2*gamma2) which maps to a1 = 0, a0 = a0 - 1. */
void orc_decompose(int32_t *a1, int32_t *a0, const int32_t *a, int n, int32_t gamma2) {
    int32_t alpha = 2 * gamma2, top = (Q - 1) / alpha;
    for (int i = 0; i < n; i++) {
        int32_t r = (int32_t)(((int64_t)a[i] % Q + Q) % Q);
        int32_t r0 = r % alpha;
        if (r0 > gamma2) r0 -= alpha;
        int32_t r1 = (r - r0) / alpha;
        if (r1 == top) { r1 = 0; r0 -= 1; }
        a1[i] = r1;
        a0[i] = r0;
    }
}

static int make_hint1(int32_t a0, int32_t a1, int32_t gamma2) {
    return (a0 > gamma2 || a0 < -gamma2 || (a0 == -gamma2 && a1 != 0)) ? 1 : 0;
}
static int32_t use_hint1(int32_t a, int hint, int32_t gamma2) {
    int32_t a1, a0, m = (Q - 1) / (2 * gamma2);
    orc_decompose(&a1, &a0, &a, 1, gamma2);
    if (!hint) return a1;
    if (a0 > 0) return a1 + 1 == m ? 0 : a1 + 1;
    return a1 == 0 ? m - 1 : a1 - 1;
}

void orc_sample_in_ball(int32_t c[N], const uint8_t seed[32], int tau) {
    orc_shake_t st;
    uint8_t buf[8], b;
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, seed, 32);
    orc_shake_squeeze(&st, buf, 8);
    uint64_t signs = 0;
    for (int i = 0; i < 8; i++) signs |= (uint64_t)buf[i] << (8 * i);
    memset(c, 0, N * sizeof(int32_t));
    for (int i = N - tau; i < N; i++) {
        do orc_shake_squeeze(&st, &b, 1); while (b > i);
        c[i] = c[b];
        c[b] = (signs & 1) ? -1 : 1;
        signs >>= 1;
    }
}

void orc_expand_mask_poly(int32_t y[N], const uint8_t rhoprime[64], uint16_t nonce, int gamma1_bits) {
    uint8_t in[66], buf[640];
    memcpy(in, rhoprime, 64);
    in[64] = (uint8_t)(nonce & 0xFF);
    in[65] = (uint8_t)(nonce >> 8);
    orc_shake256(buf, (size_t)N * (gamma1_bits + 1) / 8, in, 66);
    orc_unpack_z(y, buf, 1, gamma1_bits);
}

void orc_sample_eta_poly(int32_t s[N], const uint8_t rhoprime[64], uint16_t nonce, int eta) {
    orc_shake_t st;
    uint8_t in[66], b;
    memcpy(in, rhoprime, 64);
    in[64] = (uint8_t)(nonce & 0xFF);
    in[65] = (uint8_t)(nonce >> 8);
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, in, 66);
    int got = 0;
    while (got < N) {
        orc_shake_squeeze(&st, &b, 1);
        for (int h = 0; h < 2 && got < N; h++) {
            int t = h ? (b >> 4) : (b & 15);
            if (eta == 2) { if (t < 15) s[got++] = 2 - (t % 5); }
            else          { if (t < 9) s[got++] = 4 - t; }
        }
    }
}

static int32_t inf_norm(const int32_t *a, int n) {
    int32_t m = 0;
    for (int i = 0; i < n; i++) {
        int32_t v = centre(a[i]);
        if (v < 0) v = -v;
        if (v > m) m = v;
    }
    return m;
}

/* t = INTT(A_hat * NTT(s1)) + s2   (canonical) */
static void compute_t(int32_t *t, const orc_params_t *P, const int32_t *a_hat, const int32_t *s1, const int32_t *s2) {
    int32_t *s1h = (int32_t *)malloc((size_t)P->l * N * 4);
    memcpy(s1h, s1, (size_t)P->l * N * 4);
    orc_ntt_batch(s1h, (size_t)P->l);
    orc_matvec(t, a_hat, s1h, P->k, P->l);
    orc_invntt_batch(t, (size_t)P->k);
    for (int i = 0; i < P->k; i++) orc_add(t + i * N, t + i * N, s2 + i * N);
    free(s1h);
}

int orc_keygen_chain(int level, const uint8_t rho[32], const uint8_t *s1p, const uint8_t *s2p, uint8_t *t1p, uint8_t *t0p) {
    orc_params_t P;
    if (orc_params(&P, level)) return -1;
    int32_t *a_hat = (int32_t *)malloc((size_t)P.k * P.l * N * 4);
    int32_t s1[8 * N], s2[8 * N], t[8 * N], t1[8 * N], t0[8 * N];
    orc_expand_a(a_hat, rho, P.k, P.l);
    orc_unpack_s(s1, s1p, P.l, P.eta);
    orc_unpack_s(s2, s2p, P.k, P.eta);
    compute_t(t, &P, a_hat, s1, s2);
    orc_power2round(t1, t0, t, P.k * N);
    orc_pack_t1(t1p, t1, P.k);
    orc_pack_t0(t0p, t0, P.k);
    free(a_hat);
    return 0;
}

int orc_keygen(int level, const uint8_t xi[32], uint8_t rho[32], uint8_t key[32], uint8_t tr[32],
               uint8_t *s1p, uint8_t *s2p, uint8_t *t1p, uint8_t *t0p) {
    orc_params_t P;
    if (orc_params(&P, level)) return -1;
    uint8_t seed[128];
    orc_shake256(seed, 128, xi, 32);
    memcpy(rho, seed, 32);
    memcpy(key, seed + 96, 32);
    int32_t s1[8 * N], s2[8 * N];
    for (int j = 0; j < P.l; j++) orc_sample_eta_poly(s1 + j * N, seed + 32, (uint16_t)j, P.eta);
    for (int i = 0; i < P.k; i++) orc_sample_eta_poly(s2 + i * N, seed + 32, (uint16_t)(P.l + i), P.eta);
    orc_pack_s(s1p, s1, P.l, P.eta);
    orc_pack_s(s2p, s2, P.k, P.eta);
    orc_keygen_chain(level, rho, s1p, s2p, t1p, t0p);
    /* tr = SHAKE-256(rho || t1_packed)[0:32]  (combined_top.v:980) */
    orc_shake_t st;
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, rho, 32);
    orc_shake_absorb(&st, t1p, (size_t)P.k * 320);
    orc_shake_squeeze(&st, tr, 32);
    return 0;
}

/* Expanded signing key: ExpandA(rho) and the NTT images of s1, s2, t0 - computed once per key
   (LOAD_RHO / NTT_S1 / NTT_S2 / NTT_T0, combined_top.v:1560-1767) and reused for every message. */
struct orc_sign_ctx {
    orc_params_t P;
    int32_t *a_hat;
    int32_t s1h[8 * N], s2h[8 * N], t0h[8 * N];
};

orc_sign_ctx_t *orc_sign_prepare(int level, const uint8_t rho[32], const uint8_t *s1p, const uint8_t *s2p, const uint8_t *t0p) {
    orc_sign_ctx_t *c = (orc_sign_ctx_t *)malloc(sizeof *c);
    if (!c || orc_params(&c->P, level)) { free(c); return NULL; }
    const int k = c->P.k, l = c->P.l;
    c->a_hat = (int32_t *)malloc((size_t)k * l * N * 4);
    orc_expand_a(c->a_hat, rho, k, l);
    orc_unpack_s(c->s1h, s1p, l, c->P.eta); orc_ntt_batch(c->s1h, (size_t)l);
    orc_unpack_s(c->s2h, s2p, k, c->P.eta); orc_ntt_batch(c->s2h, (size_t)k);
    orc_unpack_t0(c->t0h, t0p, k);          orc_ntt_batch(c->t0h, (size_t)k);
    return c;
}
void orc_sign_free(orc_sign_ctx_t *c) {
    if (c) { free(c->a_hat); free(c); }
}

int orc_sign(int level, const uint8_t rho[32], const uint8_t key[32], const uint8_t tr[32],
             const uint8_t *s1p, const uint8_t *s2p, const uint8_t *t0p,
             const uint8_t *msg, size_t mlen, uint8_t *zp, uint8_t *hp, uint8_t ctilde[32]) {
    orc_sign_ctx_t *c = orc_sign_prepare(level, rho, s1p, s2p, t0p);
    if (!c) return -1;
    int r = orc_sign_msg(c, key, tr, msg, mlen, zp, hp, ctilde);
    orc_sign_free(c);
    return r;
}

int orc_sign_msg(const orc_sign_ctx_t *ctx, const uint8_t key[32], const uint8_t tr[32],
                 const uint8_t *msg, size_t mlen, uint8_t *zp, uint8_t *hp, uint8_t ctilde[32]) {
    const orc_params_t P = ctx->P;
    const int k = P.k, l = P.l;
    const int32_t *a_hat = ctx->a_hat, *s1h = ctx->s1h, *s2h = ctx->s2h, *t0h = ctx->t0h;
    int32_t y[8 * N], yh[8 * N], w[8 * N], w1[8 * N], w0[8 * N], c[N], ch[N];
    int32_t z[8 * N], tmp[8 * N], hint[8 * N];
    uint8_t mu[64], rhoprime[64], w1p[8 * 192];
    orc_shake_t st;

    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, tr, 32);
    orc_shake_absorb(&st, msg, mlen);
    orc_shake_squeeze(&st, mu, 64);
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, key, 32);
    orc_shake_absorb(&st, mu, 64);
    orc_shake_squeeze(&st, rhoprime, 64);

    int attempts = 0;
    for (unsigned kappa = 0; kappa < 1000; kappa++) {
        attempts++;
        for (int j = 0; j < l; j++) orc_expand_mask_poly(y + j * N, rhoprime, (uint16_t)(l * kappa + j), P.gamma1_bits);
        memcpy(yh, y, (size_t)l * N * 4);
        orc_ntt_batch(yh, (size_t)l);
        orc_matvec(w, a_hat, yh, k, l);
        orc_invntt_batch(w, (size_t)k);
        orc_decompose(w1, w0, w, k * N, P.gamma2);
        orc_pack_w1(w1p, w1, k, P.gamma2);
        orc_shake_init(&st, 136);
        orc_shake_absorb(&st, mu, 64);
        orc_shake_absorb(&st, w1p, (size_t)k * P.w1_bytes);
        orc_shake_squeeze(&st, ctilde, 32);
        orc_sample_in_ball(c, ctilde, P.tau);
        memcpy(ch, c, sizeof ch);
        orc_ntt(ch);
        /* z = INTT(y_hat + c_hat o s1_hat)  (combined_top.v:2011-2037, :2088-2101) */
        for (int j = 0; j < l; j++) {
            memcpy(z + j * N, yh + j * N, N * 4);
            orc_pointwise_acc(z + j * N, ch, s1h + j * N);
        }
        orc_invntt_batch(z, (size_t)l);
        if (inf_norm(z, l * N) >= P.gamma1 - P.beta) continue;
        /* c*t0 check */
        for (int i = 0; i < k; i++) orc_pointwise(tmp + i * N, ch, t0h + i * N);
        orc_invntt_batch(tmp, (size_t)k);
        int bad_ct0 = inf_norm(tmp, k * N) >= P.gamma2;
        /* w0 - c*s2 check */
        int32_t cs2[8 * N];
        for (int i = 0; i < k; i++) orc_pointwise(cs2 + i * N, ch, s2h + i * N);
        orc_invntt_batch(cs2, (size_t)k);
        for (int i = 0; i < k * N; i++) w0[i] = centre(w0[i] - centre(cs2[i]));
        if (inf_norm(w0, k * N) >= P.gamma2 - P.beta) continue;
        if (bad_ct0) continue;
        int cnt = 0;
        for (int i = 0; i < k * N; i++) {
            hint[i] = make_hint1(w0[i] + centre(tmp[i]), w1[i], P.gamma2);
            cnt += hint[i];
        }
        if (cnt > P.omega) continue;
        /* accept */
        orc_pack_z(zp, z, l, P.gamma1_bits);
        memset(hp, 0, (size_t)(P.omega + k));
        int idx = 0;
        for (int i = 0; i < k; i++) {
            for (int j = 0; j < N; j++)
                if (hint[i * N + j]) hp[idx++] = (uint8_t)j;
            hp[P.omega + i] = (uint8_t)idx;
        }
        return attempts;
    }
    return -2;
}

int orc_verify(int level, const uint8_t rho[32], const uint8_t *t1p, const uint8_t *msg, size_t mlen,
               const uint8_t *zp, const uint8_t *hp, const uint8_t ctilde[32]) {
    orc_params_t P;
    if (orc_params(&P, level)) return 1;
    const int k = P.k, l = P.l;
    int32_t z[8 * N], t1[8 * N], c[N], w[8 * N], w1[8 * N], hint[8 * N];
    uint8_t tr[32], mu[64], w1p[8 * 192], c2[32];
    orc_shake_t st;

    orc_unpack_z(z, zp, l, P.gamma1_bits);
    if (inf_norm(z, l * N) >= P.gamma1 - P.beta) return 1;
    /* hint decode with the standard well-formedness checks */
    memset(hint, 0, sizeof hint);
    int idx = 0;
    for (int i = 0; i < k; i++) {
        int end = hp[P.omega + i];
        if (end < idx || end > P.omega) return 1;
        for (int j = idx; j < end; j++) {
            if (j > idx && hp[j] <= hp[j - 1]) return 1;
            hint[i * N + hp[j]] = 1;
        }
        idx = end;
    }
    for (int j = idx; j < P.omega; j++)
        if (hp[j]) return 1;

    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, rho, 32);
    orc_shake_absorb(&st, t1p, (size_t)k * 320);
    orc_shake_squeeze(&st, tr, 32);
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, tr, 32);
    orc_shake_absorb(&st, msg, mlen);
    orc_shake_squeeze(&st, mu, 64);

    int32_t *a_hat = (int32_t *)malloc((size_t)k * l * N * 4);
    orc_expand_a(a_hat, rho, k, l);
    orc_sample_in_ball(c, ctilde, P.tau);
    orc_ntt(c);
    orc_ntt_batch(z, (size_t)l);
    orc_matvec(w, a_hat, z, k, l);
    free(a_hat);
    orc_unpack_t1(t1, t1p, k);
    for (int i = 0; i < k * N; i++) t1[i] <<= D; /* decoder.v:96-100 */
    orc_ntt_batch(t1, (size_t)k);
    for (int i = 0; i < k; i++) {
        int32_t ct1[N];
        orc_pointwise(ct1, c, t1 + i * N);
        orc_sub(w + i * N, w + i * N, ct1);
    }
    orc_invntt_batch(w, (size_t)k);
    for (int i = 0; i < k * N; i++) w1[i] = use_hint1(w[i], hint[i], P.gamma2);
    orc_pack_w1(w1p, w1, k, P.gamma2);
    orc_shake_init(&st, 136);
    orc_shake_absorb(&st, mu, 64);
    orc_shake_absorb(&st, w1p, (size_t)k * P.w1_bytes);
    orc_shake_squeeze(&st, c2, 32);
    return memcmp(c2, ctilde, 32) ? 1 : 0;
}
