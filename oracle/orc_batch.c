/*
 * orc_batch.c — pthread batch drivers over the oracle, for the "port" CPU
 * baseline in bench.py and for faster large-batch parity checks in tests/.
 * TEST/BENCH INFRASTRUCTURE ONLY (see dil_oracle.h).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "dil_oracle.h"

typedef struct {
    const orc_sign_ctx_t *ctx;
    const uint8_t *key, *tr, *msgs;
    const uint64_t *off;
    uint8_t *z, *h, *c;
    uint32_t *attempts;
    size_t zb, hb;
} sign_args_t;

typedef struct {
    int op;
    int32_t *w, *v;
    const int32_t *a_hat, *b;
    int k, l;
    size_t lo, hi;
    const sign_args_t *sg;
} job_t;

enum { OP_NTT, OP_INVNTT, OP_POINTWISE, OP_SIGNCORE, OP_MATVEC, OP_SIGN };



static void *run(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t p = j->lo; p < j->hi; p++) {
        switch (j->op) {
        case OP_NTT: orc_ntt(j->v + p * ORC_N); break;
        case OP_INVNTT: orc_invntt(j->v + p * ORC_N); break;
        case OP_POINTWISE: orc_pointwise(j->w + p * ORC_N, j->v + p * ORC_N, j->b + p * ORC_N); break;
        case OP_MATVEC: orc_matvec(j->w + p * j->k * ORC_N, j->a_hat, j->v + p * j->l * ORC_N, j->k, j->l); break;
        case OP_SIGN: {
            const sign_args_t *g = j->sg;
            int a = orc_sign_msg(g->ctx, g->key, g->tr, g->msgs + g->off[p], (size_t)(g->off[p + 1] - g->off[p]),
                                 g->z + p * g->zb, g->h + p * g->hb, g->c + p * 32);
            if (g->attempts) g->attempts[p] = (uint32_t)a;
            break;
        }
        case OP_SIGNCORE:
            orc_ntt_batch(j->v + p * j->l * ORC_N, (size_t)j->l);
            orc_matvec(j->w + p * j->k * ORC_N, j->a_hat, j->v + p * j->l * ORC_N, j->k, j->l);
            orc_invntt_batch(j->w + p * j->k * ORC_N, (size_t)j->k);
            break;
        }
    }
    return NULL;
}

static void dispatch(job_t base, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        jobs[t] = base;
        jobs[t].lo = lo;
        jobs[t].hi = hi;
        pthread_create(&th[t], NULL, run, &jobs[t]);
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void orc_ntt_batch_mt(int32_t *a, size_t n, int threads) {
    job_t j = {.op = OP_NTT, .v = a};
    dispatch(j, n, threads);
}
void orc_invntt_batch_mt(int32_t *a, size_t n, int threads) {
    job_t j = {.op = OP_INVNTT, .v = a};
    dispatch(j, n, threads);
}
void orc_pointwise_batch_mt(int32_t *c, const int32_t *a, const int32_t *b, size_t n, int threads) {
    job_t j = {.op = OP_POINTWISE, .w = c, .v = (int32_t *)a, .b = b};
    dispatch(j, n, threads);
}
void orc_matvec_batch_mt(int32_t *w, const int32_t *a_hat, const int32_t *v, int k, int l, size_t batch, int threads) {
    job_t j = {.op = OP_MATVEC, .w = w, .v = (int32_t *)v, .a_hat = a_hat, .k = k, .l = l};
    dispatch(j, batch, threads);
}
/* cfg2 sign-core: y (time domain, overwritten with NTT(y)) -> w = INTT(A_hat * NTT(y)) */
void orc_signcore_batch_mt(int32_t *w, int32_t *y, const int32_t *a_hat, int k, int l, size_t batch, int threads) {
    job_t j = {.op = OP_SIGNCORE, .w = w, .v = y, .a_hat = a_hat, .k = k, .l = l};
    dispatch(j, batch, threads);
}

/* batched deterministic signing with one key: messages concatenated, off[n+1] offsets */
void orc_sign_batch_mt(int level, const uint8_t *rho, const uint8_t *key, const uint8_t *tr, const uint8_t *s1p,
                       const uint8_t *s2p, const uint8_t *t0p, const uint8_t *msgs, const uint64_t *off, size_t n,
                       uint8_t *z, uint8_t *h, uint8_t *c, uint32_t *attempts, int threads) {
    orc_params_t P;
    if (orc_params(&P, level)) return;
    orc_sign_ctx_t *ctx = orc_sign_prepare(level, rho, s1p, s2p, t0p);   /* once per key */
    if (!ctx) return;
    sign_args_t g = {ctx, key, tr, msgs, off, z, h, c, attempts, (size_t)P.l * (size_t)P.z_bytes, (size_t)(P.omega + P.k)};
    job_t j = {.op = OP_SIGN, .sg = &g};
    dispatch(j, n, threads);
    orc_sign_free(ctx);
}
