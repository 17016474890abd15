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
    int op;
    int32_t *w, *v;
    const int32_t *a_hat, *b;
    int k, l;
    size_t lo, hi;
} job_t;

enum { OP_NTT, OP_INVNTT, OP_POINTWISE, OP_SIGNCORE, OP_MATVEC };

static void *run(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t p = j->lo; p < j->hi; p++) {
        switch (j->op) {
        case OP_NTT: orc_ntt(j->v + p * ORC_N); break;
        case OP_INVNTT: orc_invntt(j->v + p * ORC_N); break;
        case OP_POINTWISE: orc_pointwise(j->w + p * ORC_N, j->v + p * ORC_N, j->b + p * ORC_N); break;
        case OP_MATVEC: orc_matvec(j->w + p * j->k * ORC_N, j->a_hat, j->v + p * j->l * ORC_N, j->k, j->l); break;
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
