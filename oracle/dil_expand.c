/*
 * dil_expand.c — oracle ExpandA (SHAKE-128 rejection sampling) and the
 * matrix-vector product in the NTT domain.  TEST INFRASTRUCTURE ONLY.
 *
 * ExpandA:  A_hat[i][j] = RejUniform(SHAKE-128(rho || byte(j) || byte(i)))
 *   - hash input and nonce byte order     rtl_src/sampler_a_ext.v:107, :122-133
 *   - 3-byte little-endian chunks, top bit masked, accept t < Q, first 256
 *     accepted in stream order             rtl_src/rejection_a.v:67-92
 *   - row-major poly index i*l + j         rtl_src/gen_a_ext.v:130-405,
 *                                          combined_top.v:797-802
 * mat-vec: w_i = sum_{j<l} A_hat[i][j] o v_j, accumulator cleared at j == 0
 *                                          rtl_src/combined_top.v:921-958,
 *                                          :1347-1386, :1875-1913
 *   built on the MULT (multiply-accumulate) butterfly mode
 *                                          rtl_src/butterfly.v:144-150
 */
#include <stdlib.h>
#include <string.h>
#include "dil_oracle.h"

#define Q ORC_Q
#define N ORC_N

int orc_expand_a_poly(int32_t a[N], const uint8_t rho[32], int i, int j) {
    orc_shake_t c;
    uint8_t seed[34], blk[168];
    memcpy(seed, rho, 32);
    seed[32] = (uint8_t)j;
    seed[33] = (uint8_t)i;
    orc_shake_init(&c, 168);
    orc_shake_absorb(&c, seed, 34);
    int got = 0, nblocks = 0;
    while (got < N) {
        orc_shake_squeeze(&c, blk, 168); /* 168 = 56 whole 3-byte candidates */
        nblocks++;
        for (int p = 0; p < 168 && got < N; p += 3) {
            uint32_t t = (uint32_t)blk[p] | ((uint32_t)blk[p + 1] << 8) | ((uint32_t)(blk[p + 2] & 0x7F) << 16);
            if (t < Q) a[got++] = (int32_t)t;
        }
    }
    return nblocks;
}

void orc_expand_a(int32_t *a_hat, const uint8_t rho[32], int k, int l) {
    for (int i = 0; i < k; i++)
        for (int j = 0; j < l; j++) orc_expand_a_poly(a_hat + (size_t)(i * l + j) * N, rho, i, j);
}

void orc_matvec(int32_t *w, const int32_t *a_hat, const int32_t *v, int k, int l) {
    for (int i = 0; i < k; i++) {
        int32_t *wi = w + (size_t)i * N;
        for (int j = 0; j < l; j++) {
            const int32_t *aij = a_hat + (size_t)(i * l + j) * N;
            const int32_t *vj = v + (size_t)j * N;
            if (j == 0) orc_pointwise(wi, aij, vj);
            else orc_pointwise_acc(wi, aij, vj);
        }
    }
}

void orc_matvec_batch(int32_t *w, const int32_t *a_hat, const int32_t *v, int k, int l, size_t batch) {
    for (size_t b = 0; b < batch; b++) orc_matvec(w + b * k * N, a_hat, v + b * l * N, k, l);
}

/* v: [batch][l][256] (time domain if ntt_in, else NTT domain); w: [batch][k][256].
   rho: n_rho == 1 (shared) or n_rho == batch (per item). */
void orc_matvec_expand_batch(int32_t *w, const uint8_t *rho, size_t n_rho, const int32_t *v, int k, int l,
                             size_t batch, int ntt_in, int invntt_out) {
    int32_t *a_hat = (int32_t *)malloc((size_t)k * l * N * sizeof(int32_t));
    int32_t *vt = (int32_t *)malloc((size_t)l * N * sizeof(int32_t));
    for (size_t b = 0; b < batch; b++) {
        if (b == 0 || n_rho > 1) orc_expand_a(a_hat, rho + (n_rho > 1 ? b * 32 : 0), k, l);
        memcpy(vt, v + b * l * N, (size_t)l * N * sizeof(int32_t));
        if (ntt_in) orc_ntt_batch(vt, (size_t)l);
        orc_matvec(w + b * k * N, a_hat, vt, k, l);
        if (invntt_out) orc_invntt_batch(w + b * k * N, (size_t)k);
    }
    free(a_hat);
    free(vt);
}
