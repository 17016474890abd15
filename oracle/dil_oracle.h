/*
 * dil_oracle.h — CPU oracle for the Dilithium polynomial-arithmetic hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * engine: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load liboracle.so.  The product (libdilithium_b200.so)
 * never links, loads or calls it.
 *
 * What it restates (reference = GMUCERG/Dilithium, paths relative to its root):
 *   - dilithium-256/params.h:30-35        ring constants Q, N
 *   - dilithium-256/consts.cpp:64-97      zetas_barrett[] (regenerated here from
 *                                         zeta=1753, not transcribed)
 *   - dilithium-256/reference_code/ref_ntt.cpp:28-87   ntt / pointwise_barrett / invntt
 *   - dilithium-256/reference_code/ref_ntt2x2.cpp:37-145  radix-2x2 schedule
 *   - rtl_src/butterfly.v:144-164,224-237  MULT(=mul-acc) / ADD / SUB modes
 *   - rtl_src/gen_a_ext.v, sampler_a_ext.v:107-133, rejection_a.v:67-92  ExpandA
 *   - rtl_src/combined_top.v:921-958 etc. mat-vec loop order
 *   - the keygen / sign / verify data flow of rtl_src/combined_top.v (Dilithium
 *     round-3.1, deterministic signing) used to pin everything against KAT/.
 *
 * Parity pinning: tests/test_oracle_*.py check this oracle against
 *   (1) oracle/_ref (the reference's own C++ compiled from /root/reference),
 *   (2) golden vectors generated from (1) and committed under tests/golden/,
 *   (3) all 100 KAT vectors at levels 2/3/5 (tests/golden/kat_L*.npz).
 *
 * Value contract: every function here emits CANONICAL residues in [0,Q).  The
 * reference emits signed non-canonical residues and compares mod Q
 * (ref_test_ntt_ntt2x2.cpp:31-42); callers canonicalise the reference side.
 */
#ifndef DIL_ORACLE_H
#define DIL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_Q 8380417
#define ORC_N 256

/* ---- ring arithmetic (dil_arith.c) ---- */
const int32_t *orc_zetas(void);                 /* centred, index 0 == 0 */
void orc_ntt(int32_t a[ORC_N]);                 /* ref_ntt.cpp:28-47   */
void orc_invntt(int32_t a[ORC_N]);              /* ref_ntt.cpp:59-87   */
void orc_ntt2x2(int32_t a[ORC_N]);              /* ref_ntt2x2.cpp:37-82  */
void orc_invntt2x2(int32_t a[ORC_N]);           /* ref_ntt2x2.cpp:100-145 */
void orc_pointwise(int32_t c[ORC_N], const int32_t a[ORC_N], const int32_t b[ORC_N]);
void orc_pointwise_acc(int32_t c[ORC_N], const int32_t a[ORC_N], const int32_t b[ORC_N]);
void orc_add(int32_t c[ORC_N], const int32_t a[ORC_N], const int32_t b[ORC_N]);
void orc_sub(int32_t c[ORC_N], const int32_t a[ORC_N], const int32_t b[ORC_N]);
/* batched helpers (n contiguous polys) */
void orc_ntt_batch(int32_t *a, size_t n);
void orc_invntt_batch(int32_t *a, size_t n);
void orc_pointwise_batch(int32_t *c, const int32_t *a, const int32_t *b, size_t n);

/* ---- Keccak / SHAKE (dil_keccak.c; FIPS-202, restating keccak_round.vhd etc.) ---- */
typedef struct {
    uint64_t s[25];
    unsigned pos;      /* bytes absorbed into / squeezed from current block */
    unsigned rate;     /* 168 (SHAKE-128) or 136 (SHAKE-256) */
    int squeezing;
} orc_shake_t;
void orc_keccak_f1600(uint64_t s[25]);
void orc_shake_init(orc_shake_t *c, unsigned rate);
void orc_shake_absorb(orc_shake_t *c, const uint8_t *in, size_t len);
void orc_shake_squeeze(orc_shake_t *c, uint8_t *out, size_t len);
void orc_shake128(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen);
void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen);

/* ---- ExpandA + mat-vec (dil_expand.c) ---- */
/* A_hat[(i*l+j)*256 ..] = RejUniform(SHAKE128(rho || j || i)); returns max #blocks used */
int orc_expand_a_poly(int32_t a[ORC_N], const uint8_t rho[32], int i, int j);
void orc_expand_a(int32_t *a_hat, const uint8_t rho[32], int k, int l);
/* w[i] = sum_j A[i*l+j] o v[j]   (all NTT domain, canonical out) */
void orc_matvec(int32_t *w, const int32_t *a_hat, const int32_t *v, int k, int l);
/* batched: shared A (n_rho==1) or per-item A expanded on the fly from rho[item] */
void orc_matvec_batch(int32_t *w, const int32_t *a_hat, const int32_t *v, int k, int l, size_t batch);
void orc_matvec_expand_batch(int32_t *w, const uint8_t *rho, size_t n_rho, const int32_t *v,
                             int k, int l, size_t batch, int ntt_in, int invntt_out);

/* ---- scheme level (dil_scheme.c): used to pin the path against KAT/ ---- */
typedef struct {
    int level, k, l, eta, tau, gamma1_bits, omega, beta;
    int32_t gamma1, gamma2;
    int z_bytes, w1_bytes, s_bytes;  /* packed bytes per polynomial */
} orc_params_t;
int orc_params(orc_params_t *p, int level);

void orc_unpack_s(int32_t *s, const uint8_t *in, int npoly, int eta);     /* centred signed  */
void orc_pack_s(uint8_t *out, const int32_t *s, int npoly, int eta);
void orc_unpack_t1(int32_t *t1, const uint8_t *in, int npoly);           /* plain 10-bit   */
void orc_pack_t1(uint8_t *out, const int32_t *t1, int npoly);
void orc_unpack_t0(int32_t *t0, const uint8_t *in, int npoly);           /* centred signed  */
void orc_pack_t0(uint8_t *out, const int32_t *t0, int npoly);
void orc_unpack_z(int32_t *z, const uint8_t *in, int npoly, int gamma1_bits);
void orc_pack_z(uint8_t *out, const int32_t *z, int npoly, int gamma1_bits);
void orc_pack_w1(uint8_t *out, const int32_t *w1, int npoly, int32_t gamma2);
void orc_power2round(int32_t *t1, int32_t *t0, const int32_t *t, int n);  /* t canonical */
void orc_decompose(int32_t *a1, int32_t *a0, const int32_t *a, int n, int32_t gamma2);
void orc_sample_in_ball(int32_t c[ORC_N], const uint8_t seed[32], int tau);
void orc_expand_mask_poly(int32_t y[ORC_N], const uint8_t rhoprime[64], uint16_t nonce, int gamma1_bits);
void orc_sample_eta_poly(int32_t s[ORC_N], const uint8_t rhoprime[64], uint16_t nonce, int eta);

/* (rho,s1,s2) -> (t1,t0): the keygen chain that pins ExpandA+NTT+MULT-ACC+INTT+ADD
   (combined_top.v:905-1021).  s1/s2 packed as in KAT; t1/t0 packed as in KAT. */
int orc_keygen_chain(int level, const uint8_t rho[32], const uint8_t *s1p, const uint8_t *s2p,
                     uint8_t *t1p, uint8_t *t0p);
/* full keygen from seed xi (combined_top.v:754-1079) */
int orc_keygen(int level, const uint8_t xi[32], uint8_t rho[32], uint8_t key[32], uint8_t tr[32],
               uint8_t *s1p, uint8_t *s2p, uint8_t *t1p, uint8_t *t0p);
/* deterministic sign (combined_top.v:1535-2229); returns #attempts (>0) or <0 on error */
int orc_sign(int level, const uint8_t rho[32], const uint8_t key[32], const uint8_t tr[32],
             const uint8_t *s1p, const uint8_t *s2p, const uint8_t *t0p,
             const uint8_t *msg, size_t mlen,
             uint8_t *zp, uint8_t *hp, uint8_t ctilde[32]);
/* the same with the per-key work (ExpandA, NTT of s1/s2/t0) hoisted out for batches */
typedef struct orc_sign_ctx orc_sign_ctx_t;
orc_sign_ctx_t *orc_sign_prepare(int level, const uint8_t rho[32], const uint8_t *s1p, const uint8_t *s2p, const uint8_t *t0p);
int orc_sign_msg(const orc_sign_ctx_t *ctx, const uint8_t key[32], const uint8_t tr[32],
                 const uint8_t *msg, size_t mlen, uint8_t *zp, uint8_t *hp, uint8_t ctilde[32]);
void orc_sign_free(orc_sign_ctx_t *ctx);
/* verify (combined_top.v:1080-1534); returns 0 = accept, 1 = reject */
int orc_verify(int level, const uint8_t rho[32], const uint8_t *t1p,
               const uint8_t *msg, size_t mlen,
               const uint8_t *zp, const uint8_t *hp, const uint8_t ctilde[32]);

#ifdef __cplusplus
}
#endif
#endif
