/*
 * dil_keccak.c — oracle Keccak-f[1600] / SHAKE-128 / SHAKE-256 (FIPS-202).
 * TEST INFRASTRUCTURE ONLY (see dil_oracle.h).
 *
 * The reference realises these only as VHDL: one round per clock
 * (rtl_src/keccak_round.vhd, keccak_datapath.vhd:190-203), 24 round constants
 * (keccak_cons.vhd:25-33), SHAKE padding 0x1F..0x80 (keccak_bytepad.vhd:37-44),
 * rates 1344/1088 bits (keccak_pkg.vhd:14-19).  This is a plain byte-oriented
 * restatement; tests cross-check it against Python's hashlib.
 */
#include <string.h>
#include "dil_oracle.h"

static inline uint64_t rotl(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }

void orc_keccak_f1600(uint64_t A[25]) {
    /* round constants from the degree-8 LFSR x^8+x^6+x^5+x^4+1 */
    static uint64_t rc[24];
    static unsigned rho[25];
    static int ready;
    if (!ready) {
        unsigned lfsr = 1;
        for (int r = 0; r < 24; r++) {
            uint64_t c = 0;
            for (int j = 0; j < 7; j++) {
                if (lfsr & 1) c |= 1ULL << ((1u << j) - 1);
                lfsr = (lfsr << 1) ^ ((lfsr & 0x80) ? 0x171 : 0);
            }
            rc[r] = c;
        }
        /* rho offsets: walk (x,y) -> (y, 2x+3y), offset (t+1)(t+2)/2 */
        int x = 1, y = 0;
        rho[0] = 0;
        for (int t = 0; t < 24; t++) {
            rho[x + 5 * y] = ((t + 1) * (t + 2) / 2) % 64;
            int ny = (2 * x + 3 * y) % 5;
            x = y;
            y = ny;
        }
        ready = 1;
    }
    for (int r = 0; r < 24; r++) {
        uint64_t C[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t D = C[(x + 4) % 5] ^ rotl(C[(x + 1) % 5], 1);
            for (int y = 0; y < 5; y++) A[x + 5 * y] ^= D;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(A[x + 5 * y], rho[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= rc[r];
    }
}

void orc_shake_init(orc_shake_t *c, unsigned rate) {
    memset(c, 0, sizeof *c);
    c->rate = rate;
}

static inline void xor_byte(uint64_t *s, unsigned pos, uint8_t b) { s[pos >> 3] ^= (uint64_t)b << (8 * (pos & 7)); }

void orc_shake_absorb(orc_shake_t *c, const uint8_t *in, size_t len) {
    for (size_t i = 0; i < len; i++) {
        xor_byte(c->s, c->pos++, in[i]);
        if (c->pos == c->rate) {
            orc_keccak_f1600(c->s);
            c->pos = 0;
        }
    }
}

void orc_shake_squeeze(orc_shake_t *c, uint8_t *out, size_t len) {
    if (!c->squeezing) {
        xor_byte(c->s, c->pos, 0x1F);
        xor_byte(c->s, c->rate - 1, 0x80);
        orc_keccak_f1600(c->s);
        c->pos = 0;
        c->squeezing = 1;
    }
    for (size_t i = 0; i < len; i++) {
        if (c->pos == c->rate) {
            orc_keccak_f1600(c->s);
            c->pos = 0;
        }
        out[i] = (uint8_t)(c->s[c->pos >> 3] >> (8 * (c->pos & 7)));
        c->pos++;
    }
}

void orc_shake128(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen) {
    orc_shake_t c;
    orc_shake_init(&c, 168);
    orc_shake_absorb(&c, in, inlen);
    orc_shake_squeeze(&c, out, outlen);
}
void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen) {
    orc_shake_t c;
    orc_shake_init(&c, 136);
    orc_shake_absorb(&c, in, inlen);
    orc_shake_squeeze(&c, out, outlen);
}
