"""CPU model of the sparse challenge product used by sign_tail_sparse_kernel (dilithium_b200/csrc/sign_kernels.cu):
the table layout (biased nibbles per sign and alignment), the per-term offsets and the grouped nibble -> byte
accumulation, restated with numpy and checked against the plain negacyclic product c*s mod (X^256 + 1).
It pins the index arithmetic of the kernel's design; the kernel itself is checked bit-exactly on the GPU
(tests/test_gpu_sign.py: all KATs and oracle parity run through it)."""
import numpy as np
import pytest

N = 256


def negacyclic(c, s):
    out = np.zeros(N, dtype=np.int64)
    for i in np.nonzero(c)[0]:
        for j in range(N):
            k = i + j
            if k < N:
                out[k] += c[i] * s[j]
            else:
                out[k - N] -= c[i] * s[j]
    return out


def build_tables(s, eta):
    """T[sign][al][m] = eta +- e[m + al] as nibbles; e[256 + i] = s[i], e[i] = -s[i]; zero beyond the end."""
    e = np.concatenate([-s, s]).astype(np.int64)
    T = np.zeros((2, 8, 512), dtype=np.int64)
    for sg in range(2):
        for al in range(8):
            for m in range(512):
                j = m + al
                v = e[j] if j < 512 else 0
                T[sg, al, m] = eta + (-v if sg else v)
    assert T.min() >= 0 and T.max() <= 2 * eta <= 15
    return T


def term_offsets(c):
    """(sign, al, shift) per non-zero coefficient, as the kernel stores them (in nibble units here)."""
    terms = []
    for pos in np.nonzero(c)[0]:
        al = (-int(pos)) & 7
        terms.append((1 if c[pos] < 0 else 0, al, 256 - int(pos) - al))
    return terms


@pytest.mark.parametrize("tau,eta", [(39, 2), (60, 2)])
def test_sparse_product_matches_negacyclic(tau, eta):
    rng = np.random.default_rng(tau)
    group = 15 // (2 * eta)
    for _ in range(20):
        s = rng.integers(-eta, eta + 1, size=N)
        c = np.zeros(N, dtype=np.int64)
        c[rng.choice(N, size=tau, replace=False)] = rng.choice([-1, 1], size=tau)
        T = build_tables(s, eta)
        terms = term_offsets(c)
        assert len(terms) == tau
        out = np.zeros(N, dtype=np.int64)
        for lane in range(32):
            nacc = np.zeros(8, dtype=np.int64)      # eight nibble fields of one 32-bit word
            bacc = np.zeros(8, dtype=np.int64)      # byte sums (two 32-bit words in the kernel)
            for t, (sg, al, shift) in enumerate(terms):
                base = shift + 8 * lane             # multiple of 8: one aligned 4-byte load
                assert base % 8 == 0 and 0 <= base and base + 7 < 512
                nacc += T[sg, al, base:base + 8]
                assert nacc.max() <= 15             # no carry between nibbles
                if t % group == group - 1 or t == tau - 1:
                    bacc += nacc
                    nacc[:] = 0
                    assert bacc.max() <= 255        # no carry between bytes
            out[8 * lane:8 * lane + 8] = bacc - tau * eta
        assert np.array_equal(out, negacyclic(c, s))


def test_sparse_product_bytes_eta4():
    """Level 3 (eta = 4, tau = 49): biased elements 0..8 do not fit nibble sums, so the tables hold BYTES (one 8-byte load per
    term and lane), 31 terms are added per byte sum and then spread into 16-bit sums - restated with the kernel's exact word
    arithmetic (packed 32-bit adds, masks 0x00FF00FF) and offsets (sign * 4096 + al * 512 + shift bytes)."""
    tau, eta = 49, 4
    rng = np.random.default_rng(3)
    group = 255 // (2 * eta)
    assert group == 31
    for _ in range(20):
        s = rng.integers(-eta, eta + 1, size=N)
        c = np.zeros(N, dtype=np.int64)
        c[rng.choice(N, size=tau, replace=False)] = rng.choice([-1, 1], size=tau)
        T = build_tables(s, eta)                       # same element values, stored one per byte
        tab = np.zeros(2 * 8 * 512, dtype=np.uint8)    # byte image of one polynomial's table as the kernel lays it out
        for sg in range(2):
            for al in range(8):
                tab[sg * 4096 + al * 512: sg * 4096 + (al + 1) * 512] = T[sg, al]
        out = np.zeros(N, dtype=np.int64)
        for lane in range(32):
            b0 = b1 = 0
            h = [0, 0, 0, 0]
            for t, (sg, al, shift) in enumerate(term_offsets(c)):
                u = sg * 4096 + al * 512 + shift + 8 * lane
                assert u % 8 == 0 and u + 8 <= (sg * 8 + al + 1) * 512
                vx = int.from_bytes(tab[u:u + 4].tobytes(), "little")
                vy = int.from_bytes(tab[u + 4:u + 8].tobytes(), "little")
                b0 = (b0 + vx) & 0xFFFFFFFF
                b1 = (b1 + vy) & 0xFFFFFFFF
                if t % group == group - 1 or t == tau - 1:
                    assert all(((b >> (8 * i)) & 0xFF) <= 255 for b in (b0, b1) for i in range(4))
                    h[0] += b0 & 0x00FF00FF; h[1] += (b0 >> 8) & 0x00FF00FF
                    h[2] += b1 & 0x00FF00FF; h[3] += (b1 >> 8) & 0x00FF00FF
                    b0 = b1 = 0
            x = [h[0] & 0xFFFF, h[1] & 0xFFFF, h[0] >> 16, h[1] >> 16, h[2] & 0xFFFF, h[3] & 0xFFFF, h[2] >> 16, h[3] >> 16]
            out[8 * lane:8 * lane + 8] = np.array(x) - tau * eta
        assert np.array_equal(out, negacyclic(c, s))
