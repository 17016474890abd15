"""CPU models of two bit-field extraction schemes used by the CUDA kernels (no GPU needed): they pin the index arithmetic that
the kernels implement with aligned 32-bit words and funnel shifts.

* item_inputs_zpacked (csrc/matvec_core.cuh): the verification core reads coefficient c = lane + 32 r of a packed z polynomial
  (18 / 20 bits per coefficient, encoder.v:96-133 / decoder.v:89-143) from the two aligned words around its bit field.
* expand_mask_kernel's cooperative unpack (csrc/sign_kernels.cu): lane L takes the 18 / 20 bytes holding coefficients
  8 L .. 8 L + 7 as five aligned words (18-byte chunks of odd lanes start two bytes into a word)."""
import numpy as np
import pytest


def funnelshift_r(lo, hi, sh):
    return ((((hi & 0xFFFFFFFF) << 32) | (lo & 0xFFFFFFFF)) >> (sh & 31)) & 0xFFFFFFFF


def pack(vals, bits):
    acc, out = 0, bytearray()
    for i, v in enumerate(vals):
        acc |= int(v) << (bits * i)
    return np.frombuffer(acc.to_bytes(len(vals) * bits // 8, "little"), dtype=np.uint8).copy()


@pytest.mark.parametrize("bits", [18, 20])
def test_zpacked_two_word_extraction(bits):
    rng = np.random.default_rng(bits)
    vals = rng.integers(0, 1 << bits, 256)
    vals[[0, 1, 254, 255]] = [(1 << bits) - 1, 0, (1 << bits) - 1, 1]
    row = pack(vals, bits)
    rowb = 32 * bits
    assert row.size == rowb
    words = row.view("<u4")
    for lane in range(32):
        for r in range(8):
            c = lane + 32 * r
            bit = c * bits
            byte, aw = bit >> 3, (bit >> 3) >> 2
            sh = (byte & 3) * 8 + (bit & 7)
            w0 = int(words[aw])
            last = 4 * aw + 4 >= rowb
            w1 = 0 if last else int(words[aw + 1])
            if last:   # the field must then lie inside the last word: nothing beyond the row is ever needed
                assert sh + bits <= 32
            assert funnelshift_r(w0, w1, sh) & ((1 << bits) - 1) == vals[c], (bits, c)


@pytest.mark.parametrize("bits", [18, 20])
def test_expand_mask_five_word_unpack(bits):
    rng = np.random.default_rng(100 + bits)
    vals = rng.integers(0, 1 << bits, 256)
    row = np.concatenate([pack(vals, bits), np.zeros(16, np.uint8)])   # the kernel's rows carry 16 bytes of padding
    words = row.view("<u4")
    g1 = 1 << (bits - 1)
    for lane in range(32):
        off = lane * bits                                  # bytes: 8 coefficients x bits / 8
        wd = [int(words[(off & ~3) // 4 + i]) for i in range(5)]
        assert (off & ~3) + 20 <= 32 * bits                # the five words never leave the polynomial's bytes
        if bits == 18:
            s = (off & 2) * 8
            wd = [funnelshift_r(wd[i], wd[i + 1], s) for i in range(4)] + [wd[4] >> s]
        for c in range(8):
            pos = c * bits
            wi, sh = pos >> 5, pos & 31
            x = wd[wi] >> sh if sh + bits <= 32 else funnelshift_r(wd[wi], wd[wi + 1], sh)
            if sh + bits != 32:
                x &= (1 << bits) - 1
            assert g1 - x == g1 - int(vals[8 * lane + c]), (bits, lane, c)
