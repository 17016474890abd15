"""GPU parity tests (run with -m gpu on the B200 box): every C-ABI entry point against the
CPU oracle on the same seeded inputs, against the committed golden vectors produced by the
reference's own C++, and - at BASELINE.json's full sizes - through size-independent
properties (round trips, linearity, checksums).  Integer work: the bar is bit-exact."""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
Q = ol.Q
LEVELS = {2: (4, 4), 3: (6, 5), 5: (8, 7)}


@pytest.fixture(scope="module")
def eng():
    import dilithium_b200 as d
    return d.Engine(0)


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ntt_golden.npz"))


def rnd(shape, seed, signed=False):
    rng = np.random.default_rng(seed)
    lo = -Q + 1 if signed else 0
    return rng.integers(lo, Q, size=shape).astype(np.int32)


# ---- golden vectors from the reference's own compiled C++ ----
def test_golden_ntt_invntt_pointwise(eng, g):
    assert np.array_equal(eng.ntt(g["x"]), g["ntt"])
    assert np.array_equal(eng.invntt(g["x"]), g["invntt"])
    assert np.array_equal(eng.ntt2x2_ref(g["x"]), g["ntt2x2"])
    assert np.array_equal(eng.invntt2x2_ref(g["x"]), g["invntt2x2"])
    assert np.array_equal(eng.pointwise_barrett(g["x"], g["y"]), g["pointwise"])


# ---- oracle parity on seeded inputs, ragged sizes ----
@pytest.mark.parametrize("n", [1, 2, 7, 31, 32, 33, 255, 1000, 4737, 20000])
def test_ntt_invntt_vs_oracle(eng, oracle, n):
    x = rnd((n, 256), 100 + n, signed=(n % 2 == 1))
    assert np.array_equal(eng.ntt(x), oracle.ntt(x))
    assert np.array_equal(eng.invntt(x), oracle.invntt(x))


def test_empty_batches(eng):
    e = np.zeros((0, 256), dtype=np.int32)
    assert eng.ntt(e).shape == (0, 256) and eng.invntt(e).shape == (0, 256)
    assert eng.pointwise_barrett(e, e).shape == (0, 256)
    assert eng.matvec(np.zeros((16, 256), np.int32), np.zeros((0, 4, 256), np.int32), 4, 4).shape == (0, 4, 256)


def test_edge_values(eng, oracle):
    x = np.zeros((6, 256), dtype=np.int32)
    x[1] = Q - 1
    x[2] = -(Q - 1)
    x[3, ::2] = Q - 1
    x[4, 0] = 1
    x[5, 255] = -1
    assert np.array_equal(eng.ntt(x), oracle.ntt(x))
    assert np.array_equal(eng.invntt(x), oracle.invntt(x))
    out = eng.ntt(x)
    assert out.min() >= 0 and out.max() < Q


def test_elementwise_vs_oracle(eng, oracle):
    a, b, c = rnd((777, 256), 1, True), rnd((777, 256), 2, True), rnd((777, 256), 3, True)
    A, B, C = (v.astype(np.int64) for v in (a, b, c))
    assert np.array_equal(eng.pointwise_barrett(a, b), oracle.pointwise(a, b))
    assert np.array_equal(eng.pointwise_acc(c, a, b), ((C + A * B) % Q).astype(np.int32))
    assert np.array_equal(eng.add(a, b), oracle.addsub(a, b))
    assert np.array_equal(eng.sub(a, b), oracle.addsub(a, b, sub=True))


@pytest.mark.parametrize("level", [2, 3, 5])
def test_matvec_shared_vs_oracle(eng, oracle, level):
    k, l = LEVELS[level]
    a_hat = rnd((k * l, 256), 10 + level)
    v = rnd((203, l, 256), 20 + level, signed=True)
    assert np.array_equal(eng.matvec(a_hat, v, k, l), oracle.matvec(a_hat, v, k, l))


def test_matvec_generic_dims(eng, oracle):
    for (k, l) in ((1, 1), (3, 2), (8, 8)):
        a_hat = rnd((k * l, 256), 31 * k + l)
        v = rnd((17, l, 256), 7 * k + l)
        assert np.array_equal(eng.matvec(a_hat, v, k, l), oracle.matvec(a_hat, v, k, l))


@pytest.mark.parametrize("level", [2, 3, 5])
def test_expand_a_vs_oracle_on_kat_rho(eng, oracle, level):
    k, l = LEVELS[level]
    K = ol.kat(level)
    rho = K["rho"][:12]
    got = eng.expand_a(rho, k, l)
    for r in range(rho.shape[0]):
        assert np.array_equal(got[r], oracle.expand_a(rho[r], k, l)), (level, r)


@pytest.mark.parametrize("level", [2, 3, 5])
def test_signcore_vs_oracle(eng, oracle, level):
    k, l = LEVELS[level]
    a_hat = rnd((k * l, 256), 40 + level)
    y = rnd((157, l, 256), 50 + level, signed=True)
    w_ref, _ = oracle.signcore(a_hat, y, k, l, threads=4)
    assert np.array_equal(eng.signcore(a_hat, y, k, l), w_ref)
    # the three-kernel composition must agree with the fused kernel
    comp = eng.invntt(eng.matvec(a_hat, eng.ntt(y), k, l))
    assert np.array_equal(comp, w_ref)


@pytest.mark.parametrize("level", [2, 3, 5])
@pytest.mark.parametrize("ntt_in,intt_out", [(False, False), (True, False), (False, True), (True, True)])
def test_matvec_expand_shared_rho(eng, oracle, level, ntt_in, intt_out):
    k, l = LEVELS[level]
    rho = ol.kat(level)["rho"][3]
    v = rnd((41, l, 256), 60 + level)
    ref = oracle.matvec_expand(rho, v, k, l, ntt_in, intt_out)
    assert np.array_equal(eng.matvec_expand(rho, v, k, l, False, ntt_in, intt_out), ref)


@pytest.mark.parametrize("level", [2, 3, 5])
def test_matvec_expand_per_item_rho(eng, oracle, level):
    k, l = LEVELS[level]
    rho = ol.kat(level)["rho"][:37]
    v = rnd((37, l, 256), 70 + level)
    for (ni, io) in ((False, False), (True, True)):
        ref = oracle.matvec_expand(rho, v, k, l, ni, io)
        assert np.array_equal(eng.matvec_expand(rho, v, k, l, True, ni, io), ref)


# ---- KAT chain through the engine: (rho, s1, s2) -> t ----
@pytest.mark.parametrize("level", [2, 3, 5])
def test_kat_keygen_chain_through_engine(eng, oracle, level):
    """t = INTT(ExpandA(rho) * NTT(s1)) + s2 computed by the engine must reproduce t1/t0 of all
    100 KATs (Power2Round + packing done by numpy here: they are codec, not hot path)."""
    k, l = LEVELS[level]
    K = ol.kat(level)
    P = ol.PARAMS[level]
    eta = P["eta"]
    w = 3 if eta == 2 else 4

    def unpack(buf, npoly, width):
        bits = np.unpackbits(buf.reshape(100, -1), axis=1, bitorder="little").reshape(100, npoly, 256, width)
        return (bits * (1 << np.arange(width))).sum(axis=3).astype(np.int32)

    s1 = eta - unpack(K["s1"], l, w)
    s2 = eta - unpack(K["s2"], k, w)
    t = eng.matvec_expand(K["rho"], s1, k, l, per_item=True, ntt_input=True, intt_output=True)
    t = eng.add(t, s2)
    t1 = (t + (1 << 12) - 1) >> 13
    t0 = t - (t1 << 13)
    assert np.array_equal(t1, unpack(K["t1"], k, 10))
    assert np.array_equal((1 << 12) - t0, unpack(K["t0"], k, 13))


# ---- full-size properties (BASELINE cfg2: L2, B = 65536) ----
def test_full_size_roundtrip_linearity_checksum(eng, oracle):
    import torch
    B, k, l = 65536, 4, 4
    gen = torch.Generator(device="cuda").manual_seed(0x44494C32)
    y = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device="cuda", generator=gen)
    z = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device="cuda", generator=gen)
    yh = eng.ntt(y)
    assert torch.equal(eng.invntt(yh), y)                                 # round trip
    zh = eng.ntt(z)
    assert torch.equal(eng.ntt(eng.add(y, z)), eng.add(yh, zh))           # linearity
    assert int(yh.min()) >= 0 and int(yh.max()) < Q
    # in-place == out-of-place
    y2 = y.clone()
    eng.ntt(y2, out=y2)
    assert torch.equal(y2, yh)
    # sampled oracle parity + checksum of the whole batch against the oracle on a strided sample
    idx = torch.arange(0, B, 16, device="cuda")
    ys = y[idx].cpu().numpy()
    assert np.array_equal(yh[idx].cpu().numpy(), oracle.ntt(ys, threads=8))
    a_hat = torch.randint(0, Q, (k * l, 256), dtype=torch.int32, device="cuda", generator=gen)
    w = eng.signcore(a_hat, y, k, l)
    w3 = eng.invntt(eng.matvec(a_hat, yh, k, l))
    assert torch.equal(w, w3)
    w_ref, _ = oracle.signcore(a_hat.cpu().numpy(), ys, k, l, threads=8)
    assert np.array_equal(w[idx].cpu().numpy(), w_ref)
    # polynomial product property: INTT(NTT(a) o NTT(b)) with b = X  is a negacyclic shift
    xpoly = torch.zeros((1, 256), dtype=torch.int32, device="cuda")
    xpoly[0, 1] = 1
    xh = eng.ntt(xpoly).expand(B * l, 256).contiguous()
    prod = eng.invntt(eng.pointwise_barrett(yh.view(-1, 256), xh)).view(B, l, 256)
    shifted = torch.roll(y, 1, dims=2)
    shifted[:, :, 0] = (Q - shifted[:, :, 0]) % Q
    assert torch.equal(prod, shifted)
