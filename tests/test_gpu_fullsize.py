"""GPU parity at BASELINE.json's full sizes for the configurations beyond cfg2:

  cfg3  Dilithium-3 (k=6, l=5), 262 144 items: fused ExpandA -> NTT(y) -> A*y -> INTT core (shared rho and per-item rho)
        and full signing of a 262 144-message batch;
  cfg4  Dilithium-5 (k=8, l=7), one GPU's shard of the 1 M batch = 131 072 signatures, EVERY signature under its own
        public key (per-item rho: A generated on chip per item), with the 100 level-5 KAT tuples injected at fixed positions.

At these sizes the oracle cannot process every item in seconds, so each test compares a strided oracle sample bit for
bit and pins the rest of the batch through size-independent properties (linearity, fused == unfused composition,
sign -> verify round trip, determinism, tampered items rejected exactly where they were tampered)."""
import hashlib

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
Q = 8380417


@pytest.fixture(scope="module")
def eng():
    import dilithium_b200 as d
    return d.Engine(0)


def test_cfg3_full_size_fused_expand_core(eng, oracle):
    import torch
    B, k, l = 262144, 6, 5
    K = ol.kat(3)
    rho = torch.from_numpy(K["rho"][0].copy()).cuda()            # KAT-derived rho (rho_3.txt line 0)
    gen = torch.Generator(device="cuda").manual_seed(0x44494C33)
    y = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device="cuda", generator=gen)
    w = eng.matvec_expand(rho, y, k, l, ntt_input=True, intt_output=True)          # mode S: one rho for the batch
    assert int(w.min()) >= 0 and int(w.max()) < Q
    # strided oracle sample, bit for bit
    idx = torch.arange(0, B, 2048, device="cuda")
    ys = y[idx].cpu().numpy()
    assert np.array_equal(w[idx].cpu().numpy(), oracle.matvec_expand(K["rho"][0], ys, k, l, True, True))
    # fused == ExpandA -> NTT -> mat-vec -> INTT composed from the stand-alone kernels, on the whole batch
    a_hat = eng.expand_a(rho, k, l)[0].contiguous()
    assert np.array_equal(a_hat.cpu().numpy(), oracle.expand_a(K["rho"][0], k, l))
    yh = eng.ntt(y)
    w3 = eng.invntt(eng.matvec(a_hat, yh, k, l))
    assert torch.equal(w, w3)
    assert torch.equal(eng.signcore(a_hat, y, k, l), w)
    del w3, yh
    # linearity over the whole batch: core(y + y') == core(y) + core(y')
    y2 = torch.roll(y, 1, dims=0)
    w2 = torch.roll(w, 1, dims=0)
    wsum = eng.matvec_expand(rho, eng.add(y, y2), k, l, ntt_input=True, intt_output=True)
    assert torch.equal(wsum, eng.add(w, w2))
    del wsum, y2, w2
    # mode P: per-item rho = SHAKE-256(seed || item), on a 16 384-item slice; oracle on a strided sample of it
    n = 16384
    rho_p = np.stack([np.frombuffer(hashlib.shake_256(b"cfg3" + int(i).to_bytes(4, "little")).digest(32), dtype=np.uint8) for i in range(n)])
    rho_p[:100] = K["rho"]                                       # KAT-derived items
    wp = eng.matvec_expand(torch.from_numpy(rho_p).cuda(), y[:n].contiguous(), k, l, per_item=True, ntt_input=True, intt_output=True)
    for i in list(range(0, 100, 9)) + list(range(100, n, 1021)):
        ref = oracle.matvec_expand(rho_p[i], y[i:i + 1].cpu().numpy(), k, l, True, True)
        assert np.array_equal(wp[i:i + 1].cpu().numpy(), ref), i


def test_cfg3_full_size_sign(eng, oracle):
    """262 144 Dilithium-3 signatures in one batch: strided oracle verification AND bit-exact equality with the oracle's
    own signatures on a sample, attempt statistics, determinism, then batched verification of all of them."""
    import torch
    import dilithium_b200 as d
    level, n, mlen = 3, 262144, 32
    K = ol.kat(level)
    i = 11
    sk = d.SignKey(eng, level, *[K[f][i] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
    vk = d.VerifyKey(eng, level, K["rho"][i], K["t1"][i])
    gen = torch.Generator(device="cuda").manual_seed(3)
    msgs = torch.randint(0, 256, (n * mlen,), dtype=torch.uint8, device="cuda", generator=gen)
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * mlen
    outs = []
    for _ in range(2):
        z = torch.zeros((n, sk.z_bytes), dtype=torch.uint8, device="cuda")
        h = torch.zeros((n, sk.h_bytes), dtype=torch.uint8, device="cuda")
        c = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        att = torch.zeros(n, dtype=torch.int32, device="cuda")
        sk.sign_dev(msgs, off, n, z, h, c, att)
        torch.cuda.synchronize()
        outs.append((z, h, c, att))
    for a, b in zip(*outs):
        assert torch.equal(a, b)                                  # deterministic signing
    z, h, c, att = outs[0]
    # expected repetitions of Dilithium-3 are ~5.1 (the oracle signs random messages under this key with 5.13 on average;
    # BASELINE.md's 4.23 is the mean over the 100 KAT messages only)
    assert int(att.min()) >= 1 and 4.9 < float(att.float().mean()) < 5.4
    mh = msgs.view(n, mlen).cpu().numpy()
    for m in range(0, n, 8191):
        zo, ho, co, a = oracle.sign(level, *[K[f][i] for f in ("rho", "k", "tr", "s1", "s2", "t0")], mh[m].tobytes())
        assert np.array_equal(z[m].cpu().numpy(), zo) and np.array_equal(h[m].cpu().numpy(), ho)
        assert np.array_equal(c[m].cpu().numpy(), co) and int(att[m]) == a
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    vk.verify_dev(msgs, off, n, z, h, c, ok)
    assert int(ok.sum()) == n


def test_cfg4_full_shard_per_item_keys_with_kat_injection(eng, oracle):
    """One GPU's cfg4 shard: 131 072 Dilithium-5 signatures, each verified under ITS OWN public key (100 distinct
    keys in rotation, A expanded on chip per item), with the 100 level-5 KAT tuples (rho, t1, M, z, h, c~ straight
    from the reference's KAT files) injected at fixed positions."""
    import dilithium_b200 as d
    level, n, mlen = 5, 131072, 32
    K = ol.kat(level)
    rng = np.random.default_rng(54)
    key_of = np.arange(n) % 100
    msgs = [bytes(rng.integers(0, 256, mlen).astype(np.uint8)) for _ in range(n)]
    zb, hb = 7 * 640, 75 + 8
    z = np.empty((n, zb), np.uint8); h = np.empty((n, hb), np.uint8); c = np.empty((n, 32), np.uint8)
    for kidx in range(100):                                       # sign every key's share of the batch
        sk = d.SignKey(eng, level, *[K[f][kidx] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
        sel = np.nonzero(key_of == kidx)[0]
        zz, hh, cc, _ = sk.sign([msgs[j] for j in sel])
        z[sel], h[sel], c[sel] = zz, hh, cc
        sk.close()
    rho = K["rho"][key_of].copy(); t1 = K["t1"][key_of].copy()
    # KAT injection: tuple j at position 1310 * j + 7 (its own key index is j, so rho / t1 there are already KAT j's)
    pos = 1310 * np.arange(100) + 7
    assert len(set(pos)) == 100 and pos.max() < n
    for j, p in enumerate(pos):
        rho[p], t1[p] = K["rho"][j], K["t1"][j]
        msgs[p] = K["msgs"][j]
        z[p], h[p], c[p] = K["zs"][j], K["h"][j], K["c"][j]
    ok = eng.verify_multi(level, rho, t1, msgs, z, h, c)
    assert ok[pos].tolist() == [1] * 100                         # the KAT tuples yield the w that makes H(mu || w1') == c~
    assert int(ok.sum()) == n
    # strided sample against the oracle's verify (0 = accept)
    for m in range(3, n, 4099):
        assert oracle.verify(level, rho[m], t1[m], msgs[m], z[m], h[m], c[m]) == 0
    # tampering: flipped z bit / wrong key / wrong message at known positions -> rejected exactly there
    z2, t1x, msgs2 = z.copy(), t1.copy(), list(msgs)
    z2[5::1000, 77] ^= 8
    t1x[6::1000, 3] ^= 1
    for m in range(9, n, 1000):
        msgs2[m] = msgs2[m] + b"!"
    ok2 = eng.verify_multi(level, rho, t1x, msgs2, z2, h, c)
    exp = np.ones(n, np.uint8)
    exp[5::1000] = 0; exp[6::1000] = 0; exp[9::1000] = 0
    assert np.array_equal(ok2, exp)
    for m in (5, 6, 9, 1005, 2006, 3009):
        assert oracle.verify(level, rho[m], t1x[m], msgs2[m], z2[m], h[m], c[m]) != 0


def test_very_large_batch_is_signed_and_verified_in_pieces(eng):
    """Beyond 1.25 x 2^20 items dil_sign_batch_dev and dil_verify_batch_dev work in 2^20-item pieces (bounded workspace);
    1.4 M Dilithium-2 signatures: all accepted, tampered ones on both sides of the piece boundary rejected exactly."""
    import torch
    import dilithium_b200 as d
    level, n, mlen = 2, 1_400_000, 16
    K = ol.kat(level)
    sk = d.SignKey(eng, level, *[K[f][9] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
    vk = d.VerifyKey(eng, level, K["rho"][9], K["t1"][9])
    gen = torch.Generator(device="cuda").manual_seed(14)
    msgs = torch.randint(0, 256, (n * mlen,), dtype=torch.uint8, device="cuda", generator=gen)
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * mlen
    z = torch.zeros((n, sk.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.zeros((n, sk.h_bytes), dtype=torch.uint8, device="cuda")
    c = torch.zeros((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    sk.sign_dev(msgs, off, n, z, h, c, att)
    vk.verify_dev(msgs, off, n, z, h, c, ok)
    assert int(att.min()) >= 1 and int(ok.sum()) == n
    bad = torch.tensor([0, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, n - 1], device="cuda")
    z[bad, 40] ^= 1
    vk.verify_dev(msgs, off, n, z, h, c, ok)
    exp = torch.ones(n, dtype=torch.uint8, device="cuda")
    exp[bad] = 0
    assert torch.equal(ok, exp)
    sk.close(); vk.close()
