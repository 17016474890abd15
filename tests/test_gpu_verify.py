"""GPU parity of batched verification through the C ABI: every KAT signature (100 x levels 2/3/5)
is accepted, tampered inputs are rejected exactly as the oracle's verify rejects them, and a large
sign -> verify round trip on the device accepts everything."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import dilithium_b200 as d
    return d.Engine(0)


@pytest.mark.parametrize("level", [2, 3, 5])
def test_verify_all_kats_and_tampering(eng, oracle, level):
    import dilithium_b200 as d
    K = ol.kat(level)
    rng = np.random.default_rng(level)
    for i in range(100):
        vk = d.VerifyKey(eng, level, K["rho"][i], K["t1"][i])
        msg = K["msgs"][i]
        z, h, c = K["zs"][i].copy(), K["h"][i].copy(), K["c"][i].copy()
        cases = [(msg, z, h, c)]
        if i % 10 == 0:  # tampered variants
            z2 = z.copy(); z2[int(rng.integers(0, z.size))] ^= 1 << int(rng.integers(0, 8))
            c2 = c.copy(); c2[3] ^= 0x40
            h2 = h.copy(); h2[-1] = (int(h2[-1]) + 1) % 256
            h3 = h.copy(); h3[0], h3[1] = h[1], h[0]
            cases += [(msg, z2, h, c), (msg, z, h, c2), (msg, z, h2, c), (msg, z, h3, c), (msg + b"x", z, h, c)]
        ok = vk.verify([m for m, *_ in cases], np.stack([x[1] for x in cases]), np.stack([x[2] for x in cases]),
                       np.stack([x[3] for x in cases]))
        exp = [1 - oracle.verify(level, K["rho"][i], K["t1"][i], m, zz, hh, cc) for m, zz, hh, cc in cases]
        assert ok.tolist() == exp, (level, i, ok.tolist(), exp)
        assert ok[0] == 1
        vk.close()


def test_sign_then_verify_on_device(eng):
    import torch
    import dilithium_b200 as d
    level, n = 3, 20000
    K = ol.kat(level)
    sk = d.SignKey(eng, level, K["rho"][4], K["k"][4], K["tr"][4], K["s1"][4], K["s2"][4], K["t0"][4])
    vk = d.VerifyKey(eng, level, K["rho"][4], K["t1"][4])
    msgs = torch.randint(0, 256, (n * 48,), dtype=torch.uint8, device="cuda")
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 48
    z = torch.empty((n, sk.z_bytes), dtype=torch.uint8, device="cuda")
    h = torch.empty((n, sk.h_bytes), dtype=torch.uint8, device="cuda")
    c = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    att = torch.zeros(n, dtype=torch.int32, device="cuda")
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    sk.sign_dev(msgs, off, n, z, h, c, att)
    vk.verify_dev(msgs, off, n, z, h, c, ok)
    assert int(ok.sum()) == n
    # flip one bit in every 7th signature's z: exactly those must be rejected
    z2 = z.clone()
    z2[::7, 100] ^= 4
    vk.verify_dev(msgs, off, n, z2, h, c, ok)
    exp = torch.ones(n, dtype=torch.uint8, device="cuda")
    exp[::7] = 0
    assert torch.equal(ok, exp)


@pytest.mark.parametrize("level", [2, 3, 5])
def test_verify_multi_key_all_kats_in_one_batch(eng, oracle, level):
    """cfg4 style: 100 signatures under 100 different public keys verified as ONE batch (per-item rho:
    A expanded on chip per item), plus tampered copies that must be rejected."""
    K = ol.kat(level)
    msgs = list(K["msgs"])
    ok = eng.verify_multi(level, K["rho"], K["t1"], msgs, K["zs"], K["h"], K["c"])
    assert ok.tolist() == [1] * 100
    z2 = K["zs"].copy()
    z2[::3, 50] ^= 2
    t1x = K["t1"].copy()
    t1x[1::3, 7] ^= 1                      # wrong public key
    ok = eng.verify_multi(level, K["rho"], t1x, msgs, z2, K["h"], K["c"])
    exp = [0 if (i % 3 == 0 or i % 3 == 1) else 1 for i in range(100)]
    assert ok.tolist() == exp
