"""One process, engines on two GPUs: every kernel that needs a per-device launch attribute (dynamic shared
memory) must be configured on each device it runs on.  Skipped on single-GPU boxes."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_two_engines_one_process(oracle):
    import torch
    import dilithium_b200 as d
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    level = 2
    K = ol.kat(level)
    msgs = [int(i).to_bytes(4, "little") * 5 for i in range(3000)]
    rng = np.random.default_rng(3)
    y = rng.integers(0, d.Q, size=(300, 4, 256)).astype(np.int32)
    rho = rng.integers(0, 256, size=32).astype(np.uint8)
    results = []
    for dev in (0, 1):
        eng = d.Engine(dev)
        key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
        vk = d.VerifyKey(eng, level, K["rho"][0], K["t1"][0])
        z, h, c, att = [np.array(a) for a in key.sign(msgs, pinned=True)]
        assert vk.verify(msgs, z, h, c).all()
        w = eng.matvec_expand(rho, y, 4, 4, ntt_input=True, intt_output=True)
        results.append((z, h, c, att, w, eng.ntt(y)))
        key.close(); vk.close(); eng.close()
    for a, b in zip(*results):
        assert np.array_equal(a, b)
    z, h, c, att = results[0][:4]
    for m in (0, 1234, 2999):
        zo, ho, co, a = oracle.sign(level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0], msgs[m])
        assert np.array_equal(z[m], zo) and np.array_equal(h[m], ho) and np.array_equal(c[m], co) and att[m] == a
    assert np.array_equal(results[0][4], oracle.matvec_expand(rho, y, 4, 4, True, True))
