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


def test_pool_signs_like_one_engine(oracle):
    """dil_pool_*: one process, several engines (every visible GPU; on a single-GPU box two engines share device 0, which
    also exercises concurrent host threads on one device).  The pool's contiguous shards must reproduce the signatures
    of one engine bit for bit, for ragged message lengths and a batch that does not divide evenly."""
    import torch
    import dilithium_b200 as d
    level = 3
    K = ol.kat(level)
    kk = [K[f][5] for f in ("rho", "k", "tr", "s1", "s2", "t0")]
    rng = np.random.default_rng(8)
    n = 5003
    msgs = [bytes(rng.integers(0, 256, int(rng.integers(0, 70))).astype(np.uint8)) for _ in range(n)]
    eng = d.Engine(0)
    key = d.SignKey(eng, level, *kk)
    ref = key.sign(msgs)
    key.close()
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    for devs in (devices, [0], [0, 0, 0]):
        pool = d.Pool(devs)
        assert pool.size == len(devs)
        pool.load_key(level, *kk)
        got = pool.sign(msgs)
        for name, a, b in zip(("z", "h", "c", "att"), ref, got):
            assert np.array_equal(a, b), (devs, name)
        # pinned, portable outputs: every engine streams its shard's signatures into its slice of the same buffers
        zt, ht, ct, at = (torch.empty(s, dtype=t).pin_memory() for s, t in (((n, ref[0].shape[1]), torch.uint8), ((n, ref[1].shape[1]), torch.uint8),
                                                                             ((n, 32), torch.uint8), ((n,), torch.int32)))
        off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy()
        pool.sign_into(blob.ctypes.data, off.ctypes.data, n, zt.data_ptr(), ht.data_ptr(), ct.data_ptr(), at.data_ptr())
        assert np.array_equal(zt.numpy(), ref[0]) and np.array_equal(ht.numpy(), ref[1]) and np.array_equal(ct.numpy(), ref[2])
        assert np.array_equal(at.numpy().view(np.uint32), ref[3])
        # the asynchronous pair over the pool (dil_pool_sign_batch_host_begin / dil_pool_sign_batch_finish), twice over
        for _ in range(2):
            zt.zero_(); ht.zero_(); ct.zero_(); at.zero_()
            pool.sign_into_begin(blob.ctypes.data, off.ctypes.data, n, zt.data_ptr(), ht.data_ptr(), ct.data_ptr(), at.data_ptr())
            with pytest.raises(RuntimeError):   # one batch per pool key at a time
                pool.sign_into_begin(blob.ctypes.data, off.ctypes.data, n, zt.data_ptr(), ht.data_ptr(), ct.data_ptr(), at.data_ptr())
            pool.finish()
            assert np.array_equal(zt.numpy(), ref[0]) and np.array_equal(ht.numpy(), ref[1]) and np.array_equal(ct.numpy(), ref[2])
            assert np.array_equal(at.numpy().view(np.uint32), ref[3])
        pool.close()
    for m in (0, 2501, n - 1):
        zo, ho, co, a = oracle.sign(level, *kk, msgs[m])
        assert np.array_equal(ref[0][m], zo) and np.array_equal(ref[1][m], ho) and np.array_equal(ref[2][m], co) and ref[3][m] == a
