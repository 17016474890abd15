"""GPU parity of batched signing through the C ABI: bit-exact z, h, c~
against all 100 KAT vectors at levels 2/3/5 (the KATs are deterministic round-3.1 signatures),
and against the oracle's sign on random messages for one key (batch path with rejection rounds)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import dilithium_b200 as d
    return d.Engine(0)


@pytest.mark.parametrize("level", [2, 3, 5])
def test_sign_all_kats(eng, level):
    import dilithium_b200 as d
    K = ol.kat(level)
    attempts = []
    for i in range(100):
        key = d.SignKey(eng, level, K["rho"][i], K["k"][i], K["tr"][i], K["s1"][i], K["s2"][i], K["t0"][i])
        z, h, c, att = key.sign([K["msgs"][i]])
        assert np.array_equal(c[0], K["c"][i]), (level, i, "ctilde")
        assert np.array_equal(z[0], K["zs"][i]), (level, i, "z")
        assert np.array_equal(h[0], K["h"][i]), (level, i, "h")
        attempts.append(int(att[0]))
        key.close()
    exp = {2: (4.21, 17), 3: (4.23, 19), 5: (4.40, 24)}[level]
    assert abs(np.mean(attempts) - exp[0]) < 0.01 and max(attempts) == exp[1]


@pytest.mark.parametrize("level", [2, 3, 5])
def test_sign_batch_vs_oracle(eng, oracle, level):
    import dilithium_b200 as d
    K = ol.kat(level)
    i = 7
    key = d.SignKey(eng, level, K["rho"][i], K["k"][i], K["tr"][i], K["s1"][i], K["s2"][i], K["t0"][i])
    rng = np.random.default_rng(level)
    msgs = [b"", b"a", bytes(135), bytes(range(104)), bytes(rng.integers(0, 256, 3301).astype(np.uint8))]
    msgs += [bytes(rng.integers(0, 256, int(rng.integers(1, 300))).astype(np.uint8)) for _ in range(120)]
    z, h, c, att = key.sign(msgs)
    assert 1 <= key.last_rounds <= int(att.max())
    for m in range(len(msgs)):
        zo, ho, co, a = oracle.sign(level, K["rho"][i], K["k"][i], K["tr"][i], K["s1"][i], K["s2"][i], K["t0"][i], msgs[m])
        assert np.array_equal(c[m], co) and np.array_equal(z[m], zo) and np.array_equal(h[m], ho) and att[m] == a, (level, m)
        assert oracle.verify(level, K["rho"][i], K["t1"][i], msgs[m], z[m], h[m], c[m]) == 0


def test_sign_large_batch_properties(eng, oracle):
    """BASELINE-size batch (65 536 Dilithium-2 signatures, one key): every signature of a strided
    sample verifies under the oracle's verify, attempts follow the expected geometric law, and
    signing the same batch twice is bit-identical (deterministic signing)."""
    import dilithium_b200 as d
    level, n = 2, 65536
    K = ol.kat(level)
    key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
    msgs = [int(i).to_bytes(4, "little") * 8 for i in range(n)]
    z, h, c, att = key.sign(msgs)
    assert att.min() >= 1 and 3.9 < att.mean() < 4.6
    for m in range(0, n, 997):
        assert oracle.verify(level, K["rho"][0], K["t1"][0], msgs[m], z[m], h[m], c[m]) == 0, m
    z2, h2, c2, att2 = key.sign(msgs)
    assert np.array_equal(z, z2) and np.array_equal(h, h2) and np.array_equal(c, c2) and np.array_equal(att, att2)


def test_sign_streaming_host_path(eng):
    """dil_sign_batch_host with pinned output buffers streams finished signatures out round by round
    (drain_kernel into mapped host memory); results must equal the pageable-buffer path (chunked
    copy-engine transfers) bit for bit, also when the batch is split (tuning host_chunk) and for ragged sizes."""
    import dilithium_b200 as d
    for level, n in ((2, 20000), (3, 3001), (5, 1025), (2, 1), (2, 0)):
        K = ol.kat(level)
        key = d.SignKey(eng, level, K["rho"][1], K["k"][1], K["tr"][1], K["s1"][1], K["s2"][1], K["t0"][1])
        msgs = [int(i).to_bytes(4, "little") * (1 + i % 9) for i in range(n)]
        ref = key.sign(msgs)
        got = [np.array(a) for a in key.sign(msgs, pinned=True)]
        key.set_tuning(host_chunk=4096)
        got2 = [np.array(a) for a in key.sign(msgs, pinned=True)]
        key.set_tuning(host_chunk=4096, host_copy_path=True)
        got3 = [np.array(a) for a in key.sign(msgs, pinned=True)]
        key.set_tuning()
        for name, r, a, b, c in zip(("z", "h", "c", "att"), ref, got, got2, got3):
            assert np.array_equal(r, a), (level, n, name, "streaming")
            assert np.array_equal(r, b), (level, n, name, "streaming, split batch")
            assert np.array_equal(r, c), (level, n, name, "pinned buffers, copy path")
        key.close()


def test_sign_dev_split_batch(eng):
    """dil_sign_batch_dev signs very large batches in pieces (bounded workspace); with a small piece size the
    result must equal the single-piece result bit for bit."""
    import torch
    import dilithium_b200 as d
    level, n = 2, 9001
    K = ol.kat(level)
    key = d.SignKey(eng, level, K["rho"][2], K["k"][2], K["tr"][2], K["s1"][2], K["s2"][2], K["t0"][2])
    msgs = torch.randint(0, 256, (n * 24,), dtype=torch.uint8, device="cuda")
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 24
    outs = []
    for chunk in (0, 2048):
        key.set_tuning(dev_chunk=chunk)
        z = torch.zeros((n, key.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.zeros((n, key.h_bytes), dtype=torch.uint8, device="cuda")
        c = torch.zeros((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
        key.sign_dev(msgs, off, n, z, h, c, att)
        torch.cuda.synchronize()
        outs.append((z, h, c, att))
    key.set_tuning()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert int(outs[0][3].min()) >= 1
    key.close()


def test_sign_scheduler_tuning_does_not_change_results(eng):
    """The round scheduler lives on the device (RoundCtl): whatever the speculation policy, the smallest accepted
    kappa wins, so signatures and attempt counts must not depend on it."""
    import dilithium_b200 as d
    level, n = 2, 3000
    K = ol.kat(level)
    key = d.SignKey(eng, level, K["rho"][3], K["k"][3], K["tr"][3], K["s1"][3], K["s2"][3], K["t0"][3])
    msgs = [int(i).to_bytes(4, "little") * (1 + i % 5) for i in range(n)]
    ref = key.sign(msgs)
    rounds = {}
    for name, kw in (("no speculation", dict(spec_target=1)), ("two slots per item", dict(spec_target=1 << 20, spec_max=2)),
                     ("small target", dict(spec_target=512, spec_max=7)), ("always 32 slots", dict(spec_target=1 << 22, spec_max=32))):
        key.set_tuning(**kw)
        got = key.sign(msgs)
        rounds[name] = (key.last_rounds, key.last_slots)
        for f, a, b in zip(("z", "h", "c", "att"), ref, got):
            assert np.array_equal(a, b), (name, f)
    # without speculation a batch needs as many rounds as its unluckiest message needs attempts
    assert rounds["no speculation"][0] == int(ref[3].max())
    assert rounds["no speculation"][1] == int(ref[3].sum())
    assert rounds["always 32 slots"][0] <= 2
    key.close()


@pytest.mark.parametrize("level", [2, 3, 5])
def test_sign_multi_all_kats_in_one_batch(eng, oracle, level):
    """Key per signature, as rtl_tb/tb_sign_top.v:171-284 streams it: all 100 KAT vectors of a level (100 different
    keys) signed as ONE batch must reproduce the KAT signatures bit for bit; a second batch mixes keys and random
    messages (several messages per key, empty and long messages) and is checked against the oracle."""
    K = ol.kat(level)
    msgs = list(K["msgs"])
    z, h, c, att = eng.sign_multi(level, K["rho"], K["k"], K["tr"], K["s1"], K["s2"], K["t0"], msgs)
    assert np.array_equal(c, K["c"]) and np.array_equal(z, K["zs"]) and np.array_equal(h, K["h"])
    exp = {2: (4.21, 17), 3: (4.23, 19), 5: (4.40, 24)}[level]
    assert abs(att.mean() - exp[0]) < 0.01 and att.max() == exp[1]
    rng = np.random.default_rng(100 + level)
    n = 257
    kidx = rng.integers(0, 100, n)
    msgs2 = [b"", bytes(3301)] + [bytes(rng.integers(0, 256, int(rng.integers(1, 200))).astype(np.uint8)) for _ in range(n - 2)]
    z, h, c, att = eng.sign_multi(level, *[K[f][kidx] for f in ("rho", "k", "tr", "s1", "s2", "t0")], msgs2)
    for m in range(0, n, 5):
        i = kidx[m]
        zo, ho, co, a = oracle.sign(level, K["rho"][i], K["k"][i], K["tr"][i], K["s1"][i], K["s2"][i], K["t0"][i], msgs2[m])
        assert np.array_equal(c[m], co) and np.array_equal(z[m], zo) and np.array_equal(h[m], ho) and att[m] == a, (level, m)
    ok = eng.verify_multi(level, K["rho"][kidx], K["t1"][kidx], msgs2, z, h, c)
    assert ok.tolist() == [1] * n


def test_sign_batches_in_flight(eng, oracle):
    """Streaming use of the API: several key handles of the same key sign different batches at the same time from their
    own host threads (host path: the handle's own streams and the round-by-round drain into pinned buffers; device path:
    one torch stream per thread).  The engine speculates less when it sees the load (sign_api.cu spec_for); every
    batch must equal what a lone call produces, bit for bit, and a sample must match the oracle."""
    import threading
    import torch
    import dilithium_b200 as d
    level, n, T = 2, 6000, 4
    K = ol.kat(level)
    parts = [K[f][5] for f in ("rho", "k", "tr", "s1", "s2", "t0")]
    keys = [d.SignKey(eng, level, *parts) for _ in range(T)]
    batches = [[(t * n + i).to_bytes(4, "little") * (1 + (i + t) % 7) for i in range(n)] for t in range(T)]
    lone = [[np.array(a) for a in keys[0].sign(b, pinned=True)] for b in batches]
    lone_rounds = keys[0].last_rounds
    for m in range(0, n, 601):
        zo, ho, co, a = oracle.sign(level, *parts, batches[1][m])
        assert np.array_equal(lone[1][0][m], zo) and np.array_equal(lone[1][1][m], ho) and np.array_equal(lone[1][2][m], co) and lone[1][3][m] == a
    # host path, T threads x 3 calls each
    got, errs = [None] * T, []

    def host_worker(t):
        try:
            torch.cuda.set_device(0)
            for _ in range(3):
                got[t] = [np.array(a) for a in keys[t].sign(batches[t], pinned=True)]
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=host_worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    for t in range(T):
        for name, a, b in zip(("z", "h", "c", "att"), lone[t], got[t]):
            assert np.array_equal(a, b), (t, name, "host path in flight")
    # device path on T streams
    dev_in, dev_out, streams = [], [], [torch.cuda.Stream() for _ in range(T)]
    for t in range(T):
        blob = np.frombuffer(b"".join(batches[t]), dtype=np.uint8).copy()
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in batches[t]])
        dev_in.append((torch.from_numpy(blob).cuda(), torch.from_numpy(off).cuda()))
        dev_out.append((torch.zeros((n, keys[t].z_bytes), dtype=torch.uint8, device="cuda"), torch.zeros((n, keys[t].h_bytes), dtype=torch.uint8, device="cuda"),
                        torch.zeros((n, 32), dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()

    def dev_worker(t):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[t]):
                for _ in range(3):
                    keys[t].sign_dev(dev_in[t][0], dev_in[t][1], n, *dev_out[t])
            streams[t].synchronize()
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=dev_worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    for t in range(T):
        for name, a, b in zip(("z", "h", "c", "att"), lone[t], dev_out[t]):
            assert np.array_equal(a, b.cpu().numpy()), (t, name, "device path in flight")
    assert lone_rounds >= 1
    for k in keys:
        k.close()


def test_sign_async_begin_finish(eng, oracle):
    """dil_sign_batch_{dev,host}_begin + dil_sign_batch_finish: one host thread keeps a batch in flight on each of three key
    handles (different messages each), several times over; results equal the synchronous calls bit for bit, a busy handle
    refuses a second batch, and a begun batch that is never finished is completed by dil_sign_key_destroy."""
    import torch
    import dilithium_b200 as d
    level, n, T = 3, 2500, 3
    K = ol.kat(level)
    parts = [K[f][9] for f in ("rho", "k", "tr", "s1", "s2", "t0")]
    keys = [d.SignKey(eng, level, *parts) for _ in range(T)]
    batches = [[(t * n + i).to_bytes(4, "little") * (1 + (i + 2 * t) % 6) for i in range(n)] for t in range(T)]
    ref = [[np.array(a) for a in keys[0].sign(b)] for b in batches]
    zo, ho, co, a = oracle.sign(level, *parts, batches[2][17])
    assert np.array_equal(ref[2][0][17], zo) and np.array_equal(ref[2][1][17], ho) and np.array_equal(ref[2][2][17], co) and ref[2][3][17] == a
    dev_in, dev_out, host_in, host_out = [], [], [], []
    for t in range(T):
        blob = np.frombuffer(b"".join(batches[t]), dtype=np.uint8).copy()
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in batches[t]])
        host_in.append((torch.from_numpy(blob).pin_memory(), torch.from_numpy(off).pin_memory()))
        dev_in.append((host_in[t][0].cuda(), host_in[t][1].cuda()))
        shapes = ((n, keys[t].z_bytes), (n, keys[t].h_bytes), (n, 32))
        dev_out.append([torch.zeros(s, dtype=torch.uint8, device="cuda") for s in shapes] + [torch.zeros(n, dtype=torch.int32, device="cuda")])
        host_out.append([torch.zeros(s, dtype=torch.uint8).pin_memory() for s in shapes] + [torch.zeros(n, dtype=torch.int32).pin_memory()])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(T)]
    for rep in range(3):                    # device path: begin all, then finish all, from this one thread
        for t in range(T):
            with torch.cuda.stream(streams[t]):
                keys[t].sign_dev_begin(dev_in[t][0], dev_in[t][1], n, *dev_out[t])
        with pytest.raises(RuntimeError):   # a busy handle refuses a second batch
            keys[0].sign_dev(dev_in[0][0], dev_in[0][1], n, *dev_out[0])
        for t in range(T):
            keys[t].finish()
    for t in range(T):
        for name, r, g in zip(("z", "h", "c", "att"), ref[t], dev_out[t]):
            assert np.array_equal(r, g.cpu().numpy()), (t, name, "async device path")
    for rep in range(3):                    # host path: pinned messages in, signatures streamed into pinned buffers
        for t in range(T):
            keys[t].sign_host_begin(host_in[t][0], host_in[t][1], n, *host_out[t])
        for t in reversed(range(T)):
            keys[t].finish()
    for t in range(T):
        for name, r, g in zip(("z", "h", "c", "att"), ref[t], host_out[t]):
            assert np.array_equal(r, g.numpy()), (t, name, "async host path")
    with pytest.raises(RuntimeError):       # nothing in flight
        keys[1].finish()
    pageable = [torch.zeros((n, keys[0].z_bytes), dtype=torch.uint8), torch.zeros((n, keys[0].h_bytes), dtype=torch.uint8),
                torch.zeros((n, 32), dtype=torch.uint8), torch.zeros(n, dtype=torch.int32)]
    with pytest.raises(RuntimeError):       # the asynchronous host path needs pinned outputs
        keys[1].sign_host_begin(host_in[1][0], host_in[1][1], n, *pageable)
    keys[2].sign_dev_begin(dev_in[2][0], dev_in[2][1], n, *dev_out[2])   # never finished: destroy completes it
    for k in keys:
        k.close()
    torch.cuda.synchronize()
    assert np.array_equal(ref[2][0], dev_out[2][0].cpu().numpy())
