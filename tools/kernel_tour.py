"""One launch of every hot-path kernel at bench-like sizes (driver for ncu captures / compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dilithium_b200 as d
small = len(sys.argv) > 1 and sys.argv[1] == "small"
eng = d.Engine(0); Q = d.Q
B = 256 if small else 65536
for level in ((2,) if small else (2, 5)):
    k, l = d.LEVEL_DIMS[level]
    Bl = B if level == 2 else B // 2
    y = torch.randint(0, Q, (Bl, l, 256), dtype=torch.int32, device="cuda"); o = torch.empty_like(y)
    a_hat = torch.randint(0, Q, (k * l, 256), dtype=torch.int32, device="cuda")
    w = torch.empty((Bl, k, 256), dtype=torch.int32, device="cuda")
    rho = torch.randint(0, 256, (Bl, 32), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.ntt(y, out=o); eng.invntt(y, out=o); eng.pointwise_barrett(y, o, o); eng.add(y, o, o)
        eng.matvec(a_hat, y, k, l, w=w); eng.signcore(a_hat, y, k, l, w=w)
        eng.matvec_expand(rho[0], y, k, l, False, True, True, w=w)
        n = Bl // 8
        eng.matvec_expand(rho[:n], y[:n], k, l, True, True, True, w=w[:n])
        eng.expand_a(rho[:256], k, l)
torch.cuda.synchronize()
import oracle_lib as ol
K = ol.kat(2)
sk = d.SignKey(eng, 2, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
vk = d.VerifyKey(eng, 2, K["rho"][0], K["t1"][0])
msgs = [bytes([i & 255]) * 40 for i in range(64 if small else 4096)]
if len(sys.argv) > 2:   # spec_target: small values force the non-speculative round kernels on a small batch
    sk.set_tuning(spec_target=int(sys.argv[2]))
z, h, c, att = sk.sign(msgs)
assert vk.verify(msgs, z, h, c).all()
# host path with pinned outputs: finished signatures are drained round by round (drain_kernel)
zp, hp, cp, ap = sk.sign(msgs, pinned=True)
assert np.array_equal(z, zp) and np.array_equal(h, hp) and np.array_equal(c, cp) and np.array_equal(att, ap)
keys = eng.keygen(2, np.arange(64 * 32, dtype=np.uint8).reshape(64, 32))
# level 5: per-item-rho kernel with the two-warp split core (key generation and per-key verification use it)
K5 = ol.kat(5)
keys5 = eng.keygen(5, np.arange(16 * 32, dtype=np.uint8).reshape(16, 32))
sk5 = d.SignKey(eng, 5, K5["rho"][0], K5["k"][0], K5["tr"][0], K5["s1"][0], K5["s2"][0], K5["t0"][0])
m5 = msgs[:16]
z5, h5, c5, _ = sk5.sign(m5)
ok5 = eng.verify_multi(5, np.repeat(K5["rho"][:1], 16, axis=0), np.repeat(K5["t1"][:1], 16, axis=0), m5, z5, h5, c5)
assert ok5.all()
# one key per signature (per-item A in HBM, per-item tail), the fused mask core, device-resident key generation, a pool
zz, hh, cc, _ = eng.sign_multi(2, K["rho"][:24], K["k"][:24], K["tr"][:24], K["s1"][:24], K["s2"][:24], K["t0"][:24], msgs[:24])
assert eng.verify_multi(2, K["rho"][:24], K["t1"][:24], msgs[:24], zz, hh, cc).all()
if "fused" in sys.argv:   # the optional fused mask core (its producer / consumer warps synchronise through shared-memory flags)
    sk.set_tuning(fused_mask=True, spec_target=(int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0))
    zf, hf, cf, af = sk.sign(msgs)
    assert np.array_equal(z, zf) and np.array_equal(h, hf) and np.array_equal(c, cf) and np.array_equal(att, af)
kd = eng.keygen_dev(2, torch.arange(64 * 32, dtype=torch.uint8, device="cuda").reshape(64, 32))
torch.cuda.synchronize()
assert np.array_equal(kd["t1"].cpu().numpy(), keys["t1"])
pool = d.Pool([0, 0])
pool.load_key(2, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
pz, ph, pc, pa = pool.sign(msgs)
assert np.array_equal(z, pz) and np.array_equal(att, pa)
pool.close()
# three batches in flight on three handles of the same key (host path with pinned outputs: own streams, drain kernel, blocking waits)
import threading
sks = [d.SignKey(eng, 2, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0]) for _ in range(3)]
res = [None] * 3
def _worker(t):
    torch.cuda.set_device(0)
    for _ in range(2):
        res[t] = [np.array(a) for a in sks[t].sign(msgs, pinned=True)]
th = [threading.Thread(target=_worker, args=(t,)) for t in range(3)]
[x.start() for x in th]; [x.join() for x in th]
for r in res:
    assert np.array_equal(z, r[0]) and np.array_equal(h, r[1]) and np.array_equal(c, r[2]) and np.array_equal(att, r[3])
print("tour ok", eng.launch_count)
