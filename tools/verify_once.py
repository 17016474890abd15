"""One device-resident shared-key verification batch (driver for ncu captures): python tools/verify_once.py [level] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dilithium_b200 as d
level = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
eng = d.Engine(0)
K = np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz"))
sk = d.SignKey(eng, level, *[K[f][0] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
vk = d.VerifyKey(eng, level, K["rho"][0], K["t1"][0])
msgs = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
z = torch.empty((n, sk.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.empty((n, sk.h_bytes), dtype=torch.uint8, device="cuda")
c = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
sk.sign_dev(msgs, off, n, z, h, c, att)
ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    vk.verify_dev(msgs, off, n, z, h, c, ok)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    vk.verify_dev(msgs, off, n, z, h, c, ok)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"L{level} n={n}: accepted {int(ok.sum())} of {n}; {ms:.3f} ms per batch = {n / ms / 1e3:.2f} M verifies/s")
