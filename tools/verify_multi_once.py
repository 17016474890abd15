"""One device-resident per-item-key verification batch (for ncu launch lists / captures):  verify_multi_once.py [level] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dilithium_b200 as d
level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
eng = d.Engine(0)
if len(sys.argv) > 3:   # threshold of the row-streamed per-item kernel (huge value = one CTA per item)
    eng._lib.dil_diag_item_rows_threshold(int(sys.argv[3]))
K = np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz"))
sk = d.SignKey(eng, level, *[K[f][0] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
msgs = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
z = torch.empty((n, sk.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.empty((n, sk.h_bytes), dtype=torch.uint8, device="cuda")
c = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
sk.sign_dev(msgs, off, n, z, h, c, att)
rho = torch.from_numpy(np.repeat(K["rho"][:1], n, axis=0)).cuda(); t1 = torch.from_numpy(np.repeat(K["t1"][:1], n, axis=0)).cuda()
ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.verify_multi_dev(level, rho, t1, msgs, off, n, z, h, c, ok)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    eng.verify_multi_dev(level, rho, t1, msgs, off, n, z, h, c, ok)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"L{level} n={n} thr={sys.argv[3] if len(sys.argv) > 3 else 64}: accepted {int(ok.sum())} of {n}; {ms:.3f} ms per batch = {n / ms / 1e3:.3f} M verifies/s")
