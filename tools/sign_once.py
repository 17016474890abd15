"""One device-resident sign batch (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import dilithium_b200 as d
import oracle_lib as ol
level = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
eng = d.Engine(0)
K = ol.kat(level)
key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
msgs = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
z = torch.empty((n, key.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.empty((n, key.h_bytes), dtype=torch.uint8, device="cuda")
c = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
key.sign_dev(msgs, off, n, z, h, c, att)
torch.cuda.synchronize()
print("rounds", key.last_rounds)
