"""One device-resident key-generation batch (driver for ncu launch lists): python tools/keygen_once.py [level] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dilithium_b200 as d
level = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
eng = d.Engine(0)
seeds = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
for _ in range(2):
    keys = eng.keygen_dev(level, seeds)
torch.cuda.synchronize()
print("keygen ok", level, n, eng.launch_count)
