"""Experiment: do two independent sign batches on two streams overlap (Keccak-bound vs multiply-bound kernels)?"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dilithium_b200 as d
import oracle_lib as ol

level, n, T = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = d.Engine(0)
K = ol.kat(level)
ctx = []
for t in range(T):
    key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
    msgs = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
    z = torch.empty((n, key.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.empty((n, key.h_bytes), dtype=torch.uint8, device="cuda")
    c = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.append((key, msgs, off, z, h, c, att, torch.cuda.Stream()))

def run(i, reps):
    key, msgs, off, z, h, c, att, st = ctx[i]
    with torch.cuda.stream(st):
        for _ in range(reps):
            key.sign_dev(msgs, off, n, z, h, c, att)
    st.synchronize()

for i in range(T): run(i, 1)
reps = 5
t0 = time.perf_counter()
for i in range(T): run(i, reps)
seq = time.perf_counter() - t0
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(i, reps)) for i in range(T)]
[x.start() for x in th]; [x.join() for x in th]
par = time.perf_counter() - t0
print(f"n={n} x {T} batches x {reps}: sequential {T*reps*n/seq/1e6:.2f} M signs/s, concurrent {T*reps*n/par/1e6:.2f} M signs/s  "
      f"(SC_WARPS={os.environ.get('DIL_SC_WARPS','16')} TAIL_WARPS={os.environ.get('DIL_TAIL_WARPS','24')})", flush=True)
