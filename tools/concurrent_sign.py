"""Experiment: T independent sign batches in flight on T streams (one key handle, one host thread each).
usage: concurrent_sign.py <n> <T> [spec_target [spec_max [level]]]
The kernels of the late, small rejection rounds of one batch leave SMs idle (Keccak wave quantisation, latency-bound
rounds); another batch's kernels fill them.  With several batches in flight the straggler speculation can be
reduced (fewer wasted slots), which the spec_target / spec_max arguments explore."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dilithium_b200 as d
import oracle_lib as ol

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec_target = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spec_max = int(sys.argv[4]) if len(sys.argv) > 4 else 0
level = int(sys.argv[5]) if len(sys.argv) > 5 else 2
eng = d.Engine(0)
K = ol.kat(level)
ctx = []
for t in range(T):
    key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
    if spec_target or spec_max:
        key.set_tuning(spec_target=spec_target, spec_max=spec_max)
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
reps = 6
t0 = time.perf_counter()
for i in range(T): run(i, reps)
seq = time.perf_counter() - t0
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(i, reps)) for i in range(T)]
[x.start() for x in th]; [x.join() for x in th]
par = time.perf_counter() - t0
key = ctx[0][0]
print(f"L{level} n={n} x {T} in flight x {reps}, spec_target={spec_target or 'default'} spec_max={spec_max or 'default'}: one at a time "
      f"{T*reps*n/seq/1e6:.2f} M signs/s, concurrent {T*reps*n/par/1e6:.2f} M signs/s  (rounds {key.last_rounds}, slots {key.last_slots})", flush=True)
