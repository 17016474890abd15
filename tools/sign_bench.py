"""Quick signing throughput check (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import dilithium_b200 as d
import oracle_lib as ol

# usage: sign_bench.py [levels=2,3,5] [sizes=65536,262144] [key=value tuning ...]   e.g.  sign_bench.py 2 65536 unfused_mask=1
eng = d.Engine(0)
levels = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 3, 5]
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [65536, 262144]
tuning = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[3:]}
for level in levels:
    K = ol.kat(level)
    key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
    key.set_tuning(**tuning)
    for n in sizes:
        mlen = 32
        msgs = torch.randint(0, 256, (n * mlen,), dtype=torch.uint8, device="cuda")
        off = (torch.arange(n + 1, dtype=torch.int64, device="cuda") * mlen)
        z = torch.empty((n, key.z_bytes), dtype=torch.uint8, device="cuda")
        h = torch.empty((n, key.h_bytes), dtype=torch.uint8, device="cuda")
        c = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        att = torch.zeros(n, dtype=torch.int32, device="cuda")
        key.sign_dev(msgs, off, n, z, h, c, att)
        torch.cuda.synchronize()
        l0 = eng.launch_count
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            key.sign_dev(msgs, off, n, z, h, c, att)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        key.set_profile(True)
        key.sign_dev(msgs, off, n, z, h, c, att)
        torch.cuda.synchronize()
        prof = " ".join(f"{k}={v[0]:.3f}" for k, v in key.get_profile().items() if v[0] > 0)
        key.set_profile(False)
        print(f"L{level} n={n} {tuning}: {dt*1e3:.3f} ms  slots={key.last_slots} [{prof}]  {n/dt/1e6:.3f} M signs/s  rounds={key.last_rounds} mean attempts={att.float().mean().item():.3f} "
              f"launches/batch={(eng.launch_count-l0)//reps}  attempts/s={att.sum().item()/dt/1e6:.2f} M", flush=True)
