"""Time the NTT kernel launch shapes selectable with DIL_NTT_CFG (one process per shape)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
import dilithium_b200 as d
eng = d.Engine(0); Q = d.Q
n = 1 << 18
x = torch.randint(0, Q, (n, 256), dtype=torch.int32, device="cuda"); o = torch.empty_like(x)
ref = None
def t(fn, it=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
f = t(lambda: eng.ntt(x, out=o)); cs = int(o.to(torch.int64).sum().item())
i = t(lambda: eng.invntt(x, out=o)); cs2 = int(o.to(torch.int64).sum().item())
print("cfg", os.environ.get("DIL_NTT_CFG", "0"), "fwd %%.4f ms %%.3f Gpoly/s frac %%.3f | inv %%.4f ms frac %%.3f | checksums %%d %%d" %% (f, n / f / 1e6, n * 2048 / f / 1e6 / 6482.7, i, n * 2048 / i / 1e6 / 6482.7, cs, cs2))
''' % ROOT
for cfg in sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6", "7"]:
    env = dict(os.environ, DIL_NTT_CFG=cfg)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-400:], flush=True)
