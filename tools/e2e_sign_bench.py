"""End-to-end signing through dil_sign_batch_host with pinned host buffers (development aid).
Knobs are read from the environment by the library: DIL_SIGN_DRAIN=0 (chunked copy path),
DIL_DRAIN_CTAS, DIL_SIGN_CHUNK.  Also prints the raw pinned D2H copy bandwidth for comparison.
  python tools/e2e_sign_bench.py [level] [n] [steps]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dilithium_b200 as d
import oracle_lib as ol

level = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
eng = d.Engine(0)
K = ol.kat(level)
key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
mlen = 32
msgs = torch.randint(0, 256, (n * mlen,), dtype=torch.uint8).pin_memory()
offs = (torch.arange(n + 1, dtype=torch.int64) * mlen).pin_memory()
z = torch.empty((n, key.z_bytes), dtype=torch.uint8).pin_memory()
h = torch.empty((n, key.h_bytes), dtype=torch.uint8).pin_memory()
c = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
a = torch.zeros(n, dtype=torch.int32).pin_memory()
P = ctypes.c_void_p


def step():
    rc = eng._lib.dil_sign_batch_host(eng._h, key._h, P(msgs.data_ptr()), P(offs.data_ptr()), n, P(z.data_ptr()), P(h.data_ptr()),
                                      P(c.data_ptr()), P(a.data_ptr()))
    assert rc == 0


step()
t0 = time.perf_counter()
for _ in range(steps):
    step()
dt = (time.perf_counter() - t0) / steps
sig = key.z_bytes + key.h_bytes + 36
# device-resident for reference
dm, do = msgs.cuda(), offs.cuda()
dz, dh, dc, da = (torch.empty_like(t, device="cuda") for t in (z, h, c, a))
key.sign_dev(dm, do, n, dz, dh, dc, da)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    key.sign_dev(dm, do, n, dz, dh, dc, da)
torch.cuda.synchronize()
ddt = (time.perf_counter() - t0) / steps
ok = bool(torch.equal(dz.cpu(), z) and torch.equal(dh.cpu(), h) and torch.equal(dc.cpu(), c) and torch.equal(da.cpu(), a))
# raw D2H bandwidth of the copy engine into the same pinned buffer
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    z.copy_(dz, non_blocking=True)
torch.cuda.synchronize()
bw = 5 * z.numel() / (time.perf_counter() - t0) / 1e9
print(f"L{level} n={n} env={ {k: v for k, v in os.environ.items() if k.startswith('DIL_')} }: e2e {dt*1e3:.3f} ms = {n/dt/1e6:.3f} M signs/s "
      f"({n*sig/dt/1e9:.1f} GB/s of signatures) | device-resident {ddt*1e3:.3f} ms = {n/ddt/1e6:.3f} M signs/s | match={ok} | "
      f"copy-engine D2H {bw:.1f} GB/s", flush=True)
