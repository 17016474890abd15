"""Quick per-kernel timing on one GPU (development aid; the judged numbers come from bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilithium_b200 as d

PEAK = 6482.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    eng = d.Engine(0)
    Q = d.Q
    for level, B in ((2, 65536), (3, 65536), (5, 32768)):
        k, l = d.LEVEL_DIMS[level]
        y = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device="cuda")
        out = torch.empty_like(y)
        a_hat = torch.randint(0, Q, (k * l, 256), dtype=torch.int32, device="cuda")
        w = torch.empty((B, k, 256), dtype=torch.int32, device="cuda")
        rho = torch.randint(0, 256, (B, 32), dtype=torch.uint8, device="cuda")
        n = B * l
        rows = []
        def rep(name, ms, bytes_):
            med, best = ms
            rows.append(f"L{level} {name:28s} med {med:8.3f} ms  best {best:8.3f} ms  {bytes_/med/1e6:8.1f} GB/s  frac {bytes_/med/1e6/PEAK:5.3f}")
        rep("ntt (out-of-place)", timeit(lambda: eng.ntt(y, out=out)), n * 2048)
        rep("invntt (out-of-place)", timeit(lambda: eng.invntt(y, out=out)), n * 2048)
        rep("pointwise", timeit(lambda: eng.pointwise_barrett(y, out, y.new_empty(y.shape))), n * 3072)
        rep("matvec shared A", timeit(lambda: eng.matvec(a_hat, y, k, l, w=w)), B * (k + l) * 1024)
        rep("signcore fused", timeit(lambda: eng.signcore(a_hat, y, k, l, w=w)), B * (k + l) * 1024)
        def three():
            eng.ntt(y, out=out); eng.matvec(a_hat, out, k, l, w=w); eng.invntt(w, out=w)
        rep("signcore 3 kernels", timeit(three), B * (k + l) * 1024)
        rep("matvec_expand shared rho", timeit(lambda: eng.matvec_expand(rho[0], y, k, l, False, True, True, w=w)), B * (k + l) * 1024)
        Bp = B // 8
        rep("matvec_expand per-item rho", timeit(lambda: eng.matvec_expand(rho[:Bp], y[:Bp], k, l, True, True, True, w=w[:Bp]), iters=5), Bp * (k + l) * 1024)
        print("\n".join(rows), flush=True)
    # batch sweep for the NTT kernel
    for logn in (10, 14, 16, 18, 20, 21):
        n = 1 << logn
        x = torch.randint(0, Q, (n, 256), dtype=torch.int32, device="cuda")
        o = torch.empty_like(x)
        med, best = timeit(lambda: eng.ntt(x, out=o))
        print(f"ntt n=2^{logn}: med {med:.4f} ms best {best:.4f}  {n/med/1e6:.3f} Gpoly/s  frac {n*2048/med/1e6/PEAK:.3f}", flush=True)
        med, best = timeit(lambda: eng.invntt(x, out=o))
        print(f"intt n=2^{logn}: med {med:.4f} ms best {best:.4f}  {n/med/1e6:.3f} Gpoly/s  frac {n*2048/med/1e6/PEAK:.3f}", flush=True)


if __name__ == "__main__":
    main()
