#!/usr/bin/env python
"""Config sweep on one GPU (or one rank of a torchrun job):
per level and batch size, device-timed throughput of every hot-path op and of full signing,
with the achieved fraction of the measured HBM roofline.  Writes one JSON record per line.

  python tools/sweep.py [--out profiles/r1_sweep.jsonl] [--max-log2 20]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import dilithium_b200 as d

Q = d.Q


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--max-log2", type=int, default=20)
    ap.add_argument("--levels", default="2,3,5")
    args = ap.parse_args()
    eng = d.Engine(torch.cuda.current_device())
    PK = peak()
    out = open(args.out, "w")

    def emit(**kw):
        out.write(json.dumps(kw) + "\n"); out.flush()
        print(kw, flush=True)

    kat = {lvl: np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{lvl}.npz")) for lvl in (2, 3, 5)}
    for level in [int(x) for x in args.levels.split(",")]:
        k, l = d.LEVEL_DIMS[level]
        K = kat[level]
        key = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
        for lg in range(10, args.max_log2 + 1, 2):
            B = 1 << lg
            if B * (k + 2 * l) * 1024 > 60e9:
                continue
            iters = 20 if lg <= 16 else 5
            y = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device="cuda")
            o = torch.empty_like(y)
            w = torch.empty((B, k, 256), dtype=torch.int32, device="cuda")
            a_hat = torch.randint(0, Q, (k * l, 256), dtype=torch.int32, device="cuda")
            rho = torch.randint(0, 256, (B, 32), dtype=torch.uint8, device="cuda")
            n = B * l
            rec = dict(level=level, batch=B)
            t = timeit(lambda: eng.ntt(y, out=o), iters); rec.update(ntt_gpolys=n / t / 1e6, ntt_frac=n * 2048 / t / 1e6 / PK)
            t = timeit(lambda: eng.invntt(y, out=o), iters); rec.update(intt_gpolys=n / t / 1e6, intt_frac=n * 2048 / t / 1e6 / PK)
            t = timeit(lambda: eng.pointwise_barrett(y, o, o), iters); rec.update(pointwise_frac=n * 3072 / t / 1e6 / PK)
            t = timeit(lambda: eng.matvec(a_hat, y, k, l, w=w), iters); rec.update(matvec_mitems=B / t / 1e3, matvec_frac=B * (k + l) * 1024 / t / 1e6 / PK)
            t = timeit(lambda: eng.signcore(a_hat, y, k, l, w=w), iters); rec.update(signcore_mitems=B / t / 1e3, signcore_frac=B * (k + l) * 1024 / t / 1e6 / PK)
            t = timeit(lambda: eng.matvec_expand(rho[0], y, k, l, False, True, True, w=w), iters)
            rec.update(cfg3_shared_rho_mitems=B / t / 1e3, cfg3_shared_rho_frac=B * (k + l) * 1024 / t / 1e6 / PK)
            Bp = min(B, 1 << 16)
            t = timeit(lambda: eng.matvec_expand(rho[:Bp], y[:Bp], k, l, True, True, True, w=w[:Bp]), 3)
            rec.update(per_item_rho_batch=Bp, per_item_rho_mitems=Bp / t / 1e3, per_item_rho_keccak_gperms=Bp * 5 * k * l / t / 1e6,
                       per_item_rho_frac=Bp * ((k + l) * 1024 + 32) / t / 1e6 / PK)
            del y, o, w
            # full signing
            msgs = torch.randint(0, 256, (B * 32,), dtype=torch.uint8, device="cuda")
            off = torch.arange(B + 1, dtype=torch.int64, device="cuda") * 32
            z = torch.empty((B, key.z_bytes), dtype=torch.uint8, device="cuda"); h = torch.empty((B, key.h_bytes), dtype=torch.uint8, device="cuda")
            c = torch.empty((B, 32), dtype=torch.uint8, device="cuda"); att = torch.zeros(B, dtype=torch.int32, device="cuda")
            t = timeit(lambda: key.sign_dev(msgs, off, B, z, h, c, att), 3 if lg >= 16 else 10)
            rec.update(sign_msigs=B / t / 1e3, sign_ms=t, sign_rounds=key.last_rounds, mean_attempts=float(att.float().mean().item()))
            # verification of the signatures just produced: one key for the batch, and one key per signature
            vk = d.VerifyKey(eng, level, K["rho"][0], K["t1"][0])
            ok = torch.zeros(B, dtype=torch.uint8, device="cuda")
            t = timeit(lambda: vk.verify_dev(msgs, off, B, z, h, c, ok), 5 if lg >= 16 else 10)
            rec.update(verify_msigs=B / t / 1e3, verify_ms=t, verify_all_ok=bool(int(ok.sum()) == B))
            vk.close()
            emit(**rec)
            del msgs, off, z, h, c, att, ok
            torch.cuda.empty_cache()
        key.close()
        # key generation and per-key verification (host-pointer paths: they include PCIe)
        import time
        for n in (4096, 65536):
            seeds = np.random.default_rng(n).integers(0, 256, size=(n, 32)).astype(np.uint8)
            eng.keygen(level, seeds[:256])
            t0 = time.perf_counter(); keys = eng.keygen(level, seeds); dt = time.perf_counter() - t0
            emit(level=level, keygen_batch=n, keygen_mkeys_e2e=n / dt / 1e6, keygen_ms=dt * 1e3)
        n = 4096
        sk = d.SignKey(eng, level, *[keys[f][0] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
        msgs_l = [bytes([i & 255]) * 32 for i in range(n)]
        z, h, c, _ = sk.sign(msgs_l)
        rho_n = np.repeat(keys["rho"][:1], n, axis=0); t1_n = np.repeat(keys["t1"][:1], n, axis=0)
        eng.verify_multi(level, rho_n[:64], t1_n[:64], msgs_l[:64], z[:64], h[:64], c[:64])
        t0 = time.perf_counter(); ok = eng.verify_multi(level, rho_n, t1_n, msgs_l, z, h, c); dt = time.perf_counter() - t0
        emit(level=level, verify_multi_batch=n, verify_multi_msigs_e2e=n / dt / 1e6, verify_multi_ms=dt * 1e3, all_ok=bool(ok.sum() == n))
        sk.close()


if __name__ == "__main__":
    main()
