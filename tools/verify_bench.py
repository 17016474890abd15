#!/usr/bin/env python
"""cfg4-style measurement: Dilithium-5 verification, batch sharded across ranks.
Run alone (1 GPU) or under torchrun.  Two modes per rank, on the rank's shard of B signatures:
  shared key  : one public key for the shard, device-resident inputs (dil_verify_batch_dev), CUDA-event timed
  per-item key: every signature carries its own (rho, t1) (dil_verify_multi_host, host buffers, wall clock)
Prints one JSON line (rank 0) with whole-job verifications/s (max over ranks)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import dilithium_b200 as d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--batch", type=int, default=131072, help="signatures per GPU")
    ap.add_argument("--multi-batch", type=int, default=32768, help="signatures per GPU in per-item-key mode")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, v)) for k, v in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = d.Engine(local)
    level, B = args.level, args.batch
    K = np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz"))
    sk = d.SignKey(eng, level, K["rho"][0], K["k"][0], K["tr"][0], K["s1"][0], K["s2"][0], K["t0"][0])
    vk = d.VerifyKey(eng, level, K["rho"][0], K["t1"][0])
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    msgs = torch.randint(0, 256, (B * 32,), dtype=torch.uint8, generator=gen).to(dev)
    off = torch.arange(B + 1, dtype=torch.int64, device=dev) * 32
    z = torch.empty((B, sk.z_bytes), dtype=torch.uint8, device=dev); h = torch.empty((B, sk.h_bytes), dtype=torch.uint8, device=dev)
    c = torch.empty((B, 32), dtype=torch.uint8, device=dev); att = torch.zeros(B, dtype=torch.int32, device=dev)
    ok = torch.zeros(B, dtype=torch.uint8, device=dev)
    sk.sign_dev(msgs, off, B, z, h, c, att)

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(3):
        vk.verify_dev(msgs, off, B, z, h, c, ok)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        vk.verify_dev(msgs, off, B, z, h, c, ok)
    b.record(); torch.cuda.synchronize()
    ms = maxr(a.elapsed_time(b)) / args.steps
    all_ok = bool(int(ok.sum()) == B)

    # per-item public keys: the shard's signatures verified under per-signature (rho, t1) copies of the key
    Bm = min(args.multi_batch, B)
    rho_n = np.repeat(K["rho"][:1], Bm, axis=0); t1_n = np.repeat(K["t1"][:1], Bm, axis=0)
    msgs_l = msgs[:Bm * 32].cpu().numpy().reshape(Bm, 32)
    msg_list = [m.tobytes() for m in msgs_l]
    zh, hh, ch = z[:Bm].cpu().numpy(), h[:Bm].cpu().numpy(), c[:Bm].cpu().numpy()
    eng.verify_multi(level, rho_n[:256], t1_n[:256], msg_list[:256], zh[:256], hh[:256], ch[:256])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    okm = eng.verify_multi(level, rho_n, t1_n, msg_list, zh, hh, ch)
    dt = maxr(time.perf_counter() - t0)
    # the same with everything resident in HBM
    d_rho, d_t1 = torch.from_numpy(rho_n).to(dev), torch.from_numpy(t1_n).to(dev)
    okd = torch.zeros(Bm, dtype=torch.uint8, device=dev)
    for _ in range(2):
        eng.verify_multi_dev(level, d_rho, d_t1, msgs, off, Bm, z, h, c, okd)
    torch.cuda.synchronize()
    a.record()
    for _ in range(args.steps):
        eng.verify_multi_dev(level, d_rho, d_t1, msgs, off, Bm, z, h, c, okd)
    b.record(); torch.cuda.synchronize()
    ms_multi = maxr(a.elapsed_time(b)) / args.steps
    if rank == 0:
        k, l = d.LEVEL_DIMS[level]
        print(json.dumps({
            "workload": f"cfg4 verify, Dilithium-{level} (k={k},l={l})", "n_gpus": world,
            "shared_key": {"batch_per_gpu": B, "verifies_per_s": world * B / (ms * 1e-3), "ms_per_batch": ms, "all_accepted": all_ok,
                           "timing": "CUDA events, device-resident signatures, max over ranks"},
            "per_item_key": {"batch_per_gpu": Bm, "verifies_per_s": world * Bm / (ms_multi * 1e-3), "ms_per_batch": ms_multi,
                             "all_accepted": bool(int(okd.sum()) == Bm), "keccak_f_per_s": world * Bm * 5 * k * l / (ms_multi * 1e-3),
                             "timing": "CUDA events, dil_verify_multi_dev, keys and signatures resident in HBM, max over ranks"},
            "per_item_key_host": {"batch_per_gpu": Bm, "verifies_per_s": world * Bm / dt, "ms_per_batch": dt * 1e3, "all_accepted": bool(okm.all()),
                                  "timing": "host wall clock around dil_verify_multi_host from pageable numpy buffers"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
