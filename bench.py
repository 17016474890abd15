#!/usr/bin/env python
"""bench.py - headline benchmark: batched Dilithium-2 signing on B200 (BASELINE.json metric
"Dilithium-2 signs/sec at 1/2/4/8 B200; NTT polys/sec vs HBM roofline", configs[1]: Dilithium-2
(k=4,l=4), 64K-signature batch per GPU).

One step = sign one batch of 65 536 independent 32-byte messages under one key, start to finish
(mu / rho' hashing, and per rejection round: ExpandMask -> NTT(y) -> A*y -> INTT(w) -> Decompose ->
challenge hash + SampleInBall -> c*s1, c*s2, c*t0, norm checks, MakeHint; deterministic round-3.1
signing exactly as the reference's KAT vectors, rtl_src/combined_top.v mode 2).  The polynomial
arithmetic inside is the engine's hot path; the stand-alone NTT kernel and the fused cfg2 sign core
(NTT+matvec+INTT) are timed separately in the same run and reported as `roofline_ntt` /
`sign_core`.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this framework (CUDA engine)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # reference C++ arithmetic on host cores

Multi-GPU: one process per GPU (torchrun), every rank signs its own 65 536-message shard (weak
scaling); the only collective is ONE NCCL broadcast of the key material from rank 0.

`value`   signs/s over all GPUs, messages already resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same through dil_sign_batch_host: pinned host messages in, signatures back on the host
          (finished signatures are drained to the host round by round while later rounds still sign).
`roofline` the kernel class with the largest share of the step's device time (measured with CUDA
          events around every launch in a separate profiled step): HBM fraction from algorithmic bytes as
          the contract asks, plus `compute_roofline` (Keccak-f/s against the measured pure-Keccak peak)
          because that class is bound by the integer ALU pipe, not by HBM.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

Q = 8380417
SEED = 0x44494C32
MSG_BYTES = 32
LEVEL_DIMS = {2: (4, 4), 3: (6, 5), 5: (8, 7)}
LEVEL_EXTRA = {2: dict(w1=192, zb=576, hb=84), 3: dict(w1=128, zb=640, hb=61), 5: dict(w1=128, zb=640, hb=83)}
UNIT = "signs/s"


def metric_name(level):
    return f"Dilithium-{level} signs/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--level", type=int, default=2, choices=[2, 3, 5])
    ap.add_argument("--batch", type=int, default=65536, help="signatures per GPU per step (weak scaling)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(level, batch):
    k, l = LEVEL_DIMS[level]
    return (f"full deterministic sign, Dilithium-{level} (k={k},l={l}), {batch}-signature batch per GPU, one key, "
            f"{MSG_BYTES}-byte messages")


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


KECCAK_PEAK_GPS = 4.27   # G Keccak-f[1600]/s: pure-permutation micro-benchmark on one B200 (tools/keccak_pipe_bench.cu,
                         # profiles/r1b_keccak_pipe_bench.txt): the ALU-pipe speed of light for the hash kernels


def class_kernel_name(kernel_class, level):
    if kernel_class == "tail" and level == 3:   # eta = 4: the sparse products do not apply
        return "sign_tail_kernel (NTT(c), c*s2 / c*s1 / c*t0 through transforms, norm checks, MakeHint, resolve)"
    return CLASS_KERNEL[kernel_class]


def keccak_roofline(kernel_class, level, slots, ms):
    """For the Keccak-bound classes: achieved permutations/s against the measured pure-Keccak peak."""
    k, l = LEVEL_DIMS[level]
    w1 = LEVEL_EXTRA[level]["w1"]
    per_slot = {"expand_mask": l * 5,                                   # 576/640 squeezed bytes per polynomial = 5 blocks
                "challenge": (64 + k * w1) // 136 + 1 + 1}.get(kernel_class)   # absorb mu||w1 + pad, then SampleInBall
    if per_slot is None or ms <= 0:
        return None
    gps = per_slot * slots / (ms * 1e-3) / 1e9
    return {"bound": "integer ALU (Keccak-f[1600])", "unit": "G Keccak-f/s", "achieved": gps, "peak": KECCAK_PEAK_GPS,
            "frac": gps / KECCAK_PEAK_GPS, "permutations_per_slot": per_slot,
            "peak_source": "tools/keccak_pipe_bench.cu on B200 (profiles/r1b_keccak_pipe_bench.txt)"}


def ncu_traffic(kernel_class):
    """dram bytes per launch of `kernel_class` from the committed ncu capture, or None."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel.json")))
        return prof.get(kernel_class, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def kat_key(level, index=0):
    """Signing key = KAT vector `index` of the reference (committed fixture tests/golden/kat_L*.npz)."""
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz"))
    return {f: np.ascontiguousarray(d[f][index]) for f in ("rho", "k", "tr", "s1", "s2", "t0", "t1")}


def class_bytes(level):
    """Algorithmic HBM bytes per slot (one signing attempt) for each kernel class."""
    k, l = LEVEL_DIMS[level]
    x = LEVEL_EXTRA[level]
    return {
        "expand_mask": 66 + l * 1024,
        "signcore": (l + k) * 1024 + k * x["w1"],   # y in, w and the packed w1 = HighBits(w) out
        "challenge": 64 + k * x["w1"] + 256 + 32,
        "tail": 256 + 2 * l * 1024 + k * 1024 + x["hb"] + 1,
        "resolve": l * 1024 + l * x["zb"] + 2 * x["hb"] + 64 + 4,
    }


CLASS_KERNEL = {"expand_mask": "expand_mask_kernel<L,GAMMA1_BITS> (SHAKE-256 ExpandMask, one Keccak state per thread)",
                "signcore": "matvec_shared_kernel<K,L,16,0,1,1,1> (fused NTT -> A*y -> INTT -> w, packed HighBits(w))",
                "challenge": "challenge_kernel (SHAKE-256 + SampleInBall)",
                "tail": "sign_tail_sparse_kernel (sparse c*s2 / c*s1 products + norm checks; NTT(c), c*t0, MakeHint and resolve for survivors)", "resolve": "resolve_kernel",
                "init": "sign_init_kernel"}
CLASS_BOUND_NOTE = {"expand_mask": "integer ALU (Keccak): ncu alu pipe 95% busy; HBM fraction is not the limiter",
                    "challenge": "integer ALU (Keccak)", "tail": "issue slots / ALU pipe 60 % + global-load latency (sparse products); fmaheavy 39 % (survivors' c*t0 transforms)",
                    "signcore": "integer multiply pipe (fmaheavy): ncu 73% busy", "resolve": "HBM/latency"}


# --------------------------------------------------------------------------------------
# CPU arm: reference arithmetic (oracle/_ref) + scheme glue, or the pure oracle port
# --------------------------------------------------------------------------------------
def cpu_sign(level, n_msgs, threads, steps=1, warmup=0):
    import numpy as np
    import oracle_lib as ol
    K = {f: v[None, ...] for f, v in kat_key(level).items()}
    rng = np.random.default_rng(SEED)
    msgs = [bytes(rng.integers(0, 256, MSG_BYTES).astype(np.uint8)) for _ in range(n_msgs)]
    ref = ol.load_ref()
    impl, kind = (ref, "reference") if ref is not None else (ol.load(), "port")
    for _ in range(warmup):
        impl.sign_batch(level, K, 0, msgs, threads)
    dt = 0.0
    att = None
    for _ in range(steps):
        *_, att, sec = impl.sign_batch(level, K, 0, msgs, threads)
        dt += sec
    dt /= steps
    sample = (f"{n_msgs} messages x {steps} step(s), one key, split over {threads} thread(s); "
              + ("every NTT / inverse NTT / pointwise product runs the reference's own compiled ref_ntt.cpp functions "
                 "(oracle/_ref, ref_bridge.cpp); hashing, sampling, packing and the rejection loop are the oracle's C "
                 "restatement (the reference has no C/C++ sign, SURVEY.md 0.1)" if kind == "reference"
                 else "oracle port (oracle/*.c) for everything"))
    return n_msgs / dt, kind, threads, sample, dt * 1e3, float(att.mean())


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    n_msgs = 256 * cores if 256 * cores < 8192 else 8192
    steps = max(1, min(args.steps, 5))
    warm = max(0, min(args.warmup, 1))
    val, kind, threads, sample, ms, att = cpu_sign(args.level, n_msgs, cores, steps, warm)
    line = {
        "impl": "reference", "metric": metric_name(args.level), "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/int64 (signed % reduction)", "data": "synthetic",
        "config": {"workload": workload_name(args.level, args.batch), "level": args.level, "sample_messages_per_step": n_msgs,
                   "mean_attempts": att, "note": "CPU arm: host cores only, no GPU; throughput does not depend on --gpus"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max, "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------
# engine arm
# --------------------------------------------------------------------------------------
def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dilithium_b200 as d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = d.Engine(local)
    level = args.level
    k, l = LEVEL_DIMS[level]
    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one key for the whole job: rank 0 owns it, ONE broadcast ships it (rho | K | tr | s1 | s2 | t0)
    fields = ("rho", "k", "tr", "s1", "s2", "t0")
    sizes = {f: v.size for f, v in kat_key(level).items()}
    blob = torch.zeros(sum(sizes[f] for f in fields), dtype=torch.uint8, device=dev)
    if rank == 0:
        kk = kat_key(level)
        blob.copy_(torch.from_numpy(np.concatenate([kk[f] for f in fields])))
    if world > 1:
        dist.broadcast(blob, src=0)
    parts, off = {}, 0
    hb = blob.cpu().numpy()
    for f in fields:
        parts[f] = hb[off:off + sizes[f]].copy()
        off += sizes[f]
    key = d.SignKey(eng, level, *[parts[f] for f in fields])

    # per-rank synthetic messages (the rank's shard of the global batch)
    gen = torch.Generator(device="cpu").manual_seed(SEED + 1 + rank)
    msgs_host = torch.randint(0, 256, (B * MSG_BYTES,), dtype=torch.uint8, generator=gen).pin_memory()
    offs_host = (torch.arange(B + 1, dtype=torch.int64) * MSG_BYTES).pin_memory()
    msgs, offs = msgs_host.to(dev), offs_host.to(dev)
    z = torch.empty((B, key.z_bytes), dtype=torch.uint8, device=dev)
    h = torch.empty((B, key.h_bytes), dtype=torch.uint8, device=dev)
    c = torch.empty((B, 32), dtype=torch.uint8, device=dev)
    att = torch.zeros(B, dtype=torch.int32, device=dev)

    def step():
        key.sign_dev(msgs, offs, B, z, h, c, att)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    launches = eng.launch_count - l0
    ms_total = maxr(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    mean_attempts = float(att.float().mean().item())
    rounds = key.last_rounds

    # per-kernel-class device time of one step (CUDA events around every launch), outside the timed region
    key.set_profile(True)
    step()
    torch.cuda.synchronize()
    prof = {n: v for n, v in key.get_profile().items() if n != "pack_w1"}   # w1 packing is fused into the sign core
    key.set_profile(False)
    prof_total = sum(ms for ms, _ in prof.values())
    dominant = max((n for n in prof if n != "init"), key=lambda n: prof[n][0])
    cb = class_bytes(level)

    # stand-alone hot-path kernels, separately timed (north_star: "NTT polys/sec vs HBM roofline")
    def time_kernel(fn, iters=50):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    yb = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device=dev)
    outb = torch.empty_like(yb)
    wb = torch.empty((B, k, 256), dtype=torch.int32, device=dev)
    a_hat = eng.expand_a(torch.from_numpy(parts["rho"]).to(dev), k, l)[0].contiguous()
    npoly = B * l
    ntt_ms = time_kernel(lambda: eng.ntt(yb, out=outb))
    intt_ms = time_kernel(lambda: eng.invntt(yb, out=outb))
    core_ms = time_kernel(lambda: eng.signcore(a_hat, yb, k, l, w=wb))
    del yb, outb, wb

    # end to end through the host-pointer C ABI: pinned host messages in, signatures out to host memory
    z_h = torch.empty((B, key.z_bytes), dtype=torch.uint8).pin_memory()
    h_h = torch.empty((B, key.h_bytes), dtype=torch.uint8).pin_memory()
    c_h = torch.empty((B, 32), dtype=torch.uint8).pin_memory()
    a_h = torch.zeros(B, dtype=torch.int32).pin_memory()
    P = ctypes.c_void_p
    lib = eng._lib

    def e2e_step():
        rc = lib.dil_sign_batch_host(eng._h, key._h, P(msgs_host.data_ptr()), P(offs_host.data_ptr()), B, P(z_h.data_ptr()),
                                     P(h_h.data_ptr()), P(c_h.data_ptr()), P(a_h.data_ptr()))
        if rc != 0:
            raise RuntimeError("dil_sign_batch_host failed")

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_val = world * B * args.e2e_steps / e2e_s
    e2e_ok = bool(torch.equal(z_h.to(dev), z) and torch.equal(h_h.to(dev), h) and torch.equal(c_h.to(dev), c))

    if rank == 0:
        peak, peak_src = measured_peak()
        dom_ms, dom_units = prof[dominant]
        dom_bytes = cb[dominant] * dom_units
        dom_achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        sig_bytes = key.z_bytes + key.h_bytes + 32
        line = {
            "metric": metric_name(level), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32/u64 integer (Shoup/Barrett modular arithmetic, 64-bit Keccak lanes)", "data": "synthetic",
            "config": {"workload": workload_name(level, B), "level": level, "k": k, "l": l, "batch_per_gpu": B,
                       "key": "reference KAT vector 0 (tests/golden)", "mean_attempts": mean_attempts, "rejection_rounds": rounds,
                       "sharding": ("independent messages, one contiguous shard per rank; one NCCL broadcast of the key material"
                                    if world > 1 else "single GPU"),
                       "l2": "no flush: every round streams > 1 GiB of per-attempt state (y, w, c), far above the 126 MB L2",
                       "timing": "CUDA events on torch's current stream (the stream every kernel is launched on), max over ranks",
                       "kernels_per_round": "ExpandMask, sign core (+ packed HighBits), challenge, tail (+ resolve when one slot per item), "
                                            "resolve only in speculative rounds"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "step_profile_ms": {n: round(ms, 4) for n, (ms, _) in prof.items()},
            "step_profile_note": f"device time per kernel class in one step (sum {prof_total:.3f} ms of {ms_step:.3f} ms wall); "
                                 "the rest is launch latency and one 4-byte D2H + stream sync per rejection round",
            "roofline": {"kernel": class_kernel_name(dominant, level), "class": dominant, "bound": "hbm", "achieved": dom_achieved, "peak": peak,
                         "unit": "GB/s", "frac": dom_achieved / peak,
                         "traffic": (ncu_traffic(dominant) / 65536.0 * dom_units / max(rounds, 1)) if ncu_traffic(dominant) else None,
                         "traffic_note": "ncu dram bytes of a 65536-slot launch scaled to this step's average launch size",
                         "algorithmic_bytes_per_slot": cb[dominant], "slots_per_step": dom_units, "launches_per_step": rounds,
                         "avg_launch_ms": dom_ms / max(rounds, 1), "peak_source": peak_src,
                         "true_limiter": CLASS_BOUND_NOTE.get(dominant, ""),
                         "compute_roofline": keccak_roofline(dominant, level, dom_units, dom_ms)},
            "roofline_ntt": {"kernel": "ntt_tma_kernel<32,3,1,false> (stand-alone forward NTT, the north_star's named kernel)",
                             "bound": "hbm", "polys_per_s": npoly / (ntt_ms * 1e-3), "achieved": npoly * 2048 / (ntt_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": npoly * 2048 / (ntt_ms * 1e-3) / 1e9 / peak, "polys_per_launch": npoly,
                             "invntt_polys_per_s": npoly / (intt_ms * 1e-3), "invntt_frac": npoly * 2048 / (intt_ms * 1e-3) / 1e9 / peak},
            "sign_core": {"workload": f"cfg2 core alone: w = INTT(A_hat * NTT(y)), {B} items, one fused kernel",
                          "items_per_s": B / (core_ms * 1e-3), "ms": core_ms,
                          "hbm_frac": B * (k + l) * 1024 / (core_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * MSG_BYTES + (B + 1) * 8,
                    "d2h_bytes_per_step": B * (sig_bytes + 4), "steps": args.e2e_steps,
                    "api": "dil_sign_batch_host (C ABI): pinned host messages in, z/h/c~/attempts back in host memory",
                    "timing": "host wall clock around the synchronous calls, max over ranks", "matches_device_path": e2e_ok},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = 256 * cores if 256 * cores < 8192 else 8192
            v, kind, threads, sample, _, _ = cpu_sign(level, n_cpu, cores, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
