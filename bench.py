#!/usr/bin/env python
"""bench.py - headline benchmark: batched Dilithium-2 signing on B200 (BASELINE.json metric
"Dilithium-2 signs/sec at 1/2/4/8 B200; NTT polys/sec vs HBM roofline", configs[1]: Dilithium-2
(k=4,l=4), 64K-signature batch per GPU).

One step = sign one batch of 65 536 independent 32-byte messages under one key, start to finish
(mu / rho' hashing, and per rejection round: ExpandMask -> NTT(y) -> A*y -> INTT(w) -> Decompose ->
challenge hash + SampleInBall -> c*s1, c*s2, c*t0, norm checks, MakeHint; deterministic round-3.1
signing exactly as the reference's KAT vectors, rtl_src/combined_top.v mode 2).  The polynomial
arithmetic inside is the engine's hot path; the stand-alone NTT kernel and the fused cfg2 sign core
(NTT+matvec+INTT) are timed separately in the same run and reported as `roofline_ntt` / `sign_core`.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this framework (CUDA engine)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # reference C++ arithmetic on host cores

Multi-GPU: one process per GPU (torchrun), every rank signs its own 65 536-message shard (weak
scaling); the only collective is ONE NCCL broadcast of the key material from rank 0.

`value`    signs/s over all GPUs for K steps, every step one 65 536-message batch per GPU with its messages resident in HBM,
           T = 4 batches in flight per GPU (one key handle and stream per batch in flight, driven by ONE host thread through
           dil_sign_batch_dev_begin / dil_sign_batch_finish; all K steps start and finish between the two CUDA events), max
           over ranks.  `one_batch_at_a_time` is the same K steps strictly serial (--in-flight 1 makes it the headline).
`e2e`      the same through dil_sign_batch_host[_begin]: pinned host messages in, signatures back on the host
           (finished signatures are drained to the host round by round while later rounds still sign), T batches in flight,
           `e2e.one_call_at_a_time` next to it; `e2e.host_ceiling` relates it to the measured host-memory ceiling of the box.
`roofline` the kernel class with the largest share of the step's device time (CUDA events around every
           launch in a separate profiled step): HBM fraction from algorithmic bytes as the contract asks,
           plus `compute_roofline`: Keccak-f/s against the pure-Keccak rate MEASURED IN THIS RUN
           (dil_diag_keccak_dev) because that class is bound by the integer ALU pipe, not by HBM.
`configs`  key generation (row N4) and the other BASELINE configurations under the same clock: cfg3 (Dilithium-3, 262 144 items: fused
           ExpandA core with shared and per-item rho, full signing), cfg4 (Dilithium-5 verification, one GPU's
           131 072-signature shard, a key per signature, 100 KAT tuples injected and asserted), cfg5 (batch sweep
           2^10 .. 2^22 x levels 2/3/5, sharded over the ranks).  --no-configs skips them.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
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
# measured with tools/d2h_ceiling.cu on the 8-GPU box (profiles/r2_d2h_ceiling_8gpu.txt): bytes/s ALL GPUs together
# can write into host memory, barrier-synchronised, copy engine and SM stores alike
HOST_CEILING_GBS = {1: 54.4, 2: 72.2, 4: 74.8, 8: 96.6}


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
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--in-flight", type=int, default=4,
                    help="sign batches in flight per GPU in the timed loops (one key handle and stream each); 1 = strictly one at a time")
    ap.add_argument("--driver", default="async", choices=["async", "threads"],
                    help="how the batches in flight are driven: one host thread with dil_sign_batch_*_begin / dil_sign_batch_finish (default) "
                         "or one host thread per batch in flight with the synchronous calls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 sub-records")
    ap.add_argument("--sweep-max-log2", type=int, default=22)
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


def class_kernel_name(kernel_class, level):
    if kernel_class == "tail" and level == 3:   # eta = 4: the sparse products do not apply
        return "sign_tail_kernel (NTT(c), c*s2 / c*s1 / c*t0 through transforms, norm checks, MakeHint, resolve)"
    return CLASS_KERNEL[kernel_class]


def keccak_perms_per_slot(kernel_class, level):
    k, l = LEVEL_DIMS[level]
    w1 = LEVEL_EXTRA[level]["w1"]
    return {"expand_mask": l * 5,                                   # 576/640 squeezed bytes per polynomial = 5 blocks
            "challenge": (64 + k * w1) // 136 + 1 + 1}.get(kernel_class)   # absorb mu||w1 + pad, then SampleInBall


def keccak_roofline(kernel_class, level, slots, ms, peak_gps=None, peak_source=None):
    """For the Keccak-bound classes: achieved permutations/s against the pure-Keccak rate measured in this run."""
    per_slot = keccak_perms_per_slot(kernel_class, level)
    if per_slot is None or ms <= 0:
        return None
    gps = per_slot * slots / (ms * 1e-3) / 1e9
    out = {"bound": "integer ALU (Keccak-f[1600])", "unit": "G Keccak-f/s", "achieved": gps, "permutations_per_slot": per_slot}
    if peak_gps:
        out.update(peak=peak_gps, frac=gps / peak_gps, peak_source=peak_source)
    return out


def ncu_record(kernel_class):
    """The committed ncu capture of `kernel_class` (profiles/dominant_kernel.json, written by tools/make_dominant.py from
    the round's `ncu --set full` reports): dram bytes per launch and the pipe utilisations that name its limiter."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel.json"))).get(kernel_class)
    except Exception:
        return None


def ncu_traffic(kernel_class):
    r = ncu_record(kernel_class)
    return r.get("dram_bytes_per_launch") if r else None


def ncu_limiter(kernel_class):
    """Which pipe the committed ncu capture shows busiest for this class (numbers, not prose)."""
    r = ncu_record(kernel_class)
    if not r:
        return None
    pipes = {n: r[f] for n, f in (("alu", "alu_pipe_pct"), ("fmaheavy", "fmaheavy_pipe_pct"), ("issue", "issue_pct")) if f in r}
    if "dram_pct" in r:
        pipes["dram"] = r["dram_pct"]
    top = max(pipes, key=pipes.get)
    return {"top_pipe": top, "pct_busy": pipes, "source": r.get("source")}


def keccak_sass_mix():
    """Instruction mix of one Keccak-f round in the SHIPPED library (cuobjdump -sass of keccak_rate_kernel's inner loop):
    how many of its instructions issue on the integer ALU pipe (LOP3 / SHF / IADD3 ...)."""
    lib = os.path.join(ROOT, "dilithium_b200", "libdilithium_b200.so")
    try:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3dil18keccak_rate_kernelEPmj", lib], capture_output=True, text=True,
                             timeout=120).stdout
    except Exception:
        return None
    ins = re.findall(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)\s*([^;]*);", out)
    if not ins:
        return None
    addr = [int(a, 16) for a, _, _ in ins]
    loops = []
    for a, op, args in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if m and int(m.group(1), 16) < int(a, 16):
                loops.append((int(a, 16) - int(m.group(1), 16), int(m.group(1), 16), int(a, 16)))
    if not loops:
        return None
    _, lo, hi = min(loops)
    body = [op for a, (_, op, _) in zip(addr, ins) if lo <= a <= hi]
    alu = sum(1 for op in body if op.split(".")[0] in ("LOP3", "SHF", "IADD3", "VIADD", "PRMT", "SEL", "ISETP", "IMNMX", "MOV", "LEA", "VIMNMX"))
    fma = sum(1 for op in body if op.split(".")[0] in ("IMAD", "FFMA", "FMUL", "FADD"))
    return {"instructions_per_round": len(body), "alu_pipe": alu, "fma_pipe": fma, "other": len(body) - alu - fma}


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


def cpu_ntt(threads, n_polys=65536):
    """reference ntt() (compiled ref_ntt.cpp when present, else the oracle port) over n_polys polynomials on `threads` threads."""
    import numpy as np
    import oracle_lib as ol
    rng = np.random.default_rng(SEED + 5)
    a = rng.integers(0, Q, size=(n_polys, 256)).astype(np.int32)
    ref = ol.load_ref()
    t0 = time.perf_counter()
    if ref is not None:
        ref.run("ref_ntt_batch", a, threads, canon=False)
    else:
        ol.load().ntt(a, threads=threads)
    return n_polys / (time.perf_counter() - t0), "reference" if ref is not None else "port"


def host_threads():
    """Threads the CPU arm uses: the cores this process may run on."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_threads()
    n_msgs = 256 * cores if 256 * cores < 8192 else 8192
    # every step signs a bounded sample (about 0.2 s on 16 cores), so the driver's own --steps / --warmup are honoured as given
    steps = max(1, min(args.steps, 200))
    warm = max(0, min(args.warmup, 20))
    val, kind, threads, sample, ms, att = cpu_sign(args.level, n_msgs, cores, steps, warm)
    line = {
        "impl": "reference", "metric": metric_name(args.level), "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/int64 (signed % reduction)", "data": "synthetic",
        "config": {"workload": workload_name(args.level, args.batch), "level": args.level, "sample_messages_per_step": n_msgs,
                   "mean_attempts": att, "note": "CPU arm: host cores only, no GPU; throughput does not depend on --gpus; every step signs "
                                                 "a bounded sample of the workload (sample_messages_per_step)"},
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
class Ctx:
    """Per-rank state shared by the headline step and the configuration sub-records."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import dilithium_b200 as d
        self.torch, self.dist, self.d = torch, dist, d
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl engine needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = d.Engine(self.local)
        self.peak, self.peak_src = measured_peak()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_ms(self, fn, iters, warm=2):
        """Device time per call (CUDA events on torch's current stream = the stream the engine launches on), max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return self.maxr(a.elapsed_time(b)) / iters


def sign_buffers(c, key, n, gen_seed):
    torch = c.torch
    gen = torch.Generator(device="cpu").manual_seed(gen_seed)
    msgs_host = torch.randint(0, 256, (n * MSG_BYTES,), dtype=torch.uint8, generator=gen).pin_memory()
    offs_host = (torch.arange(n + 1, dtype=torch.int64) * MSG_BYTES).pin_memory()
    z = torch.empty((n, key.z_bytes), dtype=torch.uint8, device=c.dev)
    h = torch.empty((n, key.h_bytes), dtype=torch.uint8, device=c.dev)
    ct = torch.empty((n, 32), dtype=torch.uint8, device=c.dev)
    att = torch.zeros(n, dtype=torch.int32, device=c.dev)
    return msgs_host, offs_host, msgs_host.to(c.dev), offs_host.to(c.dev), z, h, ct, att


def cfg3_record(c, keccak_peak):
    """cfg3: Dilithium-3 (k=6,l=5), 262 144 items per GPU: fused ExpandA -> NTT -> A*y -> INTT core (mode S: one rho,
    mode P: rho per item) and full signing of a 262 144-message batch."""
    torch, d, eng = c.torch, c.d, c.eng
    level, B = 3, 262144
    k, l = LEVEL_DIMS[level]
    kk = kat_key(level)
    gen = torch.Generator(device=c.dev).manual_seed(SEED + 3)
    y = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device=c.dev, generator=gen)
    w = torch.empty((B, k, 256), dtype=torch.int32, device=c.dev)
    rho = torch.from_numpy(kk["rho"]).to(c.dev)
    rho_p = torch.randint(0, 256, (B, 32), dtype=torch.uint8, device=c.dev, generator=gen)
    ms_s = c.time_ms(lambda: eng.matvec_expand(rho, y, k, l, ntt_input=True, intt_output=True, w=w), 10)
    ms_p = c.time_ms(lambda: eng.matvec_expand(rho_p, y, k, l, per_item=True, ntt_input=True, intt_output=True, w=w), 3, warm=1)
    del y, w
    torch.cuda.empty_cache()
    key = d.SignKey(eng, level, *[kk[f] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
    _, _, msgs, offs, z, h, ct, att = sign_buffers(c, key, B, SEED + 31 + c.rank)
    ms_sign = c.time_ms(lambda: key.sign_dev(msgs, offs, B, z, h, ct, att), 3, warm=1)
    mean_att = float(att.float().mean().item())
    key.close()
    bytes_item = (l + k) * 1024 + 32
    out = {"workload": f"Dilithium-3 (k=6,l=5), {B} items per GPU", "n_gpus": c.world,
           "fused_expand_core_shared_rho": {"items_per_s": c.world * B / (ms_s * 1e-3), "ms": ms_s,
                                            "hbm_frac": B * bytes_item / (ms_s * 1e-3) / 1e9 / c.peak, "algorithmic_bytes_per_item": bytes_item,
                                            "keccak_f_per_launch": "5 x k*l per CTA (A is expanded once per persistent CTA, never written to HBM)"},
           "fused_expand_core_per_item_rho": {"items_per_s": c.world * B / (ms_p * 1e-3), "ms": ms_p,
                                              "hbm_frac": B * bytes_item / (ms_p * 1e-3) / 1e9 / c.peak,
                                              "keccak_f_per_s": B * 5 * k * l / (ms_p * 1e-3),
                                              "keccak_frac_of_measured_peak": B * 5 * k * l / (ms_p * 1e-3) / keccak_peak if keccak_peak else None,
                                              "bound": "integer ALU (Keccak: 5 x k*l = 150 permutations per item)"},
           "full_sign": {"signs_per_s": c.world * B / (ms_sign * 1e-3), "ms": ms_sign, "mean_attempts": mean_att}}
    return out


def cfg4_record(c, keccak_peak):
    """cfg4: Dilithium-5 verification, 1 M signatures over 8 GPUs = 131 072 per GPU, EVERY signature under its own public
    key (per-item rho: A expanded on chip per item); the 100 level-5 KAT tuples are injected at fixed positions of every
    GPU's shard and must be accepted (they yield the w that makes H(mu || w1') == c~)."""
    import numpy as np
    torch, d, eng = c.torch, c.d, c.eng
    level, n = 5, 131072
    k, l = LEVEL_DIMS[level]
    K = np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz"))
    M = np.load(os.path.join(ROOT, "tests", "golden", "kat_msgs.npz"))
    kat_off = np.concatenate([[0], np.cumsum(M["mlen"])])
    kat_msgs = [M["blob"][kat_off[i]:kat_off[i + 1]] for i in range(100)]
    key_of = np.arange(n) % 100
    pos = 1310 * np.arange(100) + 7                               # fixed KAT positions in every shard
    rng = np.random.default_rng(SEED + 4 + c.rank)
    mlen = np.full(n, MSG_BYTES, dtype=np.int64)
    mlen[pos] = M["mlen"]
    off = np.concatenate([[0], np.cumsum(mlen)]).astype(np.uint64)
    blob = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
    for j, p in enumerate(pos):
        blob[int(off[p]):int(off[p + 1])] = kat_msgs[j]
        key_of[p] = j
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(c.dev)   # noqa: E731
    d_rho, d_t1 = dv(K["rho"][key_of]), dv(K["t1"][key_of])
    d_msgs, d_off = dv(blob), dv(off.view(np.int64))
    zb, hb = l * 640, 75 + k
    z = torch.empty((n, zb), dtype=torch.uint8, device=c.dev); h = torch.empty((n, hb), dtype=torch.uint8, device=c.dev)
    ct = torch.empty((n, 32), dtype=torch.uint8, device=c.dev); att = torch.zeros(n, dtype=torch.int32, device=c.dev)
    P = ctypes.c_void_p
    skeys = [dv(K[f][key_of]) for f in ("rho", "k", "tr", "s1", "s2", "t0")]   # one secret key per signature (kept alive over the call)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = eng._lib.dil_sign_multi_dev(eng._h, level, *[P(t.data_ptr()) for t in skeys],
                                     P(d_msgs.data_ptr()), P(d_off.data_ptr()), n, P(z.data_ptr()), P(h.data_ptr()), P(ct.data_ptr()),
                                     P(att.data_ptr()), eng._stream())
    eng._check(rc, "dil_sign_multi_dev")
    torch.cuda.synchronize()
    sign_multi_s = time.perf_counter() - t0
    del skeys
    # the injected tuples are the reference's own signatures: signing them again must reproduce the KAT bytes
    kat_sig_ok = bool(np.array_equal(z[pos].cpu().numpy(), K["zs"]) and np.array_equal(h[pos].cpu().numpy(), K["h"]) and
                      np.array_equal(ct[pos].cpu().numpy(), K["c"]))
    z[torch.from_numpy(pos).to(c.dev)] = dv(K["zs"]); h[torch.from_numpy(pos).to(c.dev)] = dv(K["h"])
    ct[torch.from_numpy(pos).to(c.dev)] = dv(K["c"])
    ok = torch.zeros(n, dtype=torch.uint8, device=c.dev)
    ms = c.time_ms(lambda: eng.verify_multi_dev(level, d_rho, d_t1, d_msgs, d_off, n, z, h, ct, ok), 3, warm=1)
    okh = ok.cpu().numpy()
    kat_ok = bool(okh[pos].all())
    all_ok = bool(okh.all())
    # tampered copy: exactly the tampered positions must be rejected
    z2 = z.clone()
    z2[5::1000, 77] ^= 8
    eng.verify_multi_dev(level, d_rho, d_t1, d_msgs, d_off, n, z2, h, ct, ok)
    exp = np.ones(n, np.uint8); exp[5::1000] = 0
    tamper_ok = bool(np.array_equal(ok.cpu().numpy(), exp))
    del z2
    # one shared key for comparison (A_hat expanded once per key): this rank's signatures under KAT key 0
    kk = kat_key(level)
    sk = d.SignKey(eng, level, *[kk[f] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
    vk = d.VerifyKey(eng, level, kk["rho"], kk["t1"])
    _, _, m1, o1, z1, h1, c1, a1 = sign_buffers(c, sk, n, SEED + 41 + c.rank)
    sk.sign_dev(m1, o1, n, z1, h1, c1, a1)
    ms_shared = c.time_ms(lambda: vk.verify_dev(m1, o1, n, z1, h1, c1, ok), 5, warm=1)
    shared_ok = bool(int(ok.sum().item()) == n)
    sk.close(); vk.close()
    if not (kat_ok and all_ok and tamper_ok and shared_ok and kat_sig_ok):
        raise RuntimeError(f"cfg4 check failed: kat_ok={kat_ok} all_ok={all_ok} tamper_ok={tamper_ok} shared_ok={shared_ok} kat_sig_ok={kat_sig_ok}")
    perms = 5 * k * l + (32 + k * 320) // 136 + 1                 # ExpandA per item + tr = SHAKE-256(rho || t1)
    bytes_item = (l + k + 1 + k) * 1024 + 32
    return {"workload": f"Dilithium-5 (k=8,l=7) verification, {n} signatures per GPU ({c.world * n} over {c.world} GPU(s)), a public key per signature",
            "n_gpus": c.world, "kat_tuples_injected": 100, "kat_positions": "1310*j + 7, j = 0..99, in every GPU's shard",
            "kat_tuples_accepted": kat_ok, "all_accepted": all_ok, "tampered_rejected_exactly": tamper_ok,
            "kat_signatures_reproduced_by_sign_multi": kat_sig_ok,
            "per_item_key": {"verifies_per_s": c.world * n / (ms * 1e-3), "ms": ms, "keccak_f_per_s": n * perms / (ms * 1e-3),
                             "keccak_frac_of_measured_peak": n * perms / (ms * 1e-3) / keccak_peak if keccak_peak else None,
                             "hbm_frac": n * bytes_item / (ms * 1e-3) / 1e9 / c.peak, "algorithmic_bytes_per_item": bytes_item,
                             "bound": f"integer ALU (Keccak: {perms} permutations per item)",
                             "timing": "CUDA events around dil_verify_multi_dev (keys, messages, signatures resident in HBM), max over ranks"},
            "shared_key": {"verifies_per_s": c.world * n / (ms_shared * 1e-3), "ms": ms_shared},
            "setup_sign_multi_s": sign_multi_s}


def kat_length_record(c, key, n):
    """Messages of KAT-like lengths (the reference's vectors are 33 .. 3300 bytes, rtl_tb/tb_sign_top.v feeds mlen per message):
    the same key and batch size as the headline, lengths cycling through the 100 KAT message lengths."""
    import numpy as np
    torch = c.torch
    M = np.load(os.path.join(ROOT, "tests", "golden", "kat_msgs.npz"))
    mlen = np.resize(M["mlen"].astype(np.int64), n)
    off = np.concatenate([[0], np.cumsum(mlen)])
    gen = torch.Generator(device="cpu").manual_seed(SEED + 77 + c.rank)
    msgs = torch.randint(0, 256, (int(off[-1]),), dtype=torch.uint8, generator=gen).to(c.dev)
    offs = torch.from_numpy(off).to(c.dev)
    z = torch.empty((n, key.z_bytes), dtype=torch.uint8, device=c.dev); h = torch.empty((n, key.h_bytes), dtype=torch.uint8, device=c.dev)
    ct = torch.empty((n, 32), dtype=torch.uint8, device=c.dev); att = torch.zeros(n, dtype=torch.int32, device=c.dev)
    ms = c.time_ms(lambda: key.sign_dev(msgs, offs, n, z, h, ct, att), 5, warm=2)
    key.set_profile(True)
    key.sign_dev(msgs, offs, n, z, h, ct, att)
    torch.cuda.synchronize()
    init_ms = key.get_profile()["init"][0]
    key.set_profile(False)
    return {"workload": f"{n} messages per GPU, lengths cycling through the 100 KAT message lengths ({int(mlen.min())} .. {int(mlen.max())} bytes, "
                        f"mean {float(mlen.mean()):.0f})", "signs_per_s": c.world * n / (ms * 1e-3), "ms": ms, "message_bytes_per_batch": int(off[-1]),
            "mu_hash_ms": init_ms, "mu_hash_note": "mu = SHAKE-256(tr || M): one Keccak state per message, 1 permutation per 136 message bytes"}


def cfg5_sweep(c, max_log2):
    """cfg5: global batch B = 2^10 .. 2^22 (x4 steps) x levels 2/3/5, sharded evenly over the ranks: forward NTT of the
    batch's l*B polynomials, the fused sign core, full signing and (shared-key) verification."""
    torch, d, eng = c.torch, c.d, c.eng
    rows = []
    for level in (2, 3, 5):
        k, l = LEVEL_DIMS[level]
        kk = kat_key(level)
        sk = d.SignKey(eng, level, *[kk[f] for f in ("rho", "k", "tr", "s1", "s2", "t0")])
        vk = d.VerifyKey(eng, level, kk["rho"], kk["t1"])
        a_hat = eng.expand_a(torch.from_numpy(kk["rho"]).to(c.dev), k, l)[0].contiguous()
        for lg in range(10, max_log2 + 1, 2):
            B = 1 << lg
            b = max(B // c.world, 1)
            it = 20 if b <= 16384 else (5 if b <= 262144 else 2)
            y = torch.randint(0, Q, (b, l, 256), dtype=torch.int32, device=c.dev)
            o = torch.empty_like(y)
            w = torch.empty((b, k, 256), dtype=torch.int32, device=c.dev)
            ms_ntt = c.time_ms(lambda: eng.ntt(y, out=o), it, warm=1)
            ms_core = c.time_ms(lambda: eng.signcore(a_hat, y, k, l, w=w), it, warm=1)
            del y, o, w
            torch.cuda.empty_cache()
            _, _, msgs, offs, z, h, ct, att = sign_buffers(c, sk, b, SEED + lg)
            ms_sign = c.time_ms(lambda: sk.sign_dev(msgs, offs, b, z, h, ct, att), max(it // 2, 1), warm=1)
            ok = torch.zeros(b, dtype=torch.uint8, device=c.dev)
            ms_ver = c.time_ms(lambda: vk.verify_dev(msgs, offs, b, z, h, ct, ok), max(it // 2, 1), warm=1)
            accepted = int(ok.sum().item()) == b
            del msgs, offs, z, h, ct, att, ok
            torch.cuda.empty_cache()
            n = c.world * b
            rows.append({"level": level, "batch": n, "per_gpu": b,
                         "ntt_polys_per_s": n * l / (ms_ntt * 1e-3), "ntt_hbm_frac": b * l * 2048 / (ms_ntt * 1e-3) / 1e9 / c.peak,
                         "signcore_items_per_s": n / (ms_core * 1e-3), "signcore_hbm_frac": b * (k + l) * 1024 / (ms_core * 1e-3) / 1e9 / c.peak,
                         "signs_per_s": n / (ms_sign * 1e-3), "verifies_per_s": n / (ms_ver * 1e-3), "verify_all_accepted": accepted})
        sk.close(); vk.close()
    return {"workload": "batch sweep 2^10 .. 2^%d x levels 2/3/5, global batch sharded over %d GPU(s)" % (max_log2, c.world),
            "note": "working sets below the 126 MB L2 (small batches) run from cache: their hbm_frac is not an HBM rate",
            "rows": rows}


def keygen_record(c, keccak_peak, n=65536):
    """SURVEY.md 8f row N4 under the same clock: batched key generation (combined_top.v mode 0, tb_keygen_top.v:145-275) from
    32-byte seeds, device-resident (dil_keygen_batch_dev) at every level; the first 100 seeds of the level-2 batch are the KAT
    seeds and must reproduce the KAT keys.  Keccak work per key: 1 (seed) + (k + l) eta-sampler polynomials (>= 1-2 permutations
    each) + 5 k l (ExpandA) + tr (SHAKE-256 over rho || t1: 10 / 15 / 20 permutations)."""
    import numpy as np
    import oracle_lib as ol
    torch, eng = c.torch, c.eng
    out = {"workload": f"{n} keys per GPU from 32-byte seeds, device-resident", "levels": {}}
    for level in (2, 3, 5):
        k, l = LEVEL_DIMS[level]
        K = ol.kat(level)
        gen = torch.Generator(device="cpu").manual_seed(SEED + 40 + level + 16 * c.rank)
        seeds = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=gen)
        seeds[:100] = torch.from_numpy(np.ascontiguousarray(K["z"][:100]))   # z_*.txt holds the keygen seeds xi
        d_seeds = seeds.to(c.dev)
        keys = eng.keygen_dev(level, d_seeds)
        torch.cuda.synchronize()
        kat_ok = None
        if True:
            kat_ok = all(bool(np.array_equal(keys[f][:100].cpu().numpy(), K[f][:100])) for f in ("rho", "k", "tr", "s1", "s2", "t1", "t0"))
            if not kat_ok:
                raise RuntimeError(f"bench: keygen level {level} does not reproduce the KAT keys")
        ms = c.time_ms(lambda: eng.keygen_dev(level, d_seeds), 5, warm=1)
        perms = 1 + 2 * (k + l) + 5 * k * l + (32 + k * 320) // 136 + 1
        out["levels"][str(level)] = {"keys_per_s": c.world * n / (ms * 1e-3), "ms": ms, "kat_keys_reproduced": kat_ok,
                                     "keccak_f_per_key_approx": perms,
                                     "keccak_frac_of_measured_peak": n * perms / (ms * 1e-3) / keccak_peak}
        del keys, d_seeds
    return out


def run_engine(args):
    import numpy as np
    c = Ctx()
    torch, dist, d, eng = c.torch, c.dist, c.d, c.eng
    world, rank, local, dev = c.world, c.rank, c.local, c.dev
    level = args.level
    k, l = LEVEL_DIMS[level]
    B = args.batch

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
    msgs_host, offs_host, msgs, offs, z, h, ct, att = sign_buffers(c, key, B, SEED + 1 + rank)

    def step():
        key.sign_dev(msgs, offs, B, z, h, ct, att)

    # Batches in flight: a streaming caller keeps several 65 536-message batches going at once - one key handle (same key), one
    # stream, one host thread and one set of output buffers per batch in flight; step i runs in context i mod T.  The late, small
    # rejection rounds of one batch leave SMs idle (Keccak wave quantisation, latency-bound rounds); the other batches' kernels fill
    # them, and the engine speculates less when it sees the load (sign_api.cu spec_for).  Every step is still one full batch, and
    # all K steps start and finish inside the timed region.  The strictly serial figure is measured first and reported next to it.
    T = max(1, min(args.in_flight, args.steps))
    ctxs = [(key, z, h, ct, att, torch.cuda.Stream())]
    for _ in range(1, T):
        kx = d.SignKey(eng, level, *[parts[f] for f in fields])
        ctxs.append((kx, torch.empty_like(z), torch.empty_like(h), torch.empty_like(ct), torch.zeros_like(att), torch.cuda.Stream()))

    def run_pipelined(n_steps, start_ev=None):
        if args.driver == "async":
            # ONE host thread: step i is begun on context i mod T as soon as that context's previous step has been finished
            if start_ev is not None:
                for cx_ in ctxs:
                    cx_[5].wait_event(start_ev)
            for i in range(n_steps):
                kx, zx, hx, cx, ax, st = ctxs[i % T]
                if i >= T:
                    kx.finish()
                with torch.cuda.stream(st):
                    kx.sign_dev_begin(msgs, offs, B, zx, hx, cx, ax)
            for i in range(max(0, n_steps - T), n_steps):
                ctxs[i % T][0].finish()
            for cx_ in ctxs:
                torch.cuda.current_stream().wait_stream(cx_[5])
            return
        errs = []

        def worker(wi):
            try:
                torch.cuda.set_device(local)
                kx, zx, hx, cx, ax, st = ctxs[wi]
                with torch.cuda.stream(st):
                    if start_ev is not None:
                        st.wait_event(start_ev)
                    for _ in range(wi, n_steps, T):
                        kx.sign_dev(msgs, offs, B, zx, hx, cx, ax)
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=worker, args=(wi,)) for wi in range(T)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        for cx_ in ctxs:
            torch.cuda.current_stream().wait_stream(cx_[5])

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    c.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    c.barrier()
    serial_ms = c.maxr(ev0.elapsed_time(ev1))
    serial = {"value": world * B * args.steps / (serial_ms * 1e-3), "unit": UNIT, "ms_per_step": serial_ms / args.steps,
              "rejection_rounds": key.last_rounds, "attempt_slots_per_step": key.last_slots}
    if T > 1:
        run_pipelined(max(warm, T))
        c.barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        if T > 1:
            run_pipelined(args.steps, ev0)
        else:
            for _ in range(args.steps):
                step()
        ev1.record()
        c.barrier()
    launches = eng.launch_count - l0
    ms_total = c.maxr(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    mean_attempts = float(att.float().mean().item())
    rounds, slots = key.last_rounds, key.last_slots
    pipelined_same = all(bool(torch.equal(zx, z) and torch.equal(hx, h) and torch.equal(cx, ct)) for _, zx, hx, cx, _, _ in ctxs[1:])
    if not pipelined_same:
        raise RuntimeError("bench: batches in flight produced different signatures for the same messages")

    # per-kernel-class device time of one step on its own (CUDA events around every launch), outside the timed region
    key.set_profile(True)
    step()
    torch.cuda.synchronize()
    prof = {n: v for n, v in key.get_profile().items() if n != "pack_w1"}   # w1 packing is fused into the sign core
    key.set_profile(False)
    prof_rounds = max(key.last_rounds, 1)
    prof_total = sum(ms for ms, _ in prof.values())
    dominant = max((n for n in prof if n != "init"), key=lambda n: prof[n][0])
    cb = class_bytes(level)

    # the pure Keccak-f rate of THIS GPU in THIS run: the ALU-pipe roofline of the hash kernels
    with ClockSampler(local) as kclk:
        krates = {cps: eng.keccak_rate(ctas_per_sm=cps, perms=1500) for cps in (1, 2, 3)}
    keccak_peak = max(krates.values())
    kmix = keccak_sass_mix()
    kclock = kclk.summary()["sm_mhz"]
    keccak_model = None
    if kmix and kclock:
        # every SM sub-partition issues one ALU-pipe warp instruction per 2 cycles (16 lanes wide)
        keccak_model = eng.sm_count * 4 * (kclock * 1e6) / 2 * 32 / (kmix["alu_pipe"] * 24)
    keccak_src = ("dil_diag_keccak_dev in this run: %d SMs x {1,2,3} CTAs of 128 threads x 1500 permutations, CUDA events, best of 3; "
                  "rates %s G/s" % (eng.sm_count, {kk_: round(v / 1e9, 3) for kk_, v in krates.items()}))

    yb = torch.randint(0, Q, (B, l, 256), dtype=torch.int32, device=dev)
    outb = torch.empty_like(yb)
    wb = torch.empty((B, k, 256), dtype=torch.int32, device=dev)
    a_hat = eng.expand_a(torch.from_numpy(parts["rho"]).to(dev), k, l)[0].contiguous()
    npoly = B * l
    ntt_ms = c.time_ms(lambda: eng.ntt(yb, out=outb), 50, warm=3)
    intt_ms = c.time_ms(lambda: eng.invntt(yb, out=outb), 50, warm=3)
    core_ms = c.time_ms(lambda: eng.signcore(a_hat, yb, k, l, w=wb), 50, warm=3)
    del yb, outb, wb

    # end to end through the host-pointer C ABI: pinned host messages in, signatures out to host memory.  T batches are in
    # flight as above (T key handles of the same key, T sets of pinned buffers, one host thread each): the next batch already
    # signs while the previous one's last signatures drain, as a streaming caller would use the API.  Every step still includes
    # the H2D copy of its messages and the D2H of all its signatures; the strictly serial figure is reported next to it.
    P = ctypes.c_void_p
    lib = eng._lib
    bufs = []
    for _ in range(T):
        bufs.append((torch.empty((B, key.z_bytes), dtype=torch.uint8).pin_memory(), torch.empty((B, key.h_bytes), dtype=torch.uint8).pin_memory(),
                     torch.empty((B, 32), dtype=torch.uint8).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()))

    def e2e_call(which):
        z_h, h_h, c_h, a_h = bufs[which]
        rc = lib.dil_sign_batch_host(eng._h, ctxs[which][0]._h, P(msgs_host.data_ptr()), P(offs_host.data_ptr()), B, P(z_h.data_ptr()),
                                     P(h_h.data_ptr()), P(c_h.data_ptr()), P(a_h.data_ptr()))
        if rc != 0:
            raise RuntimeError("dil_sign_batch_host failed")

    def e2e_run(n_steps, in_flight):
        if in_flight == 1:
            for _ in range(n_steps):
                e2e_call(0)
            return
        if args.driver == "async":
            for i in range(n_steps):
                which = i % in_flight
                if i >= in_flight:
                    ctxs[which][0].finish()
                ctxs[which][0].sign_host_begin(msgs_host, offs_host, B, *bufs[which])
            for i in range(max(0, n_steps - in_flight), n_steps):
                ctxs[i % in_flight][0].finish()
            return
        errs = []

        def worker(which):
            try:
                torch.cuda.set_device(local)
                for _ in range(which, n_steps, in_flight):
                    e2e_call(which)
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=worker, args=(w,)) for w in range(in_flight)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    e2e_steps = max(args.e2e_steps, T)
    e2e_res = {}
    for in_flight in sorted({1, T}):
        e2e_run(max(2, in_flight), in_flight)
        c.barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps, in_flight)
        c.barrier()
        e2e_res[in_flight] = world * B * e2e_steps / c.maxr(time.perf_counter() - t0)
    e2e_val = e2e_res[T]
    e2e_ok = all(bool(torch.equal(zh.to(dev), z) and torch.equal(hh.to(dev), h) and torch.equal(ch.to(dev), ct)) for zh, hh, ch, _ in bufs)
    del bufs
    for kx, *_ in ctxs[1:]:
        kx.close()

    configs = None
    if not args.no_configs:
        t_cfg = time.perf_counter()
        configs = {"kat_length_messages": kat_length_record(c, key, B), "keygen": keygen_record(c, keccak_peak), "cfg3": cfg3_record(c, keccak_peak),
                   "cfg4": cfg4_record(c, keccak_peak), "cfg5": cfg5_sweep(c, args.sweep_max_log2)}
        configs["wall_s"] = time.perf_counter() - t_cfg

    if rank == 0:
        peak, peak_src = c.peak, c.peak_src
        dom_ms, dom_units = prof[dominant]
        dom_bytes = cb[dominant] * dom_units
        dom_achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        sig_bytes = key.z_bytes + key.h_bytes + 32
        traffic = ncu_traffic(dominant)
        ceiling = HOST_CEILING_GBS.get(world)
        line = {
            "metric": metric_name(level), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32/u64 integer (Shoup/Barrett modular arithmetic, 64-bit Keccak lanes)", "data": "synthetic",
            "config": {"workload": workload_name(level, B), "level": level, "k": k, "l": l, "batch_per_gpu": B,
                       "key": "reference KAT vector 0 (tests/golden)", "mean_attempts": mean_attempts,
                       "batches_in_flight": T,
                       "batches_in_flight_driver": ("one host thread: dil_sign_batch_dev_begin / dil_sign_batch_finish" if args.driver == "async"
                                                    else "one host thread per batch in flight: dil_sign_batch_dev"),
                       "batches_in_flight_note": ("every step is one full batch; step i runs on key handle / stream i mod T, all K steps "
                                                  "start and finish inside the timed region; `one_batch_at_a_time` is the strictly serial figure"),
                       "rejection_rounds": rounds,
                       "attempt_slots_per_step": slots, "host_syncs_per_step": "1 (the rejection loop runs on the device; rounds are enqueued ahead)",
                       "sharding": ("independent messages, one contiguous shard per rank; one NCCL broadcast of the key material"
                                    if world > 1 else "single GPU"),
                       "l2": "no flush: every round streams > 1 GiB of per-attempt state (y, w, c), far above the 126 MB L2",
                       "timing": ("CUDA events on torch's current stream; the batch streams wait for the start event and the current stream waits "
                                  "for every batch stream before the stop event; max over ranks"),
                       "kernels_per_round": "ExpandMask, sign core (+ packed HighBits), challenge, tail (+ resolve when one slot per item), "
                                            "resolve (returns at once unless the round speculates), plan"},
            "one_batch_at_a_time": serial,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "step_profile_ms": {n: round(ms, 4) for n, (ms, _) in prof.items()},
            "step_profile_note": (f"device time per kernel class of one step run alone (sum {prof_total:.3f} ms; one batch at a time takes "
                                  f"{serial['ms_per_step']:.3f} ms per step, {ms_step:.3f} ms with {T} in flight)"),
            "roofline": {"kernel": class_kernel_name(dominant, level), "class": dominant, "bound": "hbm", "achieved": dom_achieved, "peak": peak,
                         "unit": "GB/s", "frac": dom_achieved / peak,
                         "traffic": (traffic / 65536.0 * dom_units / prof_rounds) if traffic else None,
                         "traffic_note": "ncu dram bytes of a 65536-slot launch scaled to this step's average launch size",
                         "algorithmic_bytes_per_slot": cb[dominant], "slots_per_step": dom_units, "launches_per_step": prof_rounds,
                         "avg_launch_ms": dom_ms / prof_rounds, "peak_source": peak_src,
                         "ncu_limiter": ncu_limiter(dominant),
                         "compute_roofline": keccak_roofline(dominant, level, dom_units, dom_ms, keccak_peak / 1e9, keccak_src)},
            "roofline_step": {"what": "all kernel classes of the timed steps together: algorithmic HBM bytes per attempt slot x slots per step / step time",
                              "algorithmic_bytes_per_slot": sum(cb[n] for n in ("expand_mask", "signcore", "challenge", "tail")),
                              "slots_per_step": slots, "ms_per_step": ms_step,
                              "achieved": sum(cb[n] for n in ("expand_mask", "signcore", "challenge", "tail")) * slots / (ms_step * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s",
                              "frac": sum(cb[n] for n in ("expand_mask", "signcore", "challenge", "tail")) * slots / (ms_step * 1e-3) / 1e9 / peak,
                              "note": "the step is bound by the integer ALU pipe (Keccak) and the multiplier pipe (transforms), not by HBM (DESIGN.md 8)"},
            "keccak_peak": {"measured_g_per_s": keccak_peak / 1e9, "source": keccak_src, "sass_mix_per_round": kmix, "sm_mhz_during": kclock,
                            "alu_pipe_model_g_per_s": keccak_model / 1e9 if keccak_model else None,
                            "alu_pipe_model": "sm_count x 4 sub-partitions x clock / 2 cycles per ALU warp instruction x 32 threads / "
                                              "(ALU-pipe instructions per round x 24 rounds)"},
            "keccak_classes": {n: keccak_roofline(n, level, prof[n][1], prof[n][0], keccak_peak / 1e9, "same run") for n in ("expand_mask", "challenge")},
            "roofline_ntt": {"kernel": "ntt_tma_kernel<32,3,1,false> (stand-alone forward NTT, the north_star's named kernel)",
                             "bound": "hbm", "polys_per_s": npoly / (ntt_ms * 1e-3), "achieved": npoly * 2048 / (ntt_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": npoly * 2048 / (ntt_ms * 1e-3) / 1e9 / peak, "polys_per_launch": npoly,
                             "invntt_polys_per_s": npoly / (intt_ms * 1e-3), "invntt_frac": npoly * 2048 / (intt_ms * 1e-3) / 1e9 / peak},
            "sign_core": {"workload": f"cfg2 core alone: w = INTT(A_hat * NTT(y)), {B} items, one fused kernel",
                          "items_per_s": B / (core_ms * 1e-3), "ms": core_ms,
                          "hbm_frac": B * (k + l) * 1024 / (core_ms * 1e-3) / 1e9 / peak, "ncu_limiter": ncu_limiter("signcore")},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * MSG_BYTES + (B + 1) * 8,
                    "d2h_bytes_per_step": B * (sig_bytes + 4), "steps": e2e_steps,
                    "api": ("dil_sign_batch_host_begin / dil_sign_batch_finish (C ABI, one host thread)" if args.driver == "async" and T > 1 else
                            "dil_sign_batch_host (C ABI)") + ": pinned host messages in, z/h/c~/attempts back in host memory",
                    "batches_in_flight": T, "one_call_at_a_time": e2e_res[1],
                    "timing": "host wall clock around the calls (one key handle and pinned buffer set per batch in flight), max over ranks",
                    "matches_device_path": e2e_ok,
                    "host_ceiling": ({"aggregate_gb_per_s": ceiling, "signatures_per_s": ceiling * 1e9 / (sig_bytes + 4),
                                      "e2e_frac_of_ceiling": e2e_val * (sig_bytes + 4) / (ceiling * 1e9),
                                      "source": "profiles/r2_d2h_ceiling_8gpu.txt: barrier-synchronised D2H of all GPUs at once on the 8-GPU box, "
                                                "copy engine and SM stores alike (tools/d2h_ceiling.cu)"} if ceiling else None)},
        }
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            # the CPU leg also CHECKS the timed step's output: a strided sample of the signatures the engine just produced goes
            # through the oracle's verify (the oracle is the checker here, never the thing measured)
            import oracle_lib as ol
            orc = ol.load()
            kk0 = kat_key(level)
            mh = msgs_host.numpy().reshape(B, MSG_BYTES)
            zc, hc_, cc = z.cpu().numpy(), h.cpu().numpy(), ct.cpu().numpy()
            sample_idx = list(range(0, B, max(B // 24, 1)))
            checked_ok = all(orc.verify(level, kk0["rho"], kk0["t1"], mh[i].tobytes(), zc[i], hc_[i], cc[i]) == 0 for i in sample_idx)
            if not checked_ok:
                raise RuntimeError("bench: the oracle rejected a signature of the timed step")
            cores = host_threads()
            n_cpu = 256 * cores if 256 * cores < 8192 else 8192
            v, kind, threads, sample, _, _ = cpu_sign(level, n_cpu, cores, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
            line["cpu_baseline"]["oracle_verified_signatures_of_the_timed_step"] = len(sample_idx)
            ntt_cpu, ntt_kind = cpu_ntt(cores)
            line["cpu_baseline"]["ntt_polys_per_s"] = ntt_cpu
            line["cpu_baseline"]["ntt_sample"] = f"65536 polynomials through {'the compiled reference ntt()' if ntt_kind == 'reference' else 'the oracle port'} on {cores} thread(s)"
            if configs is not None:
                for lv in (3, 5):
                    vv, *_ = cpu_sign(lv, max(n_cpu // 4, 64), cores, steps=1, warmup=0)
                    line["cpu_baseline"][f"dilithium{lv}_signs_per_s"] = vv
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
