#!/usr/bin/env python
"""Derive profiles/dominant_kernel.json (read by bench.py for `roofline.traffic`) from the per-kernel ncu
summaries that tools/ncu_summary.py wrote:  python tools/make_dominant.py [tag]   (default tag r1b)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2d"
CLASSES = {"expand_mask": "expand_mask_kernel", "signcore": "matvec_shared_kernel", "challenge": "challenge_kernel",
           "tail": "sign_tail_sparse_kernel"}
out = {}
for cls, kern in CLASSES.items():
    path = os.path.join(ROOT, "profiles", f"{tag}_sign_{kern}.json")
    l = json.load(open(path))["launches"][0]
    mb = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(l[f"dram__bytes_{d}.sum"] * mb[l[f"dram__bytes_{d}.sum__unit"]] for d in ("read", "write"))
    out[cls] = {"kernel": l["kernel"][:120], "dram_bytes_per_launch": dram, "slots_per_launch": 65536,
                "time_us": l["gpu__time_duration.sum"],
                "alu_pipe_pct": l["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"],
                "fmaheavy_pipe_pct": l["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"],
                "issue_pct": l["sm__issue_active.avg.pct_of_peak_sustained_elapsed"],
                "dram_pct": l.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "round_tag": tag,
                "source": f"profiles/{tag}_sign_{kern}.json (ncu --set full, first launch of a 65536-message Dilithium-2 batch)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "dominant_kernel.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
