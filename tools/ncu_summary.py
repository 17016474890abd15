#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON + text file for profiles/.

usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_x  [--units-per-launch N --unit-name polys]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units, unit_name = None, "units"
    if "--units-per-launch" in sys.argv:
        units = float(sys.argv[sys.argv.index("--units-per-launch") + 1])
    if "--unit-name" in sys.argv:
        unit_name = sys.argv[sys.argv.index("--unit-name") + 1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit_row, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i])
                except ValueError:
                    d[k] = r[i]
                d[k + "__unit"] = unit_row[i]
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                v = float(r[i])
                if v >= 0.05:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))

        def tobytes(key):
            u = d.get(key + "__unit", "byte").lower()
            mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return d.get(key, 0.0) * mult
        d["dram_bytes_per_launch"] = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        if units:
            d[f"warp_instructions_per_{unit_name}"] = d.get("smsp__inst_executed.sum", 0) / units
            d[f"dram_bytes_per_{unit_name}"] = d["dram_bytes_per_launch"] / units
        launches.append(d)
    json.dump({"source": rep, "launches": launches}, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in launches:
            f.write(d["kernel"] + "\n")
            for k, v in d.items():
                if k != "kernel" and not k.endswith("__unit"):
                    u = d.get(k + "__unit", "")
                    f.write(f"  {k:75s} {v} {u}\n")
            f.write("\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
