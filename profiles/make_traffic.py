#!/usr/bin/env python
"""profiles/ncu_traffic.json from this round's `ncu --set full` summaries (profiles/summarize.py output): per bench workload
the dominant kernel's DRAM bytes per launch and the utilisation figures bench.py copies into its JSON line (`roofline.traffic`,
`ncu`).  Regenerate after every new capture:  python profiles/make_traffic.py"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# workload -> summary file of the CURRENT dominant kernel (one launch, cold cache, --clock-control none)
SOURCES = {
    "cfg1": "r2_final_cfg1_k_rc1pass.txt",
    "cfg2": "r2_final_cfg2_k_ebs_coop.txt",
    "cfg3": "r2_final_cfg3_k_dos_shade.txt",
    "cfg4": "r2_final_cfg4_k_gt_shade.txt",
    "cfg5-1gpu": "r2_final_cfg5-1gpu_k_vct_shade.txt",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WANT = {
    "gpu__time_duration.sum": "duration_ms", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct", "launch__registers_per_thread": "registers",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions_per_launch",
}


def parse(path):
    rec = {}
    for line in open(path):
        if line.startswith("== "):
            if "kernel" in rec:
                break                                   # first kernel of the file only
            rec["kernel"] = line[3:].strip()
            continue
        parts = line.split()
        if len(parts) < 2:
            continue
        name = parts[0]
        if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            rec[name.replace("__", "_").replace(".sum", "")] = int(float(parts[1]) * UNIT.get(parts[2] if len(parts) > 2 else "byte", 1.0))
        elif name in WANT:
            v = float(parts[1])
            if name == "gpu__time_duration.sum" and len(parts) > 2:
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(parts[2], 1.0)
            rec[WANT[name]] = v
    return rec


def main():
    out = {}
    for wl, f in SOURCES.items():
        p = os.path.join(HERE, f)
        if not os.path.exists(p):
            continue
        rec = parse(p)
        if "dram_bytes_read" in rec and "dram_bytes_write" in rec:
            rec["source"] = "profiles/" + f + " (ncu --set full --clock-control none, one launch, cold cache)"
            out[wl] = rec
    json.dump(out, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps({k: (v["kernel"][:40], v["dram_bytes_read"] + v["dram_bytes_write"]) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
