#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked text/JSON file under profiles/.

    python scripts/ncu_summary.py gpurun_out/r1_full.ncu-rep profiles/r1_clahe_full.md
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu summary of `{rep}`", "", "Captured with `ncu --set full --clock-control none --import-source on` under gpurun "
             "(one B200); numbers below are per launch (cold cache, serialised replays).", ""]
    summary = []
    for d in data:
        name = d[hdr.index("Kernel Name")].split("(")[0]
        lines += [f"## {name}  (launch id {d[0]})", "", "| metric | value | unit |", "|---|---|---|"]
        rec = {"kernel": name}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {d[i]} | {units[i]} |")
                rec[k] = d[i] + " " + units[i]
        lines.append("")
        summary.append(rec)
    with open(out, "w") as f:
        f.write("\n".join(lines))
    with open(out.rsplit(".", 1)[0] + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
