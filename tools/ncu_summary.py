#!/usr/bin/env python
"""`ncu -i REPORT.ncu-rep --page raw --csv` -> a small JSON list, one entry per profiled launch, with the metrics the
roofline discussion uses (B200_PROFILING.md).  Run in the build container (ncu is installed; no GPU needed to read).

    python tools/ncu_summary.py gpurun_out/ncu_r2.ncu-rep profiles/r2_ncu_full.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    res = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"Kernel Name": r[names.index("Kernel Name")], "launch": int(r[0])}
        for w in WANT:
            if w in names:
                i = names.index(w)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[f"{w} [{units[i]}]" if units[i] else w] = v
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["launch"], d["Kernel Name"][:60], {k.split(" ")[0].split(".")[0]: v for k, v in list(d.items())[2:5]})


if __name__ == "__main__":
    main()
