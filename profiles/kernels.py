#!/usr/bin/env python
"""Key metrics of every launch in an ncu report (any kernel):  python profiles/kernels.py report.ncu-rep [more.ncu-rep ...]
Duration, DRAM bytes, registers, occupancy limits, issue / warp activity, L2 hit rate and the stall reasons above 0.3
cycles per issued instruction."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]

for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {rep.split('/')[-1]}: {r[idx['Kernel Name']][:150]}")
        for w in WANT:
            if w in idx:
                print(f"   {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
        st = []
        for h in hdr:
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                v = float(r[idx[h]] or 0)
                if v > 0.3:
                    st.append((v, h.split("stalled_")[1].split("_per")[0]))
        print("   stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
