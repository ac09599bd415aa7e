#!/usr/bin/env python
"""Turns an `ncu --set full --import-source on` report of the step kernel into the committed evidence:

    python profiles/summarise.py gpurun_out/prof_step_<tag>.ncu-rep <tag> [--bucket "B0(NZ=2,NC=28,G=4)"]

writes profiles/<tag>_step_kernel_ncu_summary.csv (the raw-page metrics that matter + stall reasons + SASS opcode mix
from the source page) and updates profiles/step_kernel_traffic.json (dram bytes per launch, read by bench.py for
`roofline.traffic`).  Runs here (no GPU needed: `ncu -i` only reads the report)."""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    bucket = sys.argv[sys.argv.index("--bucket") + 1] if "--bucket" in sys.argv else "default"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    out = [["metric", "unit"] + [f"launch{i}" for i in range(len(launches))]]
    out.append(["Kernel Name", ""] + [r[hdr.index("Kernel Name")] for r in launches])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append([k, units[i]] + [r[i] for r in launches])
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            if max(float(r[i] or 0) for r in launches) > 0.1:
                out.append([h, units[i]] + [r[i] for r in launches])
    # traffic per launch
    tr = []
    for r in launches:
        t = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            t += float(r[i]) * UNIT.get(units[i], 1.0)
        tr.append(t)
    out.append(["dram_traffic_bytes_per_launch", "byte"] + [f"{t:.0f}" for t in tr])
    # SASS opcode mix of launch 0
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    try:
        h2 = next(r for r in srows if "Instructions Executed" in r)
        ie, sc = h2.index("Instructions Executed"), h2.index("Source")
        body = [(int(x[ie]), x[sc].strip()) for x in srows if len(x) > max(ie, sc) and x[ie].isdigit()]
        tot = sum(n for n, _ in body)
        cnt = collections.Counter()
        for n, s in body:
            t = s.split()
            op = t[1] if t[0].startswith("@") else t[0]
            cnt[op.split(".")[0] + (".MOV" if ".MOV" in op else "")] += n
        out.append(["sass_warp_instructions_total", "inst", str(tot)])
        for op, n in cnt.most_common(16):
            out.append([f"sass_share_{op}", "%", f"{100.0 * n / tot:.1f}"])
        hist = collections.Counter()
        for n, _ in body:
            hist[n] += 1
        for n, c in sorted(hist.items(), key=lambda kv: -kv[0] * kv[1])[:6]:
            out.append([f"sass_block_exec{n}_x{c}instr", "%", f"{100.0 * n * c / tot:.1f}"])
    except StopIteration:
        pass
    path = os.path.join(HERE, f"{tag}_step_kernel_ncu_summary.csv")
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows(out)
    tj = os.path.join(HERE, "step_kernel_traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d[bucket] = tr[0]
    d["default"] = tr[0]
    d["source"] = f"profiles/{tag}_step_kernel_ncu_summary.csv (ncu --set full, one launch, 65,536 scenarios of the 5-dim workload)"
    json.dump(d, open(tj, "w"), indent=1)
    print(open(path).read())


if __name__ == "__main__":
    main()
