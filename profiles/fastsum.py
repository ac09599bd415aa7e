#!/usr/bin/env python
"""Summary of an `ncu --set full --import-source on` capture of one kernel launch: headline metrics, stall mix per issue, and
the source lines with most executed warp instructions.
    python profiles/fastsum.py gpurun_out/x.ncu-rep [warps_per_launch] [top]
    python profiles/fastsum.py gpurun_out/x.ncu-rep --json "tz::fast_step_kernel/dense"     # record in step_kernel_traffic.json
(the --json form writes the numbers bench.py quotes as roofline.traffic / fp64_fraction / issue_fraction)"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_write.sum',
        'sm__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def record(rep, key):
    import json
    import os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]

    def val(k, scale=1.0):
        i = h.index(k)
        unit = u[i].lower()
        mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        return float(v[i]) * mult * scale
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_kernel_traffic.json")
    try:
        d = json.load(open(out))
    except Exception:
        d = {}
    d[key] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
              "fp64_pipe_frac": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.01),
              "issue_active_frac": val("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.01),
              "gpu_time_us_under_ncu": val("gpu__time_duration.sum") * ({"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u[h.index("gpu__time_duration.sum")], 1.0)),
              "registers": int(float(v[h.index("launch__registers_per_thread")])),
              "kernel": v[h.index("Kernel Name")][:160],
              "source": "ncu --set full --clock-control none, one launch, 65,536 scenarios of the 5-dim workload: " + os.path.basename(rep)}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d[key], indent=1))


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "--json":
        return record(rep, sys.argv[3])
    warps = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    print("kernel:", v[h.index("Kernel Name")][:140])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k},{u[i]},{v[i]}")
    for i, k in enumerate(h):
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(v[i] or 0) > 0.15:
            print(f"stall_per_issue {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')},{v[i]}")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    fname, hdr, out = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            smp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        out.append((inst, smp, fname, int(r[0]), r[1].strip()))
    tot = sum(o[0] for o in out) or 1
    ts = sum(o[1] for o in out) or 1
    print(f"source lines: total warp instructions {tot}, per warp {tot / warps:.0f}")
    for inst, smp, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
        print(f"{100 * inst / tot:5.1f}% ins {inst / warps:7.1f}/warp {100 * smp / ts:5.1f}% smp {f}:{ln:<4d} {src[:100]}")


if __name__ == "__main__":
    main()
