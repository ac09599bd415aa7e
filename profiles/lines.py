#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report (needs -lineinfo and --import-source on):
    python profiles/lines.py gpurun_out/prof.ncu-rep [top]
prints, for the `top` source lines with most warp-stall samples: file:line, share of samples, warp instructions
executed, dominant stall reasons, source text."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
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
            samples = int(d.get("# Samples", "0") or 0)
            inst = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        out.append((samples, inst, fname, int(r[0]), r[1].strip(), stalls))
    tot_s = sum(o[0] for o in out) or 1
    tot_i = sum(o[1] for o in out) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    agg = {}
    for s, i, f, ln, src, st in out:
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    print("stall mix:", ", ".join(f"{k} {100.0 * v / tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for s, i, f, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
        tops = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100.0 * s / tot_s:5.1f}% smp {100.0 * i / tot_i:5.1f}% ins  {f}:{ln:<4d} [{tops}]  {src[:110]}")


if __name__ == "__main__":
    main()
