#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (captured with --import-source on, -lineinfo build).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] -> file:line, samples, share, dominant stall reasons, source text"""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        lines.append((int(d["Warp Stall Sampling (All Samples)"] or 0), cur_file, int(r[0]), r[1].strip(), stalls,
                      int(d.get("Instructions Executed", "0") or 0)))
total = sum(l[0] for l in lines) or 1
print(f"# {path}: {total} warp-stall samples over {len(lines)} source lines with code")
for smp, f, ln, src, stalls, inst in sorted(lines, key=lambda x: -x[0])[:top]:
    why = " ".join(f"{k[6:]}={v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * smp / total:5.1f}%  {smp:7d}  inst={inst:9d}  {f}:{ln:<4d} {why:44s} | {src[:100]}")
