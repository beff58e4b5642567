#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (source page, CUDA view).
  python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top=25]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or not r[0].strip().isdigit(): continue
    try:
        agg.append((cur_file, int(r[0]), r[1].strip(), int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0)))
    except (ValueError, IndexError):
        pass
ti = sum(a[4] for a in agg) or 1; ts = sum(a[3] for a in agg) or 1
print("total warp-instructions %d, samples %d" % (ti, ts))
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * a[4] / ti, 100.0 * a[3] / ts, a[0], a[1], a[2][:110]))
