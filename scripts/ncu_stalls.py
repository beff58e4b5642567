#!/usr/bin/env python
"""Per-SASS-instruction stall reasons of the first kernel in an .ncu-rep (source page, SASS view).
  python scripts/ncu_stalls.py gpurun_out/x.ncu-rep [min share in %, default 0.6] [context: 0|1 prints the whole loop region]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
ins = []; last = -1
for r in rows[hi + 1:]:
    try:
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    except (ValueError, IndexError):
        continue
    if a < last: break
    last = a; ins.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in ins) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in ins) for h in stalls}
print("%d SASS instructions, %d samples; " % (len(ins), tot) + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for k, r in enumerate(ins):
    s = int(r[ix["# Samples"]] or 0)
    if s >= tot * thr / 100.0:
        st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stalls), key=lambda x: -x[1])[:3]
        print("%5d  %-52s %5.1f%%  exec %9s  %s" % (k, r[1][:52], 100.0 * s / tot, r[ix["Instructions Executed"]], " ".join("%s:%d" % x for x in st if x[1])))
