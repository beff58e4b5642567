#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (the tracked evidence).

  python scripts/ncu_summary.py launches gpurun_out/<tag>_launches.csv            > profiles/<tag>_launches.md
  python scripts/ncu_summary.py kernel   gpurun_out/<tag>_force.ncu-rep [regex]   > profiles/<tag>_force.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void\s+", "", r[ix["Kernel Name"]].split("(")[0]).replace("xnb::", "")
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised launches: compare SHARES)")
    print("source: %s ; total %.1f us over %d launches\n" % (path, tot, sum(a[0] for a in agg.values())))
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k[:70], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    print("ncu --set full summary of %s (values per launch; profiler-replayed, not bench numbers)\n" % rep)
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        print("kernel: %s  grid %s block %s" % (d["Kernel Name"].split("(")[0], d.get("Grid Size"), d.get("Block Size")))
        for k in RAW_KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        print()
    rows = ncu_csv(rep, "source")
    # the source page repeats one table per launch; take the first
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return
    hdr = rows[hi[0]]; ix = {h: i for i, h in enumerate(hdr)}
    end = hi[1] - 1 if len(hi) > 1 else len(rows)
    data = [r for r in rows[hi[0] + 1:end] if len(r) >= len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    n_inst = n_samp = n_wave = 0
    for r in data:
        for h in stalls:
            tot[h] += int(r[ix[h]] or 0)
        n_inst += int(r[ix["Instructions Executed"]] or 0); n_samp += int(r[ix["# Samples"]] or 0)
        n_wave += int(r[ix["L1 Wavefronts Shared"]] or 0)
    print("source page, first launch: %d SASS instructions, %d warp-instructions executed, %d shared-memory wavefronts, %d samples" % (len(data), n_inst, n_wave, n_samp))
    print("warp stall samples: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(n_samp, 1)) for k, v in tot.most_common(8)))
    print("\nhottest SASS instructions by samples:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
        print("  %5.1f%%  exec %9s  smem wavefronts %9s  %s" % (100.0 * int(r[ix["# Samples"]]) / max(n_samp, 1), r[ix["Instructions Executed"]], r[ix["L1 Wavefronts Shared"]], r[ix["Source"]].strip()[:80]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2])
