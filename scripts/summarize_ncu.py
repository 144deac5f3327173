#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.csv
  python scripts/summarize_ncu.py full gpurun_out/r01_gemm.ncu-rep profiles/r01_gemm_full.csv
"""
import collections
import csv
import re
import statistics
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    data = []
    for row in r:
        try:
            v = float(row[vi].replace(",", ""))
        except Exception:
            continue
        if row[ui] == "ns":
            v /= 1e3
        data.append((re.sub(r"\(.*", "", row[ki]).replace("void ", "").strip(), row[gi], v))
    step = data[len(data) // 2:]  # second half = the measured step (first half is the warm-up step)
    agg = collections.defaultdict(list)
    for k, g, v in step:
        agg[(k, g)].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "launches", "median_us", "min_us", "max_us", "total_ms", "share_pct"])
        for (k, g), v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            w.writerow([k, g, len(v), round(statistics.median(v), 2), round(min(v), 2), round(max(v), 2),
                        round(sum(v) / 1e3, 3), round(100 * sum(v) / tot, 2)])
        w.writerow(["TOTAL (serialised, cold-cache per-launch times: compare shares, not absolutes)", "", len(step), "", "", "",
                    round(tot / 1e3, 3), 100.0])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    cols = [("Kernel Name", hdr.index("Kernel Name"))] + [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([c for c, _ in cols])
        w.writerow([rows[1][i] if c != "Kernel Name" else "(unit)" for c, i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
