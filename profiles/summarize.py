"""Turns the raw captures under gpurun_out/ into the committed summaries of this directory.
  python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
  python profiles/summarize.py ncu gpurun_out/xyz.ncu-rep           > profiles/r1_xyz.md   (needs ncu on PATH)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, body = rows[hi], rows[hi + 2:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in body:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        agg[name][0] += float(r[vi].replace(",", ""))
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if v[0] / tot < 0.001:
            continue
        print("| `%s` | %d | %.0f | %.1f%% |" % (k[:90], v[1], v[0] / 1e3, 100 * v[0] / tot))
    print("\n%d launches, %.1f ms of serialised kernel time (ncu, cold cache: compare shares, not absolutes)" % (len(body), tot / 1e6))


def ncu(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("### `%s`  grid %s block %s\n" % (d.get("Kernel Name", "?")[:100], d.get("launch__grid_size"), d.get("launch__block_size")))
        print("| metric | value |\n|---|---:|")
        for k in KEYS:
            if k in d:
                print("| %s | %s |" % (k, d[k]))
        print()


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2])
