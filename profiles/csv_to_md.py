"""ncu `--page raw --csv` export -> compact markdown table (one row per captured launch)."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
print("|---|" + "---:|" * len(COLS))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"].replace("void ", "").split("(")[0].replace("segsde::", "")
    cells = []
    for k, _ in COLS:
        v = d.get(k, "")
        cells.append("%s %s" % (v, u.get(k, "")) if k.endswith(".sum") and "bytes" in k or k.startswith("gpu__time") else v)
    print("| `%s` | %s |" % (name, " | ".join(cells)))
