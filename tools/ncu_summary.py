"""Condense `ncu -i report.ncu-rep --page raw --csv` into one line per launch with the metrics the roofline uses.

usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/rNN_ncu_sweep.txt"""
import csv
import re
import sys

KEEP = [("gpu__time_duration.sum", "time_us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
        ("dram__bytes_read.sum", "dram_rd_MB", 1e-6), ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__occupancy_limit_shared_mem", "occ_smem", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct", 1)]
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
head, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


print("# kernel | " + " | ".join(n for _, n, _ in KEEP))
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    vals = []
    for m, n, scale in KEEP:
        if m not in col:
            vals.append("-")
            continue
        v = num(r[col[m]])
        u = units[col[m]]
        if m == "gpu__time_duration.sum":
            v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
            vals.append(f"{v:.1f}")
        elif "bytes" in m:
            v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append(f"{v:.1f}")
        else:
            vals.append(f"{v:.1f}")
    print(f"{name} | " + " | ".join(vals))
