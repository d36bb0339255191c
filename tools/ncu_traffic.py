"""profiles/ncu_traffic.json from an `ncu --set full` report of one compress of the benchmark field: DRAM bytes per
launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel, averaged over the kernel group bench.py reports
as its dominant kernel.

    ncu --set full --clock-control none -k regex:sweep -c 28 -o gpurun_out/X python tools/one_compress.py   (GPU box)
    python tools/ncu_traffic.py gpurun_out/X.ncu-rep 900 > profiles/ncu_traffic.json                         (here)
"""
import csv, io, json, subprocess, sys

rep, n_chunks = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
MODES = {"1": "sweep_tc_kernel<K3>", "2": "sweep_tc_kernel<K4>", "3": "sweep_tc_kernel<K4_OUT1>",
         "4": "sweep_tc_kernel<DOWN>", "5": "sweep_tc_kernel<UP>"}


def val(r, m):
    return float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)


groups = {}
for r in data:
    name = r[col["Kernel Name"]]
    if "sweep_acc_kernel<3" in name:
        key = MODES["1"]
    elif "sweep_acc_kernel<4" in name:
        key = MODES["2"]
    elif "sweep_tc_kernel<" in name:
        key = MODES[name.split("sweep_tc_kernel<")[1][0]]
    else:
        key = name.split("(")[0].split("::")[-1]
    grid = int(float(r[col["launch__grid_size"]].replace(",", "")))
    g = groups.setdefault(key, [])
    g.append({"kernel": name.split("(")[0].replace("void ", "").replace("<unnamed>::", ""), "grid": grid,
              "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
              "time_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
              {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)})
out = {"source": f"ncu --set full --clock-control none of tools/one_compress.py (24x720x1440, {n_chunks} chunks): {rep.split('/')[-1]}",
       "n_chunks": n_chunks, "kernels": {}}
for key, ls in groups.items():
    # full-resolution launches only (the stride-2 / transposed modes also run on the small inner levels)
    big = max(l["dram_read"] + l["dram_write"] for l in ls)
    sel = [l for l in ls if l["dram_read"] + l["dram_write"] > 0.25 * big]
    out["kernels"][key] = {
        "launches": len(sel), "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in sel) / len(sel),
        "dram_read_per_launch": sum(l["dram_read"] for l in sel) / len(sel),
        "dram_write_per_launch": sum(l["dram_write"] for l in sel) / len(sel),
        "time_us_under_ncu": sum(l["time_us"] for l in sel) / len(sel), "members": sorted({l["kernel"] for l in sel})}
print(json.dumps(out, indent=1))
