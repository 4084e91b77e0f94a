"""Reads an ncu report (--set full) here, without a GPU, and writes the per-kernel summary committed under profiles/:
duration, DRAM bytes, tensor-pipe activity, registers, occupancy, L2 hit rate.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv [profiles/x_traffic.json]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
if rep.endswith(".csv"):      # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): reports > 64 MiB do not travel
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def col(r, name, scale=1.0, default=""):
    i = idx.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", "")) * scale
    except ValueError:
        return default


def unit_scale(name, want):
    """ncu prints e.g. Mbyte / Gbyte / usecond / msecond: normalise."""
    u = units[idx[name]] if name in idx else ""
    table = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
             "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    return table.get(u, 1.0)


fields = ["id", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB", "dram_TBps", "tensor_pipe_active_pct",
          "sm_pct_peak", "regs", "achieved_occupancy_pct", "l2_hit_pct"]
out_rows, traffic, total = [], {}, 0.0
data = rows[2:]
if "--all" not in sys.argv:
    data = data[len(data) // 2:]      # scripts/ncu_step.py runs two identical steps: keep the second (warm) one
for n, r in enumerate(data):
    name = r[idx["Kernel Name"]]
    t = col(r, "gpu__time_duration.sum", unit_scale("gpu__time_duration.sum", "us"), 0.0)
    rd = col(r, "dram__bytes_read.sum", unit_scale("dram__bytes_read.sum", "byte"), 0.0)
    wr = col(r, "dram__bytes_write.sum", unit_scale("dram__bytes_write.sum", "byte"), 0.0)
    tp = col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0, "")
    if tp == "":
        tp = col(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", 1.0, "")
    out_rows.append([n, name.split("(")[0], r[idx["Grid Size"]] if "Grid Size" in idx else "", r[idx["Block Size"]] if "Block Size" in idx else "",
                     round(t, 1), round(rd / 1e6, 1), round(wr / 1e6, 1), round((rd + wr) / max(t, 1e-9) / 1e6, 2), tp,
                     col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), col(r, "launch__registers_per_thread"),
                     col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), col(r, "lts__t_sector_hit_rate.pct")])
    key = name.split("(")[0].split("<")[0].replace("void ", "").replace("csn::", "")
    d = traffic.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    d["launches"] += 1
    d["dram_bytes"] += rd + wr
    d["time_us"] += t
    total += rd + wr
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(fields)
    w.writerows(out_rows)
if len(sys.argv) > 3:
    for d in traffic.values():
        d["dram_bytes_per_launch"] = d["dram_bytes"] / d["launches"]
    traffic["whole_step"] = {"dram_bytes": total, "kernels": len(out_rows)}
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
print(f"{len(out_rows)} kernels, {total / 1e9:.2f} GB of DRAM traffic")
