"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active over every launch of one pass; tools/gpu_ncu2.sh)
-> per-kernel JSON (profiles/*_traffic_c2_b1.json, read by bench.py for roofline.traffic).
    python tools/traffic_summary.py gpurun_out/traffic.csv > profiles/r1_x_traffic_c2_b1.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}
per = collections.OrderedDict()
launch = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ix["ID"]].strip().isdigit():
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    val = float(r[ix["Metric Value"]].replace(",", "")) * SCALE.get(r[ix["Metric Unit"]], 1.0)
    launch.setdefault((r[ix["ID"]], name), {})[r[ix["Metric Name"]]] = val
for (_, name), m in launch.items():
    d = per.setdefault(name, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "_tw": 0.0})
    us = m.get("gpu__time_duration.sum", 0.0)
    d["launches"] += 1
    d["us"] += us
    d["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
    d["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
    d["_tw"] += us * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
out = collections.OrderedDict()
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
    d["tensor_pipe_pct_time_weighted"] = d.pop("_tw") / d["us"] if d["us"] else 0.0
    out[name] = d
print(json.dumps(out, indent=1))
