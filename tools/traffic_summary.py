"""ncu csv (gpu__time_duration, dram__bytes_read/write per launch) -> per-kernel-class DRAM traffic summary (json)."""
import collections
import csv
import json
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
for r in csv.DictReader(rows):
    name = r["Kernel Name"]
    m = re.search(r"(\w+_kernel)", name)
    k = m.group(1) if m else name[:30]
    if k == "msda_kernel":
        k = "msda_enc" if re.search(r"msda_kernel<\w+, 32", name) else "msda_dec"
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    per[k][r["Metric Name"]] += v * scale
    if (r["ID"], k) not in seen:
        seen.add((r["ID"], k))
        cnt[k] += 1
out = {}
for k, d in per.items():
    out[k] = {"launches": cnt[k], "time_us": d["gpu__time_duration.sum"], "dram_read_bytes": d["dram__bytes_read.sum"],
              "dram_write_bytes": d["dram__bytes_write.sum"],
              "dram_bytes_per_launch": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / max(1, cnt[k])}
json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one forward (workload B, batch 1, cold caches per launch)",
           "kernels": out}, open(sys.argv[2], "w"), indent=1)
for k, v in sorted(out.items(), key=lambda kv: -kv[1]["time_us"]):
    print(f"{k:32s} n={v['launches']:3d} {v['time_us']:9.1f} us  dram {v['dram_read_bytes'] / 1e6:9.1f} MB read {v['dram_write_bytes'] / 1e6:9.1f} MB written")
