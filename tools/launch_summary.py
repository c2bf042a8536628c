"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the top launches."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt, items = collections.defaultdict(float), collections.Counter(), []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000.0 if row["Metric Unit"] == "ns" else (v * 1000.0 if row["Metric Unit"] == "ms" else v)
    m = re.search(r"(\w+)(<[^>]*>)?\(", row["Kernel Name"].replace("void ", "").replace("egtr::<unnamed>::", ""))
    short = (m.group(1) + (m.group(2) or "")) if m else row["Kernel Name"][:40]
    tot[short] += v
    cnt[short] += 1
    items.append((v, short, row.get("Grid Size", ""), len(items)))
T = sum(tot.values())
print(f"total {T:.1f} us over {len(items)} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  {k}")
print(f"--- top {top} launches (us, kernel, grid, launch index)")
for v, s, g, i in sorted(items, reverse=True)[:top]:
    print(f"{v:9.1f} {s:40s} grid={g} #{i}")
