"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name:
python tools/ncu_launch_agg.py <launches.csv> [N]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:70]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print(f"{100*v/tot:5.1f}%  {v/1e3:9.3f} ms  n={c:4d}  avg={v/c:9.1f} us  {name}")
