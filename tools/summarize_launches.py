"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share of GPU time."""
import csv, sys, collections, re
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
    rows.append((int(r["ID"]), r["Kernel Name"], us))
rows = rows[skip:]
if len(sys.argv) > 3 and sys.argv[3] == "laststep":
    idx = [i for i, r in enumerate(rows) if "adam_kernel" in r[1]]
    if len(idx) >= 2:
        rows = rows[idx[-2] + 1: idx[-1] + 1]
agg = collections.OrderedDict()
for _, name, us in rows:
    name = re.sub(r"\(.*", "", name)
    name = name[:90]
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1; c[1] += us
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total GPU time {tot/1e3:.3f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}%  x{n:<5d} {name}")
