"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel.
usage: launch_summary.py X.csv [first_id last_id]    (ids select one synthesis call out of the whole program)"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    rows.append((int(r["ID"]), r["Kernel Name"], us))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (rows[0][0], rows[-1][0])
sel = [r for r in rows if lo <= r[0] <= hi]
agg = defaultdict(list)
for _, name, us in sel:
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    agg[name].append(us)
tot = sum(sum(v) for v in agg.values())
print(f"# total {tot / 1000:.2f} ms over {len(sel)} launches (ids {lo}..{hi})")
print("kernel,launches,sum_ms,share,median_us,max_us")
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v = sorted(v)
    print(f"{name},{len(v)},{sum(v) / 1000:.3f},{sum(v) / tot:.4f},{v[len(v) // 2]:.1f},{v[-1]:.1f}")
