"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count,
total and share (cold-cache, serialised: compare SHARES, not absolutes)."""
import csv, sys, collections, re
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # launches to skip (warm-up)
take = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v_us))
rows = rows[skip:skip + take]
agg = collections.OrderedDict()
for n, v in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, total {tot/1e3:.2f} ms")
print(f"{'kernel':70s} {'n':>6s} {'ms':>10s} {'share':>7s}")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:6d} {v/1e3:10.3f} {100*v/tot:6.1f}%")
