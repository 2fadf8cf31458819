"""Print every launch of an ncu gpu__time_duration list (name, grid, us) and per-kernel totals."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
seq = []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        name = d["Kernel Name"].split("(")[0].replace("void ", "")[:40]
        seq.append((name, d.get("Grid Size", ""), v))
        agg.setdefault(name, []).append(v)
for name, grid, v in seq:
    if v > 15:
        print(f"{name:42s} {grid:18s} {v:8.1f}")
tot = sum(sum(v) for v in agg.values())
print("---- totals", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:42s} n={len(v):3d} sum={sum(v):8.1f} max={max(v):7.1f} {sum(v)/tot:.3f}")
