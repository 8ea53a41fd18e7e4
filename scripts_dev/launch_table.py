"""launches.csv (ncu --metrics gpu__time_duration.sum --csv) -> markdown share table.  Usage: launch_table.py csv out.md "title" """
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", "")); u = r[mu]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
out = [f"# {sys.argv[3]}", "", "Cold-cache, serialised per-launch times: compare SHARES, not absolutes.", "", "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
out.append(f"| **total** | {n} | {tot:.1f} | | |")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
