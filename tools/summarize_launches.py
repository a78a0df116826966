"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].replace("<unnamed>::", "").split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(t for _, t in agg.values())
print(f"| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {t/1e3:.1f} | {t/n/1e3:.2f} | {100*t/tot:.1f}% |")
print(f"| total | {sum(n for n,_ in agg.values())} | {tot/1e3:.1f} | | |")
