"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    n = r[4].split("(")[0][:64]; t = float(r[-1].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot / 1e6:.2f} ms of kernel time")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:66s} n={c:5d} total={t / 1e3:10.1f} us avg={t / c / 1e3:8.1f} us {100 * t / tot:5.1f}%")
