"""Aggregate an ncu CSV launch list with gpu__time_duration.sum + dram bytes per kernel name (cold-cache, serialised:
compare SHARES and bytes, not absolute times)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    lid, name, metric, val = r[0], r[4], r[-3], float(r[-1].replace(",", ""))
    per.setdefault(lid, {"name": name.split("(")[0][:70]})[metric] = val
agg = collections.OrderedDict()
for v in per.values():
    a = agg.setdefault(v["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += v.get("gpu__time_duration.sum", 0.0); a[2] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(per)} launches, {tot / 1e6:.1f} ms of kernel time (per-launch times are cold-cache and serialised: compare SHARES)")
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:72s} n={c:5d} total={t / 1e6:8.2f} ms {100 * t / tot:5.1f}%  avg={t / c / 1e3:8.1f} us  dram/launch={b / c / 1e6:8.2f} MB  dram GB/s={b / max(t, 1):7.0f}")
