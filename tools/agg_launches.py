"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, io, sys
def agg(path, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        tot[name][0] += 1; tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"{path}: {sum(v[0] for v in tot.values())} launches, {total/1e3:.2f} ms of kernel time")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {k:70s} n={v[0]:5d} total={v[1]/1e3:9.3f} ms avg={v[1]/v[0]:9.1f} us {100*v[1]/total:5.1f}%")
if __name__ == "__main__":
    for p in sys.argv[1:]:
        agg(p)
