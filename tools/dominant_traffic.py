"""profiles/r02_dominant_kernel_traffic.json from an ncu launch list (gpu__time_duration.sum + dram bytes per launch):
DRAM bytes per launch of gemm_split_kernel over the UNet launches of the captured steps; the K-means score GEMMs (the
launches directly followed by km_assign_tc_kernel) are excluded.  Usage: python tools/dominant_traffic.py launches.csv[.gz] out.json"""
import collections, csv, gzip, json, sys
src, dst = sys.argv[1], sys.argv[2]
fh = gzip.open(src, "rt") if src.endswith(".gz") else open(src)
per = collections.OrderedDict()
for r in csv.reader(fh):
    if len(r) > 10 and r[0].isdigit():
        per.setdefault(r[0], {"name": r[4]})[r[-3]] = float(r[-1].replace(",", ""))
launches = list(per.values())
tot_b, tot_t, n = 0.0, 0.0, 0
for i, v in enumerate(launches):
    if "gemm_split_kernel" not in v["name"]:
        continue
    if i + 1 < len(launches) and "km_assign_tc_kernel" in launches[i + 1]["name"]:
        continue
    tot_b += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
    tot_t += v.get("gpu__time_duration.sum", 0.0)
    n += 1
out = {"workload": "c2", "kernel": "gemm_split_kernel", "dram_bytes_per_launch": tot_b / max(n, 1), "launches": n,
       "avg_launch_us_under_ncu": tot_t / max(n, 1) / 1e3,
       "source": f"{src}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                 "eager steps of `bench.py --workload c2`; all gemm_split_kernel launches of the UNet (Linear + implicit-GEMM conv), "
                 "the K-means score GEMMs (the launches followed by km_assign_tc_kernel) excluded"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
