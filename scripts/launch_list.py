"""Aggregate the per-launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file raw.csv <command>`
into kernel, block, grid, launches, total_ns, share.  usage: launch_list.py raw.csv out.csv"""
import collections, csv, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    key = (name, r[ix["Block Size"]], r[ix["Grid Size"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values()) or 1.0
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "block", "grid", "launches", "total_ns", "share"])
    for (name, blk, grd), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, blk, grd, n, int(ns), round(ns / total, 4)])
print(open(sys.argv[2]).read())
