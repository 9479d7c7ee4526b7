"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by kernel:
python tools/launch_summary.py gpurun_out/rNN_launches.csv [--all]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
iname, imet, ival, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(int(r[iid]), {"name": r[iname], "grid": r[ig], "block": r[ib]})[r[imet]] = float(r[ival].replace(",", ""))


def short(n):
    n = re.sub(r"\(.*", "", n).replace("<unnamed>::", "").replace("void ", "")
    return re.sub(r"<.*", "", n)


agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for k, v in d.items():
    a = agg[short(v["name"])]
    a[0] += v["gpu__time_duration.sum"]
    a[1] += 1
    a[2] += v.get("dram__bytes_read.sum", 0)
    a[3] += v.get("dram__bytes_write.sum", 0)
tot = sum(a[0] for a in agg.values())
print(f"{len(d)} launches, {tot / 1e6:.3f} ms of kernel time (ncu: serialised, cold caches)")
print(f"{'kernel':34s} {'n':>4s} {'us':>9s} {'share':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:34s} {a[1]:4d} {a[0] / 1e3:9.1f} {a[0] / tot * 100:5.1f}% {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f}")
if "--all" in sys.argv:
    for k, v in d.items():
        print(k, f"{short(v['name']):28s} {v['grid']:>16s} {v['block']:>12s} {v['gpu__time_duration.sum'] / 1e3:8.1f} us  "
                 f"rd {v.get('dram__bytes_read.sum', 0) / 1e6:7.1f} wr {v.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB")
