"""Instruction mix of one kernel from an `ncu --page source --csv` dump: python tools/ncu_mix.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
    if not m:
        continue
    op = m.group(2)
    n = int(r[ie])
    ops[op] += n
    samp[op] += int(r[isamp])
    tot += n
print("total warp instructions", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {n:11d} {100 * n / tot:5.1f}%  stall samples {samp[op]}")
