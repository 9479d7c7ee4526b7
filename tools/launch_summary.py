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

# --traffic-json PATH: DRAM bytes (read + write) per CALL of each C-ABI entry point, for bench.py's roofline.traffic.
# A call of ick_mha_bwd is one dK/dV launch plus either a recomputing dQ launch or rowdot + dQ-from-dS; a call of
# ick_wgrad_group_tc one grouped wgrad + one reduce, ...  (counting kernel, kernels of the family, entry points that share them)
FAMILY = {
    "ick_mha_bwd": ("bwd_tc_kernel", ["bwd_tc_kernel", "tb_rowdot_kernel", "bwd_fused_kernel", "fb_rowdot_kernel", "bwd_dq_pkernel", "bwd_dkv_pkernel",
                                       "bwd_dq_ds_kernel", "rowdot_kernel", "bwd_dq_kernel", "bwd_dkv_kernel"], []),
    "ick_mha_fwd": ("fwd_pkernel", ["fwd_pkernel", "fwd_kernel", "fwd_tc_kernel"], []),
    "ick_gemm_tn_tc": ("gemm_tn_tc_kernel", ["gemm_tn_tc_kernel"], ["ick_gemm_tn_tc_dual"]),
    "ick_wgrad_group_tc": ("wgrad_group_tc_kernel", ["wgrad_group_tc_kernel", "wgrad_group_reduce_kernel"], []),
    "ick_wgrad_tc": ("wgrad_tc_kernel", ["wgrad_tc_kernel", "wgrad_reduce_kernel", "bias_grad_kernel"], []),
    "ick_add_ln_fwd": ("add_ln_fwd_fast_kernel", ["add_ln_fwd_fast_kernel", "add_ln_fwd_kernel"], ["ick_add_ln_fwd_dual"]),
    "ick_add_ln_bwd": ("add_ln_bwd_fast_kernel", ["add_ln_bwd_fast_kernel", "add_ln_bwd_kernel"], ["ick_add_ln_bwd_dual"]),
    "ick_adam_step": ("adam_kernel", ["adam_kernel"], []),
    "ick_ce_fwd_bwd": ("ce_kernel", ["ce_kernel"], []),
    "ick_pointer_bwd": ("pointer_bwd_mma_kernel", ["pointer_bwd_mma_kernel"], []),
    "ick_pointer_fwd": ("pointer_fwd_mma_kernel", ["pointer_fwd_mma_kernel"], []),
}
if "--traffic-json" in sys.argv:
    import json

    out = {}
    for fam, (counter, kernels, aliases) in FAMILY.items():
        if counter not in agg:
            continue
        calls = agg[counter][1]
        present = [k for k in kernels if k in agg]
        rec = {"dram_bytes_per_call": sum(agg[k][2] + agg[k][3] for k in present) / calls, "calls": calls,
               "kernel_us_per_call": sum(agg[k][0] for k in present) / 1e3 / calls}
        out[fam] = rec
        for a in aliases:
            out[a] = dict(rec, note=f"same kernels as {fam}: average over both entry points")
    json.dump({"source": sys.argv[1], "per_call": out}, open(sys.argv[sys.argv.index("--traffic-json") + 1], "w"), indent=1)
