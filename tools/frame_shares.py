#!/usr/bin/env python
"""Launch list of one steady frame (ncu --metrics gpu__time_duration.sum CSV) -> markdown share table + ordered list.
usage: python tools/frame_shares.py launches.csv [frame_index]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if "compact_mask" in r["Kernel Name"]]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frame = rows[idx[k]:idx[k + 1]]


def us(r):
    return float(r["Metric Value"].replace(",", "")) / (1000 if r["Metric Unit"] == "ns" else 1)


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+(<[0-9, a-z]+>)?)", n)
    return (m.group(1) if m else n)[:70]


tot = sum(us(r) for r in frame)
agg = collections.OrderedDict()
for r in frame:
    a = agg.setdefault(short(r["Kernel Name"]), [0.0, 0])
    a[0] += us(r)
    a[1] += 1
OURS = ("bc::", "conv_igemm", "conv_stem", "ew_fused", "head_", "tma_move", "maxpool_halo", "spp_", "stem_pack", "compact_mask",
        "scatter_kernel", "gather_kernel", "copy_blocks", "policy_features", "info_gain")
ours = sum(v[0] for n, v in agg.items() if any(n.startswith(o) or ("bc::" + o) in n for o in OURS))
print(f"frame total {tot:.0f} us over {len(frame)} kernels (ncu, serialised, cold cache); "
      f"this repo's kernels (`bc::*`) = {100 * ours / tot:.0f} % of it\n")
print("| share | us/frame | launches | kernel |\n|---:|---:|---:|---|")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {100 * t / tot:.1f} % | {t:.1f} | {c} | `{n}` |")
print("\nIn launch order:\n\n| us | grid | block | kernel |\n|---:|---|---|---|")
for r in frame:
    print(f"| {us(r):.2f} | {r['Grid Size']} | {r['Block Size']} | `{short(r['Kernel Name'])}` |")
