#!/usr/bin/env python
"""Per-kernel counts of the tcgen05 / TMA SASS mnemonics in libblockcopy_sm100.so -> markdown (profiles/rNN_sass_excerpt.md).
usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200", "blockcopy", "_lib", "libblockcopy_sm100.so")
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
rows = []
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    c = lambda pat: len(re.findall(pat, part))  # noqa: E731
    rows.append((name, c(r"UTCHMMA"), c(r"LDTM"), c(r"UTMALDG"), c(r"UTMASTG"), c(r"UTCBAR"), c(r"SYNCS\."), len(part.splitlines())))
print("# SASS evidence per tensor-core / TMA kernel of libblockcopy_sm100.so (sm_100a)\n")
print("`cuobjdump -sass blockcopy/_lib/libblockcopy_sm100.so`, instruction counts per kernel (`tools/sass_excerpt.py`).  `UTCHMMA` ="
      " tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` / `UTMASTG` = cp.async.bulk.tensor load / "
      "store (TMA), `UTCBAR` = tcgen05.commit -> mbarrier, `SYNCS` = mbarrier operations.\n")
print("| kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG | UTCBAR | SYNCS | SASS lines |\n|---|---:|---:|---:|---:|---:|---:|---:|")
tot = collections.Counter()
for r in sorted(rows, key=lambda r: -r[1] - r[3]):
    if r[1] + r[2] + r[3] + r[4] == 0:
        continue
    n = subprocess.run(["c++filt", r[0]], capture_output=True, text=True).stdout.strip()
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    print(f"| `{n}` | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | {r[6]} | {r[7]} |")
    for k, v in zip("abcde", r[1:6]):
        tot[k] += v
print(f"| **total** | {tot['a']} | {tot['b']} | {tot['c']} | {tot['d']} | {tot['e']} | | |")
