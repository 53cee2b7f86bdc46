"""A/B of bc_conv_igemm dispatch variants on the characteristic SwiftNet shapes (experiments behind the selection
rules in csrc/bc_conv.cu): runs tools/conv_bench.py in sub-processes under different BC_CONV_* / BC_SPLIT_* settings.
usage: python tools/conv_variants.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [
    {},
    {"BC_CONV_NTILE": "64"},
    {"BC_CONV_NTILE": "64", "BC_SPLIT_TARGET": "148"},
    {"BC_SPLIT_TARGET": "148"},
    {"BC_SPLIT_TARGET": "100"},
    {"BC_SPLIT_TARGET": "80"},
    {"SPLIT": "0"},
    {"BC_CONV_PERSIST": "2"},
]
for v in VARIANTS:
    env = dict(os.environ, **v)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "conv_bench.py")], capture_output=True, text=True,
                       env=env, cwd=ROOT, timeout=600)
    print(v, (r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1], flush=True)
