"""rl_semseg (bench.bench_rl): policy inference / training frames on the native kernels vs torch (cuDNN graph replays)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.argv = ["bench.py"]
sys.path.insert(0, "."); sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
import bench
args = bench.parse_args()
r = bench.bench_rl(args, torch.device("cuda", 0))
r.pop("windows_ms", None); r.pop("note", None)
print(json.dumps(r))
'''
for fused, flag in (("0", "0"), ("1", "0"), ("1", "1"), ("0", "0"), ("1", "0"), ("1", "1")):
    env = dict(os.environ, BLOCKCOPY_POLICY_FUSED=fused, BLOCKCOPY_POLICY_FUSED_TRAINING=flag)
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    print("fused_inference =", fused, "fused_training =", flag, (r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip()[-400:] if r.returncode else "", flush=True)
