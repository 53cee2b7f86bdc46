"""bench.conv_microbench at config-4 batch (8 streams, E = 320) under the conv dispatch switches."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
import bench
dev = torch.device("cuda", 0)
r = bench.conv_microbench(dev, bench.load_peaks(), reps=20, sets=2, E=int(sys.argv[1]), images=int(sys.argv[2]))
print({k.split("(")[1][:-1]: (round(v["us"], 1), round(v["frac_of_tensor_peak"], 3)) for k, v in r.items()})
'''
for E, images in ((40, 1), (320, 8), (160, 4), (80, 2)):
    for env in ({"BC_CONV_MULTICAST": "3"}, {"BC_CONV_MULTICAST": "1"}, {"BC_CONV_MULTICAST": "0"}):
        r = subprocess.run([sys.executable, "-c", CHILD, str(E), str(images)], capture_output=True, text=True,
                           env=dict(os.environ, **env), cwd=ROOT, timeout=600)
        print(f"E={E} {env}:", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
