import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, torch
sys.path.insert(0, "."); sys.path.insert(0, "blockcopy-video-processing-pytorch_b200")
from blockcopy import _C
dev = torch.device("cuda", 0)
def run(C, H, W, BS, E, G, sets, reps=100):
    g = torch.Generator().manual_seed(0)
    cells = torch.randperm(G, generator=g)[:E].sort().values.to(torch.int32).to(dev)
    fmt = torch.channels_last
    planes = [torch.randn(1, C, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    padded = [torch.empty(E, C, BS + 2, BS + 2, device=dev, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    tiles = [torch.empty(E, C, BS, BS, device=dev, dtype=torch.float16).contiguous(memory_format=fmt) for _ in range(sets)]
    res = {}
    for name, fn, nbytes in (("gather_halo", lambda i: _C.gather_halo(padded[i % sets], planes[i % sets], cells, E, BS, 1), 2 * E * C * (BS + 2) ** 2 * 2),
                             ("gather", lambda i: _C.gather(tiles[i % sets], planes[i % sets], cells, E), 2 * E * C * BS * BS * 2),
                             ("scatter", lambda i: _C.scatter(tiles[i % sets], planes[i % sets], cells, E), 2 * E * C * BS * BS * 2)):
        for i in range(sets): fn(i)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(reps): fn(i)
        gr.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        res[name] = round(nbytes / us * 1e-3)
    return res
print(json.dumps({"cfg2": run(128, 256, 512, 32, 38, 128, 8), "cfg5": run(128, 512, 1024, 32, 154, 512, 2)}))
'''
for box in (8, 16, 32):
    for stages in (4, 6):
        for ctas in (1, 2, 4):
            env = dict(os.environ, BC_TMA_BOX_KB=str(box), BC_TMA_STAGES=str(stages), BC_TMA_CTAS_PER_SM=str(ctas))
            r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
            print(f"box={box}KB stages={stages} ctas/SM<={ctas}:", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
