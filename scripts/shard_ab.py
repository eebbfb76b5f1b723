#!/usr/bin/env python
"""One-GPU experiment: cost of one rank's share of a strong-scaled batch, contiguous slice vs spatial shard."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import PointCloudIndex, synth
n_pts, n_q, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
stream = torch.cuda.current_stream().cuda_stream
base, half = synth.forest_cloud(min(5_000_000, n_pts), seed=3, variant="L", return_half=True)
tb = torch.from_numpy(base).to(dev)
tiles = -(-n_pts // len(base)); side = int(np.ceil(np.sqrt(tiles)))
g = torch.Generator(device=dev).manual_seed(7)
parts = [tb + torch.tensor([(t % side) * 2 * half, (t // side) * 2 * half, 0.0], device=dev) + (torch.rand(tb.shape, device=dev, generator=g) - 0.5) * 0.1 for t in range(tiles)]
t_pts = torch.cat(parts)[:n_pts].contiguous(); del parts
ext = torch.tensor([2 * half * side, 2 * half * side, 3.4], device=dev); lo = torch.tensor([-half, -half, 0.6], device=dev)
q = (torch.rand((n_q, 3), device=dev, generator=g) * ext + lo).contiguous()
ix = PointCloudIndex(max_points=n_pts, device=0, stream=stream)
ix.build(t_pts)
ix.profile(True)
def run(qq, label):
    ts = []
    for _ in range(3):
        ix.nearest(qq); ts.append(ix.last_batch_ms())
    a, b = ts[-1]
    print(f"{label:34s} order {a:8.3f} ms  search {b:8.3f} ms  total {a + b:8.3f}", flush=True)
    return a + b
T = run(q, f"full batch ({n_q})")
run(q[: n_q // G].contiguous(), f"contiguous 1/{G} slice")
for r in range(G):
    ix.batch_shard(r, G)
    run(q, f"spatial share rank {r}/{G}")
ix.batch_shard(0, 1)
print(f"ideal 1/{G} of full: {T / G:.3f} ms")
# the same shares timed the way bench_configs.py does (CUDA events around the whole Python call, incl. output allocation)
for r in (0, G // 2):
    ix.batch_shard(r, G)
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ix.nearest(q); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"spatial share rank {r}/{G}: whole call {np.median(ts[1:]):.3f} ms (events around ix.nearest)", flush=True)
ix.batch_shard(0, 1)
