#!/usr/bin/env python
"""Where do ordered warp packets start to pay?  Radius and nearest batches of 100 k .. 10 M queries on the 1M-point map:
one thread per query on the unordered batch (PC_QUERY_UNSORTED) vs ordering pass + packet kernels (PC_QUERY_SORTED).
Device buffers, CUDA events around the whole call, median of 8."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
stream = torch.cuda.current_stream().cuda_stream
pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
t_pts = torch.from_numpy(pts).to(dev)
P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
os.environ["PC_COOP_MAX_BATCH"] = "0"
ix = PointCloudIndex(max_points=len(pts), stream=stream)
ix.build(t_pts)
L = ix._L
print(f"{'m':>9s} {'kind':>8s} {'thread/query unsorted':>22s} {'ordered packets':>16s}   ms")
for m in (100_000, 200_000, 400_000, 800_000, 1_600_000, 3_200_000, 10_000_000):
    q = torch.from_numpy(synth.rrt_queries(m, half, seed=5)).to(dev)
    out = torch.empty(m, dtype=torch.float32, device=dev)
    oi = torch.empty(m, dtype=torch.int32, device=dev)
    for kind in ("radius", "nearest"):
        row = []
        for flags in (2, 4):
            ts = []
            for _ in range(11):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if kind == "radius":
                    L.pc_radius_batch(ix._h, C.c_void_p(q.data_ptr()), m, 3, 1, flags, C.byref(P), C.c_void_p(out.data_ptr()), None)
                else:
                    L.pc_nearest_batch(ix._h, C.c_void_p(q.data_ptr()), m, 3, 1, flags, C.c_void_p(oi.data_ptr()), C.c_void_p(out.data_ptr()))
                e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            row.append(float(np.median(ts[3:])))
        print(f"{m:9d} {kind:>8s} {row[0]:22.4f} {row[1]:16.4f}", flush=True)
ix.close()
