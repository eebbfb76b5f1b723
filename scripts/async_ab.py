#!/usr/bin/env python
"""A/B: K radius batches back to back, PC_DEVICE (one stream) vs PC_DEVICE_ASYNC (three internal streams)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
ix = PointCloudIndex(max_points=len(pts), stream=torch.cuda.current_stream().cuda_stream)
ix.build(torch.from_numpy(pts).to(dev))
qs = [torch.from_numpy(synth.rrt_queries(M, half, seed=1000 + k)).to(dev) for k in range(3)]
outs = [torch.empty(M, dtype=torch.float32, device=dev) for _ in range(3)]
P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
L = ix._L
ref = None
for space, name in ((1, "PC_DEVICE"), (3, "PC_DEVICE_ASYNC")):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        K = 12
        for k in range(K):
            assert L.pc_radius_batch(ix._h, C.c_void_p(qs[k % 3].data_ptr()), M, 3, space, 0, C.byref(P), C.c_void_p(outs[k % 3].data_ptr()), None) == 0
        ix.sync()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res = [o.clone() for o in outs]
    if ref is None: ref = res
    same = all(bool((a == b).all().item()) for a, b in zip(ref, res))
    print(f"{name:16s} {ms:7.3f} ms/batch  {M / ms / 1e6:6.3f} Gq/s  same={same}")
