#!/usr/bin/env python
"""Diagnostic: distribution of how many of a packet's 64 queries wanted each visited node (build with -DPC_STATS)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import _lib, synth, PcRadiusParams
out = os.path.join(ROOT, "gpurun_out", "libpcindex_stats.so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-DPC_STATS",
                "-o", out, os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu"), "-lcudart", "-ldl"], check=True, capture_output=True)
_lib.LIB_PATH = out
L = _lib.load()
os.environ["PC_QUERY_KERNEL"] = "4"
from pointcloudtraj_b200 import PointCloudIndex
pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
q = torch.from_numpy(synth.rrt_queries(10_000_000, half, seed=1000)).cuda()
ix = PointCloudIndex(max_points=len(pts))
ix.build(torch.from_numpy(pts).cuda())
P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
h = (C.c_ulonglong * 65)()
L.pc_stats_read(h, 1)
r = ix.radius(q, P)
torch.cuda.synchronize()
L.pc_stats_read(h, 0)
a = np.array(list(h), dtype=np.float64)
tot = a.sum()
print("visits", int(tot), "mean interested queries per visit", (a * np.arange(65)).sum() / tot)
cum = np.cumsum(a) / tot
for k in (0, 1, 2, 4, 8, 16, 32, 48, 63, 64):
    print(f"  <= {k:2d} interested: {cum[k] * 100:5.1f}% of visits")
print("share of (query, visit) pairs that were wanted:", (a * np.arange(65)).sum() / (tot * 64))
