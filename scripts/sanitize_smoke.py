#!/usr/bin/env python
"""Small pass over every kernel of libpcindex (for compute-sanitizer; one tool per gpurun call):
   python scripts/sanitize_smoke.py && compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import PC_QUERY_SORTED, PC_QUERY_UNSORTED, PC_RADIUS_FULL_NN, PcRadiusParams, PointCloudIndex, synth  # noqa: E402

ix = PointCloudIndex(max_points=0)
P = PcRadiusParams.make(0.25, 1.5, 12.0, (0.0, 0.0, 2.0))
for n in (0, 1, 5, 1000, 33_333):
    pts = synth.uniform_cloud(n, half=6.0, seed=n)
    q = synth.rrt_queries(3001, 8.0, seed=n + 1)
    ix.build(pts)
    for flags in (PC_QUERY_UNSORTED, PC_QUERY_SORTED):
        i, d = ix.nearest(q, flags=flags)
        r = ix.radius(q, P, flags=flags)
        r2, i2 = ix.radius(q, P, flags=flags | PC_RADIUS_FULL_NN, want_idx=True)
    off, lst = ix.range(q[:500], 0.8)
    tr = synth.bezier_trajectories(40, 5.0, seed=3)
    fh, mr, ns = ix.clearance(tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"], P, horizon=3.0)
    print(n, int((i >= 0).sum()), int(off[-1]), int(ns.sum()))
big = synth.forest_cloud(150_000, seed=2)
ix.build(big)
q = synth.rrt_queries(70_000, 16.0, seed=9)
i, d = ix.nearest(q)
r = ix.radius(q, P)
ix.close()
print("sanitize smoke done")
