#!/usr/bin/env python
"""C1's range batch twice (200k points, 100k queries, r = 1 m) for `ncu -k regex:pc_range` captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import PointCloudIndex, synth  # noqa: E402

pts1, half1 = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
ix = PointCloudIndex(max_points=len(pts1))
ix.build(pts1)
q1 = synth.rrt_queries(100_000, half1, seed=2)
for _ in range(2):
    off, idx = ix.range(q1, 1.0)
print("range hits", int(off[-1]))
ix.close()
