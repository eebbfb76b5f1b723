#!/usr/bin/env python
"""Runs each kernel family once on a representative batch (for `ncu -k regex:<kernel>` captures of the kernels the bench
does not reach): clearance (1M-point map, 2000 trajectories x up to 1000 samples), range (C1: 200k points, 100k queries,
r = 1 m), the small-batch lane-group kernel (50k unordered radius queries on the 1M map), and two ordered 10M-query radius
batches (ordering pass + packet kernel)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth  # noqa: E402

P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
ix = PointCloudIndex(max_points=len(pts))
ix.build(pts)
tr = synth.bezier_trajectories(2000, half * 0.9, seed=4)
for _ in range(2):
    fh, mr, ns = ix.clearance(tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"], P, dt=0.02, horizon=20.0)
print("clearance samples", int(ns.sum()))
q50 = synth.rrt_queries(50_000, half, seed=3)
for _ in range(2):
    r = ix.radius(q50, P)
q10 = synth.rrt_queries(10_000_000, half, seed=1000)
for _ in range(2):
    r = ix.radius(q10, P)
pts1, half1 = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
ix.build(pts1)
q1 = synth.rrt_queries(100_000, half1, seed=2)
for _ in range(2):
    off, idx = ix.range(q1, 1.0)
print("range hits", int(off[-1]))
ix.close()
