#!/usr/bin/env python
"""Wall-clock latency of pc_clearance_batch (host buffers, blocking call) for small numbers of trajectories -- the planner
checks ONE trajectory per call (sim_planning_demo.cpp:729-781)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth  # noqa: E402

pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
ix = PointCloudIndex(max_points=len(pts))
ix.build(pts)
P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
print(f"{'traj':>6s} {'horizon':>8s} {'samples':>9s} {'us/call':>9s}")
for n in (1, 16, 256, 2000, 10000):
    tr = synth.bezier_trajectories(n, half * 0.9, seed=4)
    a = (tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"], P)
    for hz in (2.0, 20.0):
        for _ in range(3):
            fh, mr, ns = ix.clearance(*a, dt=0.02, horizon=hz)
        reps = 20 if n <= 2000 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            ix.clearance(*a, dt=0.02, horizon=hz)
        us = (time.perf_counter() - t0) / reps * 1e6
        print(f"{n:6d} {hz:8.1f} {int(ns.sum()):9d} {us:9.1f}")
ix.close()
