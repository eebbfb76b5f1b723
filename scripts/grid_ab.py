#!/usr/bin/env python
"""Grid vs tree on the same batches (run on the GPU box): bounded radius queries answered by the voxel-grid ring search of
grid_kernels.cuh (PC_GRID=1, several cell sizes) and by the warp-packet walk of the prefix-split tree, on
  (a) the bench workload (1M-point forest, 10M uniform in-box samples, clean_demo parameters),
  (b) a free-space-heavy batch (the same map, samples between 4 and 8 m height: above most pillars),
  (c) the simulation.launch parameters (search_margin 0, max_radius 5: a 5 m bound) on a 2M-sample batch.
Prints search-kernel and ordering times (CUDA events inside the library) and checks that the results are identical."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pointcloudtraj_b200 import _lib, synth
    from pointcloudtraj_b200._lib import PcRadiusParams
    L = _lib.load()
    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    t_pts = torch.from_numpy(pts).to(dev)
    cases = []
    q = synth.rrt_queries(10_000_000, half, seed=1000)
    cases.append(("bench workload, bound 1.75 m", q, PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))))
    qf = synth.rrt_queries(10_000_000, half, seed=1001, z=(4.0, 8.0))
    cases.append(("free-space heavy (z in 4..8 m), bound 1.75 m", qf, PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))))
    cases.append(("simulation.launch params, bound 5 m, 2M samples", q[:2_000_000].copy(), PcRadiusParams.make(0.0, 5.0, 30.0, (0.0, 0.0, 2.0))))
    variants = [("tree (default)", {})] + [(f"grid, cell {c} m", {"PC_GRID": "1", "PC_GRID_CELL": str(c)}) for c in (0.3, 0.5, 0.9, 1.75)]
    print("# 1M-point forest J; ms per batch, median of 4; search = the kernel that answers the queries, order = curve ordering of the batch")
    for name, q, P in cases:
        t_q = torch.from_numpy(q).to(dev)
        M = len(q)
        t_r = torch.empty(M, dtype=torch.float32, device=dev)
        ref = None
        print(f"## {name}")
        for vname, env in variants:
            for k in ("PC_GRID", "PC_GRID_CELL"):
                os.environ.pop(k, None)
            os.environ.update(env)
            h = C.c_void_p()
            assert L.pc_index_create(C.byref(h), 0, len(pts), C.c_void_p(stream)) == 0
            assert L.pc_index_build(h, C.c_void_p(t_pts.data_ptr()), len(pts), 3, 1) == 0, L.pc_last_error(h)
            L.pc_profile_enable(h, 1)
            order, search = [], []
            for _ in range(5):
                assert L.pc_radius_batch(h, C.c_void_p(t_q.data_ptr()), M, 3, 1, 0, C.byref(P), C.c_void_p(t_r.data_ptr()), None) == 0
                a, b = C.c_float(), C.c_float()
                L.pc_profile_last_batch(h, C.byref(a), C.byref(b))
                order.append(a.value); search.append(b.value)
            torch.cuda.synchronize()
            if ref is None:
                ref = t_r.clone()
            same = bool((t_r == ref).all().item())
            L.pc_index_destroy(h)
            print(f"{vname:24s} order {np.median(order[1:]):7.3f}  search {np.median(search[1:]):8.3f}  identical {same}", flush=True)
        del t_q, t_r


if __name__ == "__main__":
    main()
