#!/usr/bin/env python
"""A/B of libpcindex variants (build/variants/lib_*.so, see scripts/variants_ab.py build) on the range workloads:
C1 (200 k points, 100 k queries, r = 1 m) and 50 k queries on the 1 M-point map; device buffers, CUDA events, median of 7."""
import ctypes as C
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pointcloudtraj_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    cases = []
    for n_pts, n_q, seed in ((200_000, 100_000, 6), (1_000_000, 50_000, 1)):
        pts, half = synth.forest_cloud(n_pts, seed=seed, variant="J", return_half=True)
        q = synth.rrt_queries(n_q, half, seed=seed + 1)
        cases.append((torch.from_numpy(pts).to(dev), torch.from_numpy(q).to(dev)))
    libs = [("default", os.path.join(ROOT, "pointcloudtraj_b200", "libpcindex.so"))] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))]
    print(f"{'variant':24s} {'c1 ms':>8s} {'1M/50k ms':>10s} same")
    ref = None
    for name, path in libs:
        _lib._lib = None
        _lib.LIB_PATH = path
        L = _lib.load()
        row, outs = [], []
        for t_pts, t_q in cases:
            h = C.c_void_p()
            assert L.pc_index_create(C.byref(h), 0, t_pts.shape[0], C.c_void_p(stream)) == 0
            assert L.pc_index_build(h, C.c_void_p(t_pts.data_ptr()), t_pts.shape[0], 3, 1) == 0
            m = t_q.shape[0]
            t_off = torch.empty(m + 1, dtype=torch.int64, device=dev)
            t_r = torch.ones(1, dtype=torch.float64, device=dev)
            assert L.pc_range_batch(h, C.c_void_p(t_q.data_ptr()), m, 3, 1, C.c_void_p(t_r.data_ptr()), 1, C.c_void_p(t_off.data_ptr()), None, 0) == 0
            torch.cuda.synchronize()
            total = int(t_off[-1].item())
            t_lst = torch.empty(total, dtype=torch.int32, device=dev)
            ms = []
            for _ in range(9):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                assert L.pc_range_batch(h, C.c_void_p(t_q.data_ptr()), m, 3, 1, C.c_void_p(t_r.data_ptr()), 1, C.c_void_p(t_off.data_ptr()), C.c_void_p(t_lst.data_ptr()), total) == 0
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            row.append(float(np.median(ms[2:])))
            outs.append((t_off.cpu().numpy().copy(), t_lst.cpu().numpy().copy()))
            L.pc_index_destroy(h)
        same = True
        if ref is None:
            ref = outs
        else:
            same = all((a[0] == b[0]).all() and (a[1] == b[1]).all() for a, b in zip(ref, outs))
        print(f"{name:24s} {row[0]:8.3f} {row[1]:10.3f} {same}")


if __name__ == "__main__":
    main()
