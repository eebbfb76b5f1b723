#!/usr/bin/env python
"""Timing of pc_expand_batch (device-generated expansion batches) on the bench cloud: wall clock per call for several batch
sizes, pageable vs pinned output buffer.  Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel times.
    python scripts/expand_profile.py [--once K]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import PcRadiusParams, PcSampler, PointCloudIndex, synth  # noqa: E402
from pointcloudtraj_b200 import _lib as L  # noqa: E402


def main():
    once = int(sys.argv[sys.argv.index("--once") + 1]) if "--once" in sys.argv else 0
    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    start = (0.0, 0.0, 2.0)
    ix = PointCloudIndex(max_points=len(pts), device=0)
    ix.build(pts)
    nodes = PointCloudIndex(max_points=1 << 16, device=0)
    rng = np.random.default_rng(77)
    n_nodes = 4096
    nc = np.column_stack([rng.uniform(-25, 25, n_nodes), rng.uniform(-25, 25, n_nodes), rng.uniform(0.7, 4.0, n_nodes)])
    nr = rng.uniform(0.6, 1.25, n_nodes).astype(np.float32)
    nv = np.ones(n_nodes, np.uint8)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, start)
    smp = PcSampler.make(start, (0.8 * half, 0.5 * half, 2.0), (-half, half, -half, half, 0.0, 4.0), 30.0, 0.6, 0.3, 0.1)
    lib = ix._L
    ns = L.PcNodeSet(n_nodes, nc.ctypes.data, nr.ctypes.data, nv.ctypes.data)
    kmax = 10_000_000
    pinned = lib.pc_host_alloc(32 * kmax)
    pageable = np.empty(kmax, dtype=np.dtype(L.PC_CANDIDATE_DTYPE))
    pageable["radius"] = 0            # touch
    cnt, st = C.c_int64(0), C.c_uint32(0)

    def call(k, out_ptr):
        rc = lib.pc_expand_batch(ix._h, nodes._h, C.byref(ns), C.byref(smp), C.byref(P), 0.0, 0.6, k, C.c_void_p(out_ptr), k, C.byref(cnt), C.byref(st))
        assert rc == 0, lib.pc_last_error(ix._h)

    if once:
        call(once, pinned)
        call(once, pinned)
        print("once", once, cnt.value)
        return
    print(f"{'k':>10s} {'pageable ms':>12s} {'pinned ms':>10s} {'candidates':>11s}")
    for k in (512, 4096, 65536, 1_000_000, 10_000_000):
        row = []
        for ptr in (pageable.ctypes.data, pinned):
            for _ in range(3):
                call(k, ptr)
            reps = 20 if k <= 65536 else 5
            t0 = time.perf_counter()
            for _ in range(reps):
                call(k, ptr)
            row.append((time.perf_counter() - t0) / reps * 1e3)
        print(f"{k:10d} {row[0]:12.3f} {row[1]:10.3f} {cnt.value:11d}")
    lib.pc_host_free(pinned)


if __name__ == "__main__":
    main()
