#!/usr/bin/env python
"""A/B of compile-time variants of libpcindex on the bench workload.

    python scripts/variants_ab.py build "<name>=<nvcc -D flags>" ...     # HERE (no GPU): cross-compiles build/variants/lib_<name>.so
    python scripts/variants_ab.py run [--queries N]                      # on the GPU box: times every built variant

`run` measures, per variant: index build (1M points, 300k points), ordering pass and search kernel of a radius batch and of an
unbounded nearest batch (CUDA events inside the library, median of 5), and checks that all variants return identical results.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, defs = spec.partition("=")
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared"] + defs.split() + ["-o", out, os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu"),
                                                                   "-lcudart", "-ldl"]
        subprocess.run(cmd, check=True)
        print("built", out, defs)


def run(queries=10_000_000):
    import torch
    from pointcloudtraj_b200 import _lib, synth
    from pointcloudtraj_b200._lib import PcRadiusParams
    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    q = synth.rrt_queries(queries, half, seed=1000)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    t_pts, t_q = torch.from_numpy(pts).to(dev), torch.from_numpy(q).to(dev)
    t_frame = t_pts[:300_000].contiguous()
    M = len(q)
    t_r = torch.empty(M, dtype=torch.float32, device=dev)
    t_i = torch.empty(M, dtype=torch.int32, device=dev)
    t_d = torch.empty(M, dtype=torch.float32, device=dev)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
    ref = None
    print(f"# {M} queries, 1M-point forest J; ms, median of 5")
    print(f"{'variant':28s} {'build1M':>8s} {'build300k':>9s} {'r_order':>8s} {'r_search':>9s} {'n_order':>8s} {'n_search':>9s} same")
    libs = [("default", os.path.join(ROOT, "pointcloudtraj_b200", "libpcindex.so"))] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(VDIR, "lib_*.so")))]
    for name, path in libs:
        _lib._lib = None
        _lib.LIB_PATH = path
        L = _lib.load()
        h = C.c_void_p()
        assert L.pc_index_create(C.byref(h), 0, len(pts), C.c_void_p(stream)) == 0
        ms = C.c_float()
        bf, b1 = [], []
        for _ in range(6):
            assert L.pc_index_build(h, C.c_void_p(t_frame.data_ptr()), 300_000, 3, 1) == 0
            L.pc_index_last_build_ms(h, C.byref(ms)); bf.append(ms.value)
        for _ in range(6):
            assert L.pc_index_build(h, C.c_void_p(t_pts.data_ptr()), len(pts), 3, 1) == 0
            L.pc_index_last_build_ms(h, C.byref(ms)); b1.append(ms.value)
        L.pc_profile_enable(h, 1)
        res, detail = {}, {}
        for kind in ("r", "n"):
            order, search = [], []
            for _ in range(6):
                if kind == "r":
                    rc = L.pc_radius_batch(h, C.c_void_p(t_q.data_ptr()), M, 3, 1, 0, C.byref(P), C.c_void_p(t_r.data_ptr()), None)
                else:
                    rc = L.pc_nearest_batch(h, C.c_void_p(t_q.data_ptr()), M, 3, 1, 0, C.c_void_p(t_i.data_ptr()), C.c_void_p(t_d.data_ptr()))
                assert rc == 0, L.pc_last_error(h)
                a, b = C.c_float(), C.c_float()
                L.pc_profile_last_batch(h, C.byref(a), C.byref(b))
                order.append(a.value); search.append(b.value)
                if hasattr(L, "pc_profile_last_order_detail"):
                    d3 = (C.c_float * 3)()
                    if L.pc_profile_last_order_detail(h, d3) == 0:
                        detail[kind] = tuple(d3)
            res[kind] = (float(np.median(order[1:])), float(np.median(search[1:])))
        torch.cuda.synchronize()
        if ref is None:
            ref = (t_r.clone(), t_i.clone(), t_d.clone())
        same = bool((t_r == ref[0]).all().item() and (t_i == ref[1]).all().item() and (t_d == ref[2]).all().item())
        L.pc_index_destroy(h)
        print(f"{name:28s} {np.median(b1[1:]):8.3f} {np.median(bf[1:]):9.3f} {res['r'][0]:8.3f} {res['r'][1]:9.3f} {res['n'][0]:8.3f} {res['n'][1]:9.3f} {same}"
              + "".join(f"  {k}: clear {v[0]:.3f} keys {v[1]:.3f} sort {v[2]:.3f}" for k, v in detail.items()), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[2] == "--queries" else 10_000_000)
