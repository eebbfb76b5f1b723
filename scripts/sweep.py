#!/usr/bin/env python
"""Sweep the query-kernel variants on the bench workload (run on the GPU box).

Builds libpcindex variants for several leaf sizes (nvcc -DPC_LEAF=n), then for every combination of
(leaf size, kernel variant, batch-ordering key width, refill threshold) times pc_radius_batch and
pc_nearest_batch on the C2 workload with CUDA events and checks that all variants return identical results.

    python scripts/sweep.py [--queries 10000000] [--leaves 2,4,8,16] > gpurun_out/sweep.txt
"""
import argparse
import ctypes as C
import itertools
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_variant(leaf, defs=""):
    tag = defs.replace("-D", "").replace("=", "").replace(" ", "_")
    out = os.path.join(ROOT, "gpurun_out", f"libpcindex_leaf{leaf}{tag}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", f"-DPC_LEAF={leaf}"] + defs.split() + ["-o", out,
           os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu"), "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--leaves", default="2,4,8,16")
    ap.add_argument("--kernels", default="1,2")
    ap.add_argument("--bits", default="16,24,32")
    ap.add_argument("--idle", default="4,8,16")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--defs", default="", help="extra nvcc -D flags, e.g. '-DPC_PACKET_ORDER=0'")
    args = ap.parse_args()
    import torch
    from pointcloudtraj_b200 import _lib, synth
    from pointcloudtraj_b200._lib import PcRadiusParams

    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    q = synth.rrt_queries(args.queries, half, seed=1000)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    t_pts = torch.from_numpy(pts).to(dev)
    t_q = torch.from_numpy(q).to(dev)
    M = len(q)
    t_r = torch.empty(M, dtype=torch.float32, device=dev)
    t_i = torch.empty(M, dtype=torch.int32, device=dev)
    t_d = torch.empty(M, dtype=torch.float32, device=dev)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
    ref_r = ref_i = None
    print(f"# {M} queries, 1M-point forest J; times in ms (median of {args.reps}); r = pc_radius_batch, n = pc_nearest_batch")
    print(f"{'leaf':>4s} {'kern':>4s} {'bits':>4s} {'idle':>4s} {'build':>7s} {'r_order':>8s} {'r_search':>9s} {'r_Gq/s':>7s} {'n_order':>8s} {'n_search':>9s} {'n_Gq/s':>7s} ok")
    for leaf in [int(v) for v in args.leaves.split(",")]:
        path = build_variant(leaf, args.defs)
        _lib._lib = None
        _lib.LIB_PATH = path
        L = _lib.load()
        for kern, bits in itertools.product([int(v) for v in args.kernels.split(",")], [int(v) for v in args.bits.split(",")]):
            idles = [int(v) for v in args.idle.split(",")] if kern == 2 else [0]
            for idle in idles:
                os.environ["PC_QUERY_KERNEL"] = str(kern)
                os.environ["PC_SORT_BITS"] = str(bits)
                os.environ["PC_MIN_IDLE"] = str(max(idle, 1))
                h = C.c_void_p()
                assert L.pc_index_create(C.byref(h), 0, len(pts), C.c_void_p(stream)) == 0
                assert L.pc_index_build(h, C.c_void_p(t_pts.data_ptr()), len(pts), 3, 1) == 0
                ms = C.c_float()
                L.pc_index_last_build_ms(h, C.byref(ms))
                L.pc_profile_enable(h, 1)
                res = {}
                for name in ("r", "n"):
                    order, search = [], []
                    for _ in range(args.reps + 1):
                        if name == "r":
                            rc = L.pc_radius_batch(h, C.c_void_p(t_q.data_ptr()), M, 3, 1, 0, C.byref(P), C.c_void_p(t_r.data_ptr()), None)
                        else:
                            rc = L.pc_nearest_batch(h, C.c_void_p(t_q.data_ptr()), M, 3, 1, 0, C.c_void_p(t_i.data_ptr()), C.c_void_p(t_d.data_ptr()))
                        assert rc == 0, L.pc_last_error(h)
                        a, b = C.c_float(), C.c_float()
                        L.pc_profile_last_batch(h, C.byref(a), C.byref(b))
                        order.append(a.value)
                        search.append(b.value)
                    res[name] = (float(np.median(order[1:])), float(np.median(search[1:])))
                torch.cuda.synchronize()
                if ref_r is None:
                    ref_r, ref_i = t_r.clone(), t_i.clone()
                ok = bool((t_r == ref_r).all().item() and (t_i == ref_i).all().item())
                L.pc_index_destroy(h)
                g = lambda o, s: M / ((o + s) * 1e-3) / 1e9
                print(f"{leaf:4d} {kern:4d} {bits:4d} {idle:4d} {ms.value:7.3f} {res['r'][0]:8.3f} {res['r'][1]:9.3f} {g(*res['r']):7.3f} "
                      f"{res['n'][0]:8.3f} {res['n'][1]:9.3f} {g(*res['n']):7.3f} {ok}", flush=True)


if __name__ == "__main__":
    main()
