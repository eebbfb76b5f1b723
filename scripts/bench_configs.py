#!/usr/bin/env python
"""Measure the BASELINE.json configurations other than the headline one (bench.py covers C2) -- run on the GPU box.

    python scripts/bench_configs.py c1            # 200k-point forest, 100k RRT* samples (the CPU-runnable case)
    python scripts/bench_configs.py c2small       # C2's own batch: 1M-point map, 50k samples (latency of one planner batch)
    python scripts/bench_configs.py c3            # LiDAR stream: 300k-point frames, rebuild + 1M radius queries per frame
    python scripts/bench_configs.py c4            # 10k trajectories x 1k samples against a 5M-point cloud
    torchrun --nproc-per-node N scripts/bench_configs.py c5 [--points 100000000 --queries 100000000]
                                                  # replicated index (one ncclBroadcast), queries sharded over N GPUs (strong)

Every config prints one JSON line with device-side timings (CUDA events) and a parity spot check against an fp64
brute force in the reference's operation order (torch on the GPU; independent of the library).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pointcloudtraj_b200 import PC_RADIUS_FULL_NN, PcRadiusParams, PointCloudIndex, synth  # noqa: E402

CLEAN = dict(search_margin=0.25, max_radius=1.5, sample_range=30.0)


def brute_check(t_pts, t_q, idx, d2, n_check=64):
    """fp64 brute force (reference operation order) for the first n_check queries: index and float32 d2 must match."""
    P = t_pts[:, :3].double()
    ok = True
    for k in range(min(n_check, t_q.shape[0])):
        qk = t_q[k, :3].double()
        dx, dy, dz = P[:, 0] - qk[0], P[:, 1] - qk[1], P[:, 2] - qk[2]
        e = (dx * dx + dy * dy) + dz * dz
        m = e.min()
        first = int(torch.nonzero(e == m)[0])            # lowest index among exact minimisers
        ok = ok and first == int(idx[k]) and float(m.float()) == float(d2[k])
    return ok


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def setup_stream(dev):
    torch.cuda.set_device(dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    return torch.cuda.current_stream().cuda_stream


def c1_like(name, n_pts, n_q, seed):
    dev = torch.device("cuda", 0)
    stream = setup_stream(dev)
    pts, half = synth.forest_cloud(n_pts, seed=seed, variant="J", return_half=True)
    q = synth.rrt_queries(n_q, half, seed=seed + 1)
    t_pts, t_q = torch.from_numpy(pts).to(dev), torch.from_numpy(q).to(dev)
    ix = PointCloudIndex(max_points=n_pts, device=0, stream=stream)
    build_ms, _ = timed(lambda: ix.build(t_pts))
    P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN)
    nn_ms, _ = timed(lambda: ix.nearest(t_q))
    rad_ms, _ = timed(lambda: ix.radius(t_q, P))
    idx, d2 = ix.nearest(t_q)
    torch.cuda.synchronize()
    # host path (what a planner with host buffers sees), wall clock
    ix.radius(q, P)                      # first call: the lane staging buffers are allocated
    t0 = time.perf_counter()
    for _ in range(5):
        ix.radius(q, P)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    # fixed-radius range queries (kd_nearest_range3), r = 1 m, host buffers, wall clock incl. both passes and copies
    ix.range(q, 1.0)                     # first call allocates scratch and the list staging buffer
    t0 = time.perf_counter()
    off, lst = ix.range(q, 1.0)
    range_ms = (time.perf_counter() - t0) * 1e3
    # the same with device buffers (queries, offsets and lists stay on the GPU), CUDA events
    t_off = torch.empty(n_q + 1, dtype=torch.int64, device=dev)
    t_lst = torch.empty(max(int(off[-1]), 1), dtype=torch.int32, device=dev)
    t_r = torch.ones(1, dtype=torch.float64, device=dev)       # PC_DEVICE: the range array is a device pointer too

    def range_dev():
        rc = ix._L.pc_range_batch(ix._h, C.c_void_p(t_q.data_ptr()), n_q, 3, 1, C.c_void_p(t_r.data_ptr()), 1, C.c_void_p(t_off.data_ptr()),
                                  C.c_void_p(t_lst.data_ptr()), t_lst.numel())
        assert rc == 0
    range_dev_ms, _ = timed(range_dev)
    same = bool((t_off.cpu().numpy() == off).all() and (t_lst.cpu().numpy()[: int(off[-1])] == lst).all())
    print(json.dumps({"config": name, "points": n_pts, "queries": n_q, "index_build_ms": build_ms,
                      "range_r1_ms_device_buffers": range_dev_ms, "range_r1_device_qps": n_q / range_dev_ms * 1e3,
                      "range_device_matches_host": same, "range_kernel": os.environ.get("PC_QUERY_KERNEL", "default (warp per query)"),
                      "range_r1_ms_host_buffers": range_ms, "range_r1_qps": n_q / range_ms * 1e3, "range_r1_mean_hits": float(off[-1]) / n_q,
                      "nearest_ms": nn_ms, "nearest_qps": n_q / nn_ms * 1e3, "radius_ms": rad_ms, "radius_qps": n_q / rad_ms * 1e3,
                      "radius_host_buffers_ms": host_ms, "parity_spot_check": brute_check(t_pts, t_q, idx, d2)}))
    ix.close()


def c3():
    """10 Hz LiDAR stream: every frame = the points within 20 m of a sensor moving 0.3 m/frame through a larger map,
    padded/truncated to 300k points; full index rebuild + 1M radius queries per frame."""
    dev = torch.device("cuda", 0)
    stream = setup_stream(dev)
    pts, half = synth.forest_cloud(1_500_000, seed=4, variant="J", return_half=True)
    t_all = torch.from_numpy(pts).to(dev)
    n_frame, n_q, frames = 300_000, 1_000_000, 12
    ix = PointCloudIndex(max_points=n_frame, device=0, stream=stream)
    P0 = dict(CLEAN)
    per_frame, builds, queries, ok = [], [], [], True
    for f in range(frames):
        sensor = torch.tensor([-10.0 + 0.3 * f, -5.0, 2.0], device=dev)
        d = ((t_all - sensor) ** 2).sum(1)
        near = torch.nonzero(d <= 20.0 ** 2)[:, 0]
        if near.numel() >= n_frame:
            sel = near[:n_frame]
        else:                                            # pad with the next nearest points
            sel = torch.argsort(d)[:n_frame]
        frame = t_all[sel].contiguous()
        q = torch.from_numpy(synth.rrt_queries(n_q, 20.0, seed=100 + f)).to(dev) + torch.tensor([sensor[0].item(), sensor[1].item(), 0.0], device=dev)
        P = PcRadiusParams.make(start=tuple(sensor.tolist()), **P0)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ix.build(frame)
        e1.record()
        r = ix.radius(q, P)
        e2.record()
        torch.cuda.synchronize()
        builds.append(e0.elapsed_time(e1)); queries.append(e1.elapsed_time(e2)); per_frame.append(e0.elapsed_time(e2))
        if f == frames - 1:
            idx, d2 = ix.nearest(q[:2000])
            torch.cuda.synchronize()
            ok = brute_check(frame, q[:2000], idx, d2, 48)
    print(json.dumps({"config": "c3_lidar_stream", "frame_points": n_frame, "queries_per_frame": n_q, "frames": frames,
                      "index_build_ms_per_frame": float(np.median(builds[2:])), "radius_queries_ms_per_frame": float(np.median(queries[2:])),
                      "frame_ms": float(np.median(per_frame[2:])), "frame_budget_ms_at_10Hz": 100.0,
                      "parity_spot_check": ok}))
    ix.close()


def c4():
    dev = torch.device("cuda", 0)
    stream = setup_stream(dev)
    pts, half = synth.forest_cloud(5_000_000, seed=2, variant="J", return_half=True)
    t_pts = torch.from_numpy(pts).to(dev)
    ix = PointCloudIndex(max_points=len(pts), device=0, stream=stream)
    build_ms, _ = timed(lambda: ix.build(t_pts), reps=3, warm=1)
    n_traj = 10_000
    tr = synth.bezier_trajectories(n_traj, half * 0.9, seed=5, seg_range=(8, 8), T_range=(2.5, 3.0))   # >= 20 s each
    P = PcRadiusParams.make(start=(0, 0, 2), search_margin=0.25, max_radius=1.5, sample_range=-1.0)
    args = (tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"], P)
    ix.clearance(*args, horizon=20.0)                   # warm-up
    t0 = time.perf_counter()
    fh, mr, ns = ix.clearance(*args, horizon=20.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    total = int(ns.sum())
    print(json.dumps({"config": "c4_clearance", "points": len(pts), "trajectories": n_traj, "samples_total": total,
                      "samples_per_traj_median": int(np.median(ns)), "index_build_ms": build_ms,
                      "clearance_call_ms_host_buffers": wall_ms, "samples_per_s": total / wall_ms * 1e3,
                      "colliding_trajectories": int((fh >= 0).sum())}))
    ix.close()


def c5(n_pts, n_q, shard):
    import torch.distributed as dist
    from pointcloudtraj_b200.dist import Replicator, env_rank, shard_range
    rank, world, local = env_rank()
    dev = torch.device("cuda", local)
    stream = setup_stream(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ix = PointCloudIndex(max_points=n_pts, device=local, stream=stream)
    meta = torch.zeros(2, dtype=torch.float64, device=dev)
    build_ms = 0.0
    t_pts = None
    if rank == 0:
        # tile a 5M-point forest map until it has n_pts points (fresh jitter per tile): ~200 points / m^2 everywhere
        base, half = synth.forest_cloud(min(5_000_000, n_pts), seed=3, variant="L", return_half=True)
        tb = torch.from_numpy(base).to(dev)
        tiles = -(-n_pts // len(base))
        side = int(np.ceil(np.sqrt(tiles)))
        g = torch.Generator(device=dev).manual_seed(7)
        parts = []
        for t in range(tiles):
            off = torch.tensor([(t % side) * 2 * half, (t // side) * 2 * half, 0.0], device=dev)
            parts.append(tb + off + (torch.rand(tb.shape, device=dev, generator=g) - 0.5) * 0.1)
        t_pts = torch.cat(parts)[:n_pts].contiguous()
        del parts
        meta[0], meta[1] = half, side
        build_ms, _ = timed(lambda: ix.build(t_pts), reps=3, warm=1)
    bcast_ms = 0.0
    if world > 1:
        rep = Replicator(rank, world, local)
        dist.broadcast(meta, 0)
        rep.broadcast(ix, 0)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rep.broadcast(ix, 0); e1.record(); torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    half, side = float(meta[0]), int(meta[1])
    ext = torch.tensor([2 * half * side, 2 * half * side, 3.4], device=dev)
    lo = torch.tensor([-half, -half, 0.6], device=dev)
    answered = n_q
    if shard == "spatial" and world > 1:
        # every rank holds the SAME batch and answers its own stretch of the Hilbert curve (pc_batch_shard)
        g = torch.Generator(device=dev).manual_seed(1000)
        q = (torch.rand((n_q, 3), device=dev, generator=g) * ext + lo).contiguous()
        ix.batch_shard(rank, world)
    else:
        # every rank answers a contiguous slice of the batch (pc_shard_range)
        b, e = shard_range(n_q, rank, world)
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        q = (torch.rand((e - b, 3), device=dev, generator=g) * ext + lo).contiguous()
    if world > 1:
        dist.barrier()
    nn_ms, _ = timed(lambda: ix.nearest(q), reps=3, warm=1)
    if shard == "spatial" and world > 1:
        idx_all, _ = ix.nearest(q)
        cnt = (idx_all != PointCloudIndex.NOT_MINE_IDX).sum().to(torch.float64)
        dist.all_reduce(cnt)
        answered = int(cnt.item())
        ix.batch_shard(0, 1)
    t = torch.tensor([nn_ms], dtype=torch.float64, device=dev)
    per_rank = [nn_ms]
    if world > 1:
        allt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(v.item()) for v in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        idx, d2 = ix.nearest(q[:4096])
        torch.cuda.synchronize()
        ok = brute_check(t_pts, q[:4096], idx, d2, 24)
        print(json.dumps({"config": "c5_scaling", "n_gpus": world, "points": n_pts, "queries_total": n_q, "scaling": "strong",
                          "sharding": shard if world > 1 else "none", "queries_answered_over_ranks": answered,
                          "index_build_ms": build_ms, "index_broadcast_ms": bcast_ms,
                          "index_bytes": int(ix.view().n_nodes * 64 + (ix.view().n_points + 4) * 16),
                          "nearest_ms_max_over_ranks": float(t.item()), "nearest_ms_per_rank": [round(v, 3) for v in per_rank], "nearest_qps": n_q / float(t.item()) * 1e3,
                          "parity_spot_check": ok}))
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c2small", "c3", "c4", "c5"])
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--shard", default="spatial", choices=["spatial", "contiguous"])
    a = ap.parse_args()
    if a.config == "c1":
        c1_like("c1_200k_points_100k_queries", 200_000, 100_000, 6)
    elif a.config == "c2small":
        c1_like("c2_1M_points_50k_samples", 1_000_000, 50_000, 1)
    elif a.config == "c3":
        c3()
    elif a.config == "c4":
        c4()
    else:
        c5(a.points, a.queries, a.shard)
