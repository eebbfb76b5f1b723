#!/usr/bin/env python
"""One-GPU emulation of the strong-scaling runs: the time every rank of G needs for its share of ONE batch, measured on a single
GPU by switching the rank of one handle (pc_batch_shard) -- what bench.py's `strong` key and scripts/bench_configs.py's C5
measure as the max over ranks.  usage: shard_sim.py bench|benchnn|c5 G [n_queries]   (bench: radius batches on the 1 M map; benchnn: unbounded nearest there)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import _lib
if os.environ.get("SHARD_SIM_DEFS"):      # A/B builds of the library: SHARD_SIM_DEFS="-DPC_SHARD_CELL_QUERIES=512"
    import subprocess
    defs = os.environ["SHARD_SIM_DEFS"]
    out = os.path.join(ROOT, "gpurun_out", "libsim_" + "".join(c if c.isalnum() else "_" for c in defs) + ".so")
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"] + defs.split() +
                   ["-o", out, os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu"), "-lcudart", "-ldl"], check=True, capture_output=True)
    _lib.LIB_PATH = out
from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth

what, G = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
stream = torch.cuda.current_stream().cuda_stream
if what in ("bench", "benchnn"):
    n_q = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000 * G
    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    t_pts = torch.from_numpy(pts).to(dev)
    q = torch.cat([torch.from_numpy(synth.rrt_queries(min(10_000_000, n_q - o), half, seed=1000 + o // 10_000_000)).to(dev) for o in range(0, n_q, 10_000_000)])
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0)) if what == "bench" else None
else:
    n_pts = 100_000_000
    n_q = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
    base, half = synth.forest_cloud(5_000_000, seed=3, variant="L", return_half=True)
    tb = torch.from_numpy(base).to(dev)
    tiles = -(-n_pts // len(base)); side = int(np.ceil(np.sqrt(tiles)))
    g = torch.Generator(device=dev).manual_seed(7)
    parts = [tb + torch.tensor([(t % side) * 2 * half, (t // side) * 2 * half, 0.0], device=dev) + (torch.rand(tb.shape, device=dev, generator=g) - 0.5) * 0.1 for t in range(tiles)]
    t_pts = torch.cat(parts)[:n_pts].contiguous(); del parts
    ext = torch.tensor([2 * half * side, 2 * half * side, 3.4], device=dev); lo = torch.tensor([-half, -half, 0.6], device=dev)
    q = (torch.rand((n_q, 3), device=dev, generator=g) * ext + lo).contiguous()
    P = None
ix = PointCloudIndex(max_points=len(t_pts), device=0, stream=stream)
ix.build(t_pts)
ix.profile(True)


def call(qq):
    return ix.radius(qq, P) if P is not None else ix.nearest(qq)


def run(qq, label):
    ts, parts = [], None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(qq); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); parts = ix.last_batch_ms()
    t = float(np.median(ts[1:]))
    d3 = (C.c_float * 3)()
    det = f"  order = clear {d3[0]:.3f} + key/count {d3[1]:.3f} + sort/scan+scatter {d3[2]:.3f}" if ix._L.pc_profile_last_order_detail(ix._h, d3) == 0 else ""
    print(f"{label:32s} call {t:8.3f} ms   (order {parts[0]:7.3f}  search {parts[1]:7.3f}){det}", flush=True)
    return t


print(f"# {what}: {len(t_pts)} points, {n_q} queries, G={G}, PC_SHARD_EXACT={os.environ.get('PC_SHARD_EXACT', '1')}")
T = run(q, "full batch, one GPU")
ts = run(q[: n_q // G].contiguous(), f"contiguous 1/{G} slice")
worst = 0.0
ranks = [int(os.environ["SHARD_SIM_RANK"])] if "SHARD_SIM_RANK" in os.environ else range(G)
for r in ranks:
    ix.batch_shard(r, G)
    out = call(q)
    own = int((~torch.isnan(out)).sum().item()) if P is not None else int((out[0] != PointCloudIndex.NOT_MINE_IDX).sum().item())
    worst = max(worst, run(q, f"pc_batch_shard rank {r}/{G} ({own / n_q:.4f})"))
ix.batch_shard(0, 1)
print(f"ideal {T / G:.3f} ms; slices {ts:.3f} ms = {T / G / ts:.1%}; pc_batch_shard max over ranks {worst:.3f} ms = {T / G / worst:.1%}")
