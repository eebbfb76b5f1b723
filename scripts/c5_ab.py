#!/usr/bin/env python
"""A/B helper: C5-like run (large tiled cloud, unbounded nearest) with alternative builds of the library."""
import ctypes as C, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointcloudtraj_b200 import _lib, synth
n_pts, n_q = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
stream = torch.cuda.current_stream().cuda_stream
base, half = synth.forest_cloud(min(5_000_000, n_pts), seed=3, variant="L", return_half=True)
tb = torch.from_numpy(base).to(dev)
tiles = -(-n_pts // len(base)); side = int(np.ceil(np.sqrt(tiles)))
g = torch.Generator(device=dev).manual_seed(7)
parts = [tb + torch.tensor([(t % side) * 2 * half, (t // side) * 2 * half, 0.0], device=dev) + (torch.rand(tb.shape, device=dev, generator=g) - 0.5) * 0.1 for t in range(tiles)]
t_pts = torch.cat(parts)[:n_pts].contiguous(); del parts
ext = torch.tensor([2 * half * side, 2 * half * side, 3.4], device=dev); lo = torch.tensor([-half, -half, 0.6], device=dev)
q = (torch.rand((n_q, 3), device=dev, generator=g) * ext + lo).contiguous()
oi = torch.empty(n_q, dtype=torch.int32, device=dev); od = torch.empty(n_q, dtype=torch.float32, device=dev)
ref = None
for defs in sys.argv[3:]:
    out = os.path.join(ROOT, "gpurun_out", "libab_" + defs.replace("-D", "").replace("=", "").replace(" ", "_") + ".so")
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"] + defs.split() +
                   ["-o", out, os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu"), "-lcudart", "-ldl"], check=True, capture_output=True)
    _lib._lib = None; _lib.LIB_PATH = out; L = _lib.load()
    h = C.c_void_p(); assert L.pc_index_create(C.byref(h), 0, n_pts, C.c_void_p(stream)) == 0
    assert L.pc_index_build(h, C.c_void_p(t_pts.data_ptr()), n_pts, 3, 1) == 0
    L.pc_profile_enable(h, 1)
    ts = []
    for _ in range(3):
        assert L.pc_nearest_batch(h, C.c_void_p(q.data_ptr()), n_q, 3, 1, 0, C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr())) == 0
        a, b = C.c_float(), C.c_float(); L.pc_profile_last_batch(h, C.byref(a), C.byref(b)); ts.append((a.value, b.value))
    torch.cuda.synchronize()
    if ref is None: ref = oi.clone()
    print(f"{defs:40s} order {ts[-1][0]:8.3f} ms  search {ts[-1][1]:9.3f} ms  {n_q / sum(ts[-1]) / 1e6:8.3f} Gq/s  same={bool((oi == ref).all().item())}", flush=True)
    L.pc_index_destroy(h)
