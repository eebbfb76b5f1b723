"""GPU test of the RRT* expansion driver (include/pc_rrt.hpp): replay parity and speculative batches."""
import os
import subprocess

import numpy as np
import pytest

from pointcloudtraj_b200 import synth
from rrt_common import read_records, validate_corridor, write_input

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replay_parity_and_batched_growth(tmp_path):
    tmp = str(tmp_path)
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)      # the clean_demo-sized map
    write_input(os.path.join(tmp, "in.bin"), pts, half, max_iter=20_000, K=512)
    exe = os.path.join(tmp, "rrt_client")
    odir, ldir = os.path.join(ROOT, "oracle"), os.path.join(ROOT, "pointcloudtraj_b200")
    subprocess.run(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c", "rrt_client.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-L", odir, "-loracle", "-L", ldir, "-lpcindex", f"-Wl,-rpath,{odir}", f"-Wl,-rpath,{ldir}"], check=True, capture_output=True)
    p = subprocess.run([exe, os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.returncode, p.stderr)
    gpu, cpu, bat = read_records(os.path.join(tmp, "out.bin"), 3)
    # replay mode: the GPU radius provider and the CPU oracle drive the SAME planner logic to bit-identical corridors
    assert gpu["k"] == cpu["k"] >= 2 and gpu["nodes"] == cpu["nodes"] and gpu["cloud_queries"] == cpu["cloud_queries"]
    assert (gpu["path"] == cpu["path"]).all() and (gpu["radius"] == cpu["radius"]).all()
    validate_corridor(gpu, pts)
    # speculative batches: a valid corridor built from far fewer (batched) radius calls
    validate_corridor(bat, pts, float_centres=True)
    print(f"corridor growth, 20k iterations: one query per iteration GPU {gpu['ms']:.1f} ms, CPU oracle {cpu['ms']:.1f} ms; "
          f"batches of 512: {bat['ms']:.1f} ms ({bat['nodes']} nodes, {bat['k']} spheres)")
