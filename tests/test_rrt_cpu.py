"""CPU test of the restated RRT* expansion driver (include/pc_rrt.hpp) with the oracle as radius provider."""
import os
import subprocess

from pointcloudtraj_b200 import synth
from rrt_common import read_records, validate_corridor, write_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expansion_drivers_build_valid_corridors(tmp_path):
    tmp = str(tmp_path)
    pts, half = synth.forest_cloud(60_000, seed=6, variant="J", return_half=True)
    write_input(os.path.join(tmp, "in.bin"), pts, max(half, 12.0), max_iter=6000, K=128)
    exe = os.path.join(tmp, "rrt_cpu_check")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c", "rrt_cpu_check.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-L", odir, "-loracle", f"-Wl,-rpath,{odir}"], check=True, capture_output=True)
    p = subprocess.run([exe, os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.returncode, p.stderr)
    seq, bat = read_records(os.path.join(tmp, "out.bin"), 2)
    validate_corridor(seq, pts)
    validate_corridor(bat, pts)
    assert seq["cloud_queries"] > 3000 and bat["cloud_queries"] > 3000
    assert seq["nodes"] > 50 and bat["nodes"] > 50
