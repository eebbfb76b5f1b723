"""CPU test of the restated RRT* expansion driver (include/pc_rrt.hpp) with the oracle as radius provider."""
import os
import subprocess

import numpy as np

from pointcloudtraj_b200 import synth
from rrt_common import blocked_cloud, read_records, validate_corridor, write_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "rrt_cpu_check")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c", "rrt_cpu_check.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-L", odir, "-loracle", f"-Wl,-rpath,{odir}"], check=True, capture_output=True)
    return exe


def _run(exe, tmp, n_rec):
    p = subprocess.run([exe, os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.returncode, p.stderr)
    return read_records(os.path.join(tmp, "out.bin"), n_rec)


def _same_tree(a, b):
    return a["k"] == b["k"] and a["nodes"] == b["nodes"] and (a["path"] == b["path"]).all() and (a["radius"] == b["radius"]).all()


def test_expansion_drivers_build_valid_corridors(tmp_path):
    tmp = str(tmp_path)
    pts, half = synth.forest_cloud(60_000, seed=6, variant="J", return_half=True)
    half = max(half, 12.0)
    exe = _build(tmp)
    write_input(os.path.join(tmp, "in.bin"), pts, half, max_iter=6000, K=128)
    seq, bat, bat_nn, bat_dev = _run(exe, tmp, 4)
    # the device-batch provider contract (engine state in, candidates + engine state out) reproduces the buffer-style batches
    assert _same_tree(bat_dev, bat_nn) and bat_nn["nodes"] > 50
    validate_corridor(seq, pts)
    validate_corridor(bat, pts)
    assert seq["cloud_queries"] > 3000 and bat["cloud_queries"] > 3000
    assert seq["nodes"] > 50 and bat["nodes"] > 50

    # a new cloud message with obstacles ON a middle sphere of each corridor: SafeRegionEvaluate (corridor_finder.cpp:817-936)
    # must drop those spheres; whatever corridor is reported afterwards is valid against the NEW cloud
    blocked = [seq["path"][seq["k"] // 2], bat["path"][bat["k"] // 2]]
    pts2 = blocked_cloud(pts, blocked)
    write_input(os.path.join(tmp, "in.bin"), pts, half, max_iter=6000, K=128, pts2=pts2, refine_iter=2000)
    recs = _run(exe, tmp, 12)
    for phase in range(3):
        assert _same_tree(recs[9 + phase], recs[6 + phase])
    for grow, ev, ref, old, b in ((recs[0], recs[1], recs[2], seq, blocked[0]), (recs[3], recs[4], recs[5], bat, blocked[1])):
        assert grow["k"] == old["k"] and (grow["path"] == old["path"]).all()          # deterministic first phase
        assert ev["cloud_queries"] >= grow["cloud_queries"] + old["k"] - 1            # every non-root path node was re-queried
        for rec in (ev, ref):
            if rec["k"]:
                assert not (rec["path"] == b).all(axis=1).any()                       # the blocked sphere is gone
                validate_corridor(rec, pts2)
        assert ref["k"] >= 2                                                          # refinement finds a way around the new obstacle
