"""GPU test of the RRT* driver (include/pc_rrt.hpp): replay parity and speculative batches, for corridor growth
(SafeRegionExpansion) and for the re-validation of the corridor against a new cloud message (SafeRegionEvaluate / Refine)."""
import os
import subprocess

import numpy as np
import pytest

from pointcloudtraj_b200 import synth
from rrt_common import blocked_cloud, read_records, validate_corridor, write_input

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("rrt"))
    exe = os.path.join(tmp, "rrt_client")
    odir, ldir = os.path.join(ROOT, "oracle"), os.path.join(ROOT, "pointcloudtraj_b200")
    subprocess.run(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c", "rrt_client.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-L", odir, "-loracle", "-L", ldir, "-lpcindex", f"-Wl,-rpath,{odir}", f"-Wl,-rpath,{ldir}"], check=True, capture_output=True)

    def run(n_rec, *args, **kw):
        write_input(os.path.join(tmp, "in.bin"), *args, **kw)
        p = subprocess.run([exe, os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, (p.returncode, p.stderr)
        print(p.stderr.strip())
        return read_records(os.path.join(tmp, "out.bin"), n_rec)
    return run


def _same(a, b):
    return (a["k"] == b["k"] and a["nodes"] == b["nodes"] and a["cloud_queries"] == b["cloud_queries"]
            and (a["path"] == b["path"]).all() and (a["radius"] == b["radius"]).all())


def _same_tree(a, b):
    return a["k"] == b["k"] and a["nodes"] == b["nodes"] and (a["path"] == b["path"]).all() and (a["radius"] == b["radius"]).all()


def test_replay_parity_and_batched_growth(client):
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)      # the clean_demo-sized map
    gpu, cpu, bat, bat_nodes, bat_range, bat_dev, bat_all = client(7, pts, half, max_iter=20_000, K=512)
    # replay mode: the GPU radius provider and the CPU oracle drive the SAME planner logic to bit-identical corridors
    assert gpu["k"] >= 2 and _same(gpu, cpu)
    validate_corridor(gpu, pts)
    # speculative batches: a valid corridor built from far fewer (batched) radius calls
    validate_corridor(bat, pts, float_centres=True)
    # ... and with the batch's nearest-vertex queries answered on the GPU too (the client checks them against the CPU node tree)
    validate_corridor(bat_nodes, pts, float_centres=True)
    # ... and with treeRewire's 2 x radius neighbourhoods answered for the whole batch by one pc_range_batch on the snapshot
    # index: the SAME corridor as with the CPU node tree, bit for bit
    assert _same(bat_range, bat)
    # ... and with the snapshot phase of every batch generated and answered on the device (pc_expand_batch): the same tree and
    # corridor as the buffer-API batches with GPU nearest-vertex answers, bit for bit
    assert _same_tree(bat_dev, bat_nodes)
    # ... and with both: device-generated batches + the rewire neighbourhoods from the GPU -- only the sequential insertion is
    # left on the host; the same tree again
    assert _same_tree(bat_all, bat_dev)
    print(f"corridor growth, 20k iterations: one query per iteration GPU {gpu['ms']:.1f} ms, CPU oracle {cpu['ms']:.1f} ms; "
          f"batches of 512: {bat['ms']:.1f} ms ({bat['nodes']} nodes, {bat['k']} spheres); "
          f"+ node-tree nearest on the GPU: {bat_nodes['ms']:.1f} ms ({bat_nodes['nodes']} nodes); "
          f"+ node-tree range on the GPU instead: {bat_range['ms']:.1f} ms; device-generated batches: {bat_dev['ms']:.1f} ms; "
          f"device-generated batches + node-tree range on the GPU: {bat_all['ms']:.1f} ms")

    # a second cloud message with new obstacles ON a middle sphere of each corridor: index rebuild, SafeRegionEvaluate
    # (batched re-query of the path nodes), refinement -- again bit-identical between the GPU and the oracle provider
    blocked = [gpu["path"][gpu["k"] // 2], bat["path"][bat["k"] // 2]]
    pts2 = blocked_cloud(pts, blocked)
    recs = client(21, pts, half, max_iter=20_000, K=512, pts2=pts2, refine_iter=5000)
    g, c, b, bn, br, bd, ba = recs[0:3], recs[3:6], recs[6:9], recs[9:12], recs[12:15], recs[15:18], recs[18:21]
    for phase in range(3):
        assert _same(br[phase], b[phase])                                               # GPU node-tree range == CPU node tree
        assert _same_tree(bd[phase], bn[phase])                                         # device-generated batches == buffer-API batches
        assert _same_tree(ba[phase], bd[phase])                                         # ... with the GPU range provider on top
    for rec in bn[1:]:
        if rec["k"]:
            validate_corridor(rec, pts2, float_centres=True)
    assert _same(g[0], gpu) and _same(b[0], bat)                                        # deterministic first phase
    for phase in range(3):
        assert _same(g[phase], c[phase])
    for rec, centre, fc in ((g[1], blocked[0], False), (g[2], blocked[0], False), (b[1], blocked[1], True), (b[2], blocked[1], True)):
        if rec["k"]:
            assert not np.isclose(rec["path"], centre, atol=1e-6).all(axis=1).any()     # the blocked sphere is gone
            validate_corridor(rec, pts2, float_centres=fc)
    assert g[2]["k"] >= 2 and b[2]["k"] >= 2                                            # a way around the new obstacle
    print(f"re-validation on a new cloud: evaluate GPU {g[1]['ms']:.2f} ms (incl. index rebuild) / oracle {c[1]['ms']:.2f} ms, "
          f"refine+evaluate cycles GPU {g[2]['ms']:.1f} ms / batched {b[2]['ms']:.1f} ms")
