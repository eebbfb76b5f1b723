"""CPU check of the radix-tree node function (pointcloudtraj_b200/csrc/lbvh.cuh) that the planned prefix-split index build
uses: compiled for the host, compared with a recursive top-down construction."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radix_tree_nodes_match_recursive_construction(tmp_path):
    exe = str(tmp_path / "lbvh_check")
    subprocess.run(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c", "lbvh_check.cpp")], check=True, capture_output=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "lbvh ok" in p.stdout


def test_planned_index_layout_and_walk_are_exact(tmp_path):
    """Records laid out as the planned CUDA build will write them (child references in the .w words, leaves of <= 4 points,
    boxes fitted bottom-up in arbitrary node order) and walked by pc_lbvh_nearest: exact nearest point, lowest index among
    ties, on a jittered forest, a tie-heavy lattice forest, tiny clouds and coincident points."""
    import numpy as np
    from pointcloudtraj_b200 import synth
    exe = str(tmp_path / "lbvh_index_check")
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "c", "lbvh_index_check.cpp")],
                   check=True, capture_output=True)
    cases = []
    pts, half = synth.forest_cloud(40_000, seed=6, variant="J", return_half=True)
    cases.append((pts[:, :3], synth.rrt_queries(600, half, seed=3)))
    pts, half = synth.forest_cloud(30_000, seed=6, variant="L", return_half=True)
    q = synth.rrt_queries(600, half, seed=4)
    q[::2] = np.round(q[::2] / 0.05) * 0.05
    cases.append((pts[:, :3], q))
    rng = np.random.default_rng(0)
    for n in (1, 2, 4, 5, 9, 33):
        cases.append((rng.normal(size=(n, 3)), rng.normal(size=(100, 3))))
    cases.append((np.tile([[1.5, -2.0, 0.75]], (300, 1)), rng.normal(size=(50, 3))))
    for k, (p, q) in enumerate(cases):
        fp, fq = str(tmp_path / f"p{k}.bin"), str(tmp_path / f"q{k}.bin")
        np.ascontiguousarray(p, np.float32).tofile(fp)
        np.ascontiguousarray(q, np.float32).tofile(fq)
        r = subprocess.run([exe, fp, fq], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "mismatches=0" in r.stdout, (k, r.stdout, r.stderr)


def test_packet_walk_over_the_planned_index_is_exact(tmp_path):
    """The 64-query packet walk (shared stack, any-query-wants-it tests, vote of the first 32 queries, leaf and inner children
    mixed under one node) emulated on the CPU over the planned index and over the implicit tree at HEAD: both exact on a
    tie-heavy lattice forest, bounded and unbounded."""
    import numpy as np
    from pointcloudtraj_b200 import synth
    exe = str(tmp_path / "lbvh_packet_check")
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "c", "lbvh_packet_check.cpp")],
                   check=True, capture_output=True)
    pts, half = synth.forest_cloud(20_000, seed=6, variant="L", return_half=True)
    q = synth.rrt_queries(2000, half, seed=3)
    q[::2] = np.round(q[::2] / 0.05) * 0.05
    fp, fq = str(tmp_path / "p.bin"), str(tmp_path / "q.bin")
    np.ascontiguousarray(pts[:, :3], np.float32).tofile(fp)
    np.ascontiguousarray(q, np.float32).tofile(fq)
    for bound in ("1.75", "0"):
        for extra in ([], ["implicit"]):
            r = subprocess.run([exe, fp, fq, bound] + extra, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0 and "mismatches=0" in r.stdout, (bound, extra, r.stdout, r.stderr)


def test_frontier_walks_over_the_planned_index_are_exact(tmp_path):
    """The warp-per-query frontier walks (nearest with 32 and 8 lanes per query, and the range walk with the records' per-leaf
    point counts) emulated on the CPU over the planned index: exact nearest points and exact range sets."""
    import numpy as np
    from pointcloudtraj_b200 import synth
    exe = str(tmp_path / "lbvh_frontier_check")
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "c", "lbvh_frontier_check.cpp")],
                   check=True, capture_output=True)
    rng = np.random.default_rng(0)
    cases = []
    pts, half = synth.forest_cloud(20_000, seed=6, variant="L", return_half=True)
    q = synth.rrt_queries(400, half, seed=3)
    q[::2] = np.round(q[::2] / 0.05) * 0.05
    cases.append((pts[:, :3], q, "1.75", "1.0"))
    cases.append((pts[:, :3], q, "0", "0.5"))
    for n in (1, 4, 5, 33):
        cases.append((rng.normal(size=(n, 3)), rng.normal(size=(40, 3)), "0", "1.5"))
    cases.append((np.tile([[1.5, -2.0, 0.75]], (200, 1)), rng.normal(size=(20, 3)), "0", "3.0"))
    for k, (p, qq, bound, rad) in enumerate(cases):
        fp, fq = str(tmp_path / f"p{k}.bin"), str(tmp_path / f"q{k}.bin")
        np.ascontiguousarray(p, np.float32).tofile(fp)
        np.ascontiguousarray(qq, np.float32).tofile(fq)
        r = subprocess.run([exe, fp, fq, bound, rad], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout.count("mismatches=0") == 2, (k, r.stdout, r.stderr)
