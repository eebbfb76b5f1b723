"""Generate tests/golden/*.npz from the UNMODIFIED reference kd-tree.

Run in the container that has /root/reference (`make -C oracle` builds oracle/_ref/libkdtree_ref.so
from the reference's own Utils/kdtree/src/kdtree.c).  Every output array below is produced by the
reference library's kd_insert3 / kd_nearest3 / kd_nearest_range3 through oracle/ref_harness.c;
inputs are stored alongside so the fixtures do not depend on a PRNG implementation.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pointcloudtraj_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(name, pts, q, order, range_q, r):
    ref = oracle.KdReference().build(pts, order)
    idx, d2 = ref.nearest(q, nthreads=1)
    off, lst = ref.range(range_q, r, nthreads=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pts=pts, q=q, order=order.astype(np.int64),
                        nn_idx=idx.astype(np.int32), nn_d2=d2, range_q=range_q, range_r=np.float64(r),
                        range_off=off, range_idx=lst.astype(np.int32))
    print(f"{name}: {len(pts)} pts, {len(q)} queries, {len(range_q)} range queries ({off[-1]} hits)")


def main():
    oracle.build()
    assert oracle.have_reference(), "needs oracle/_ref/libkdtree_ref.so (container with /root/reference)"
    # 1. lattice forest (tie-heavy): half of the queries snapped to the res/2 lattice
    pts, half = synth.forest_cloud(6000, seed=6, variant="L", return_half=True)
    q = synth.rrt_queries(3000, half, seed=0, lattice_frac=0.5)
    order = np.random.default_rng(11).permutation(len(pts))
    run_case("forest_lattice", pts, q, order, q[:300], 0.6)
    # 2. jittered forest (tie-free)
    pts, half = synth.forest_cloud(6000, seed=1, variant="J", return_half=True)
    q = synth.rrt_queries(3000, half, seed=2)
    order = np.random.default_rng(12).permutation(len(pts))
    run_case("forest_jitter", pts, q, order, q[:300], 0.8)
    # 3. uniform control, generation-order insertion
    pts = synth.uniform_cloud(5000, half=8.0, seed=3)
    q = synth.rrt_queries(2000, 8.0, seed=4, z=(0.0, 8.0))
    run_case("uniform", pts, q, np.arange(len(pts)), q[:200], 0.7)
    # 4. known-answer behaviours of the reference: insertion-order ties, duplicates, range boundary
    kat = {}
    ref = oracle.KdReference().build(np.array([[1, 0, 0], [-1, 0, 0]], np.float32))
    kat["tie_two_first_inserted_wins"] = ref.nearest(np.zeros((1, 3), np.float32))[0]
    ref = oracle.KdReference().build(np.array([[1, 0, 0], [-1, 0, 0]], np.float32), np.array([1, 0]))
    kat["tie_two_reversed_insert"] = ref.nearest(np.zeros((1, 3), np.float32))[0]
    ref = oracle.KdReference().build(np.array([[2, 2, 2]] * 3, np.float32))
    kat["three_duplicates"] = ref.nearest(np.array([[2.5, 2, 2]], np.float32))[0]
    bp = np.array([[0.5, 5, 5], [0.5, 0, 0]], np.float32)
    ref = oracle.KdReference().build(bp)
    kat["range_boundary_left"] = ref.range(np.array([[-0.5, 0, 0]], np.float32), 1.0)[0]
    kat["range_boundary_right"] = ref.range(np.array([[1.5, 0, 0]], np.float32), 1.0)[0]
    ref = oracle.KdReference().build(np.zeros((0, 3), np.float32))
    i, d = ref.nearest(np.zeros((2, 3), np.float32))
    kat["empty_idx"], kat["empty_d2"] = i, d
    kat["empty_range"] = ref.range(np.zeros((2, 3), np.float32), 1.0)[0]
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)
    print({k: v.tolist() for k, v in kat.items()})


if __name__ == "__main__":
    main()
