"""GPU edge cases: degenerate clouds, far / non-finite queries, argument validation, error reporting."""
import ctypes as C

import numpy as np
import pytest

import oracle
from pointcloudtraj_b200 import (PC_QUERY_SORTED, PC_QUERY_UNSORTED, PC_RADIUS_FULL_NN, PcError, PcRadiusParams,
                                 PointCloudIndex, synth)
from pointcloudtraj_b200 import _lib as L
from parity import check_lowest_index_everywhere

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ix():
    h = PointCloudIndex(max_points=0, device=0)      # grows on demand: exercises the arena re-allocation path
    yield h
    h.close()


def test_all_points_identical(ix):
    pts = np.tile(np.array([[1.5, -2.0, 0.75]], np.float32), (5000, 1))
    q = synth.rrt_queries(40_000, 4.0, seed=1)
    ix.build(pts)
    for flags in (PC_QUERY_UNSORTED, PC_QUERY_SORTED):
        idx, d2 = ix.nearest(q, flags=flags)
        assert (idx == 0).all()                                  # 5000 exact ties: lowest index
        assert (d2 == oracle.pair_d2(pts, q, np.zeros(len(q), np.int64)).astype(np.float32)).all()
    off, lst = ix.range(q[:50], 100.0)
    assert (np.diff(off) == 5000).all() and (lst[:5000] == np.arange(5000)).all()


def test_collinear_and_planar_clouds(ix):
    rng = np.random.default_rng(3)
    line = np.zeros((20_000, 3), np.float32)
    line[:, 0] = rng.uniform(-50, 50, len(line))                  # all on the x axis: boxes degenerate in y and z
    plane = np.zeros((20_000, 3), np.float32)
    plane[:, :2] = rng.uniform(-5, 5, (len(plane), 2))
    q = synth.rrt_queries(33_000, 6.0, seed=2, z=(-3, 3))
    for pts in (line, plane):
        ix.build(pts)
        idx, d2 = ix.nearest(q)
        check_lowest_index_everywhere(pts, q[:3000], idx[:3000])
        bi, bd, _ = oracle.brute_nearest(pts, q[:3000])
        assert (d2[:3000] == bd.astype(np.float32)).all()


def test_far_and_huge_queries(ix):
    pts = synth.uniform_cloud(10_000, half=5.0, seed=4)
    ix.build(pts)
    q = np.array([[1e6, 0, 0], [-1e6, 1e6, -1e6], [0, 0, 1e9], [3e4, -3e4, 2.0]], np.float32)
    q = np.concatenate([q, synth.rrt_queries(40_000, 5.0, seed=5)])      # mixed into an ordered batch
    idx, d2 = ix.nearest(q)
    bi, bd, _ = oracle.brute_nearest(pts, q[:64])
    assert (idx[:64] == bi).all() and (d2[:64] == bd.astype(np.float32)).all()
    P = PcRadiusParams.make(0.25, 1.5, -1.0, (0, 0, 0))                   # sensing-range early-out disabled
    r = ix.radius(q[:4], P)
    assert (r == np.float32(1.5)).all()                                   # far from everything: clamped


def test_non_finite_queries_do_not_poison_their_neighbours(ix):
    pts = synth.uniform_cloud(20_000, half=5.0, seed=6)
    ix.build(pts)
    q = synth.rrt_queries(50_000, 5.0, seed=7)
    ref_idx, ref_d2 = ix.nearest(q)
    bad = q.copy()
    rows = np.arange(0, len(q), 97)
    bad[rows, 0] = np.nan
    bad[rows[::2], 1] = np.inf
    idx, d2 = ix.nearest(bad)
    good = np.ones(len(q), bool)
    good[rows] = False
    assert (idx[good] == ref_idx[good]).all() and (d2[good] == ref_d2[good]).all()
    assert (idx[rows[1::2]] == -1).all()                                   # NaN query: no neighbour (documented)


def test_argument_validation_and_error_strings(ix):
    lib = ix._L
    pts = synth.uniform_cloud(100, seed=1)
    ix.build(pts)
    q = synth.rrt_queries(10, 5.0, seed=1)
    out = np.empty(10, np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    assert lib.pc_nearest_batch(ix._h, vp(q), 10, 5, 0, 0, vp(out), None) == L.PC_EINVAL         # stride must be 3 or 4
    assert b"stride" in lib.pc_last_error(ix._h)
    assert lib.pc_nearest_batch(ix._h, vp(q), -1, 3, 0, 0, vp(out), None) == L.PC_EINVAL
    assert lib.pc_nearest_batch(ix._h, None, 10, 3, 0, 0, vp(out), None) == L.PC_EINVAL
    assert lib.pc_nearest_batch(ix._h, vp(q), 10, 3, 7, 0, vp(out), None) == L.PC_EINVAL         # unknown memory space
    assert lib.pc_index_build(ix._h, vp(pts), 100, 2, 0) == L.PC_EINVAL
    assert lib.pc_radius_batch(ix._h, vp(q), 10, 3, 0, 0, None, vp(out), None) == L.PC_EINVAL      # params required
    assert lib.pc_nearest_batch(ix._h, vp(q), 0, 3, 0, 0, None, None) == L.PC_OK                   # empty batch is legal
    assert ix.size == 100                                                                         # failed calls changed nothing
    h = C.c_void_p()
    assert lib.pc_index_create(C.byref(h), 99, 10, None) == L.PC_EINVAL                            # no such device
    with pytest.raises(PcError):
        ix.clearance([0, 1], [4], [-1.0], [0, 15], np.zeros(15), PcRadiusParams.make())            # T <= 0


def test_arena_growth_and_shrink(ix):
    sizes = [10, 70_000, 300, 2_100_000, 5]
    for n in sizes:
        pts = synth.uniform_cloud(n, half=6.0, seed=n)
        q = synth.rrt_queries(2000, 6.0, seed=n + 1)
        ix.build(pts)
        idx, _ = ix.nearest(q)
        check_lowest_index_everywhere(pts, q[:300], idx[:300])
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0, 0, 2))
    r, i = ix.radius(q, P, flags=PC_RADIUS_FULL_NN, want_idx=True)
    assert (i == idx).all()


def test_non_finite_points_in_the_cloud_are_never_returned(ix):
    pts = synth.uniform_cloud(30_000, half=5.0, seed=8)
    dirty = pts.copy()
    bad = np.arange(0, len(pts), 113)
    dirty[bad[0::3], 0] = np.nan
    dirty[bad[1::3], 1] = np.inf
    dirty[bad[2::3], 2] = -np.inf
    q = synth.rrt_queries(40_000, 5.0, seed=9)
    ix.build(dirty)
    idx, d2 = ix.nearest(q)
    clean = np.ones(len(pts), bool)
    clean[bad] = False
    keep = np.nonzero(clean)[0]
    bi, bd, _ = oracle.brute_nearest(pts[keep], q[:4000])          # brute force over the finite points only
    assert (idx[:4000] == keep[bi]).all() and (d2[:4000] == bd.astype(np.float32)).all()
    assert np.isin(idx, keep).all() and np.isfinite(d2).all()


def test_tiny_host_batches_take_the_mapped_memory_path_with_identical_results(monkeypatch):
    """PC_HOST calls of <= 4096 queries (the planner's one-query-at-a-time radiusSearch) run as one kernel on mapped pinned
    memory; a handle with that path switched off (staged copies) must give the same bits, also across the size threshold."""
    pts, half = synth.forest_cloud(80_000, seed=3, variant="J", return_half=True)
    q = synth.rrt_queries(4200, half, seed=9)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))
    fast = PointCloudIndex(max_points=len(pts), device=0)
    monkeypatch.setenv("PC_TINY_BATCH_QUERIES", "0")
    staged = PointCloudIndex(max_points=len(pts), device=0)
    monkeypatch.delenv("PC_TINY_BATCH_QUERIES")
    try:
        fast.build(pts); staged.build(pts)
        ko = oracle.KdOracle().build(pts)
        o_idx, o_d2 = ko.nearest(q)
        for m in (1, 2, 31, 129, 4096, 4097, 4200):
            n0 = fast.launches()
            r1, i1 = fast.radius(q[:m], P, want_idx=True)
            launches = fast.launches() - n0
            r2, i2 = staged.radius(q[:m], P, want_idx=True)
            assert (r1 == r2).all() and (i1 == i2).all()
            assert launches == 1 or m > 4096                      # one kernel, no ordering pass
            j1, d1 = fast.nearest(q[:m])
            assert fast.nearest(q[:m], want_idx=False)[0] is None
            j2, d2 = staged.nearest(q[:m])
            assert (j1 == j2).all() and (d1 == d2).all() and (j1 == o_idx[:m]).all() and (d1 == o_d2[:m].astype(np.float32)).all()
        # stride-4 queries, outputs partly disabled, empty index
        q4 = np.zeros((100, 4), np.float32); q4[:, :3] = q[:100]
        assert (fast.radius(q4, P) == staged.radius(q[:100], P)).all()
        fast.build(np.zeros((0, 3), np.float32))
        r, i = fast.radius(q[:5], P, want_idx=True)
        assert (r == np.float32(1.25)).all() and (i == -1).all()
    finally:
        fast.close(); staged.close()


@pytest.mark.parametrize("group", [8, 16])
def test_small_batch_kernel_group_widths_agree(monkeypatch, group):
    """The small-batch kernel puts 32, 16 or 8 lanes on a query depending on the batch size; forced widths must return the
    same bits as the default choice -- on a tie-heavy lattice cloud (lowest index among exact ties) and on small trees."""
    monkeypatch.setenv("PC_COOP_GROUP", str(group))
    forced = PointCloudIndex(max_points=0, device=0)
    monkeypatch.delenv("PC_COOP_GROUP")
    auto = PointCloudIndex(max_points=0, device=0)
    P = PcRadiusParams.make(0.0, 5.0, 30.0, (0.0, 0.0, 2.0))            # simulation.launch parameters: long free-space searches
    try:
        pts, half = synth.forest_cloud(120_000, seed=6, variant="L", return_half=True)
        q = synth.rrt_queries(20_000, half, seed=4)
        q[::2] = np.round(q[::2] / 0.05) * 0.05                          # half of the queries snapped to the lattice: exact ties
        for n in (120_000, 5, 9, 40, 130, 700):
            forced.build(pts[:n]); auto.build(pts[:n])
            for m in (1, 333, 4096, 20_000):
                i1, d1 = forced.nearest(q[:m]); i2, d2 = auto.nearest(q[:m])
                assert (i1 == i2).all() and (d1 == d2).all()
                r1, j1 = forced.radius(q[:m], P, want_idx=True); r2, j2 = auto.radius(q[:m], P, want_idx=True)
                assert (r1 == r2).all() and (j1 == j2).all()
        forced.build(pts[:30_000])
        check_lowest_index_everywhere(pts[:30_000], q[:3000], forced.nearest(q[:3000])[0])     # independent of the default kernel
    finally:
        forced.close(); auto.close()
