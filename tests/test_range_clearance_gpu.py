"""GPU parity tests for pc_range_batch (kd_nearest_range3) and pc_clearance_batch (checkSafeTrajectory)."""
import numpy as np
import pytest

import oracle
from pointcloudtraj_b200 import PcError, PcRadiusParams, PointCloudIndex, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ix():
    h = PointCloudIndex(max_points=1 << 18, device=0)
    yield h
    h.close()


def _brute_range(pts, q, r):
    """Exact fp64 sets {i : d2(i) <= r*r} in the reference's operation order, ascending index."""
    P = pts.astype(np.float64)
    out = []
    r = np.broadcast_to(np.asarray(r, np.float64), (len(q),))
    for k in range(len(q)):
        d = P - q[k].astype(np.float64)
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        out.append(np.nonzero(d2 <= r[k] * r[k])[0])
    return out


def _lists(off, idx):
    return [idx[off[k]:off[k + 1]] for k in range(len(off) - 1)]


@pytest.mark.parametrize("name", ["forest_lattice", "forest_jitter", "uniform"])
def test_range_golden(ix, golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    pts, q, r = g["pts"], g["range_q"], float(g["range_r"])
    ix.build(pts)
    off, idx = ix.range(q, r)
    gpu = _lists(off, idx)
    ref = _lists(g["range_off"], g["range_idx"])
    brute = _brute_range(pts, q, r)
    for k in range(len(q)):
        assert (np.diff(gpu[k]) > 0).all()                        # ascending original index, no duplicates
        assert (gpu[k] == brute[k]).all() if len(gpu[k]) == len(brute[k]) else False
        # the reference may miss points at exactly `range` across a split plane (kdtree.c:283), never more
        assert np.isin(ref[k], gpu[k]).all()
    if name != "forest_lattice":
        assert (off == g["range_off"]).all()                      # tie-free inputs: identical sets
        assert all((np.sort(ref[k]) == gpu[k]).all() for k in range(len(q)))


def test_range_known_answers(ix):
    # points (0.5,5,5), (0.5,0,0); the point at exactly range 1 is found from BOTH sides here
    ix.build(np.array([[0.5, 5, 5], [0.5, 0, 0]], np.float32))
    off, idx = ix.range(np.array([[-0.5, 0, 0], [1.5, 0, 0], [9, 9, 9]], np.float32), 1.0)
    assert off.tolist() == [0, 1, 2, 2] and idx.tolist() == [1, 1]
    # empty index, empty batch
    ix.build(np.zeros((0, 3), np.float32))
    off, idx = ix.range(np.zeros((3, 3), np.float32), 2.0)
    assert off.tolist() == [0, 0, 0, 0] and idx.size == 0
    off, idx = ix.range(np.zeros((0, 3), np.float32), 2.0)
    assert off.tolist() == [0]


def test_range_forest_vs_oracle_and_capacity(ix):
    pts, half = synth.forest_cloud(120_000, seed=6, variant="J", return_half=True)
    q = synth.rrt_queries(3000, half, seed=2)
    radii = np.random.default_rng(0).uniform(0.2, 3.0, len(q))          # planner uses 2 * radius <= 3 m
    ix.build(pts)
    off, idx = ix.range(q, radii)
    ko = oracle.KdOracle().build(pts, np.random.default_rng(1).permutation(len(pts)))
    roff, ridx = ko.range(q, radii)
    assert (off == roff).all()
    gl, rl = _lists(off, idx), _lists(roff, ridx)
    assert all((np.sort(rl[k]) == gl[k]).all() for k in range(len(q)))
    assert off[-1] > 100_000                                            # a real workload, not an empty check
    # duplicates of the last point pad the last leaf: they must not be reported twice
    last = pts[-1:]
    o2, i2 = ix.range(last, 1e-3)
    assert (np.diff(i2) > 0).all() and (len(pts) - 1) in i2
    # capacity overflow -> PC_ECAP, offsets still complete
    with pytest.raises(PcError) as e:
        ix.range(q, radii, cap=1000)
    assert e.value.code == -4


def test_range_long_lists(ix):
    """Lists beyond the in-warp sort (1024 hits) are sorted by one CTA each (up to 32768 hits), beyond that by the slow path:
    every size class gives the brute-force set in ascending order."""
    pts, half = synth.forest_cloud(150_000, seed=6, variant="J", return_half=True)
    ix.build(pts)
    q = synth.rrt_queries(24, half, seed=5)
    radii = np.array([0.5, 2.0, 4.0, 6.0, 9.0, 12.0, 40.0, 5.0] * 3)
    off, idx = ix.range(q, radii)
    brute = _brute_range(pts, q, radii)
    sizes = np.diff(off)
    assert (sizes > 32768).any() and ((sizes > 1024) & (sizes <= 32768)).any() and (sizes <= 1024).any()
    for k in range(len(q)):
        assert len(brute[k]) == sizes[k] and (idx[off[k]:off[k + 1]] == brute[k]).all()


# ---- clearance ---------------------------------------------------------------------------------------
def _clearance_oracle(ko, P, tr, t_now, horizon, dt=0.02):
    first, order, T, off, coef = tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"]
    res = []
    for t in range(len(first) - 1):
        segs = range(first[t], first[t + 1])
        ld = 3 * (int(order[first[t]:first[t + 1]].max()) + 1) if len(segs) else 1
        mat = np.zeros((len(segs), ld))
        for r_, s in enumerate(segs):
            c = coef[off[s]:off[s + 1]]
            mat[r_, : len(c)] = c
        res.append(ko.check_safe_trajectory(P, order[first[t]:first[t + 1]], T[first[t]:first[t + 1]], mat,
                                            t_now=float(t_now[t]), stop_time=horizon, dt=dt, cap=8192))
    return res


@pytest.mark.parametrize("horizon,n_traj", [(2.0, 300), (20.0, 60)])
def test_clearance_vs_oracle(ix, horizon, n_traj):
    pts, half = synth.forest_cloud(150_000, seed=6, variant="J", return_half=True)
    ix.build(pts)
    tr = synth.bezier_trajectories(n_traj, half * 0.9, seed=3)
    rng = np.random.default_rng(5)
    t_now = np.where(rng.uniform(size=n_traj) < 0.5, 0.0, rng.uniform(0, 4.0, n_traj))
    start = (0.0, 0.0, 2.0)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, start)
    fh, mr, ns = ix.clearance(tr["traj_first_seg"], tr["seg_order"], tr["seg_T"], tr["seg_coef_off"], tr["coef"], P,
                              t_now=t_now, dt=0.02, horizon=horizon)
    ko = oracle.KdOracle().build(pts, np.random.default_rng(1).permutation(len(pts)))
    ref = _clearance_oracle(ko, oracle.RadiusParams.make(0.25, 1.5, 30.0, start), tr, t_now, horizon)
    r_ns = np.array([r["n_samples"] for r in ref])
    r_fh = np.array([r["first_hit"] for r in ref])
    r_mr = np.array([r["min_radius"] for r in ref])
    assert (ns == r_ns).all()                                     # the repeated-addition time walk is reproduced exactly
    assert ns.max() >= min(int(horizon / 0.02) - 1, 400) and (r_fh >= 0).any() and (r_fh < 0).any()
    # exact: the kernel's powers are the correctly rounded u^j (double-double), like libm's pow in the oracle, so the float32
    # sample positions, every radius and therefore the first colliding sample are identical (DESIGN.md "Clearance")
    fin = np.isfinite(r_mr)                                         # trajectories without samples report +inf
    assert (np.isinf(mr) == ~fin).all()
    assert (mr[fin] == r_mr[fin].astype(np.float32)).all()
    assert (fh == r_fh).all()


def test_clearance_edge_cases(ix):
    pts = np.array([[5.0, 0, 1.0]], np.float32)
    ix.build(pts)
    P = PcRadiusParams.make(0.25, 1.5, 30.0, (0, 0, 1))
    T, n = 4.0, 4
    xs = np.linspace(0, 8, n + 1) / T
    coef = np.concatenate([xs, np.zeros(n + 1), np.full(n + 1, 1.0 / T)])
    first = np.array([0, 1, 1], np.int32)                          # second trajectory has no segments
    fh, mr, ns = ix.clearance(first, [n], [T], [0, len(coef)], coef, P, t_now=[0.0, 0.0], horizon=10.0)
    ko = oracle.KdOracle().build(pts)
    ref = ko.check_safe_trajectory(oracle.RadiusParams.make(0.25, 1.5, 30.0, (0, 0, 1)), [n], [T], coef[None, :], 0.0, 10.0)
    assert ns.tolist() == [ref["n_samples"], 0] == [200, 0]
    assert fh.tolist() == [ref["first_hit"], -1]
    assert mr[0] == np.float32(ref["min_radius"]) and np.isinf(mr[1])
    # empty cloud: every sample gets max_radius - search_margin (corridor_finder.cpp:118-120), no collision
    ix.build(np.zeros((0, 3), np.float32))
    fh, mr, ns = ix.clearance(first[:2], [n], [T], [0, len(coef)], coef, P, horizon=1.0)
    assert fh.tolist() == [-1] and mr.tolist() == [1.25]
    # bad arguments are rejected, not executed
    with pytest.raises(PcError):
        ix.clearance(first[:2], [13], [T], [0, len(coef)], coef, P)
    with pytest.raises(PcError):
        ix.clearance(first[:2], [n], [T], [0, len(coef)], coef, P, dt=0.0)


def test_sphere_gather_matches_brute_force(ix):
    """pc_sphere_gather: the LiDAR-mode observation (camera_sensor.cpp:133-145) = all points within max_dist of the sensor."""
    pts, half = synth.forest_cloud(250_000, seed=4, variant="J", return_half=True)
    ix.build(pts)
    for center, radius in (((1.0, -2.0, 2.0), 20.0), ((half, half, 0.0), 6.5), ((0.0, 0.0, 100.0), 5.0), ((0.0, 0.0, 2.0), 1e4)):
        got = ix.sphere_gather(center, radius)
        c = np.array(center, np.float32).astype(np.float64)
        d = pts.astype(np.float64) - c
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        want = np.nonzero(d2 <= radius * radius)[0]
        assert got.dtype == np.int32 and (got == want).all() if len(got) == len(want) else False
    ix.build(np.zeros((0, 3), np.float32))
    assert ix.sphere_gather((0, 0, 0), 5.0).size == 0


def test_clearance_flat_path_equals_warp_per_trajectory_path(ix, tmp_path):
    """The flat path (schedule / eval / finish kernels) and the one-warp-per-trajectory kernel (PC_CLEARANCE_FLAT=0, also the
    path for schedules beyond 2^27 samples) return identical arrays; the latter runs in a subprocess (the switch is read once)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from pointcloudtraj_b200 import PcRadiusParams, PointCloudIndex, synth\n"
        "pts, half = synth.forest_cloud(150_000, seed=6, variant='J', return_half=True)\n"
        "tr = synth.bezier_trajectories(700, half * 0.9, seed=9)\n"
        "tn = np.where(np.arange(700) % 3 == 0, 0.7, 0.0)\n"
        "P = PcRadiusParams.make(0.25, 1.5, 30.0, (0.0, 0.0, 2.0))\n"
        "ix = PointCloudIndex(max_points=len(pts), device=0); ix.build(pts)\n"
        "fh, mr, ns = ix.clearance(tr['traj_first_seg'], tr['seg_order'], tr['seg_T'], tr['seg_coef_off'], tr['coef'], P, t_now=tn, dt=0.02, horizon=7.0)\n"
        "np.savez(sys.argv[1], fh=fh, mr=mr, ns=ns)\n")
    outs = []
    for flat in ("1", "0"):
        path = str(tmp_path / f"clr{flat}.npz")
        env = dict(os.environ, PC_CLEARANCE_FLAT=flat)
        p = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        outs.append(np.load(path))
    a, b = outs
    assert (a["ns"] == b["ns"]).all() and (a["fh"] == b["fh"]).all() and (a["mr"] == b["mr"]).all()
    assert a["ns"].max() > 300 and (a["fh"] >= 0).any() and (a["fh"] < 0).any()
