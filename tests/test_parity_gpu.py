"""GPU parity tests: libpcindex.so (through the C ABI) against the CPU oracle on the same seeded inputs,
against the committed golden vectors of the unmodified reference, and -- at full BASELINE sizes --
through size-independent properties."""
import numpy as np
import pytest

import oracle
from pointcloudtraj_b200 import (PC_QUERY_SORTED, PC_QUERY_UNSORTED, PC_RADIUS_BOUNDED, PC_RADIUS_FULL_NN,
                                 PcRadiusParams, PointCloudIndex, synth)
from parity import check_lowest_index_everywhere, check_nearest

pytestmark = pytest.mark.gpu

CLEAN_DEMO = dict(search_margin=0.25, max_radius=1.5, sample_range=30.0)     # clean_demo.launch:31-34
SIMULATION = dict(search_margin=0.0, max_radius=5.0, sample_range=30.0)      # simulation.launch:24-27


@pytest.fixture(scope="module")
def ix():
    h = PointCloudIndex(max_points=1 << 20, device=0)
    yield h
    h.close()


def _oracle(pts, seed=0):
    return oracle.KdOracle().build(pts, np.random.default_rng(seed).permutation(len(pts)))


# ---- golden vectors of the unmodified reference ----------------------------------------------------
@pytest.mark.parametrize("name", ["forest_lattice", "forest_jitter", "uniform"])
def test_golden_nearest(ix, golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    ix.build(g["pts"])
    for flags in (PC_QUERY_UNSORTED, PC_QUERY_SORTED):
        idx, d2 = ix.nearest(g["q"], flags=flags)
        check_nearest(g["pts"], g["q"], idx, d2, g["nn_idx"].astype(np.int64), g["nn_d2"])


def test_known_answers(ix, golden_dir):
    z = np.zeros((1, 3), np.float32)
    # two equidistant points: lowest index (documented tie rule; the reference returns the first inserted)
    ix.build(np.array([[1, 0, 0], [-1, 0, 0]], np.float32))
    assert ix.nearest(z)[0].tolist() == [0]
    ix.build(np.array([[-1, 0, 0], [1, 0, 0]], np.float32))
    assert ix.nearest(z)[0].tolist() == [0]
    # three duplicates -> index 0
    ix.build(np.array([[2, 2, 2]] * 3, np.float32))
    i, d = ix.nearest(np.array([[2.5, 2, 2]], np.float32))
    assert i.tolist() == [0] and d.tolist() == [0.25]
    # query coincident with a point
    pts = synth.uniform_cloud(1000, seed=5)
    ix.build(pts)
    i, d = ix.nearest(pts[123:124])
    assert i.tolist() == [123] and d.tolist() == [0.0]


def test_empty_and_single(ix):
    q = synth.rrt_queries(100, 5.0, seed=1)
    ix.build(np.zeros((0, 3), np.float32))
    assert ix.size == 0
    i, d = ix.nearest(q)
    assert (i == -1).all() and np.isinf(d).all()                                  # kd_nearest3 -> NULL
    P = PcRadiusParams.make(**CLEAN_DEMO)
    r, i = ix.radius(q, P, want_idx=True)
    assert (r == np.float32(1.5 - 0.25)).all() and (i == -1).all()                # corridor_finder.cpp:118-120
    i, d = ix.nearest(np.zeros((0, 3), np.float32))
    assert i.shape == (0,)
    one = np.array([[0.5, -0.25, 1.0]], np.float32)
    ix.build(one)
    i, d = ix.nearest(q)
    assert (i == 0).all()
    ref = oracle.pair_d2(one, q, np.zeros(len(q), np.int64))
    assert (d == ref.astype(np.float32)).all()


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 257, 1000, 4097])
def test_small_sizes_every_tree_shape(ix, n):
    pts = synth.uniform_cloud(n, half=4.0, seed=n)
    q = synth.rrt_queries(512, 5.0, seed=n + 1, z=(-1.0, 9.0))
    ix.build(pts)
    idx, d2 = ix.nearest(q)
    check_lowest_index_everywhere(pts, q, idx)
    ridx, rd2 = _oracle(pts).nearest(q)
    check_nearest(pts, q, idx, d2, ridx, rd2)


# ---- seeded clouds vs the oracle (C1-sized) ----------------------------------------------------------
@pytest.mark.parametrize("variant,lat", [("L", 0.5), ("J", 0.0)])
def test_forest_200k(ix, variant, lat):
    pts, half = synth.forest_cloud(200_000, seed=6, variant=variant, return_half=True)
    q = synth.rrt_queries(100_000, half, seed=0, lattice_frac=lat)
    ix.build(pts)
    idx, d2 = ix.nearest(q)
    ridx, rd2 = _oracle(pts).nearest(q)
    n_tied = check_nearest(pts, q, idx, d2, ridx, rd2)
    if variant == "J":
        assert n_tied == 0 and (idx == ridx).all()       # tie-free: bit-exact indices
    idx2, d22 = ix.nearest(q, flags=PC_QUERY_UNSORTED)
    assert (idx2 == idx).all() and (d22 == d2).all()     # batch ordering does not change results


def test_uniform_cloud_and_strides(ix):
    pts = synth.uniform_cloud(150_000, half=20.0, seed=9)
    q = synth.rrt_queries(50_000, 22.0, seed=10, z=(-1, 9))
    ix.build(pts)
    idx, d2 = ix.nearest(q)
    ridx, rd2 = _oracle(pts).nearest(q)
    check_nearest(pts, q, idx, d2, ridx, rd2)
    # pcl::PointXYZ layout: x, y, z, pad at 16-byte stride, for both cloud and queries
    pts4 = np.concatenate([pts, np.full((len(pts), 1), np.nan, np.float32)], 1)
    q4 = np.concatenate([q, np.ones((len(q), 1), np.float32)], 1)
    ix.build(pts4)
    idx4, d24 = ix.nearest(q4)
    assert (idx4 == idx).all() and (d24 == d2).all()


@pytest.mark.parametrize("params,seed", [(CLEAN_DEMO, 0), (SIMULATION, 1)])
def test_radius_batch(ix, params, seed):
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
    q = synth.rrt_queries(60_000, half * 2.5, seed=seed)       # beyond the map: exercises the sensing-range early-out
    start = (1.0, -2.0, 2.0)
    ix.build(pts)
    P = PcRadiusParams.make(start=start, **params)
    r_ref, i_ref = _oracle(pts).radius_batch(oracle.RadiusParams.make(start=start, **params), q)
    r32 = r_ref.astype(np.float32)
    # unbounded search: the index is the true nearest point wherever a query was issued
    r, i = ix.radius(q, P, flags=PC_RADIUS_FULL_NN, want_idx=True)
    assert (r == r32).all()
    assert (i == i_ref).all()
    # bounded search (default): same radii bit for bit; index only where the radius is not clamped
    r, i = ix.radius(q, P, flags=PC_RADIUS_BOUNDED, want_idx=True)
    assert (r == r32).all()
    clamped = ~(r_ref < params["max_radius"])
    assert (i[clamped] == -1).all() and (i[~clamped] == i_ref[~clamped]).all()
    assert (r_ref == params["max_radius"] - params["search_margin"]).any() and clamped.any() and (~clamped).any()
    if params["search_margin"] > 0:
        assert (r_ref < 0).any()                                 # collisions exercised
    col = ix.check_traj_pt_col(q, P)
    assert (col == (r_ref < 0)).all()                            # checkTrajPtCol, corridor_finder.cpp:412-416


def test_device_pointers_match_host(ix):
    torch = pytest.importorskip("torch")
    pts, half = synth.forest_cloud(100_000, seed=3, variant="J", return_half=True)
    q = synth.rrt_queries(70_000, half, seed=4)
    ix.build(pts)
    idx_h, d2_h = ix.nearest(q)
    tp = torch.from_numpy(pts).cuda()
    tq = torch.from_numpy(q).cuda()
    h2 = PointCloudIndex(max_points=0, device=0, stream=torch.cuda.current_stream().cuda_stream)
    h2.build(tp)
    idx_d, d2_d = h2.nearest(tq)
    P = PcRadiusParams.make(**CLEAN_DEMO)
    r_d = h2.radius(tq, P)
    torch.cuda.synchronize()
    assert (idx_d.cpu().numpy() == idx_h).all() and (d2_d.cpu().numpy() == d2_h).all()
    assert (r_d.cpu().numpy() == ix.radius(q, P)).all()
    h2.close()


def test_rebuild_reuses_handle(ix):
    # LiDAR-stream pattern (C3): the same handle is rebuilt every frame
    for frame in range(4):
        pts = synth.uniform_cloud(30_000 + 1000 * frame, half=10.0, seed=100 + frame)
        q = synth.rrt_queries(5000, 10.0, seed=frame)
        ix.build(pts)
        assert ix.size == len(pts)
        idx, _ = ix.nearest(q)
        check_lowest_index_everywhere(pts, q, idx)


# ---- full-size properties (C2: 1M-point map) ---------------------------------------------------------
def test_one_million_points_properties(ix):
    pts, half = synth.forest_cloud(1_000_000, seed=1, variant="J", return_half=True)
    ix.build(pts)
    assert ix.size == 1_000_000
    # (1) every point is its own nearest neighbour at distance 0 (covers every leaf, checks the permutation)
    sel = np.random.default_rng(0).permutation(len(pts))[:300_000]
    idx, d2 = ix.nearest(pts[sel])
    assert (d2 == 0).all()
    same = idx == sel
    if not same.all():      # exact duplicates in the cloud: lowest index wins
        assert (pts[idx[~same]] == pts[sel[~same]]).all() and (idx[~same] < sel[~same]).all()
    # (2) oracle on a sample, (3) batch order independence, (4) radius == epilogue of nearest
    q = synth.rrt_queries(200_000, half, seed=7)
    idx, d2 = ix.nearest(q)
    sub = slice(0, 20_000)
    ridx, rd2 = _oracle(pts).nearest(q[sub])
    check_nearest(pts, q[sub], idx[sub], d2[sub], ridx, rd2)
    idx_u, d2_u = ix.nearest(q, flags=PC_QUERY_UNSORTED)
    assert (idx_u == idx).all() and (d2_u == d2).all()
    P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN_DEMO)
    r = ix.radius(q, P)
    exact = oracle.pair_d2(pts, q, idx.astype(np.int64))
    expect = np.minimum(np.sqrt(exact) - 0.25, 1.5)
    far = np.sqrt(((q.astype(np.float64) - np.array([0, 0, 2.0])) ** 2).sum(1)) > 31.5
    expect[far] = 1.25
    assert (r == expect.astype(np.float32)).all()


def test_five_million_points_64bit_keys():
    """C4-sized cloud: above 4 Mi points the build switches to 63-bit Hilbert keys (uint64 radix sort, 5 passes)."""
    pts, half = synth.forest_cloud(5_000_000, seed=2, variant="J", return_half=True)
    h = PointCloudIndex(max_points=len(pts), device=0)
    try:
        h.build(pts)
        assert h.size == len(pts)
        v = h.view()
        assert v.n_nodes == len(pts) - 1 and v.root == 0 and v.records and v.points == v.records + 64 * v.n_nodes
        assert np.allclose(np.array(v.bbox_lo[:]), pts.min(0)) and np.allclose(np.array(v.bbox_hi[:]), pts.max(0))
        sel = np.random.default_rng(1).permutation(len(pts))[:400_000]
        idx, d2 = h.nearest(pts[sel])
        assert (d2 == 0).all() and (idx == sel).all()            # every sampled point finds itself
        q = synth.rrt_queries(300_000, half, seed=3)
        idx, d2 = h.nearest(q)
        bi, bd, ties = oracle.brute_nearest(pts, q[:1500])       # independent fp64 brute force
        assert (idx[:1500] == bi).all() and (d2[:1500] == bd.astype(np.float32)).all()
        idx_u, d2_u = h.nearest(q, flags=PC_QUERY_UNSORTED)
        assert (idx_u == idx).all() and (d2_u == d2).all()
        P = PcRadiusParams.make(start=(0, 0, 2), **SIMULATION)
        r = h.radius(q, P, flags=PC_RADIUS_FULL_NN)
        exact = oracle.pair_d2(pts, q, idx.astype(np.int64))
        expect = np.minimum(np.sqrt(exact) - 0.0, 5.0)
        far = np.sqrt(((q.astype(np.float64) - np.array([0, 0, 2.0])) ** 2).sum(1)) > 35.0
        expect[far] = 5.0
        assert (r == expect.astype(np.float32)).all()
    finally:
        h.close()


def test_async_host_batches_match_blocking_calls(ix):
    """PC_HOST_ASYNC: batches enqueued back to back on the internal streams, one wait; same results as blocking calls,
    and a rebuild issued while batches are in flight waits for them."""
    torch = pytest.importorskip("torch")
    pts, half = synth.forest_cloud(120_000, seed=5, variant="J", return_half=True)
    ix.build(pts)
    P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN_DEMO)
    batches = [torch.from_numpy(synth.rrt_queries(90_000 + 1000 * k, half, seed=20 + k)).pin_memory() for k in range(5)]
    outs = [torch.full((len(b),), -7.0).pin_memory() for b in batches]
    for b, o in zip(batches, outs):
        ix.radius_async(b.numpy(), o.numpy(), P)
    ix.build(pts[::-1].copy())                     # must not disturb the batches still in flight
    ix.sync()
    ix.build(pts)
    for b, o in zip(batches, outs):
        assert (o.numpy() == ix.radius(b.numpy(), P)).all()


@pytest.mark.parametrize("n_ranks", [2, 8])
def test_spatial_batch_sharding(ix, n_ranks):
    """pc_batch_shard: every rank gets the same batch and answers its stretch of the Hilbert curve.  Emulated on one GPU
    by switching the rank of one handle: every query is answered by exactly one rank, with the unsharded result, and
    the shares are balanced."""
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
    q = synth.rrt_queries(400_000, half, seed=11)
    P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN_DEMO)
    ix.build(pts)
    ref_idx, ref_d2 = ix.nearest(q)
    ref_r = ix.radius(q, P)
    owners = np.zeros(len(q), np.int32)
    shares = []
    got_idx = np.full(len(q), -9, np.int32)
    got_r = np.full(len(q), np.nan, np.float32)
    try:
        for r in range(n_ranks):
            ix.batch_shard(r, n_ranks)
            idx, d2 = ix.nearest(q)
            mine = idx != PointCloudIndex.NOT_MINE_IDX
            assert (np.isnan(d2) == ~mine).all()
            owners += mine
            shares.append(int(mine.sum()))
            got_idx[mine] = idx[mine]
            assert (d2[mine] == ref_d2[mine]).all()
            rad = ix.radius(q, P)
            have = ~np.isnan(rad)
            got_r[have] = rad[have]
    finally:
        ix.batch_shard(0, 1)
    assert (owners == 1).all()                                       # a partition of the batch
    assert (got_idx == ref_idx).all() and (got_r == ref_r).all()
    # balanced: the cells are dealt by a hash, so the spread shrinks with the number of cells per rank (about 250 / n_ranks
    # occupied cells for this small batch; the bench's 10^7-query batches have tens of thousands per rank)
    assert max(shares) <= (1.25 if n_ranks == 2 else 1.4) * len(q) / n_ranks
    idx, _ = ix.nearest(q)
    assert (idx == ref_idx).all()                                     # sharding switched off again


def test_shard_ownership_is_rank_invariant_right_after_an_async_build():
    """Two replicas built independently from device memory: one is queried at once (its host copy of the bounding box has
    not arrived yet), the other after a sync.  The owner of every query is computed on the device from rank-invariant inputs,
    so the two still partition the batch."""
    torch = pytest.importorskip("torch")
    pts, half = synth.forest_cloud(300_000, seed=6, variant="J", return_half=True)
    q = torch.from_numpy(synth.rrt_queries(500_000, half, seed=12)).cuda()
    t_pts = torch.from_numpy(pts).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    a = PointCloudIndex(max_points=len(pts), stream=stream)
    b = PointCloudIndex(max_points=len(pts), stream=stream)
    try:
        b.build(t_pts)
        b.sync()
        b.batch_shard(1, 2)
        a.batch_shard(0, 2)
        a.build(t_pts)                       # asynchronous: returns before the bounding box reaches the host
        ia, _ = a.nearest(q)
        ib, _ = b.nearest(q)
        torch.cuda.synchronize()
        mine_a = (ia != PointCloudIndex.NOT_MINE_IDX).cpu().numpy()
        mine_b = (ib != PointCloudIndex.NOT_MINE_IDX).cpu().numpy()
        assert (mine_a ^ mine_b).all()
        assert 0.4 < mine_a.mean() < 0.6
    finally:
        a.close()
        b.close()


def test_async_device_batches(ix):
    """PC_DEVICE_ASYNC: device batches rotate over the internal streams; results valid after sync and equal to PC_DEVICE."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    from pointcloudtraj_b200 import _lib as L
    pts, half = synth.forest_cloud(150_000, seed=8, variant="J", return_half=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        h = PointCloudIndex(max_points=len(pts), device=0, stream=s.cuda_stream)
        h.build(torch.from_numpy(pts).cuda())
        P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN_DEMO)
        qs = [torch.from_numpy(synth.rrt_queries(100_000 + 7 * k, half, seed=40 + k)).cuda() for k in range(6)]
        ref = [h.radius(q, P) for q in qs]
        outs = [torch.full((len(q),), -5.0, device="cuda") for q in qs]
        for q, o in zip(qs, outs):
            rc = h._L.pc_radius_batch(h._h, C.c_void_p(q.data_ptr()), len(q), 3, L.PC_DEVICE_ASYNC, 0, C.byref(P), C.c_void_p(o.data_ptr()), None)
            assert rc == 0
        h.sync()
        torch.cuda.synchronize()
        for r, o in zip(ref, outs):
            assert (r == o).all().item()
        h.close()


def test_batch_shard_exact_share_sizing(ix):
    """Large PC_DEVICE batches in pc_batch_shard mode read the size of the rank's share back and size the sort and search
    launches by it (pc_share_size): the shares still partition the batch with the unsharded results -- on the radix-ordered
    path (nearest), the binned path (bounded radius), and when a rank's share is empty or the whole batch."""
    torch = pytest.importorskip("torch")
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
    q = torch.from_numpy(synth.rrt_queries(1_300_000, half, seed=13)).cuda()      # >= PC_SHARD_EXACT_MIN (1 Mi)
    P = PcRadiusParams.make(start=(0, 0, 2), **CLEAN_DEMO)
    ix.build(pts)
    ref_idx, ref_d2 = ix.nearest(q)
    ref_r = ix.radius(q, P)
    n_ranks = 3
    owners = torch.zeros(len(q), dtype=torch.int32, device="cuda")
    got_idx = torch.full((len(q),), -9, dtype=torch.int32, device="cuda")
    got_r = torch.full((len(q),), float("nan"), device="cuda")
    try:
        for r in range(n_ranks):
            ix.batch_shard(r, n_ranks)
            idx, d2 = ix.nearest(q)
            mine = idx != PointCloudIndex.NOT_MINE_IDX
            owners += mine
            got_idx[mine] = idx[mine]
            assert bool((d2[mine] == ref_d2[mine]).all())
            rad = ix.radius(q, P)
            have = ~torch.isnan(rad)
            got_r[have] = rad[have]
        assert bool((owners == 1).all())
        assert bool((got_idx == ref_idx).all()) and bool((got_r == ref_r).all())
        # one cell holds the whole batch: one rank owns everything, the others nothing (zero-size launches are skipped)
        one = q[:1].repeat(1_100_000, 1).contiguous()
        ix.batch_shard(0, 1)
        want_idx, _ = ix.nearest(one)
        total = 0
        for r in range(4):
            ix.batch_shard(r, 4)
            idx, _ = ix.nearest(one)
            mine = idx != PointCloudIndex.NOT_MINE_IDX
            k = int(mine.sum().item())
            assert k in (0, len(one))
            total += k
            if k:
                assert bool((idx == want_idx).all())
        assert total == len(one)
    finally:
        ix.batch_shard(0, 1)


def test_incoherent_packets_walk_alone(ix):
    """A dense batch with a hole: the queries fill two boxes at opposite ends of the map, so the curve order jumps between
    them and the packets cut across the jump hold queries tens of metres apart.  An unbounded packet walk of such a packet can
    run for milliseconds; pc_query_packet_kernel puts it aside and pc_query_deferred_kernel answers its queries one by one --
    with the same results as the unordered batch and as the reference kd-tree."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    pts, half = synth.forest_cloud(200_000, seed=6, variant="J", return_half=True)
    rng = np.random.default_rng(31)
    n = 150_011                                      # not a multiple of the packet size
    a = rng.uniform([-half, -half, 0.6], [-half / 2, -half / 2, 4.0], (n, 3))
    b = rng.uniform([half / 2, half / 2, 0.6], [half, half, 4.0], (n, 3))
    q = np.concatenate([a, b]).astype(np.float32)
    q = q[rng.permutation(len(q))]
    ix.build(pts)
    tq = torch.from_numpy(q).cuda()
    idx_s, d2_s = ix.nearest(tq, flags=PC_QUERY_SORTED)
    deferred = C.c_int64(-1)
    assert ix._L.pc_profile_last_deferred_packets(ix._h, C.byref(deferred)) == 0
    assert deferred.value >= 1, "the batch was built to have packets across the gap"
    assert deferred.value < 200, "only the packets at the jumps of the order are put aside"
    idx_u, d2_u = ix.nearest(tq, flags=PC_QUERY_UNSORTED)
    assert bool((idx_s == idx_u).all()) and bool((d2_s == d2_u).all())
    ko = oracle.KdOracle().build(pts, np.random.default_rng(0).permutation(len(pts)))
    sub = rng.choice(len(q), 20_000, replace=False)
    ridx, rd2 = ko.nearest(q[sub])
    check_nearest(pts, q[sub], idx_s.cpu().numpy()[sub], d2_s.cpu().numpy()[sub], ridx, rd2)
    # the unbounded radius search (PC_RADIUS_FULL_NN) takes the same route
    P = PcRadiusParams.make(start=(0, 0, 2), search_margin=0.25, max_radius=1.5, sample_range=-1.0)
    rad_s = ix.radius(tq, P, flags=PC_RADIUS_FULL_NN | PC_QUERY_SORTED)
    assert ix._L.pc_profile_last_deferred_packets(ix._h, C.byref(deferred)) == 0 and deferred.value >= 1
    rad_u = ix.radius(tq, P, flags=PC_RADIUS_FULL_NN | PC_QUERY_UNSORTED)
    assert bool((rad_s == rad_u).all())
    rrad, _ = ko.radius_batch(oracle.RadiusParams.make(0.25, 1.5, -1.0, (0, 0, 2)), q[sub])
    assert (rad_s.cpu().numpy()[sub] == rrad.astype(np.float32)).all()
