"""Shared helpers of the RRT* driver tests: input file, output parsing, corridor validation against the oracle."""
import struct

import numpy as np

import oracle

PRM = (0.6, 0.25, 1.5, 30.0)          # safety_margin, search_margin, max_radius, sensing range (clean_demo.launch:31-34)
START, GOAL = (-10.0, -10.0, 2.0), (9.0, 9.0, 2.0)


def write_input(path, pts, half, max_iter, K, pts2=None, refine_iter=0):
    """pts2: an optional SECOND cloud message -> the clients also run evaluate() and refine() + evaluate() on it."""
    box = (-half, half, -half, half, 0.0, 4.0)
    with open(path, "wb") as f:
        f.write(struct.pack("<qqq4d3d3d6d2d", len(pts), max_iter, K, *PRM, *START, *GOAL, *box, 0.3, 0.1))
        f.write(np.ascontiguousarray(pts[:, :3], np.float32).tobytes())
        if pts2 is not None:
            f.write(struct.pack("<qq", len(pts2), refine_iter))
            f.write(np.ascontiguousarray(pts2[:, :3], np.float32).tobytes())


def blocked_cloud(pts, centres, seed=0):
    """The first cloud plus a small blob of new obstacle points ON each given sphere centre (a newly observed obstacle)."""
    rng = np.random.default_rng(seed)
    blobs = [np.asarray(c, np.float64) + np.concatenate([np.zeros((1, 3)), rng.normal(0, 0.05, (63, 3))]) for c in centres]
    return np.concatenate([pts[:, :3].astype(np.float32)] + [b.astype(np.float32) for b in blobs])


def read_records(path, n_rec):
    raw = open(path, "rb").read()
    off, out = 0, []
    for _ in range(n_rec):
        k, nodes, cq = struct.unpack_from("<qqq", raw, off)
        ms, = struct.unpack_from("<d", raw, off + 24)
        off += 32
        p = np.frombuffer(raw, np.float64, 3 * k, off).reshape(k, 3); off += 24 * k
        r = np.frombuffer(raw, np.float64, k, off); off += 8 * k
        out.append(dict(k=k, nodes=nodes, cloud_queries=cq, ms=ms, path=p, radius=r))
    assert off == len(raw)
    return out


def validate_corridor(rec, pts, float_centres=False):
    """A corridor is a chain of obstacle-free spheres from the start to the goal."""
    p, r = rec["path"], rec["radius"]
    assert rec["k"] >= 2, "no corridor"
    assert np.allclose(p[0], START)                                               # the root sphere sits on the start
    assert np.linalg.norm(p[-1] - np.array(GOAL)) + 0.1 < r[-1]                   # checkEnd (corridor_finder.cpp:418-426)
    gaps = np.linalg.norm(np.diff(p, axis=0), axis=1)
    assert (gaps <= r[:-1] + r[1:]).all()                                         # consecutive spheres intersect
    assert (r >= np.float32(PRM[0])).all() and (r <= PRM[2]).all()                # safety_margin <= radius <= max_radius
    ko = oracle.KdOracle().build(pts)
    P = oracle.RadiusParams.make(PRM[1], PRM[2], PRM[3], START)
    for c, rad in zip(p, r):                                                      # every radius is the exact radiusSearch value
        cc = c.astype(np.float32).astype(np.float64) if float_centres else c
        assert np.float32(ko.radius_search(P, cc)[0]) == np.float32(rad)
