"""Synthetic clouds, RRT* sample queries and Bezier trajectories for tests and bench.

Everything here is numpy on the host; PRNG = ``numpy.random.default_rng(seed)`` (PCG64).

* :func:`forest_cloud` is modelled on the reference's random-forest map
  (Planner/src/map_generator.cpp:16-125): axis-aligned square pillars whose half-diagonal is
  ``w ~ U[0.6, 2.0]`` and height ``h ~ U[1, 8]`` (rounded to ``res``), centres uniform in the map,
  rejected when overlapping an earlier pillar (:40-49) or too close to start/goal (:36-39); four
  walls sampled on the ``res`` lattice for ``h_i = 1 .. h/res - 1`` and two caps (:97-125).
  Attempt density 0.133 / m^2 = 120 attempts on the 30 m x 30 m map of clean_demo.launch:163-175.
  Variant "L" keeps the lattice-exact coordinates (tie-heavy parity stress), variant "J" adds
  U(-res/2, res/2) float32 jitter per coordinate (LiDAR-like, tie-free throughput headline).
* :func:`rrt_queries` follows genSample's uniform branch (Planner/src/corridor_finder.cpp:333-359,
  ranges set at :64-70): uniform in the map box, z in [z_l + safety_margin, z_h] = [0.6, 4.0]
  (clean_demo.launch:27-28,31), stored float32 as the planner casts them (:122-125).
* :func:`bezier_trajectories` makes piecewise Bezier curves in the layout of `_PolyCoeff`
  (Planner/src/sim_planning_demo.cpp:715-727: row = [x | y | z] blocks of order+1 scaled
  control points; position = T_i * B(t / T_i)).
"""
from __future__ import annotations

import math

import numpy as np

ATTEMPTS_PER_M2 = 120.0 / 900.0  # clean_demo.launch:163-175 (120 attempts on +-15 m)
POINTS_PER_M2 = 200.0            # saturated density measured in SURVEY 8(d)


def _pillar_points(cx, cy, w, h, res):
    """Lattice cells (int32 (k,3)) of one pillar: 4 walls + bottom/top caps."""
    hc = int(round(round(h / res) * res / res))
    cor = []
    for phi in (45.0, 135.0, 225.0, 315.0, 405.0):
        cor.append((int(round((cx + w * math.cos(math.pi / 180.0 * phi)) / res)),
                    int(round((cy + w * math.sin(math.pi / 180.0 * phi)) / res))))
    parts = []
    hs = np.arange(1, hc, dtype=np.int32)
    for k in range(4):
        (x1, y1), (x2, y2) = cor[k], cor[k + 1]
        if x1 == x2:
            step = 1 if y1 < y2 else -1
            ys = np.arange(y1, y2, step, dtype=np.int32)
            yy, hh = np.meshgrid(ys, hs, indexing="ij")
            parts.append(np.stack([np.full(yy.size, x1, np.int32), yy.ravel(), hh.ravel()], 1))
        elif y1 == y2:
            step = 1 if x1 < x2 else -1
            xs = np.arange(x1, x2, step, dtype=np.int32)
            xx, hh = np.meshgrid(xs, hs, indexing="ij")
            parts.append(np.stack([xx.ravel(), np.full(xx.size, y1, np.int32), hh.ravel()], 1))
    (x1, y1), (x2, y2) = cor[0], cor[2]
    xs = np.arange(min(x1, x2), max(x1, x2) + 1, dtype=np.int32)
    ys = np.arange(min(y1, y2), max(y1, y2) + 1, dtype=np.int32)
    xx, yy = np.meshgrid(xs, ys, indexing="ij")
    for level in (0, hc):
        parts.append(np.stack([xx.ravel(), yy.ravel(), np.full(xx.size, level, np.int32)], 1))
    return np.concatenate(parts, 0)


def forest_map(half, seed=6, res=0.1, start=(-10.0, -10.0), goal=(9.0, 9.0), attempts=None):
    """All lattice cells (int32 (k,3)) of a forest on [-half, half]^2. Returns (cells, n_pillars)."""
    rng = np.random.default_rng(seed)
    if attempts is None:
        attempts = int(round(ATTEMPTS_PER_M2 * (2 * half) ** 2))
    px = np.empty(attempts); py = np.empty(attempts); pw = np.empty(attempts)
    n_acc = 0
    parts = []
    draws = rng.uniform(size=(attempts, 4))
    for k in range(attempts):
        x = -half + 2 * half * draws[k, 0]
        y = -half + 2 * half * draws[k, 1]
        w = 0.6 + 1.4 * draws[k, 2]
        h = 1.0 + 7.0 * draws[k, 3]
        if (x - start[0]) ** 2 + (y - start[1]) ** 2 < 2 + w * w or (x - goal[0]) ** 2 + (y - goal[1]) ** 2 < 2 + w * w:
            continue
        if n_acc and np.any((px[:n_acc] - x) ** 2 + (py[:n_acc] - y) ** 2 < (pw[:n_acc] + w) ** 2):
            continue
        px[n_acc], py[n_acc], pw[n_acc] = x, y, w
        n_acc += 1
        parts.append(_pillar_points(x, y, w, h, res))
    cells = np.concatenate(parts, 0) if parts else np.zeros((0, 3), np.int32)
    return cells, n_acc


def forest_cloud(n_points, seed=6, res=0.1, variant="J", return_half=False):
    """Exactly `n_points` float32 points of a cluttered forest, seeded shuffle + truncate.

    The square map is enlarged (at constant pillar-attempt density) until it yields >= n_points.
    """
    half = max(5.0, 0.5 * math.sqrt(1.12 * n_points / POINTS_PER_M2))
    while True:
        cells, _ = forest_map(half, seed=seed, res=res)
        if cells.shape[0] >= n_points:
            break
        half *= 1.08
    rng = np.random.default_rng(seed + 1_000_003)
    sel = rng.permutation(cells.shape[0])[:n_points]
    pts = (cells[sel].astype(np.float64) * res).astype(np.float32)
    if variant == "J":
        pts += rng.uniform(-0.5 * res, 0.5 * res, size=pts.shape).astype(np.float32)
    elif variant != "L":
        raise ValueError("variant must be 'L' or 'J'")
    pts = np.ascontiguousarray(pts)
    return (pts, half) if return_half else pts


def uniform_cloud(n_points, half=16.0, seed=0, z=(0.0, 8.0)):
    rng = np.random.default_rng(seed)
    p = np.empty((n_points, 3), np.float32)
    p[:, 0] = rng.uniform(-half, half, n_points)
    p[:, 1] = rng.uniform(-half, half, n_points)
    p[:, 2] = rng.uniform(z[0], z[1], n_points)
    return p


def rrt_queries(m, half, seed=0, z=(0.6, 4.0), lattice_frac=0.0, res=0.1):
    """Uniform in-box RRT* samples (float32 (m,3)); `lattice_frac` of them snapped to res/2."""
    rng = np.random.default_rng(seed + 7_000_003)
    q = np.empty((m, 3), np.float64)
    q[:, 0] = rng.uniform(-half, half, m)
    q[:, 1] = rng.uniform(-half, half, m)
    q[:, 2] = rng.uniform(z[0], z[1], m)
    if lattice_frac > 0:
        snap = rng.uniform(size=m) < lattice_frac
        q[snap] = np.round(q[snap] / (0.5 * res)) * (0.5 * res)
    return np.ascontiguousarray(q.astype(np.float32))


def bezier_trajectories(n_traj, half, seed=0, seg_range=(3, 8), order_range=(4, 8), z=(0.6, 4.0),
                        step_range=(1.0, 3.0), T_range=(0.5, 3.0)):
    """Random piecewise Bezier trajectories.

    Returns dict with CSR arrays:
      traj_first_seg int32[n_traj+1], seg_order int32[S], seg_T float64[S],
      seg_coef_off int64[S+1], coef float64[total] (per segment [x|y|z] blocks of order+1
      SCALED control points c = p / T so that position = T * B(u), as the reference stores them).
    Consecutive segments share their end/start control point (C0 continuity).
    """
    rng = np.random.default_rng(seed + 9_000_011)
    n_seg = rng.integers(seg_range[0], seg_range[1] + 1, size=n_traj)
    first = np.zeros(n_traj + 1, np.int32)
    np.cumsum(n_seg, out=first[1:])
    S = int(first[-1])
    order = rng.integers(order_range[0], order_range[1] + 1, size=S).astype(np.int32)
    T = rng.uniform(T_range[0], T_range[1], size=S)
    off = np.zeros(S + 1, np.int64)
    np.cumsum(3 * (order + 1), out=off[1:])
    coef = np.empty(int(off[-1]), np.float64)
    lo = np.array([-half, -half, z[0]]); hi = np.array([half, half, z[1]])
    for t in range(n_traj):
        p = rng.uniform(lo, hi)
        for s in range(first[t], first[t + 1]):
            n = int(order[s])
            ctrl = np.empty((n + 1, 3))
            ctrl[0] = p
            for j in range(1, n + 1):
                d = rng.normal(size=3)
                d[2] *= 0.3
                d *= rng.uniform(step_range[0], step_range[1]) / (np.linalg.norm(d) * n + 1e-12) * 2.0
                p = np.clip(p + d, lo, hi)
                ctrl[j] = p
            coef[off[s]:off[s + 1]] = (ctrl / T[s]).T.ravel()
    return dict(traj_first_seg=first, seg_order=order, seg_T=T, seg_coef_off=off, coef=coef)
