"""CPU parity oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package, and there only as the checker (or as the reported CPU baseline),
never as the product path.  ``pointcloudtraj_b200`` never imports it.

Two backends with one interface:

* :class:`KdOracle`     -- this repo's C restatement (``kd_oracle.c`` / ``planner_oracle.c``).
* :class:`KdReference`  -- the UNMODIFIED reference ``Utils/kdtree`` (``kdtree.c``) compiled by
  ``oracle/Makefile`` from ``/root/reference`` into ``oracle/_ref/libkdtree_ref.so`` and driven
  through its public ``kd_*`` API by ``ref_harness.c``.

Reference anchors: Utils/kdtree/src/kdtree.c:112-126 (kd_create), :244-251 (kd_insert3),
:493-500 (kd_nearest3), :595-602 (kd_nearest_range3), :613-650 (kd_res_*).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libkdtree_ref.so")

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference is present) _ref/libkdtree_ref.so."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_reference() -> bool:
    return os.path.exists(_REF_PATH)


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _as_f32_rows(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("expected an (n, 3) or (n, 4) float32 array")
    return a


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.kdo_create.restype = C.c_void_p
        L.kdo_free.argtypes = [C.c_void_p]
        L.kdo_clear.argtypes = [C.c_void_p]
        L.kdo_size.argtypes = [C.c_void_p]
        L.kdo_size.restype = C.c_int64
        L.kdo_depth.argtypes = [C.c_void_p]
        L.kdo_depth.restype = C.c_int64
        L.kdo_insert3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int64]
        L.kdo_build.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _i64p]
        L.kdo_nearest3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _i64p, _f64p, _f64p]
        L.kdo_nearest_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _i64p, _f64p, C.c_int]
        L.kdo_range3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, _i64p, C.c_int64]
        L.kdo_range3.restype = C.c_int64
        L.kdo_range_count_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, C.c_int, _i64p, C.c_int]
        L.kdo_range_fill_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, C.c_int, _i64p, _i64p, C.c_int]
        L.kdo_brute_nearest_batch.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, C.c_int64, C.c_int64,
                                              _i64p, _f64p, _i32p, C.c_int]
        L.kdo_pair_d2.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, _i64p, C.c_int64, _f64p]
        L.po_radius_search.argtypes = [C.c_void_p, C.c_void_p, _f64p, _i64p]
        L.po_radius_search.restype = C.c_double
        L.po_radius_batch.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, _i64p]
        L.po_binomial.argtypes = [C.c_int, C.c_int]
        L.po_binomial.restype = C.c_double
        L.po_bezier_pos.argtypes = [_f64p, C.c_int, C.c_double, _f64p]
        L.po_check_safe_trajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, _i32p, _f64p, _f64p, C.c_int64,
                                               C.c_double, C.c_double, C.c_double, C.c_int64, _f32p, _f64p,
                                               _i64p, _f64p]
        L.po_check_safe_trajectory.restype = C.c_int64
        L.po_gen_samples.argtypes = [C.c_void_p, C.c_int64, _f64p]
        L.po_steer.argtypes = [_f64p, _f64p, C.c_float, _f64p]
        L.po_steer_batch.argtypes = [_f64p, C.c_int64, _f64p, _f32p, _i32p, _f64p]
        _lib = L
    return _lib


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_PATH):
            raise FileNotFoundError(
                f"{_REF_PATH} missing: run `make -C oracle` in the container that has /root/reference")
        L = C.CDLL(_REF_PATH)
        L.refh_create.restype = C.c_void_p
        L.refh_free.argtypes = [C.c_void_p]
        L.refh_clear.argtypes = [C.c_void_p]
        L.refh_build.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _i64p]
        L.refh_max_threads.restype = C.c_int
        L.refh_nearest_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _i64p, _f64p, C.c_int]
        L.refh_range3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, _i64p, C.c_int64]
        L.refh_range3.restype = C.c_int64
        L.refh_range_count_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, C.c_int, _i64p, C.c_int]
        L.refh_range_fill_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, C.c_int, _i64p, _i64p, C.c_int]
        L.refh_radius_batch.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f64p, _f64p, C.c_int]
        _ref = L
    return _ref


class RadiusParams(C.Structure):
    """Mirror of po_radius_params / pc_radius_params (safeRegionRrtStar::setParam,
    Planner/src/corridor_finder.cpp:17-23; start_pt from setStartPt :43-50)."""
    _fields_ = [("search_margin", C.c_double), ("max_radius", C.c_double),
                ("sample_range", C.c_double), ("start", C.c_double * 3)]

    @classmethod
    def make(cls, search_margin=0.25, max_radius=1.5, sample_range=30.0, start=(0.0, 0.0, 0.0)):
        p = cls()
        p.search_margin, p.max_radius, p.sample_range = search_margin, max_radius, sample_range
        p.start[0], p.start[1], p.start[2] = [float(v) for v in start]
        return p


class Sampler(C.Structure):
    """Mirror of po_sampler / pc_sampler: genSample's state (corridor_finder.cpp:333-358) -- the minstd_rand0 engine state and the
    bounds setPt gives the uniform distributions (:52-91)."""
    _fields_ = [("engine_state", C.c_uint32), ("reserved", C.c_uint32), ("goal_ratio", C.c_double), ("inlier_ratio", C.c_double),
                ("end_pt", C.c_double * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3), ("in_lo", C.c_double * 3), ("in_hi", C.c_double * 3)]

    @classmethod
    def make(cls, start, end, box, sample_range, safety_margin, inlier_ratio, goal_ratio, engine_state=1):
        """The state after setParam(safety_margin, ., ., sample_range) + setPt(start, end, *box, ...); engine_state 1 is what
        default_random_engine(0) starts from."""
        s = cls()
        s.engine_state, s.goal_ratio, s.inlier_ratio = int(engine_state), float(goal_ratio), float(inlier_ratio)
        xl, xh, yl, yh, zl, zh = [float(v) for v in box]
        for a in range(3):
            s.end_pt[a] = float(end[a])
        s.lo[0], s.hi[0], s.lo[1], s.hi[1], s.lo[2], s.hi[2] = xl, xh, yl, yh, zl + safety_margin, zh
        s.in_lo[0], s.in_hi[0] = start[0] - sample_range, start[0] + sample_range
        s.in_lo[1], s.in_hi[1] = start[1] - sample_range, start[1] + sample_range
        s.in_lo[2], s.in_hi[2] = zl + safety_margin, zh
        return s


def gen_samples(sampler: Sampler, k):
    """The next k samples of the stream (float64 (k, 3)); sampler.engine_state is advanced."""
    out = np.empty((int(k), 3))
    lib().po_gen_samples(C.byref(sampler), int(k), _ptr(out, _f64p))
    return out


def steer(samples, node_coord, node_radius, nearest):
    """genNewNode's centres for samples[j] steered from node nearest[j] (float64 (k, 3))."""
    samples = np.ascontiguousarray(samples, np.float64); node_coord = np.ascontiguousarray(node_coord, np.float64)
    node_radius = np.ascontiguousarray(node_radius, np.float32); nearest = np.ascontiguousarray(nearest, np.int32)
    out = np.empty_like(samples)
    lib().po_steer_batch(_ptr(samples, _f64p), len(samples), _ptr(node_coord, _f64p), _ptr(node_radius, _f32p), _ptr(nearest, _i32p), _ptr(out, _f64p))
    return out


class _KdBase:
    """Shared batch interface over a tree handle."""
    _prefix = ""

    def __init__(self):
        self._L = None
        self._h = None
        self.n = 0

    # -- building ---------------------------------------------------------------------------
    def build(self, xyz, order=None):
        """kd_clear + n x kd_insert3(x, y, z, (void*)i) in `order` (default 0..n-1)."""
        xyz = _as_f32_rows(xyz)
        self._xyz = xyz
        o = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
        rc = getattr(self._L, self._prefix + "build")(self._h, _ptr(xyz, _f32p), xyz.shape[0], xyz.shape[1], _ptr(o, _i64p))
        if rc != 0:
            raise MemoryError("kd build failed")
        self.n = xyz.shape[0]
        return self

    # -- nearest ----------------------------------------------------------------------------
    def nearest(self, q, nthreads=0):
        """Returns (idx int64[m], d2 float64[m]); idx -1 / d2 inf on an empty tree."""
        q = _as_f32_rows(q)
        m = q.shape[0]
        idx = np.empty(m, dtype=np.int64)
        d2 = np.empty(m, dtype=np.float64)
        getattr(self._L, self._prefix + "nearest_batch")(self._h, _ptr(q, _f32p), m, q.shape[1],
                                                         _ptr(idx, _i64p), _ptr(d2, _f64p), nthreads)
        return idx, d2

    # -- range ------------------------------------------------------------------------------
    def range(self, q, r, nthreads=0):
        """kd_nearest_range3 per query; returns (offsets int64[m+1], idx int64[total]) with each
        list in the reference result set's iteration order."""
        q = _as_f32_rows(q)
        m = q.shape[0]
        r = np.ascontiguousarray(np.atleast_1d(r), dtype=np.float64)
        scalar = 1 if r.shape[0] == 1 else 0
        counts = np.empty(m, dtype=np.int64)
        getattr(self._L, self._prefix + "range_count_batch")(self._h, _ptr(q, _f32p), m, q.shape[1],
                                                             _ptr(r, _f64p), scalar, _ptr(counts, _i64p), nthreads)
        off = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        out = np.empty(max(int(off[-1]), 1), dtype=np.int64)
        getattr(self._L, self._prefix + "range_fill_batch")(self._h, _ptr(q, _f32p), m, q.shape[1],
                                                            _ptr(r, _f64p), scalar, _ptr(off, _i64p), _ptr(out, _i64p), nthreads)
        return off, out[: int(off[-1])]


class KdOracle(_KdBase):
    """This repo's restatement (kd_oracle.c)."""
    _prefix = "kdo_"

    def __init__(self):
        super().__init__()
        self._L = lib()
        self._h = C.c_void_p(self._L.kdo_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.kdo_free(self._h)
            self._h = None

    def insert3(self, x, y, z, payload):
        rc = self._L.kdo_insert3(self._h, x, y, z, payload)
        self.n += 1
        return rc

    def depth(self):
        return int(self._L.kdo_depth(self._h))

    # planner-side restatements -------------------------------------------------------------
    def radius_search(self, params: RadiusParams, p):
        """safeRegionRrtStar::radiusSearch on one double-precision point."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        idx = C.c_int64(-1)
        r = self._L.po_radius_search(self._h, C.byref(params), _ptr(p, _f64p), C.byref(idx))
        return r, idx.value

    def radius_batch(self, params: RadiusParams, q):
        q = _as_f32_rows(q)
        m = q.shape[0]
        out = np.empty(m, dtype=np.float64)
        idx = np.empty(m, dtype=np.int64)
        self._L.po_radius_batch(self._h, C.byref(params), _ptr(q, _f32p), m, q.shape[1], _ptr(out, _f64p), _ptr(idx, _i64p))
        return out, idx

    def check_safe_trajectory(self, params: RadiusParams, order, T, coef, t_now, stop_time, dt=0.02, cap=4096):
        """checkSafeTrajectory full walk. Returns dict(first_hit, n_samples, min_radius, pts, radius)."""
        order = np.ascontiguousarray(order, dtype=np.int32)
        T = np.ascontiguousarray(T, dtype=np.float64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n_seg = order.shape[0]
        ld = coef.shape[1] if coef.ndim == 2 and n_seg > 0 else 0
        pts = np.zeros((cap, 3), dtype=np.float32)
        rad = np.zeros(cap, dtype=np.float64)
        ns = C.c_int64(0)
        rmin = C.c_double(0.0)
        fh = self._L.po_check_safe_trajectory(self._h, C.byref(params), n_seg, _ptr(order, _i32p), _ptr(T, _f64p),
                                              _ptr(coef, _f64p), ld, t_now, stop_time, dt, cap,
                                              _ptr(pts, _f32p), _ptr(rad, _f64p), C.byref(ns), C.byref(rmin))
        n = min(ns.value, cap)
        return dict(first_hit=int(fh), n_samples=int(ns.value), min_radius=rmin.value, pts=pts[:n], radius=rad[:n])


class KdReference(_KdBase):
    """The unmodified reference Utils/kdtree through ref_harness.c."""
    _prefix = "refh_"

    def __init__(self):
        super().__init__()
        self._L = ref_lib()
        self._h = C.c_void_p(self._L.refh_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.refh_free(self._h)
            self._h = None

    def max_threads(self):
        return int(self._L.refh_max_threads())

    def radius_batch(self, params: RadiusParams, q, nthreads=0):
        q = _as_f32_rows(q)
        m = q.shape[0]
        out = np.empty(m, dtype=np.float64)
        pv = np.array([params.search_margin, params.max_radius, params.sample_range,
                       params.start[0], params.start[1], params.start[2]], dtype=np.float64)
        self._L.refh_radius_batch(self._h, _ptr(q, _f32p), m, q.shape[1], _ptr(pv, _f64p), _ptr(out, _f64p), nthreads)
        return out


def brute_nearest(xyz, q, nthreads=0):
    """Exact fp64 minimum in the reference's operation order, lowest index on ties.
    Returns (idx int64[m], d2 float64[m], n_ties int32[m])."""
    xyz = _as_f32_rows(xyz)
    q = _as_f32_rows(q)
    m = q.shape[0]
    idx = np.empty(m, dtype=np.int64)
    d2 = np.empty(m, dtype=np.float64)
    ties = np.empty(m, dtype=np.int32)
    lib().kdo_brute_nearest_batch(_ptr(xyz, _f32p), xyz.shape[0], xyz.shape[1], _ptr(q, _f32p), m, q.shape[1],
                                  _ptr(idx, _i64p), _ptr(d2, _f64p), _ptr(ties, _i32p), nthreads)
    return idx, d2, ties


def pair_d2(xyz, q, idx):
    """fp64 d2 between point idx[k] and query k in the reference's operation order."""
    xyz = _as_f32_rows(xyz)
    q = _as_f32_rows(q)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty(idx.shape[0], dtype=np.float64)
    lib().kdo_pair_d2(_ptr(xyz, _f32p), xyz.shape[1], _ptr(q, _f32p), q.shape[1], _ptr(idx, _i64p), idx.shape[0], _ptr(out, _f64p))
    return out


def bezier_pos(coef_row, order, u):
    coef_row = np.ascontiguousarray(coef_row, dtype=np.float64)
    out = np.zeros(3, dtype=np.float64)
    lib().po_bezier_pos(_ptr(coef_row, _f64p), order, u, _ptr(out, _f64p))
    return out


def binomial(n, k):
    return lib().po_binomial(n, k)


# ---- the compiled, UNMODIFIED reference planner (oracle/_ref/libplanner_ref.so) ---------------------------------------
_PLANNER_PATH = os.path.join(_HERE, "_ref", "libplanner_ref.so")
_planner = None


def have_planner_reference() -> bool:
    return os.path.exists(_PLANNER_PATH)


def planner_lib():
    global _planner
    if _planner is None:
        if not os.path.exists(_PLANNER_PATH):
            raise FileNotFoundError(f"{_PLANNER_PATH} missing: run `make -C oracle` in the container that has /root/reference")
        L = C.CDLL(_PLANNER_PATH)
        d = C.c_double
        L.rp_create.argtypes = [d, d, d, d]
        L.rp_set_input.argtypes = [_f32p, C.c_longlong, C.c_longlong]
        L.rp_set_pt.argtypes = [_f64p, _f64p, d, d, d, d, d, d, d, C.c_int, d, d]
        L.rp_set_start_pt.argtypes = [_f64p, _f64p]
        L.rp_expand.argtypes = [d]
        L.rp_refine.argtypes = [d]
        L.rp_reset_root.argtypes = [_f64p]
        L.rp_get_path.argtypes = [_f64p, _f64p, C.c_int]
        L.rp_get_tree.argtypes = [_f64p, C.c_int]
        L.rp_stats.argtypes = [C.POINTER(C.c_longlong)]
        L.rp_gen_samples.argtypes = [C.c_longlong, _f64p]
        L.rp_radius_search.argtypes = [_f64p]
        L.rp_radius_search.restype = d
        L.rp_radius_batch.argtypes = [_f64p, C.c_longlong, _f64p]
        L.rp_check_traj_pt_col.argtypes = [_f64p]
        L.rp_bezier_pos.argtypes = [C.c_int, _f64p, d, _f64p]
        L.rp_check_safe_trajectory.argtypes = [C.c_int, _i32p, _f64p, _f64p, C.c_longlong, d, d, _f32p, C.c_longlong,
                                               C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_float)]
        _planner = L
    return _planner


class PlannerReference:
    """The reference's safeRegionRrtStar (Planner/src/corridor_finder.cpp, unmodified) and checkSafeTrajectory /
    getPosFromBezier (Planner/src/sim_planning_demo.cpp:715-781, unmodified) behind oracle/ref_planner_harness.cpp.
    ONE planner object per process (checkSafeTrajectory names a global): creating a new instance re-creates it."""

    def __init__(self, safety_margin=0.6, search_margin=0.25, max_radius=1.5, sample_range=30.0):
        self._L = planner_lib()
        self._L.rp_create(safety_margin, search_margin, max_radius, sample_range)

    def set_input(self, xyz):
        xyz = _as_f32_rows(xyz)
        self._L.rp_set_input(_ptr(xyz, _f32p), xyz.shape[0], xyz.shape[1])
        return self

    def reset(self):
        self._L.rp_reset()

    def set_pt(self, start, end, box, local_range, max_iter, sample_portion, goal_portion):
        s = np.ascontiguousarray(start, np.float64); e = np.ascontiguousarray(end, np.float64)
        self._L.rp_set_pt(_ptr(s, _f64p), _ptr(e, _f64p), *[float(v) for v in box], float(local_range), int(max_iter),
                          float(sample_portion), float(goal_portion))

    def set_start_pt(self, start, end):
        s = np.ascontiguousarray(start, np.float64); e = np.ascontiguousarray(end, np.float64)
        self._L.rp_set_start_pt(_ptr(s, _f64p), _ptr(e, _f64p))

    def expand(self, n_iter):
        self._L.rp_expand(float(n_iter))

    def refine(self, n_iter):
        self._L.rp_refine(float(n_iter))

    def evaluate(self):
        self._L.rp_evaluate()

    def path(self, cap=4096):
        p = np.zeros((cap, 3)); r = np.zeros(cap)
        n = self._L.rp_get_path(_ptr(p, _f64p), _ptr(r, _f64p), cap)
        return p[:n].copy(), r[:n].copy()

    def tree(self, cap=1 << 20):
        """(n, 7): x, y, z, radius, g, parent index in the same list (-1), valid."""
        t = np.zeros((cap, 7))
        n = self._L.rp_get_tree(_ptr(t, _f64p), cap)
        return t[:n].copy()

    def stats(self):
        s = (C.c_longlong * 6)()
        self._L.rp_stats(s)
        return dict(nodes=s[0], path_exists=bool(s[1]), cloud_queries=s[2], rebuilds=s[3], warnings=s[4], global_navi=bool(s[5]))

    def gen_samples(self, k):
        """k calls of the unmodified genSample on the planner's own engine: float64 (k, 3)."""
        out = np.empty((int(k), 3))
        self._L.rp_gen_samples(int(k), _ptr(out, _f64p))
        return out

    def radius_search(self, p):
        p = np.ascontiguousarray(p, np.float64)
        return self._L.rp_radius_search(_ptr(p, _f64p))

    def radius_batch(self, pts):
        """radiusSearch for every row of pts (float64 (m, 3): the planner passes Vector3d)."""
        pts = np.ascontiguousarray(pts, np.float64)
        out = np.empty(pts.shape[0])
        self._L.rp_radius_batch(_ptr(pts, _f64p), pts.shape[0], _ptr(out, _f64p))
        return out

    def bezier_pos(self, coef_row, order, u):
        c = np.ascontiguousarray(coef_row, np.float64)
        out = np.zeros(3)
        self._L.rp_bezier_pos(int(order), _ptr(c, _f64p), float(u), _ptr(out, _f64p))
        return out

    def check_safe_trajectory(self, order, T, coef, t_now, stop_time, cap=8192):
        """Returns (collides, pts float32 (k, 3), k): the reference's return value and the sample points it visited.
        self.last_searched / self.last_min_d2: how many of them reached the cloud query, and the smallest float32 squared
        distance those saw (logged by the stand-in pcl::search::KdTree)."""
        order = np.ascontiguousarray(order, np.int32); T = np.ascontiguousarray(T, np.float64)
        coef = np.ascontiguousarray(coef, np.float64)
        pts = np.zeros((cap, 3), np.float32)
        n, ns, md = C.c_longlong(0), C.c_longlong(0), C.c_float(0)
        hit = self._L.rp_check_safe_trajectory(order.shape[0], _ptr(order, _i32p), _ptr(T, _f64p), _ptr(coef, _f64p),
                                               coef.shape[1] if coef.ndim == 2 else 0, float(t_now), float(stop_time),
                                               _ptr(pts, _f32p), cap, C.byref(n), C.byref(ns), C.byref(md))
        self.last_searched, self.last_min_d2 = int(ns.value), np.float32(md.value)
        return bool(hit), pts[: min(n.value, cap)].copy(), int(n.value)
