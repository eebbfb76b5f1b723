#!/usr/bin/env python
"""CPU experiment: node visits of the bounded nearest-obstacle search on different bounding-box trees (see tree_quality.c).

    python scripts/tree_quality.py [--points 1000000] [--queries 100000]
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloudtraj_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=100_000)
    a = ap.parse_args()
    pts, half = synth.forest_cloud(a.points, seed=1, variant="J", return_half=True)
    q = synth.rrt_queries(a.queries * 2, half, seed=1000)
    q = q[np.linalg.norm(q.astype(np.float64) - np.array([0, 0, 2.0]), axis=1) <= 31.5][: a.queries]   # the searched share
    with tempfile.TemporaryDirectory() as tmp:
        pts[:, :3].astype(np.float32).tofile(os.path.join(tmp, "p.bin"))
        q.astype(np.float32).tofile(os.path.join(tmp, "q.bin"))
        exe = os.path.join(tmp, "tree_quality")
        subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "scripts", "tree_quality.c"), "-lm"], check=True)
        for radius in ("1.75", "0"):
            subprocess.run([exe, os.path.join(tmp, "p.bin"), os.path.join(tmp, "q.bin"), radius], check=True)


if __name__ == "__main__":
    main()
