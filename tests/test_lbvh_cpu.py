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
