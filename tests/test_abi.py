"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, and
refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pointcloudtraj_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_are_exported():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "pc_index.h")).read()
    declared = set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", header))
    declared -= {"pc_radius_params", "pc_traj", "pc_index_view"}
    assert declared == set(L.ABI_SYMBOLS), declared ^ set(L.ABI_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libpcindex.so does not export {name}"


def test_struct_layouts_match_header():
    assert C.sizeof(L.PcRadiusParams) == 6 * 8
    assert C.sizeof(L.PcTraj) == 16
    assert C.sizeof(L.PcIndexView) == 3 * 8 + 2 * 8 + 6 * 4


def test_shard_range_partitions():
    from pointcloudtraj_b200 import shard_range
    for m in (0, 1, 7, 100, 10**8 + 3):
        for g in (1, 2, 4, 8):
            spans = [shard_range(m, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_version_string():
    assert b"sm_100a" in L.load().pc_version()


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from pointcloudtraj_b200 import PcError, PointCloudIndex
    with pytest.raises(PcError) as e:
        PointCloudIndex(1000)
    assert "no CPU fallback" in str(e.value)


def test_null_handles_are_rejected():
    lib = L.load()
    assert lib.pc_index_sync(None) == L.PC_EINVAL
    assert lib.pc_index_size(None) == 0
    assert lib.pc_index_build(None, None, 0, 3, 0) == L.PC_EINVAL
    assert lib.pc_nearest_batch(None, None, 0, 3, 0, 0, None, None) == L.PC_EINVAL
