"""CPU checks of the host-side mirrors: the kd_* compat header is valid C, the C++ corridor wrapper is valid C++,
and libpcindex.so exports every function they declare."""
import os
import re
import subprocess

from pointcloudtraj_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def test_compat_header_is_valid_c_and_exported():
    src = '#include "pc_kdtree_compat.h"\n#include "pc_index.h"\nint main(void){return 0;}\n'
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", INC, "-x", "c", "-"], input=src,
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    header = open(os.path.join(INC, "pc_kdtree_compat.h")).read().split("#ifdef PC_KDTREE_COMPAT_RENAME")[0]
    names = set(re.findall(r"\b(pckd_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 30
    lib = L.load()
    for n in sorted(names):
        assert hasattr(lib, n), f"libpcindex.so does not export {n}"


def test_rename_macros_cover_the_reference_api():
    # a client written against kd_* compiles unchanged with the rename switch (syntax only: no GPU here)
    p = subprocess.run(["gcc", "-std=c99", "-DUSE_PCINDEX", "-fsyntax-only", "-I", INC,
                        os.path.join(ROOT, "tests", "c", "kd_client.c")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_corridor_wrapper_is_valid_cpp():
    src = '#include "pc_corridor.hpp"\nint main(){ return 0; }\n'
    p = subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only", "-I", INC, "-x", "c++", "-"], input=src,
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
