#!/usr/bin/env python
"""Registers / stack / spills / shared memory of every kernel of libpcindex (nvcc -Xptxas -v, no GPU needed).
    python scripts/ptxas_report.py [filter-substring]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v", "-c", "-o", "/tmp/pc_ptxas.o",
                      os.path.join(ROOT, "pointcloudtraj_b200", "csrc", "pc_index.cu")], capture_output=True, text=True)
txt = subprocess.run(["c++filt"], input=out.stderr + out.stdout, capture_output=True, text=True).stdout
name, st = None, ("0", "0")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for l in txt.split("\n"):
    m = re.search(r"Compiling entry function '(.*)' for", l)
    if m:
        name = m.group(1)[:64]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", l)
    if m:
        st = m.group(1), m.group(2)
    m = re.search(r"Used (\d+) registers", l)
    if m and name and flt in name:
        sm = re.search(r"(\d+) bytes smem", l)
        print(f"{name:66s} regs={m.group(1):>3s} stack={st[0]:>4s} spill={st[1]:>3s} smem={sm.group(1) if sm else 0}")
