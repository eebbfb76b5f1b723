#!/usr/bin/env python
"""Per-SASS-block breakdown of a kernel from an ncu capture (--import-source on): consecutive instructions with the same
execution count form a block; prints blocks sorted by issued instructions with their stall samples.
    python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys


def main(path, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split per kernel
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    k = kernels[int(which)]
    hdr = k["rows"][0]
    ia, isrc, iexe, isam = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in k["rows"][1:] if len(r) > iexe and r[iexe].isdigit()]
    total = sum(int(r[iexe]) for r in body)
    tsam = sum(int(r[isam]) for r in body)
    print(f"# {k['name'][:100]}")
    print(f"# {len(body)} SASS instructions, {total} warp instructions executed, {tsam} stall samples")
    blocks, start = [], 0
    for i in range(1, len(body) + 1):
        if i == len(body) or body[i][iexe] != body[start][iexe]:
            blk = body[start:i]
            blocks.append((start, len(blk), int(blk[0][iexe]), sum(int(r[iexe]) for r in blk), sum(int(r[isam]) for r in blk), blk))
            start = i
    print(f"{'first':>6s} {'len':>4s} {'exec/inst':>11s} {'issued':>12s} {'share':>7s} {'samples':>8s} {'s.share':>7s}  mnemonics")
    for b in sorted(blocks, key=lambda b: -b[3])[:24]:
        mn = {}
        for r in b[5]:
            m = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
            m = m.split(".")[0]
            mn[m] = mn.get(m, 0) + 1
        top = " ".join(f"{m}x{c}" for m, c in sorted(mn.items(), key=lambda kv: -kv[1])[:8])
        print(f"{b[0]:6d} {b[1]:4d} {b[2]:11d} {b[3]:12d} {b[3] / total * 100:6.2f}% {b[4]:8d} {b[4] / max(tsam, 1) * 100:6.2f}%  {top}")


if __name__ == "__main__":
    main(*sys.argv[1:])
