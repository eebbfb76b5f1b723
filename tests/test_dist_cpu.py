"""CPU tests of the N>1 host logic: two gloo ranks shard a batch with pc_shard_range, each answers its slice, the
slices are gathered and must equal the single-rank answer.  The per-rank backend here is the CPU oracle -- injected
by the TEST to exercise the sharding / rendezvous code of pointcloudtraj_b200.dist; the product path never uses it."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import oracle
    from pointcloudtraj_b200 import dist as pcd, synth, _lib

    rank, world, _ = pcd.env_rank()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = synth.uniform_cloud(3000, half=5.0, seed=1)
    q = synth.rrt_queries(1001, 5.0, seed=2)            # not divisible by 2: ragged shards
    ko = oracle.KdOracle().build(pts)
    b, e, (idx, d2) = pcd.sharded_call(lambda part: ko.nearest(part), q, rank, world)
    full_idx = pcd.gather_slices(idx, len(q), rank, world)
    full_d2 = pcd.gather_slices(d2, len(q), rank, world)
    ref_idx, ref_d2 = ko.nearest(q)
    assert (full_idx == ref_idx).all() and (full_d2 == ref_d2).all()
    # the NCCL unique id travels from rank 0 to every rank through the process group
    uid = pcd.exchange_unique_id(rank, world, lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    # empty batch and batch smaller than the world
    for m in (0, 1):
        b2, e2, _ = pcd.sharded_call(lambda part: ko.nearest(part), q[:m], rank, world)
        got = [None] * world
        dist.all_gather_object(got, (b2, e2))
        assert got[0][0] == 0 and got[-1][1] == m and all(got[i][1] == got[i + 1][0] for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok", b, e)
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_bench_reference_arm_other_ranks_exit_quietly():
    """`bench.py --impl reference` under torchrun: ranks other than 0 exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
