"""CPU test of the N>1 path (world_size 2, gloo): independent frame pairs sharded over ranks, no data-path collective,
final gather of the feature lists.  The per-rank compute is stood in by the CPU oracle (this is a test)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pyfeaturetrack_b200 import synth
    from pyfeaturetrack_b200.shard import shard_range, gather_features
    from oracle import klt_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_units, n = 5, 40
    p = O.Params(nPyramidLevels=2, subsampling=2, max_residue=10.0)
    local = {}
    for u in shard_range(n_units, rank, world):
        a, b = synth.frame_pair(120, 160, seed=50 + u)
        sel = O.select_good_features(p, a, n)
        local[u] = O.track_features(p, a, b, *sel)[:3]
    x, y, v = gather_features(local, n_units, n, dist)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), x=x, y=y, v=v)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from pyfeaturetrack_b200 import synth
    from oracle import klt_oracle as O
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    p = O.Params(nPyramidLevels=2, subsampling=2, max_residue=10.0)
    for u in range(5):
        a, b = synth.frame_pair(120, 160, seed=50 + u)
        sel = O.select_good_features(p, a, 40)
        x, y, v = O.track_features(p, a, b, *sel)[:3]
        assert np.array_equal(got["x"][u], x) and np.array_equal(got["y"][u], y) and np.array_equal(got["v"][u], v)


def test_bench_collectives_are_rank_symmetric(tmp_path):
    """bench.py's N > 1 control flow on the CPU (tests/_bench_flow_driver.py: real gloo collectives, stubbed GPU library): every
    rank issues the same collectives in the same order, rank 0 prints the one JSON line, nobody hangs.  (A barrier that only
    rank 0 reached -- after the others had left -- hung every N > 1 run of an earlier version of the script.)"""
    import json
    import subprocess
    world = 2
    port = 29700 + os.getpid() % 2000
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   KLT_FLOW_OUT=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_bench_flow_driver.py")], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=240))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    logs = [open(os.path.join(str(tmp_path), "collectives_%d.txt" % r)).read().split() for r in range(world)]
    assert len(logs[0]) > 20 and all(l == logs[0] for l in logs[1:]), [len(l) for l in logs]
    lines = [l for l in outs[0][0].splitlines() if l.strip()]
    assert len(lines) == 1 and not outs[1][0].strip()            # exactly one JSON line, from rank 0
    d = json.loads(lines[0])
    assert d["n_gpus"] == world and d["scaling"] == "weak" and "gather" in d and "sequence" in d
    assert "h2d_ceiling_gbps_per_gpu" in d["e2e"] and "mix_ceiling" in d["roofline"]
