"""The N>1 path on CPU: world_size-2 gloo; block partition + gather of the result records."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = importlib.import_module("2dhmsr_b200")
sharding = importlib.import_module("2dhmsr_b200.sharding")


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 4096, 1048576):
        for ws in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, ws, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import helpers
    from oracle import oracle  # the CPU stand-in for the per-rank stepper in this CPU-only test

    robots = helpers.phase_robots(n, seed=4)
    loco = helpers.locomotion(1.0)
    b, e = sharding.shard_range(n, rank, ws)
    local, _ = oracle.run(loco.compile(robots[b:e]))
    full = sharding.gather_results(local, n)
    dist.barrier()
    if rank == 0:
        q.put(full.tobytes())
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import oracle

    n = 5  # ragged over 2 ranks (3 + 2)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = np.frombuffer(q.get(timeout=120), dtype=H.RESULT_DTYPE)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref, _ = oracle.run(helpers.locomotion(1.0).compile(helpers.phase_robots(n, seed=4)))
    assert got.tobytes() == ref.tobytes()
