"""World-size-2 (gloo, CPU) test of the sharded map matching exchange: every rank searches its shard,
records are all-gathered and merged -- the N>1 data path of bench.py, with the oracle standing in for the
per-rank kernel (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q_out):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orb = importlib.import_module("cooperative-orb-slam_b200")
    synth = importlib.import_module("cooperative-orb-slam_b200.synth")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.descriptors(8000, seed=11)
    q, m, _ = synth.query_set(m, nq=300, seed=12)
    m[4100] = m[17]; q[1] = m[17]                       # a tie across the shard boundary
    lo = len(m) * rank // world; hi = len(m) * (rank + 1) // world
    shard = np.ascontiguousarray(m[lo:hi])
    i1, d1, i2, d2 = (np.zeros(len(q), np.int32) for _ in range(4))
    oracle_lib.lib().orc_knn2_full(q.ctypes.data, len(q), shard.ctypes.data, len(shard), lo, i1.ctypes.data, d1.ctypes.data,
                                   i2.ctypes.data, d2.ctypes.data, 1)
    rec = torch.from_numpy(np.stack([d1, i1, d2, i2], 1).astype(np.int32))
    parts = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec)
    parts = np.ascontiguousarray(torch.stack(parts).numpy())
    out = np.zeros((len(q), 4), np.int32)
    assert orb.lib().orbm_merge_top2_host(parts.ctypes.data, world, len(q), out.ctypes.data) == 0
    if rank == 0:
        j1, e1, j2, e2 = (np.zeros(len(q), np.int32) for _ in range(4))
        oracle_lib.lib().orc_knn2_full(q.ctypes.data, len(q), m.ctypes.data, len(m), 0, j1.ctypes.data, e1.ctypes.data,
                                       j2.ctypes.data, e2.ctypes.data, 1)
        q_out.put(bool(np.array_equal(out, np.stack([e1, j1, e2, j2], 1))))
    dist.destroy_process_group()


def test_sharded_match_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q_out.get(timeout=5) is True
