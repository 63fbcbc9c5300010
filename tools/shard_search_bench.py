#!/usr/bin/env python3
"""Sharded 2000 x 1M map search alone (no extraction): the fused peer-memory exchange (orbm_knn2_exchange_device) timed on every
rank's stream, max over ranks, and its records compared with the unsharded search of rank 0's own full copy of the map.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_search_bench.py
ORBCUDA_SHARE_BOUND=1 switches the cross-rank pruning bounds on (A/B; off by default).  SEARCHES_IN_FLIGHT=k (default 1) keeps k independent
searches in flight on k streams, each with its own peer buffers: the exchange of one search (a cross-GPU wait) overlaps the tensor
core kernel of the next -- the throughput a server sees that matches the key frames of several agents."""
import ctypes as C
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
NQ, NM = 2000, 1000000
m = synth.descriptors(NM, seed=1234); q = synth.descriptors(NQ, seed=99)
lo, hi = NM * rank // world, NM * (rank + 1) // world
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
out = torch.empty((NQ, 4), dtype=torch.int32, device=dev); ref = torch.empty_like(out)


def gather_bytes(b):
    t = torch.tensor(list(b), dtype=torch.uint8, device=dev)
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    return [bytes(x.cpu().tolist()) for x in g]


K = int(os.environ.get("SEARCHES_IN_FLIGHT", "1"))
peers = [orb.PeerExchange(NQ, rank, world, local, gather_bytes) for _ in range(K)]
peer = peers[0]
cur = torch.cuda.current_stream()
streams = [cur] if K == 1 else [torch.cuda.Stream(dev) for _ in range(K)]
outs = [torch.empty((NQ, 4), dtype=torch.int32, device=dev) for _ in range(K)]
out = outs[0]
L = orb.lib()
L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), NQ, C.c_void_p(d_m.data_ptr()), NM, 0, C.c_void_p(ref.data_ptr()), 5, C.c_void_p(cur.cuda_stream))


calls = 0


def step():
    global calls
    k = calls % K
    calls += 1
    peers[k].knn2(d_q.data_ptr(), NQ, d_m[lo:hi].data_ptr(), hi - lo, lo, outs[k].data_ptr(), 5, streams[k].cuda_stream)


def fork():
    for st in streams:
        if st is not cur:
            st.wait_stream(cur)


def join():
    for st in streams:
        if st is not cur:
            cur.wait_stream(st)



torch.cuda.synchronize()
fork()
for _ in range(20):
    step()
join()
torch.cuda.synchronize(); dist.barrier()
assert all(torch.equal(o, ref) for o in outs), "sharded result differs from the single search"
res = []
for rep in range(3):
    n = 500
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    fork()
    for _ in range(n):
        step()
    join()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.append(float(t.item()))
assert all(torch.equal(o, ref) for o in outs) and all(p.error() == 0 for p in peers)
if rank == 0:
    print("world %d share_bound=%s searches_in_flight=%d: %.4f ms per search (runs: %s), records == single search" %
          (world, os.environ.get("ORBCUDA_SHARE_BOUND", "0 (default)"), K, min(res), ", ".join("%.4f" % r for r in res)))
for p in peers:
    p.close()
dist.destroy_process_group()
