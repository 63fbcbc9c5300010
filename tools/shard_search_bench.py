#!/usr/bin/env python3
"""Sharded 2000 x 1M map search alone (no extraction): the fused peer-memory exchange (orbm_knn2_exchange_device) timed on every
rank's stream, max over ranks, and its records compared with the unsharded search of rank 0's own full copy of the map.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_search_bench.py
ORBCUDA_SHARE_BOUND=0 switches the cross-rank pruning bounds off (A/B)."""
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


peer = orb.PeerExchange(NQ, rank, world, local, gather_bytes)
cur = torch.cuda.current_stream()
L = orb.lib()
L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), NQ, C.c_void_p(d_m.data_ptr()), NM, 0, C.c_void_p(ref.data_ptr()), 5, C.c_void_p(cur.cuda_stream))


def step():
    peer.knn2(d_q.data_ptr(), NQ, d_m[lo:hi].data_ptr(), hi - lo, lo, out.data_ptr(), 5, cur.cuda_stream)


for _ in range(20):
    step()
torch.cuda.synchronize(); dist.barrier()
assert torch.equal(out, ref), "sharded result differs from the single search"
res = []
for rep in range(3):
    n = 500
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.append(float(t.item()))
assert torch.equal(out, ref) and peer.error() == 0
if rank == 0:
    print("world %d share_bound=%s: %.4f ms per search (runs: %s), records == single search" %
          (world, os.environ.get("ORBCUDA_SHARE_BOUND", "1"), min(res), ", ".join("%.4f" % r for r in res)))
peer.close()
dist.destroy_process_group()
