#!/usr/bin/env python3
"""Per-search time of the default 2-NN chain (CTA-pair tcgen05 kernel + split merge) against the map size: the intercept is the
fixed cost per search (launches, operand set-up, pipeline fill, drain, merge) that limits the sharded search at N = 8."""
import ctypes as C
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
dev = torch.device("cuda", 0)
m = synth.descriptors(1000000, seed=1234); q = synth.descriptors(2000, seed=99)
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
rec = torch.empty((2000, 4), dtype=torch.int32, device=dev)
L = orb.lib()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for nm in (2304, 23040, 62500, 125000, 250000, 500000, 1000000):
    for _ in range(20):
        rc = L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), nm, 0, C.c_void_p(rec.data_ptr()), variant, None)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300):
        L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), nm, 0, C.c_void_p(rec.data_ptr()), variant, None)
    e1.record(); torch.cuda.synchronize()
    print("variant %d  map %8d  %.4f ms per search  (rc %d)" % (variant, nm, e0.elapsed_time(e1) / 300, rc))
