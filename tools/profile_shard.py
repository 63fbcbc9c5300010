#!/usr/bin/env python3
"""One rank's share of the 8-GPU map search (2000 queries x 125 000 descriptors, CTA-pair kernel + split merge), 20 times:
launch sequence for `ncu --metrics gpu__time_duration.sum` (what the 89 us per search at N = 8 are made of)."""
import ctypes as C
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
dev = torch.device("cuda", 0)
m = synth.descriptors(125000, seed=1234); q = synth.descriptors(2000, seed=99)
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
rec = torch.empty((2000, 4), dtype=torch.int32, device=dev)
L = orb.lib()
for _ in range(20):
    rc = L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), 125000, 0, C.c_void_p(rec.data_ptr()), 5, None)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), 125000, 0, C.c_void_p(rec.data_ptr()), 5, None)
e1.record(); torch.cuda.synchronize()
print("ok", rc, int(rec[:, 0].sum().item()), "%.4f ms per search" % (e0.elapsed_time(e1) / 200))
