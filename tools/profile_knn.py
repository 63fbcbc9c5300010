#!/usr/bin/env python3
"""2000 x 1M 2-NN search, all kernel variants (for ncu).  usage: profile_knn.py [variant ...]"""
import ctypes as C
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
variants = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 3, 4, 5]
dev = torch.device("cuda", 0)
m = synth.descriptors(1000000, seed=1234)
q = synth.descriptors(2000, seed=99)
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
rec = torch.empty((2000, 4), dtype=torch.int32, device=dev)
L = orb.lib()
for rep in range(2):
    for variant in variants:
        rc = L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), 1000000, 0, C.c_void_p(rec.data_ptr()), variant, None)
        torch.cuda.synchronize()
        print("variant", variant, rc, int(rec[:, 0].sum().item()))
