#!/usr/bin/env python3
"""Per-launch timing distribution of the 2000 x 1M 2-NN kernels (CUDA events around every call)."""
import ctypes as C
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
variants = [int(v) for v in sys.argv[1:]] or [3, 4]
dev = torch.device("cuda", 0)
m = synth.descriptors(1000000, seed=1234)
q = synth.descriptors(2000, seed=99)
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
rec = torch.empty((2000, 4), dtype=torch.int32, device=dev)
L = orb.lib()
cur = torch.cuda.current_stream(dev)
for rep in range(3):
    for variant in variants:
        ts = []
        for i in range(40):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            rc = L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), 1000000, 0, C.c_void_p(rec.data_ptr()), variant,
                                    C.c_void_p(cur.cuda_stream))
            e1.record(cur)
            assert rc == 0
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = np.array(ts)
        print("rep", rep, "variant", variant, "min %.3f med %.3f p90 %.3f max %.3f ms" % (ts.min(), np.median(ts), np.percentile(ts, 90), ts.max()), flush=True)
