#!/usr/bin/env python3
"""Small fixed launch sequence for ncu: 3 batches of 64 frames (12 launches each: 7 pyramid, fast score,
blur, cell nms, quadtree, describe) followed by one 2000 x 1M 2-NN search (2 launches)."""
import ctypes as C
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")

B, W, H = 64, 640, 480
dev = torch.device("cuda", 0)
frames = np.stack([synth.frame(s, W, H) for s in range(16)])
frames = np.concatenate([frames] * 4)
d_frames = torch.from_numpy(frames).to(dev)
ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
cap = ex.max_keypoints(W, H)
d_k = torch.empty((B, cap, 7), dtype=torch.int32, device=dev)
d_d = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_c = torch.empty((B,), dtype=torch.int32, device=dev)
for it in range(3):
    ex.extract_batch_device(d_frames.data_ptr(), B, W, H, W, W * H, d_k.data_ptr(), d_d.data_ptr(), cap, d_c.data_ptr())
    ex.wait()
m = synth.descriptors(1000000, seed=1234)
q = synth.descriptors(2000, seed=99)
d_m = torch.from_numpy(m).to(dev); d_q = torch.from_numpy(q).to(dev)
rec = torch.empty((2000, 4), dtype=torch.int32, device=dev)
L = orb.lib()
rc = L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), 2000, C.c_void_p(d_m.data_ptr()), 1000000, 0, C.c_void_p(rec.data_ptr()), 0, None)
torch.cuda.synchronize()
print("ok", rc, int(d_c.sum().item()), int(rec[:, 0].sum().item()))
