#!/usr/bin/env python3
"""Single-frame latency of ORBextractor.__call__ (wall clock) and its per-stage CUDA-event breakdown (GPU box)."""
import importlib, sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
img = synth.frame(1, 640, 480)
ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(20): ex(img)
t0 = time.perf_counter()
for _ in range(200): ex(img)
print("wall ms/frame", (time.perf_counter() - t0) / 200 * 1e3)
ex.set_profiling(True) if hasattr(ex, "set_profiling") else None
for _ in range(5): ex(img)
print(ex.stage_times())
pin = orb.PinnedArray((480, 640), np.uint8); pin.array[:] = img
t0 = time.perf_counter()
for _ in range(200): ex(pin.array)
print("wall ms/frame pinned input", (time.perf_counter() - t0) / 200 * 1e3)
