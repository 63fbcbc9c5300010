#!/usr/bin/env python3
"""Golden vectors for the frame-side oracle: cv2.undistortPoints (cv2 4.13.0) on the calibrations the reference's
example settings use (TUM1/TUM2/TUM3 with distortion, 4 and 5 coefficients), called as Frame.cc:428 calls it
(R = none, P = K, float32 in/out).  Writes tests/golden/undistort_cv2.npz.  Run in the build container (needs cv2)."""
import os
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALS = {
    # Examples/Monocular/TUM1.yaml, TUM2.yaml (5 coefficients), a 4-coefficient variant, and a strong barrel case
    "tum1": ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    "tum2": ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
    "four": ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    "barrel": ((300.0, 310.0, 320.0, 240.0), (-0.35, 0.12, 0.001, -0.002, -0.02)),
}
out = {}
rng = np.random.default_rng(2024)
for name, (k, d) in CALS.items():
    K = np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float32)
    D = np.array(d, np.float32).reshape(-1, 1)
    pts = np.concatenate([rng.uniform([0, 0], [640, 480], (4000, 2)),
                          np.array([[0, 0], [640, 0], [0, 480], [640, 480], [k[2], k[3]]], np.float64),
                          rng.integers(0, 640, (500, 2)).astype(np.float64) * 1.2 ** rng.integers(0, 8, (500, 1))]).astype(np.float32)
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
    out[name + "_K"] = np.array(k, np.float32)
    out[name + "_D"] = np.array(d, np.float32)
    out[name + "_pts"] = pts
    out[name + "_und"] = und.astype(np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "undistort_cv2.npz"), **out)
print("cv2", cv2.__version__, {k: v.shape for k, v in out.items() if k.endswith("_und")})
