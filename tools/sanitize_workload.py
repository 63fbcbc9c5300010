#!/usr/bin/env python3
"""Small pass over every kernel of liborbcuda for compute-sanitizer (memcheck / initcheck / racecheck), checked against the
oracle as it goes: extraction (batch, single frame, strided pinned upload, device batch), all six 2-NN variants + the device
ratio test, the batched stereo search, the frame-side kernels and every candidate loop (the reference-checked golden calls).
    compute-sanitizer --tool memcheck python tools/sanitize_workload.py"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
import oracle_lib                      # noqa: E402
import test_matcher_golden as G        # noqa: E402

which = set(sys.argv[1:]) or {"extract", "knn", "stereo", "loops"}
if "extract" in which:
    frames = np.stack([synth.frame(s, 640, 480) for s in range(2)])
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=2)
    kps, desc, cnt = ex.extract_batch(frames)
    ok, od = oracle_lib.OracleExtractor(1000, trig_mode=1).extract(frames[1])
    assert cnt[1] == len(ok) and kps[1, :cnt[1]].tobytes() == ok.tobytes() and np.array_equal(desc[1, :cnt[1]], od)
    k1, d1 = orb.ORBextractor(1000, 1.2, 8, 20, 7)(frames[0])
    assert len(k1) == cnt[0]
    kit = synth.frame(3, 1241, 376)
    k2, d2 = orb.ORBextractor(2000, 1.2, 8, 20, 7)(kit)
    ok2, od2 = oracle_lib.OracleExtractor(2000, trig_mode=1).extract(kit)
    assert k2.tobytes() == ok2.tobytes() and np.array_equal(d2, od2)
    print("extract ok", int(cnt.sum()), len(k2), flush=True)
if "knn" in which:
    m = synth.descriptors(6000, seed=1); q, m, _ = synth.query_set(m, nq=300, seed=2)
    i1, d1, d2 = (np.zeros(300, np.int32) for _ in range(3))
    oracle_lib.lib().orc_knn2(q.ctypes.data, 300, m.ctypes.data, len(m), 0, i1.ctypes.data, d1.ctypes.data, d2.ctypes.data, 2)
    for v in range(6):
        bi, bd, sd, si = orb.ORBmatcher().knn2(q, m, variant=v)
        assert np.array_equal(bi, i1) and np.array_equal(bd, d1) and np.array_equal(sd, d2), v
    dq = torch.from_numpy(q).cuda(); dm = torch.from_numpy(m).cuda()
    match = torch.zeros(300, dtype=torch.int32, device="cuda")
    assert orb.lib().orbm_knn2_ratio_device(C.c_void_p(dq.data_ptr()), 300, C.c_void_p(dm.data_ptr()), len(m), 0, 0.7, 50, 0, None,
                                            C.c_void_p(match.data_ptr()), 5, None) == 0
    torch.cuda.synchronize()
    print("knn ok", int((match >= 0).sum().item()), flush=True)
if "stereo" in which:
    left, right = synth.stereo_pair(1, 752, 480)
    el = orb.ORBextractor(1200, 1.2, 8, 20, 7); er = orb.ORBextractor(1200, 1.2, 8, 20, 7)
    kl, dl = el(left); kr, dr = er(right)
    ur, dep, n = orb.compute_stereo_matches(el, er, kl, dl, kr, dr, 40.0, 40.0 / 458.0)
    print("stereo ok", n, flush=True)
if "loops" in which:
    seen = set()
    for call, fn, a in G._calls():
        if fn in seen and fn != "window_best":
            continue
        seen.add(fn)
        G._check(call, fn, a, G._cuda(fn, a, orb))
    print("loops ok", sorted(seen), flush=True)
