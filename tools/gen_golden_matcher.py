#!/usr/bin/env python3
"""Generates tests/golden/matcher_ref.npz: the POD inputs and outputs of every C-ABI call the drop-in ORBmatcher class made
while a small scene was driven through the REFERENCE's Frame / KeyFrame / MapPoint objects -- kept only after the whole
scenario's results were checked against the reference's own ORBmatcher.cc (oracle/_ref/libmatchref.so).  The outputs in
the file are therefore the reference's; tests/test_matcher_golden.py replays the inputs through the oracle (CPU) and
through liborbcuda (GPU) anywhere, without /root/reference.  Needs oracle/_ref built (make -C oracle)."""
import glob
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

rec = tempfile.mkdtemp(prefix="orbm_rec_")
os.environ["ORBM_RECORD_DIR"] = rec          # read once by the library at its first recorded call

import matchref_lib as mr                    # noqa: E402
import test_matcher_ref as T                 # noqa: E402

blobs, index = {}, []
for case in ((11, True, False), (12, True, True)):
    sc = mr.Scene(case[0], distortion=case[1], n_points=330, n_clutter=70, stereo_frac=0.35 if case[2] else 0.0)
    obs = [sc.observe(0), sc.observe(1, angle_offset=-20.0), sc.observe(2, angle_offset=12.0)]
    ref = T._run("ref", sc, obs)
    before = set(glob.glob(os.path.join(rec, "*.npy")))
    got = T._run("shim_cpu", sc, obs)
    T._same(ref, got)                        # the recorded outputs ARE the reference's results
    assert ref["bow_kf_f", 0.7, True][0] > 30 and ref["fuse"] > 3 and ref["sim3"][0] > 10, "scene too small to exercise the loops"
    for path in sorted(set(glob.glob(os.path.join(rec, "*.npy"))) - before):
        call, fn, arg = os.path.basename(path)[:-4].split(".", 2)
        a = np.load(path)
        h = "b" + hashlib.sha1(a.tobytes() + str(a.dtype).encode() + str(a.shape).encode()).hexdigest()[:12]
        blobs[h] = a
        index.append(["%d_%s" % (case[0], call), fn, arg, h])
out = os.path.join(ROOT, "tests", "golden", "matcher_ref.npz")
np.savez_compressed(out, index=np.frombuffer(json.dumps(index).encode(), np.uint8), **blobs)
calls = sorted(set((c, f) for c, f, _, _ in index))
print("%d calls (%s), %d unique arrays -> %s (%.0f KB)" % (len(calls), ", ".join(sorted(set(f for _, f in calls))), len(blobs), out,
                                                            os.path.getsize(out) / 1e3))
