#!/usr/bin/env python3
"""Randomised parity soak of the frame side, every projection / window search and the BoW transform (GPU box).
usage: soak_frame.py [seconds] [seed]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as oracle  # noqa: E402
import test_frame as tf      # noqa: E402  (scene generators)
import test_bow as tb        # noqa: E402
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")

CALS = [((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
        ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
        ((500.0, 500.0, 320.0, 240.0), (0.0, 0.0, 0.0, 0.0)),
        ((300.0, 310.0, 320.0, 240.0), (-0.35, 0.12, 0.001, -0.002, -0.02))]


def eq3(a, b, what, ctx):
    assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (what, ctx)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    n = 0
    while time.time() < t_end:
        s = int(rng.integers(0, 1 << 30))
        nf = int(rng.integers(1, 4000)); npts = int(rng.integers(1, 5000)); crowded = bool(rng.random() < 0.3)
        K, D = CALS[int(rng.integers(0, len(CALS)))]
        K = np.array(K, np.float32); D = np.array(D, np.float32)
        keys, desc_f, sf, r2 = tf._scene(oracle, synth, nf, 0, s, K, D, crowded)
        fr = orb.FrameFeatures(keys, K, D, 640, 480)
        assert fr.keys_un.tobytes() == oracle.undistort_keypoints(keys, K, D).tobytes(), ("undistort", s)
        ptr, idx = oracle.assign_grid(fr.keys_un, fr.bounds)
        assert np.array_equal(fr.cell_ptr, ptr) and np.array_equal(fr.cell_idx[:ptr[-1]], idx[:ptr[-1]]), ("grid", s)
        ur = np.where(r2.random(nf) < 0.5, fr.keys_un["x"] - r2.uniform(0, 30, nf), -1).astype(np.float32)
        occ = (r2.random(nf) < r2.choice([0.0, 0.1, 0.5])).astype(np.uint8)
        th = float(r2.choice([1.0, 3.0, 7.0, 15.0])); ratio = float(r2.choice([0.6, 0.8, 0.9]))
        ctx = (s, nf, npts, crowded, th, ratio)
        mp, dm = tf._map_points(oracle, fr.keys_un, desc_f, npts, r2)
        eq3(orb.search_by_projection_frame(fr, desc_f, ur, occ, sf, mp, dm, th=th, nnratio=ratio),
            oracle.search_by_projection_frame(fr.keys_un, desc_f, ur, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, mp, dm, th, ratio), "frame", ctx)
        pts, dp = tf._proj_points(oracle, fr.keys_un, desc_f, npts, r2)
        for check in (True, False):
            d = int(r2.integers(0, 3))
            eq3(orb.search_by_projection_last_frame(fr, desc_f, ur, occ, sf, pts, dp, th, direction=d, check_orientation=check),
                oracle.search_by_projection_last_frame(fr.keys_un, desc_f, ur, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th, d, check), "last", ctx)
            od = int(r2.choice([50, 64, 100]))
            eq3(orb.search_by_projection_keyframe(fr, desc_f, occ, sf, pts, dp, th, od, check_orientation=check),
                oracle.search_by_projection_keyframe(fr.keys_un, desc_f, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th, od, check), "kf", ctx)
        eq3(orb.search_by_projection_sim3(fr, desc_f, occ, sf, pts, dp, th), oracle.search_by_projection_sim3(fr.keys_un, desc_f, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th), "sim3", ctx)
        sig = (1.0 / (sf * sf)).astype(np.float32)
        for gated in (True, False):
            a = orb.window_best_match(fr, desc_f, sf, pts, dp, th, u_right=ur if gated else None, inv_level_sigma2=sig if gated else None)
            b = oracle.window_best_match(fr.keys_un, desc_f, ur if gated else None, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, sig if gated else None, pts, dp, th)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), ("window", ctx, gated)
        k1, d1, k2, d2, prev = tf._init_pair(oracle, synth, max(nf, 2), s + 1, crowded)
        Kz = np.array([500, 500, 320, 240], np.float32); Dz = np.zeros(4, np.float32)
        f2 = orb.FrameFeatures(k2, Kz, Dz, 640, 480)
        win = int(r2.choice([30, 100]))
        for check in (True, False):
            rm, rxy, rn = oracle.search_for_initialization(k1, d1, f2.keys_un, d2, f2.cell_ptr, f2.cell_idx, f2.bounds, prev, win, 0.9, check)
            gm, gxy, gn = orb.search_for_initialization(k1, d1, f2, d2, prev, win, 0.9, check)
            assert gn == rn and np.array_equal(gm, rm) and np.array_equal(gxy, rxy), ("init", ctx)
        if n % 5 == 0:
            k = int(r2.choice([4, 10, 17])); L = int(r2.integers(2, 6)); lu = int(r2.integers(0, 5))
            voc = tb.make_vocabulary(k, L, seed=s & 0xffff)
            q = tb._queries(voc, int(r2.integers(10, 2500)), s & 0xfff)
            v = orb.ORBVocabulary(voc["child_ptr"], voc["child_idx"], voc["node_desc"], voc["word_id"], voc["weight"], L)
            w, nd, wt = v.transform_features(q, lu)
            rw, rn_, rwt = oracle.bow_transform(q, voc, lu)
            assert np.array_equal(w, rw) and np.array_equal(nd, rn_) and np.array_equal(wt.view(np.uint64), rwt.view(np.uint64)), ("bow", ctx)
            (bw, bv), fv = v.transform(q, lu)
            (obw, obv), ofv = oracle.bow_vectors(rw, rn_, rwt)
            assert np.array_equal(bw, obw) and np.array_equal(bv.view(np.uint64), obv.view(np.uint64)) and all(np.array_equal(x, y) for x, y in zip(fv, ofv)), ("bowvec", ctx)
            v.close()
        n += 1
    print("frame soak ok: %d random scenes x (undistort, grid, 6 searches, init%s), seed %d" % (n, ", BoW every 5th", seed))


if __name__ == "__main__":
    main()
