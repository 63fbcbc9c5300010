#!/usr/bin/env python3
"""CPU model of the round-based window search (csrc/frame.cu: window_first_round_kernel / window_search_kernel) against the
oracle's sequential SearchByProjection(Frame, map points) (oracle/frame_oracle.cc, R21/src/ORBmatcher.cc:45-130).

The reference walks the points in order; the kernels resolve them in rounds.  This script restates the ROUND RULE alone -- no CUDA:
  * an unresolved point i walks the candidates that no EARLIER point has taken (a feature remembers its taker), finds the two
    smallest (distance, walk position) keys B and S and its accept decision, and claims every free candidate within the threshold;
  * i is final when neither B nor S carries a claim of an earlier point; decisions are applied after all walks of the round.
and compares the outcome with the oracle on random scenes (the rule first shipped with a plain "taken" flag instead of the taker
index; tools/soak_frame.py found the scene where that is wrong, `--flag-rule` reproduces it).
usage: model_window_rounds.py [seed] [seconds] [--flag-rule]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

TH_HIGH = 100
INF = 1 << 30
_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)
CALS = [((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
        ((500.0, 500.0, 320.0, 240.0), (0.0, 0.0, 0.0, 0.0))]


def candidates(oracle, keys_un, ptr, idx, bounds, sf, ur, mp, dm, desc_f, th):
    """Per point: the walk of GetFeaturesInArea with the static filters applied -> [(feature, distance, octave)] in walk order."""
    out = []
    for i in range(len(mp)):
        m = mp[i]
        if not m["in_view"]:
            out.append(None); continue
        r = np.float32(2.5 if float(m["view_cos"]) > 0.998 else 4.0)
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        r = np.float32(r * sf[m["level"]])
        c = oracle.features_in_area(keys_un, ptr, idx, bounds, float(m["proj_x"]), float(m["proj_y"]), float(r), int(m["level"]) - 1, int(m["level"]))
        c = [int(f) for f in c if not (ur[f] > 0 and abs(np.float32(m["proj_xr"]) - ur[f]) > r)]
        if c:
            d = _POP[dm[i][None, :] ^ desc_f[c]].sum(1)
            out.append(list(zip(c, d.tolist(), keys_un["octave"][c].tolist())))
        else:
            out.append([])
    return out


def run_rounds(cands, occ, blocks, nf, ratio, flag_rule=False):
    """The kernels' rule.  Returns (feature -> point, point -> feature, rounds)."""
    npts = len(cands)
    taker = np.where(occ.astype(bool), -1, INF).astype(np.int64)
    fp = -np.ones(nf, np.int64); pf = -np.ones(npts, np.int64)
    resolved = np.array([c is None for c in cands])
    rounds = 0
    while not resolved.all():
        rounds += 1
        claim = np.full(nf, INF); tent = {}
        for i in np.nonzero(~resolved)[0]:
            k1 = k2 = (256, INF, -1, -1)                          # (distance, position, feature, octave)
            for pos, (f, d, o) in enumerate(cands[i]):
                if (taker[f] < INF) if flag_rule else (taker[f] < i):
                    continue
                if d <= TH_HIGH:
                    claim[f] = min(claim[f], i)
                k = (d, pos, f, o)
                if k < k1:
                    k2 = k1; k1 = k
                elif k < k2:
                    k2 = k
            if k1[0] > TH_HIGH:
                resolved[i] = True; continue
            accept = not (k1[3] == k2[3] and np.float32(k1[0]) > np.float32(ratio) * np.float32(k2[0]))
            tent[i] = (k1[2], k2[2], accept)
        taken = set()
        for i, (b, s, accept) in tent.items():
            if claim[b] < i or (s >= 0 and claim[s] < i):
                continue
            resolved[i] = True
            if accept:
                assert b not in taken, "two points that became final in one round took the same feature"
                taken.add(b); fp[b] = i; pf[i] = b
                if blocks[i]:
                    taker[b] = i
    return fp, pf, rounds


def check_scene(oracle, synth, tf, rng, max_features=900, max_points=2500, th=None, crowded=None, flag_rule=False):
    s = int(rng.integers(0, 1 << 30)); nf = int(rng.integers(1, max_features)); npts = int(rng.integers(1, max_points))
    crowded = bool(rng.random() < 0.5) if crowded is None else crowded
    K, D = CALS[int(rng.integers(0, len(CALS)))]
    K = np.array(K, np.float32); D = np.array(D, np.float32)
    keys, desc_f, sf, r2 = tf._scene(oracle, synth, nf, 0, s, K, D, crowded)
    keys_un = oracle.undistort_keypoints(keys, K, D); bounds = oracle.image_bounds(640, 480, K, D)
    ptr, idx = oracle.assign_grid(keys_un, bounds)
    ur = np.where(r2.random(nf) < 0.5, keys_un["x"] - r2.uniform(0, 30, nf), -1).astype(np.float32)
    occ = (r2.random(nf) < r2.choice([0.0, 0.1, 0.5])).astype(np.uint8)
    th = float(r2.choice([1.0, 3.0, 7.0, 15.0])) if th is None else th
    ratio = float(r2.choice([0.6, 0.8, 0.9]))
    mp, dm = tf._map_points(oracle, keys_un, desc_f, npts, r2)
    ref_fp, ref_pf, _ = oracle.search_by_projection_frame(keys_un, desc_f, ur, occ, ptr, idx, bounds, sf, mp, dm, th, ratio)
    cands = candidates(oracle, keys_un, ptr, idx, bounds, sf, ur, mp, dm, desc_f, th)
    fp, pf, rounds = run_rounds(cands, occ, mp["obs_positive"].astype(bool), nf, ratio, flag_rule)
    return bool(np.array_equal(fp, ref_fp) and np.array_equal(pf, ref_pf)), rounds, (s, nf, npts, crowded, th, ratio)


def main():
    import oracle_lib as oracle
    import test_frame as tf
    synth = importlib.import_module("cooperative-orb-slam_b200.synth")
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flag_rule = "--flag-rule" in sys.argv
    rng = np.random.default_rng(int(args[0]) if args else 1)
    t_end = time.time() + (float(args[1]) if len(args) > 1 else 60.0)
    n = bad = 0; most = 0
    while time.time() < t_end:
        ok, rounds, ctx = check_scene(oracle, synth, tf, rng, flag_rule=flag_rule)
        n += 1; most = max(most, rounds)
        if not ok:
            bad += 1
            print("differs from the oracle:", ctx)
            if not flag_rule:
                raise SystemExit(1)
    print("window-search round model: %d scenes, %d differ from the oracle, at most %d rounds%s" % (n, bad, most, " (plain taken-flag rule)" if flag_rule else ""))


if __name__ == "__main__":
    main()
