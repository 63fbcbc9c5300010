#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle mismatch report (debug aid; run on the GPU box)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orbcuda
import oracle_lib as O
synth = importlib.import_module("cooperative-orb-slam_b200.synth")


def main():
    for (w, h, nf) in [(640, 480, 1000), (1241, 376, 2000)]:
        img = synth.frame(3, w, h)
        ex = orbcuda.ORBextractor(nf, 1.2, 8, 20, 7)
        ex.set_profiling(True)
        t0 = time.time(); kps, desc = ex(img); t1 = time.time()
        print("config", w, h, nf, "n", len(kps), "wall ms %.2f" % ((t1 - t0) * 1e3), ex.stage_times())
        oe = O.OracleExtractor(nf, trig_mode=1)
        okps, odesc = oe.extract(img)
        for l in range(8):
            a, b = ex.pyramid(l, True), oe.pyramid(l, True)
            bl = oe.blurred(l)
            nb = int((ex.blurred(l) != bl).sum()) if bl is not None else -1
            gx, gy, gs = ex.candidates(l); ox, oy, os_ = oe.candidates(l)
            same = len(gx) == len(ox) and np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_)
            # score map vs oracle scores at oracle candidate positions
            sc = ex.scores(l)
            sm = int((sc[oy.astype(int) + 16, ox.astype(int) + 16] != os_).sum()) if len(ox) else 0
            print("  L%d %dx%d pyr_mism %d blur_mism %d cand gpu %d oracle %d same %s score_mism_at_cands %d" %
                  (l, a.shape[1] - 38, a.shape[0] - 38, int((a != b).sum()), nb, len(gx), len(ox), same, sm))
            if not same and len(gx) and len(ox):
                so = set(zip(ox.tolist(), oy.tolist(), os_.tolist())); sg = set(zip(gx.tolist(), gy.tolist(), gs.tolist()))
                print("     only gpu %d only oracle %d ; first diff order idx %s" % (len(sg - so), len(so - sg),
                      next((i for i in range(min(len(gx), len(ox))) if (gx[i], gy[i], gs[i]) != (ox[i], oy[i], os_[i])), None)))
        print("  kps gpu %d oracle %d" % (len(kps), len(okps)))
        n = min(len(kps), len(okps))
        for f in kps.dtype.names:
            print("   field %s ndiff %d" % (f, int((kps[f][:n] != okps[f][:n]).sum())))
        if len(kps) == len(okps):
            print("   desc rows differing", int((desc != odesc).any(1).sum()), "max angle diff", float(np.abs(kps["angle"] - okps["angle"]).max()))
    # batch timing
    frames = np.stack([synth.frame(s) for s in range(32)])
    ex = orbcuda.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=32)
    ex.set_profiling(True)
    for it in range(3):
        t0 = time.time(); out = ex.extract_batch(frames); t1 = time.time()
        print("batch32 wall ms %.2f" % ((t1 - t0) * 1e3), {k: round(v, 3) for k, v in ex.stage_times().items()}, "counts", out[2][:4])


if __name__ == "__main__":
    main()
