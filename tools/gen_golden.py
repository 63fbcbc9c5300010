#!/usr/bin/env python3
"""Generate tests/golden/*.npz.

  primitives.npz  -- outputs of cv2 4.13.0 (the pinned OpenCV) for resize /
                     copyMakeBorder / GaussianBlur / FAST / fastAtan2 on small seeded inputs.
  extractor_*.npz -- keypoints + descriptors produced by the REFERENCE's own ORBextractor.cc
                     (compiled verbatim from /root/reference over oracle/cvshim -> oracle/_ref)
                     on seeded synthetic frames.
Run here (needs cv2 and oracle/_ref); the fixtures travel to the GPU box.
"""
import importlib, os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
G = os.path.join(ROOT, "tests", "golden")


def primitives():
    assert cv2.__version__ == "4.13.0", cv2.__version__
    rng = np.random.default_rng(20261018)
    img = synth.frame(11, 200, 160)
    noise = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    out = {"img": img, "noise": noise}
    out["resize_167x133"] = cv2.resize(img, (167, 133), interpolation=cv2.INTER_LINEAR)
    out["resize_noise_67x53"] = cv2.resize(noise, (67, 53), interpolation=cv2.INTER_LINEAR)
    out["border19"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    out["blur"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for name, im in (("img", img), ("noise", noise)):
        for th in (20, 7):
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                 type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            k = det.detect(im)
            out["fast_%s_%d" % (name, th)] = np.array([(kp.pt[0], kp.pt[1], kp.response) for kp in k], np.float32).reshape(-1, 3)
    yx = rng.integers(-30000, 30000, (4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [5, -5]]
    out["atan2_yx"] = yx
    out["atan2"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(G, "primitives.npz"), **out)


def extractor():
    cases = {"640x480_nf1000_seed0": (640, 480, 1000, 0, False),
             "640x480_nf1000_lowtex_seed3": (640, 480, 1000, 3, True),
             "1241x376_nf2000_seed1": (1241, 376, 2000, 1, False),
             "752x480_nf1200_seed2": (752, 480, 1200, 2, False)}
    for name, (w, h, nf, seed, low) in cases.items():
        r = O.RefExtractor(nf)
        img = synth.frame(seed, w, h, low_texture=low)
        kps, desc = r.extract(img)
        np.savez_compressed(os.path.join(G, "extractor_%s.npz" % name), kps=kps, desc=desc,
                            params=np.array([w, h, nf, seed, int(low)]),
                            img_sha=np.frombuffer(__import__("hashlib").sha256(img.tobytes()).digest(), np.uint8))


if __name__ == "__main__":
    os.makedirs(G, exist_ok=True)
    primitives()
    extractor()
    print(sorted(os.listdir(G)))
