#!/usr/bin/env python3
"""Randomised parity soak (GPU box): random image sizes, contents and extractor parameters through the CUDA path and the
CPU oracle, compared bit for bit; random 2-NN problems through every kernel variant.  usage: soak.py [seconds] [seed]"""
import importlib
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")


def random_image(rng, w, h):
    kind = rng.integers(0, 6)
    if kind == 0:
        return synth.frame(int(rng.integers(0, 1 << 30)), w, h)
    if kind == 1:
        return synth.frame(int(rng.integers(0, 1 << 30)), w, h, low_texture=True)
    if kind == 2:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)                       # white noise: every pixel a corner candidate
    if kind == 3:
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)                # flat + a few rectangles: sparse corners, empty cells
        for _ in range(int(rng.integers(0, 12))):
            x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
            img[y0:y0 + int(rng.integers(4, 60)), x0:x0 + int(rng.integers(4, 60))] = int(rng.integers(0, 256))
        return img
    if kind == 4:
        yy, xx = np.mgrid[0:h, 0:w]                                               # checkerboard with a random period: many equal scores (ties)
        p = int(rng.integers(3, 24))
        return ((((xx // p) + (yy // p)) & 1) * int(rng.integers(30, 255))).astype(np.uint8)
    base = synth.frame(int(rng.integers(0, 1 << 30)), w, h).astype(np.int16)     # quantised texture: plateaus
    q = int(rng.integers(8, 64))
    return ((base // q) * q).clip(0, 255).astype(np.uint8)


def one_extraction(rng):
    w = int(rng.integers(96, 1400)); h = int(rng.integers(96, 900))
    h = min(h, int(1.8 * w))      # taller than 2:1 means zero quadtree roots: undefined in the reference (it indexes an empty vector, :552-560)
    nlevels = int(rng.integers(1, 9)); sf = float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5]))
    while min(w, h) / sf ** (nlevels - 1) < 64 and nlevels > 1:
        nlevels -= 1
    nfeat = int(rng.choice([50, 300, 500, 1000, 2000, 3000])); ini = int(rng.choice([20, 12, 30])); mn = int(rng.choice([7, 5, 12]))
    mn = min(mn, ini)
    img = random_image(rng, w, h)
    return (w, h, nfeat, sf, nlevels, ini, mn), img


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    n_ext = n_kp = n_knn = n_rej = n_batch = 0
    pool = ThreadPoolExecutor(8)
    while time.time() < t_end:
        cases = [one_extraction(rng) for _ in range(8)]

        def cpu(c):
            (w, h, nfeat, sf, nl, ini, mn), img = c
            try:
                return oracle_lib.OracleExtractor(nfeat, sf, nl, ini, mn, trig_mode=1).extract(img, cap=30000)
            except RuntimeError:
                return None      # a level whose quadtree would have zero roots (aspect < 1:2): the reference divides by zero (:543-545)
        refs = list(pool.map(cpu, cases))
        for c, ref in zip(cases, refs):
            (w, h, nfeat, sf, nl, ini, mn), img = c
            if ref is None:
                try:
                    orb.ORBextractor(nfeat, sf, nl, ini, mn)(img)
                    raise AssertionError(("the CUDA path accepted what the reference cannot compute", c[0]))
                except orb.OrbCudaError:
                    n_rej += 1
                continue
            rk, rd = ref
            try:
                k, d = orb.ORBextractor(nfeat, sf, nl, ini, mn)(img)
            except orb.OrbCudaError as exc:
                # the CUDA path refuses every size with a zero-root level up front; the oracle (like the reference) only
                # trips over it when that level actually has corners
                assert "zero quadtree roots" in str(exc), exc
                n_rej += 1
                continue
            if len(rk) == 0:
                assert len(k) == 0, c[0]
            else:
                assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd), ("extraction differs", c[0], len(k), len(rk))
            n_ext += 1; n_kp += len(rk)
            if n_ext % 2 == 0:
                # the same image as a batch of five frames: batches of >= 4 frames run the throughput variants of the kernels
                # (14-row FAST strips, 16-row blur strips, 64-row resize tiles), single frames the latency variants
                ex = orb.ORBextractor(nfeat, sf, nl, ini, mn, max_width=w, max_height=h, max_batch=5)
                bk, bd, bc = ex.extract_batch(np.stack([img] * 5))
                for f in range(5):
                    assert bc[f] == len(rk), ("batch count differs", c[0], f)
                    if len(rk):
                        assert bk[f, :bc[f]].tobytes() == rk.tobytes() and np.array_equal(bd[f, :bc[f]], rd), ("batch extraction differs", c[0], f)
                ex.close()
                n_batch += 1
        # matcher
        nq = int(rng.integers(1, 3000)); nm = int(rng.integers(0, 150000))
        m = synth.descriptors(max(nm, 1), seed=int(rng.integers(0, 1 << 30)))[:nm]
        q = synth.descriptors(nq, seed=int(rng.integers(0, 1 << 30)))
        if nm > 4:
            dup = rng.integers(0, nm, 6)
            m[dup[1:]] = m[dup[0]]; q[0] = m[dup[0]]; q[nq // 2] = m[dup[0]] ^ np.uint8(1)
        base = orb.ORBmatcher().knn2(q, m, index_base=int(rng.integers(0, 1 << 20)), variant=0)
        i1, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
        for variant in (1, 2, 3, 4, 5):
            got = orb.ORBmatcher().knn2(q, m, index_base=0, variant=variant)
            ref0 = orb.ORBmatcher().knn2(q, m, index_base=0, variant=0)
            assert all(np.array_equal(a, b) for a, b in zip(got, ref0)), ("2-NN variant differs", variant, nq, nm)
        oracle_lib.lib().orc_knn2(q.ctypes.data, nq, np.ascontiguousarray(m).ctypes.data, nm, 0, i1.ctypes.data, d1.ctypes.data, d2.ctypes.data, 8)
        assert np.array_equal(ref0[0], i1) and np.array_equal(ref0[1], d1) and np.array_equal(ref0[2], d2), ("2-NN differs from the oracle", nq, nm)
        n_knn += 1
    print("soak ok: %d extractions (%d key points, %d rejected by both; %d of them again as batches of 5), %d 2-NN problems x 6 variants, seed %d"
          % (n_ext, n_kp, n_rej, n_batch, n_knn, seed))


if __name__ == "__main__":
    main()
