"""GPU parity tests of the extraction path: every stage of liborbcuda (through the C ABI) against the
CPU oracle on the same seeded synthetic frames.  Bit-exact bar for pyramid, blur, FAST candidates,
quadtree selection, key points; angles bit-exact (tolerance 1e-3 deg allowed by north_star);
descriptors bit-exact against the oracle in trig_mode=1 and >= 99.9 % against trig_mode=0."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [(640, 480, 1000), (1241, 376, 2000), (752, 480, 1200)]


@pytest.fixture(scope="module")
def orb():
    import orbcuda
    if orbcuda.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return orbcuda


def _extract_both(orb, oracle, synth, w, h, nf, seed, low=False, trig_mode=1):
    img = synth.frame(seed, w, h, low_texture=low)
    ex = orb.ORBextractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    oe = oracle.OracleExtractor(nf, trig_mode=trig_mode)
    okps, odesc = oe.extract(img)
    return img, ex, kps, desc, oe, okps, odesc


@pytest.mark.parametrize("cfg", CONFIGS)
def test_tables(orb, oracle, cfg):
    w, h, nf = cfg
    ex = orb.ORBextractor(nf, 1.2, 8, 20, 7)
    t = oracle.OracleExtractor(nf).tables()
    assert np.array_equal(ex.GetScaleFactors(), t["sf"])
    assert np.array_equal(ex.GetInverseScaleFactors(), t["isf"])
    assert np.array_equal(ex.GetScaleSigmaSquares(), t["s2"])
    assert np.array_equal(ex.GetInverseScaleSigmaSquares(), t["is2"])
    assert np.array_equal(ex.mnFeaturesPerLevel, t["nfeat"])
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == float(np.float32(1.2))


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("low", [False, True])
def test_stages_bit_exact(orb, oracle, synth, cfg, low):
    w, h, nf = cfg
    img, ex, kps, desc, oe, okps, odesc = _extract_both(orb, oracle, synth, w, h, nf, seed=3, low=low)
    report = []
    for l in range(8):
        assert ex.level_size(l) == oe.level_size(l)
        a, b = ex.pyramid(l, True), oe.pyramid(l, True)
        report.append(("pyr", l, int((a != b).sum())))
        ob = oe.blurred(l)
        if ob is not None:
            report.append(("blur", l, int((ex.blurred(l) != ob).sum())))
        # dense FAST score map vs the oracle's per-pixel cornerScore on the same (oracle) pyramid
        gx, gy, gs = ex.candidates(l)
        ox, oy, os_ = oe.candidates(l)
        same = len(gx) == len(ox) and np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_)
        report.append(("cand", l, 0 if same else max(1, abs(len(gx) - len(ox)))))
    bad = [r for r in report if r[2]]
    assert not bad, bad
    assert len(kps) == len(okps), (len(kps), len(okps))
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], okps[f]), f
    assert np.array_equal(kps["angle"], okps["angle"]), float(np.abs(kps["angle"] - okps["angle"]).max())
    assert np.array_equal(desc, odesc), int((desc != odesc).any(1).sum())


def test_score_map_matches_corner_score(orb, oracle, synth):
    img = synth.frame(9)
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    ex(img)
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for l in (0, 3, 7):
        pad = np.ascontiguousarray(ex.pyramid(l, True))
        sc = ex.scores(l)
        h, w = sc.shape
        ys = rng.integers(19, h - 19, 3000); xs = rng.integers(19, w - 19, 3000)
        for y, x in zip(ys, xs):
            ref = L.orc_fast_score(int(pad.ctypes.data) + (int(y) + 19) * int(pad.strides[0]) + (int(x) + 19), int(pad.strides[0]), 7)
            assert sc[y, x] == ref, (l, x, y, sc[y, x], ref)
        assert sc[:19].sum() == 0 and sc[:, :19].sum() == 0 and sc[h - 19:].sum() == 0 and sc[:, w - 19:].sum() == 0


def test_reference_trig_budget(orb, oracle, synth):
    """Against the reference's own libm cosf/sinf: >= 99.9 % of descriptors bit-identical, angles within 1e-3 deg."""
    tot = diff = 0
    for seed in range(4):
        img, ex, kps, desc, oe, okps, odesc = _extract_both(orb, oracle, synth, 640, 480, 1000, seed, trig_mode=0)
        assert len(kps) == len(okps)
        assert np.abs(kps["angle"] - okps["angle"]).max() <= 1e-3
        tot += len(desc); diff += int((desc != odesc).any(1).sum())
    assert diff <= 1e-3 * tot, (diff, tot)


def test_golden_reference_fixtures(orb, synth, golden_dir):
    """Committed outputs of the reference's own ORBextractor.cc (tools/gen_golden.py)."""
    for f in sorted(glob.glob(os.path.join(golden_dir, "extractor_*.npz"))):
        g = np.load(f)
        w, h, nf, seed, low = [int(v) for v in g["params"]]
        img = synth.frame(seed, w, h, low_texture=bool(low))
        kps, desc = orb.ORBextractor(nf, 1.2, 8, 20, 7)(img)
        gk = g["kps"]
        assert len(kps) == len(gk), f
        for fld in ("x", "y", "size", "response", "octave", "class_id", "angle"):
            assert np.array_equal(kps[fld], gk[fld]), (f, fld)
        nd = int((desc != g["desc"]).any(1).sum())
        assert nd <= max(1, int(1e-3 * len(gk))), (f, nd)


def test_batch_equals_single(orb, synth):
    frames = np.stack([synth.frame(s) for s in range(6)])
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=6)
    kps, desc, cnt = ex.extract_batch(frames)
    single = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    for i in range(len(frames)):
        k1, d1 = single(frames[i])
        assert cnt[i] == len(k1)
        assert kps[i, :cnt[i]].tobytes() == k1.tobytes()
        assert np.array_equal(desc[i, :cnt[i]], d1)
    # strided input (ROI-like row stride) gives the same result
    big = np.zeros((480, 700), np.uint8); big[:, :640] = frames[0]
    k2, d2 = single(big[:, :640])
    k1, d1 = single(frames[0])
    assert k2.tobytes() == k1.tobytes() and np.array_equal(d1, d2)


def test_strided_pinned_and_unaligned_device_input(orb, synth):
    """A tight 1241-byte row stride from pinned host memory (one block copy + on-device re-pitch) and an unaligned device
    batch both equal the plain call."""
    import torch
    W, H, B = 1241, 376, 3
    frames = np.stack([synth.frame(s, W, H) for s in range(B)])
    single = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    ref = [single(frames[i]) for i in range(B)]
    ex = orb.ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    cap = ex.max_keypoints(W, H)
    pin = orb.PinnedArray((B, H, W), np.uint8); pin.array[...] = frames
    pk = orb.PinnedArray((B, cap), orb.KP_DTYPE); pd = orb.PinnedArray((B, cap, 32), np.uint8); pc = orb.PinnedArray((B,), np.int32)
    ex.extract_batch_async(pin.array, pk.array, pd.array, pc.array); ex.wait()
    for i in range(B):
        n = int(pc.array[i])
        assert n == len(ref[i][0]) and pk.array[i, :n].tobytes() == ref[i][0].tobytes() and np.array_equal(pd.array[i, :n], ref[i][1])
    # padded pinned rows (stride 1300) with a gap between frames
    pin2 = orb.PinnedArray((B, H + 5, 1300), np.uint8); pin2.array[:, :H, :W] = frames
    ex.extract_batch_async(pin2.array[:, :H, :W], pk.array, pd.array, pc.array); ex.wait()
    for i in range(B):
        n = int(pc.array[i])
        assert n == len(ref[i][0]) and pk.array[i, :n].tobytes() == ref[i][0].tobytes() and np.array_equal(pd.array[i, :n], ref[i][1])
    # device batch with an odd row stride
    d = torch.from_numpy(frames).cuda()
    dk = torch.zeros((B, cap, 7), dtype=torch.int32, device="cuda"); dd = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    dc = torch.zeros((B,), dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d.data_ptr(), B, W, H, W, W * H, dk.data_ptr(), dd.data_ptr(), cap, dc.data_ptr()); ex.wait()
    dk = dk.cpu().numpy().view(orb.KP_DTYPE).reshape(B, cap); dd = dd.cpu().numpy(); dc = dc.cpu().numpy()
    for i in range(B):
        n = int(dc[i])
        assert n == len(ref[i][0]) and dk[i, :n].tobytes() == ref[i][0].tobytes() and np.array_equal(dd[i, :n], ref[i][1])
    for p in (pin, pin2, pk, pd, pc):
        p.free()


def test_degenerate_inputs(orb, oracle):
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d is None
    k, d = ex(np.full((480, 640), 77, np.uint8))
    assert len(k) == 0 and d is None
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    k, d = ex(noise)
    ok, od = oracle.OracleExtractor(1000, trig_mode=1).extract(noise)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    with pytest.raises(orb.OrbCudaError):
        ex(np.zeros((60, 60), np.uint8))   # smaller than one FAST cell at the top level: the reference divides by zero


def test_octree_standalone(orb, oracle):
    rng = np.random.default_rng(11)
    for trial in range(12):
        w = int(rng.integers(100, 1300)); h = int(rng.integers(100, 500))
        n = int(rng.integers(1, 6000)); N = int(rng.integers(1, 500))
        # unique integer positions
        pos = rng.choice(w * h, size=min(n, w * h), replace=False)
        x = (pos % w).astype(np.int16); y = (pos // w).astype(np.int16)
        s = rng.integers(7, 60, len(x)).astype(np.uint8)
        if h > 2 * w:
            continue
        ref = oracle.distribute_octtree(x, y, s, 16, 16 + w, 16, 16 + h, N)
        got = orb.distribute_octtree(x, y, s, 16, 16 + w, 16, 16 + h, N)
        assert np.array_equal(ref, got), (trial, w, h, n, N, len(ref), len(got))


PARAM_CASES = [(640, 480, 2000, 1.2, 8, 20, 7),      # the 2x monocular-initialisation extractor (R21 Tracking.cc:125)
               (640, 480, 500, 1.2, 8, 20, 7),
               (640, 480, 1000, 1.1, 8, 20, 7),
               (800, 600, 1500, 1.5, 5, 25, 10),
               (640, 480, 1000, 1.2, 4, 12, 5),
               (1920, 1080, 3000, 1.2, 8, 20, 7),
               (376, 413, 800, 1.2, 6, 20, 7)]         # taller than wide: one quadtree root


@pytest.mark.parametrize("case", PARAM_CASES)
def test_other_parameters(orb, oracle, synth, case):
    w, h, nf, sf, nl, ini, mn = case
    img = synth.frame(5, w, h)
    kps, desc = orb.ORBextractor(nf, sf, nl, ini, mn)(img)
    okps, odesc = oracle.OracleExtractor(nf, sf, nl, ini, mn, trig_mode=1).extract(img, cap=40000)
    assert len(kps) == len(okps)
    assert kps.tobytes() == okps.tobytes()
    assert np.array_equal(desc, odesc)


def test_large_batch_equals_single(orb, oracle, synth):
    """Batches of >= 16 frames run the quadtree with the small (256-thread) CTA; single frames with the 512-thread one."""
    rng = np.random.default_rng(5)
    frames = [synth.frame(200 + s) for s in range(14)] + [synth.frame(300 + s, low_texture=True) for s in range(4)]
    frames += [rng.integers(0, 256, (480, 640), dtype=np.uint8), np.full((480, 640), 128, np.uint8)]
    frames = np.stack(frames)
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=len(frames))
    kps, desc, cnt = ex.extract_batch(frames)
    oe = oracle.OracleExtractor(1000, trig_mode=1)
    for i in range(len(frames)):
        ok, od = oe.extract(frames[i])
        assert cnt[i] == len(ok), i
        assert kps[i, :cnt[i]].tobytes() == ok.tobytes() and np.array_equal(desc[i, :cnt[i]], od), i


def test_stream_of_frames_kitti_shape(orb, oracle, synth):
    """BASELINE config 2 shape: 1241x376, nFeatures=2000 -- a short stream through one handle, batched."""
    frames = np.stack([synth.frame(100 + s, 1241, 376) for s in range(8)])
    ex = orb.ORBextractor(2000, 1.2, 8, 20, 7, max_width=1241, max_height=376, max_batch=4)
    oe = oracle.OracleExtractor(2000, trig_mode=1)
    for b0 in (0, 4):
        kps, desc, cnt = ex.extract_batch(frames[b0:b0 + 4])
        for i in range(4):
            ok, od = oe.extract(frames[b0 + i])
            assert cnt[i] == len(ok)
            assert kps[i, :cnt[i]].tobytes() == ok.tobytes() and np.array_equal(desc[i, :cnt[i]], od)


def test_size_change_and_reuse(orb, oracle, synth):
    """One handle, alternating image sizes (geometry is rebuilt; results must not depend on history)."""
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    for (w, h, seed) in [(640, 480, 1), (752, 480, 2), (640, 480, 3), (320, 240, 4), (640, 480, 1)]:
        img = synth.frame(seed, w, h)
        k, d = ex(img)
        ok, od = oracle.OracleExtractor(1000, trig_mode=1).extract(img)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (w, h, seed)


def test_concurrent_handles(orb, oracle, synth):
    """Two extractors used from two threads at once (the stereo Frame constructor, R21 Frame.cc:80-83)."""
    import threading
    left, right = synth.stereo_pair(3, 752, 480)
    res = {}

    def run(name, img):
        ex = orb.ORBextractor(1200, 1.2, 8, 20, 7)
        for _ in range(5):
            res[name] = ex(img)

    th = [threading.Thread(target=run, args=("l", left)), threading.Thread(target=run, args=("r", right))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for name, img in (("l", left), ("r", right)):
        ok, od = oracle.OracleExtractor(1200, trig_mode=1).extract(img)
        assert res[name][0].tobytes() == ok.tobytes() and np.array_equal(res[name][1], od)
