"""Matcher / frame parity against the REFERENCE's own code: the same scene is pushed through
   (A) the reference's ORBmatcher.cc on the reference's Frame / KeyFrame / MapPoint objects (oracle/_ref/libmatchref.so), and
   (B) the product's C++ drop-in ORBmatcher class on the same objects -- backed by the CPU restatement here (pins the oracle
       that every CUDA parity test compares against), and by liborbcuda on the GPU box (-m gpu: the drop-in itself, executed).
Every output must be identical: match lists, map-point assignments, the map state after Fuse, representative descriptors."""
import numpy as np
import pytest

import matchdata
import matchref_lib as mr

pytestmark = pytest.mark.skipif(not mr.available("ref"), reason="oracle/_ref/libmatchref.so not built (needs /root/reference once)")


def _f12(T1, T2, K):
    """LocalMapping::ComputeF12 (R21/src/LocalMapping.cc:561-577) in float32"""
    R1, t1, R2, t2 = T1[:3, :3], T1[:3, 3], T2[:3, :3], T2[:3, 3]
    R12 = (R1 @ R2.T).astype(np.float32)
    t12 = (-R1 @ R2.T @ t2 + t1).astype(np.float32)
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    Ki = np.linalg.inv(Km).astype(np.float32)
    return (Ki.T @ tx @ R12 @ Ki).astype(np.float32)


def _build(flavour, sc, obs):
    """frames f0..f2 (+ copies for the searches that write into a frame), key frames k0, k1 and the map."""
    w = mr.World(flavour, sc.K, sc.D, sc.bf, sc.th_depth, sc.cols, sc.rows)
    out = {}
    f = []
    for v in range(3):
        kps, desc, owner, ur, dep = obs[v]
        fi = w.frame_from_features(kps, desc, ur if sc.stereo_frac > 0 else None, dep if sc.stereo_frac > 0 else None)
        w.frame_set_pose(fi, sc.T[v])
        w.frame_set_featvec(fi, matchdata.featvec(desc))
        f.append(fi)
    k0, k1 = w.keyframe(f[0]), w.keyframe(f[1])
    # map: points seen by k0, about half of them by k1 as well; a second group seen by k1 only
    own0, own1 = obs[0][2], obs[1][2]
    idx1_of = {int(p): i for i, p in enumerate(own1) if p >= 0}
    rng = np.random.Generator(np.random.PCG64(99))
    mps = []
    for i0, p in enumerate(own0):
        if p < 0 or rng.random() > 0.6:
            continue
        m = w.mappoint(sc.P[p], k0); w.observe(m, k0, i0)
        if int(p) in idx1_of and rng.random() < 0.5:
            w.observe(m, k1, idx1_of[int(p)])
        w.mappoint_update(m)
        mps.append(m)
    only1 = []
    for i1, p in enumerate(own1):
        if p < 0 or rng.random() > 0.25:
            continue
        if w.keyframe_mappoints(k1)[i1] >= 0:
            continue
        m = w.mappoint(sc.P[p], k1); w.observe(m, k1, i1); w.mappoint_update(m)
        only1.append(m)
    return w, f, k0, k1, mps, only1


def _run(flavour, sc, obs):
    w, f, k0, k1, mps, only1 = _build(flavour, sc, obs)
    R = {}
    g = w.frame_get(f[0])
    R["kps_un"] = g["kps_un"].tobytes(); R["bounds"] = g["bounds"].copy(); R["cell_ptr"] = g["cell_ptr"]; R["cell_idx"] = g["cell_idx"]
    R["kf_bounds"] = w.keyframe_bounds(k0)
    R["area"] = [w.frame_features_in_area(f[0], 300.0 + 7 * i, 200.0 + 3 * i, 25.0 + i, -1 if i % 3 else 0, -1 if i % 2 else 3) for i in range(12)]
    R["kf_area"] = [w.keyframe_features_in_area(k0, 1.5 + 60 * i, 2.5 + 38 * i, 30.0) for i in range(12)]
    R["distinctive"] = np.stack([w.mappoint_get(m)["desc"] for m in mps[:200]])
    for ratio, ori in ((0.7, True), (0.75, False), (0.9, True)):
        R["bow_kf_f", ratio, ori] = w.search_by_bow_kf_f(k0, f[2], ratio, ori)
        R["bow_kf_kf", ratio, ori] = w.search_by_bow_kf_kf(k0, k1, ratio, ori)
    F12 = _f12(sc.T[0], sc.T[1], sc.K)
    for only_stereo in (False, True):
        for ori in (False, True):
            R["tri", only_stereo, ori] = w.search_for_triangulation(k0, k1, F12, only_stereo, 0.6, ori)
    # Tracking::SearchLocalPoints on a frame that already holds a few points
    kps2, desc2, own2, ur2, dep2 = obs[2]
    def fresh(view):
        kps, desc, owner, ur, dep = obs[view]
        fi = w.frame_from_features(kps, desc, ur if sc.stereo_frac > 0 else None, dep if sc.stereo_frac > 0 else None)
        w.frame_set_pose(fi, sc.T[view])
        return fi
    for th in (1.0, 3.0):
        c = fresh(2)
        for i in range(0, len(kps2), 9):
            w.frame_set_mappoint(c, i, mps[i % len(mps)])
        R["local", th] = w.search_by_projection_local(c, mps + only1, th, 0.8)
    # the same search with windows that span a large part of the frame and a strict ratio: hundreds of points compete for each
    # feature, the outcome of a point hangs on which features the points before it took (the order-exact rounds of the CUDA path)
    for th, ratio in ((15.0, 0.6), (15.0, 0.9)):
        c = fresh(2)
        for i in range(0, len(kps2), 9):
            w.frame_set_mappoint(c, i, mps[i % len(mps)])
        R["local_wide", th, ratio] = w.search_by_projection_local(c, (mps + only1) * 2, th, ratio)
    # TrackWithMotionModel: last = frame 1 carrying its key frame's points (some flagged outliers)
    last = fresh(1)
    kf1_mps = w.keyframe_mappoints(k1)
    for i, m in enumerate(kf1_mps):
        if m >= 0:
            w.frame_set_mappoint(last, i, int(m), outlier=(i % 17 == 0))
    for mono in ((True, False) if sc.stereo_frac > 0 else (True,)):
        for th in (7.0, 15.0):
            for ori in (True, False):
                c = fresh(2)
                R["last", mono, th, ori] = w.search_by_projection_last(c, last, th, mono, 0.9, ori)
    # Relocalization: project k0's points into frame 2
    for th, od in ((10.0, 100), (3.0, 64)):
        c = fresh(2)
        for i in range(0, len(kps2), 11):
            w.frame_set_mappoint(c, i, mps[(3 * i) % len(mps)])
        R["reloc", th] = w.search_by_projection_kf(c, k0, mps[::7], th, od, 0.9, True)
    # LoopClosing::ComputeSim3: points of the map into k1 under a similarity
    S = sc.T[1].copy(); S[:3, :] *= np.float32(1.04)
    n_bow, m12 = w.search_by_bow_kf_kf(k1, k0, 0.75, True)
    R["sim3proj"] = w.search_by_projection_sim3(k1, S, mps, m12, 10)
    # monocular initialisation between frames 0 and 1
    g0 = w.frame_get(f[0])
    prev = np.stack([g0["kps_un"]["x"], g0["kps_un"]["y"]], 1)
    for ori in (True, False):
        R["init", ori] = w.search_for_initialization(f[0], f[1], prev, 100, 0.9, ori)
    # SearchBySim3 k0 <-> k1 with the true relative pose
    T01 = sc.T[0] @ np.linalg.inv(sc.T[1])
    R["sim3"] = w.search_by_sim3(k0, k1, np.full(w.keyframe_n(k0), -1, np.int32), 1.0, T01[:3, :3], T01[:3, 3], 7.5)
    pre = w.search_by_bow_kf_kf(k0, k1, 0.75, True)[1]
    R["sim3_pre"] = w.search_by_sim3(k0, k1, pre, 1.0, T01[:3, :3], T01[:3, 3], 7.5)
    # Fuse (mutates the map): LoopClosing flavour first, then LocalMapping's
    R["fuse_sim3"] = w.fuse_sim3(k1, sc.T[1], mps, 4.0)
    R["fuse_sim3_state"] = (w.keyframe_mappoints(k1), [tuple(sorted(w.mappoint_get(m).items(), key=lambda kv: kv[0])[1:3]) for m in mps[:50]])
    R["fuse"] = w.fuse(k0, only1 + [-1] + mps[:40], 3.0)
    R["fuse_k0"] = w.keyframe_mappoints(k0)
    R["fuse_state"] = [(s["observations"], s["bad"], s["replaced"], s["desc"].tobytes()) for s in (w.mappoint_get(m) for m in only1 + mps)]
    w.close()
    return R


def _same(a, b, key=""):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), key
        for k in a:
            _same(a[k], b[k], "%s/%s" % (key, k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), key
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, "%s[%d]" % (key, i))
    elif isinstance(a, np.ndarray):
        assert np.array_equal(a, b), "%s: %d of %d entries differ" % (key, int((np.asarray(a) != np.asarray(b)).sum()), a.size)
    else:
        assert a == b, (key, a, b)


def _scene(seed, distortion, stereo):
    sc = mr.Scene(seed, distortion=distortion, stereo_frac=0.35 if stereo else 0.0)
    obs = [sc.observe(0), sc.observe(1, angle_offset=-20.0), sc.observe(2, angle_offset=12.0)]
    return sc, obs


CASES = [(1, True, False), (2, False, False), (3, True, True)]


@pytest.mark.skipif(not mr.available("shim_cpu"), reason="oracle/_ref/libmatchshim_cpu.so not built")
@pytest.mark.parametrize("case", CASES)
def test_restatement_equals_reference_matcher(case):
    """the CPU restatement behind the drop-in class == the reference's ORBmatcher.cc, on the reference's own objects"""
    sc, obs = _scene(*case)
    ref = _run("ref", sc, obs)
    got = _run("shim_cpu", sc, obs)
    # the scene must exercise the loops, not pass vacuously
    assert ref["bow_kf_f", 0.7, True][0] > 150 and ref["bow_kf_kf", 0.75, False][0] > 100
    assert ref["tri", False, False][0] > 30 and ref["local", 3.0][0] > 200 and ref["last", True, 15.0, True][0] > 150
    assert ref["reloc", 10.0][0] > 100 and ref["sim3proj"][0] > 50 and ref["init", True][0] > 100
    assert ref["sim3"][0] > 50 and ref["fuse_sim3"][0] > 50 and ref["fuse"] > 20
    _same(ref, got)


@pytest.mark.gpu
@pytest.mark.skipif(not mr.available("shim_cuda"), reason="oracle/_ref/libmatchshim_cuda.so not built")
@pytest.mark.parametrize("case", CASES)
def test_cuda_dropin_equals_reference_matcher(case):
    """the product's C++ drop-in (shim/ORBmatcher.cc -> liborbcuda) == the reference's ORBmatcher.cc, executed"""
    sc, obs = _scene(*case)
    _same(_run("ref", sc, obs), _run("shim_cuda", sc, obs))


@pytest.mark.parametrize("flavour", ["shim_cpu"])
def test_descriptor_distance(flavour):
    if not mr.available(flavour):
        pytest.skip("not built")
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 32), dtype=np.uint8); b = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    wr = mr.lib("ref"); ws = mr.lib(flavour)
    import ctypes as C
    for L in (wr, ws):
        L.mh_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    for i in range(64):
        ref = wr.mh_descriptor_distance(a[i].ctypes.data, b[i].ctypes.data)
        assert ref == ws.mh_descriptor_distance(a[i].ctypes.data, b[i].ctypes.data) == int(np.unpackbits(a[i] ^ b[i]).sum())


@pytest.mark.gpu
@pytest.mark.skipif(not mr.available("shim_cuda"), reason="oracle/_ref/libmatchshim_cuda.so not built")
def test_cuda_dropin_frames_from_images(synth):
    """The reference's Frame constructors (Frame.cc:63-119, :176-233) on top of the product's ORBextractor class (CUDA) against the
    same constructors on the reference's own extractor: key points, descriptors, undistortion, grid, and the stereo matches --
    once computed by the reference's ComputeStereoMatches on the drop-in's mvImagePyramid mirror and once by the GPU search."""
    K = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    # monocular, TUM shape
    img = synth.frame(5, 640, 480)
    got = {}
    for fl in ("ref", "shim_cuda"):
        w = mr.World(fl, K, D, 40.0, 35.0, 640, 480, nfeatures=1000)
        got[fl] = w.frame_get(w.frame_from_image(img))
        w.close()
    a, b = got["ref"], got["shim_cuda"]
    assert len(a["kps"]) == len(b["kps"]) > 900
    assert a["kps"].tobytes() == b["kps"].tobytes() and a["kps_un"].tobytes() == b["kps_un"].tobytes()
    assert np.array_equal(a["cell_ptr"], b["cell_ptr"]) and np.array_equal(a["cell_idx"], b["cell_idx"])
    assert int((a["desc"] != b["desc"]).any(1).sum()) <= max(1, len(a["desc"]) // 1000)
    # stereo, EuRoC shape
    left, right = synth.stereo_pair(3, 752, 480)
    mbf, fx = 47.90639384423901, 458.654
    res = {}
    for fl in ("ref", "shim_cuda"):
        w = mr.World(fl, K, np.zeros(4, np.float32), mbf, 35.0, 752, 480, nfeatures=1200)
        f = w.frame_from_stereo(left, right, mbf / fx)
        res[fl] = w.frame_get(f)
        if fl == "shim_cuda":
            res["accel"] = w.frame_stereo_accel(f)
        w.close()
    a, b = res["ref"], res["shim_cuda"]
    assert a["kps"].tobytes() == b["kps"].tobytes() and a["kps_right"].tobytes() == b["kps_right"].tobytes()
    same_desc = np.array_equal(a["desc"], b["desc"]) and np.array_equal(a["desc_right"], b["desc_right"])
    assert (a["u_right"] >= 0).sum() > 100
    if same_desc:
        assert np.array_equal(a["u_right"], b["u_right"]) and np.array_equal(a["depth"], b["depth"])
        m, ur, dep = res["accel"]
        assert m == int((a["u_right"] >= 0).sum()) and np.array_equal(ur, a["u_right"]) and np.array_equal(dep, a["depth"])
    else:      # a descriptor at a rounding boundary (<= 0.1 %): the match lists may differ in those few features
        assert (a["u_right"] != b["u_right"]).sum() <= 5


@pytest.mark.gpu
@pytest.mark.skipif(not mr.available("shim_cuda"), reason="oracle/_ref/libmatchshim_cuda.so not built")
def test_cuda_dropin_single_frame_calls(synth):
    """operator() called frame after frame through the C++ class (the reference's real calling pattern): the replayed CUDA graph
    and the one-synchronisation pyramid mirror give the same key points as the first (plain) call, and the latency is reported."""
    K = np.array([517.3, 516.5, 318.6, 255.3], np.float32)
    w = mr.World("shim_cuda", K, np.zeros(4, np.float32), 40.0, 35.0, 640, 480, nfeatures=1000)
    imgs = [synth.frame(s, 640, 480) for s in range(3)]
    first = [w.frame_get(w.frame_from_image(im))["kps"].tobytes() for im in imgs]
    again = [w.frame_get(w.frame_from_image(im))["kps"].tobytes() for im in imgs]      # these run through the captured graph
    assert first == again
    wr = mr.World("ref", K, np.zeros(4, np.float32), 40.0, 35.0, 640, 480, nfeatures=1000)
    assert [wr.frame_get(wr.frame_from_image(im))["kps"].tobytes() for im in imgs] == first
    ms_on, n = w.extract_latency_ms(imgs[0], 200, True)
    ms_off, _ = w.extract_latency_ms(imgs[0], 200, False)
    ms_ref, _ = wr.extract_latency_ms(imgs[0], 5, True)
    print("\nshim operator() per frame: %.3f ms with the mvImagePyramid mirror, %.3f ms without; reference CPU %.1f ms (%d key points)"
          % (ms_on, ms_off, ms_ref, n))
    assert n > 900 and ms_off <= ms_on
    w.close(); wr.close()
