"""Frame-side steps (undistortion, image bounds, 64x48 grid, GetFeaturesInArea) and the projection search that consumes
them.  CPU part: the oracle against cv2 golden vectors and against brute-force properties.
GPU part (-m gpu): the CUDA path against the oracle, bit for bit."""
import os
import numpy as np
import pytest

CALS = ("tum1", "tum2", "four", "barrel")


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "undistort_cv2.npz"))


def _keys(oracle, n, seed, w=640, h=480, nlevels=8):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, oracle.KP_DTYPE)
    lvl = rng.integers(0, nlevels, n)
    s = 1.2 ** lvl
    # extractor-like coordinates: integers at the level, scaled up to level 0
    k["x"] = (rng.integers(16, (w / s - 16).astype(int).clip(17)) * s).astype(np.float32)
    k["y"] = (rng.integers(16, (h / s - 16).astype(int).clip(17)) * s).astype(np.float32)
    k["size"] = (31 * s).astype(np.float32); k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["response"] = rng.integers(7, 200, n).astype(np.float32); k["octave"] = lvl; k["class_id"] = -1
    return k


def _scene(oracle, synth, nf, nmp, seed, K, dist, crowded=False):
    """A frame with nf features and nmp projected map points around (some of) them."""
    rng = np.random.default_rng(seed)
    keys = _keys(oracle, nf, seed)
    if crowded:      # many points fight for the same few features: forces many dependent rounds
        keys["x"] = (200 + rng.integers(0, 60, nf)).astype(np.float32); keys["y"] = (200 + rng.integers(0, 40, nf)).astype(np.float32)
    desc_f = synth.descriptors(nf, seed=seed + 1)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    return keys, desc_f, sf, rng


def _map_points(oracle, keys_un, desc_f, nmp, rng, noise_bits=20):
    nf = len(keys_un)
    mp = np.zeros(nmp, oracle.MPV_DTYPE)
    src = rng.integers(0, nf, nmp)
    mp["proj_x"] = keys_un["x"][src] + rng.normal(0, 2.0, nmp).astype(np.float32)
    mp["proj_y"] = keys_un["y"][src] + rng.normal(0, 2.0, nmp).astype(np.float32)
    mp["proj_xr"] = mp["proj_x"] - rng.uniform(0, 30, nmp).astype(np.float32)
    mp["view_cos"] = rng.choice(np.array([0.9, 0.9975, 0.998, 0.9985, 1.0], np.float32), nmp)
    mp["level"] = np.clip(keys_un["octave"][src] + rng.integers(0, 2, nmp), 0, 7)
    mp["in_view"] = rng.random(nmp) < 0.9
    mp["obs_positive"] = rng.random(nmp) < 0.85
    dm = desc_f[src].copy()
    flips = rng.integers(0, 256, (nmp, noise_bits))
    for i in range(nmp):
        for b in flips[i][: rng.integers(0, noise_bits + 1)]:
            dm[i, b >> 3] ^= np.uint8(1 << (b & 7))
    far = rng.random(nmp) < 0.1      # some points look like nothing in the frame
    dm[far] = rng.integers(0, 256, (int(far.sum()), 32), dtype=np.uint8)
    return mp, dm


# ------------------------------------------------------------------------------------------------ CPU: oracle
@pytest.mark.parametrize("cal", CALS)
def test_oracle_undistort_matches_cv2_golden(oracle, golden_dir, cal):
    g = _golden(golden_dir)
    out = oracle.undistort_points(g[cal + "_pts"], g[cal + "_K"], g[cal + "_D"])
    assert np.array_equal(out.view(np.uint32), g[cal + "_und"].view(np.uint32))      # bit-exact with cv2 4.13.0


def test_oracle_undistort_identity_and_bounds(oracle, golden_dir):
    g = _golden(golden_dir)
    k = _keys(oracle, 500, 3)
    K = g["tum1_K"]
    same = oracle.undistort_keypoints(k, K, np.zeros(5, np.float32))
    assert same.tobytes() == k.tobytes()
    assert np.array_equal(oracle.image_bounds(640, 480, K, np.zeros(4, np.float32)), np.array([0, 640, 0, 480], np.float32))
    b = oracle.image_bounds(640, 480, K, g["tum1_D"])
    c = oracle.undistort_points(np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32), K, g["tum1_D"])
    assert b[0] == min(c[0, 0], c[2, 0]) and b[1] == max(c[1, 0], c[3, 0]) and b[2] == min(c[0, 1], c[1, 1]) and b[3] == max(c[2, 1], c[3, 1])
    un = oracle.undistort_keypoints(k, K, g["tum1_D"])
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(un[f], k[f])
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), oracle.undistort_points(np.stack([k["x"], k["y"]], 1), K, g["tum1_D"]))


def test_oracle_grid_and_area_properties(oracle, golden_dir):
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    k = oracle.undistort_keypoints(_keys(oracle, 3000, 5), K, D)
    b = oracle.image_bounds(640, 480, K, D)
    ptr, idx = oracle.assign_grid(k, b)
    winv = np.float32(64) / np.float32(b[1] - b[0]); hinv = np.float32(48) / np.float32(b[3] - b[2])
    px = np.round((k["x"] - b[0]) * winv).astype(int); py = np.round((k["y"] - b[2]) * hinv).astype(int)      # PosInGrid: round
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert ptr[-1] == inside.sum() and np.all(np.diff(ptr) >= 0)
    for c in np.flatnonzero(np.diff(ptr))[:400]:
        members = idx[ptr[c]:ptr[c + 1]]
        assert np.all(np.diff(members) > 0)                              # push_back order
        assert np.all(px[members] * 48 + py[members] == c)
    rng = np.random.default_rng(1)
    for _ in range(300):
        x, y, r = np.float32(rng.uniform(-20, 660)), np.float32(rng.uniform(-20, 500)), np.float32(rng.uniform(1, 40))
        lo, hi = int(rng.integers(-1, 7)), int(rng.integers(-1, 8))
        got = oracle.features_in_area(k, ptr, idx, b, x, y, r, lo, hi)
        ok = (np.abs(k["x"] - x) < r) & (np.abs(k["y"] - y) < r) & inside
        if lo > 0 or hi >= 0:
            ok &= k["octave"] >= lo
            if hi >= 0:
                ok &= k["octave"] <= hi
        assert set(got) <= set(np.flatnonzero(ok)) and len(set(got)) == len(got)
        cells = px[got] * 48 + py[got]
        assert np.all(np.diff(cells) >= 0)                               # ix-major, then iy, then push_back order


def test_oracle_projection_search_is_order_dependent_greedy(oracle, synth, golden_dir):
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, 600, 0, 11, K, D)
    k = oracle.undistort_keypoints(keys, K, D); b = oracle.image_bounds(640, 480, K, D)
    ptr, idx = oracle.assign_grid(k, b)
    mp, dm = _map_points(oracle, k, desc_f, 900, rng)
    ur = np.full(len(k), -1, np.float32); occ = np.zeros(len(k), np.uint8)
    fp, pf, n = oracle.search_by_projection_frame(k, desc_f, ur, occ, ptr, idx, b, sf, mp, dm, 1.0, 0.8)
    assert n == (pf >= 0).sum() > 100
    # a feature keeps the LAST point that took it; a point with observations blocks it for the later ones
    for f in np.flatnonzero(fp >= 0):
        takers = np.flatnonzero(pf == f)
        assert fp[f] == takers.max()
        assert np.all(mp["obs_positive"][takers[:-1]] == 0)
    assert np.all(pf[mp["in_view"] == 0] == -1)


# ------------------------------------------------------------------------------------------------ GPU: CUDA vs oracle
@pytest.fixture(scope="module")
def orb():
    import orbcuda
    if orbcuda.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return orbcuda


@pytest.mark.gpu
@pytest.mark.parametrize("cal", CALS)
def test_gpu_undistort_bounds_grid(orb, oracle, golden_dir, cal):
    g = _golden(golden_dir)
    K, D = g[cal + "_K"], g[cal + "_D"]
    pts = g[cal + "_pts"]
    k = np.zeros(len(pts), oracle.KP_DTYPE); k["x"] = pts[:, 0]; k["y"] = pts[:, 1]; k["octave"] = np.arange(len(pts)) % 8
    fr = orb.FrameFeatures(k, K, D, 640, 480)
    und = np.stack([fr.keys_un["x"], fr.keys_un["y"]], 1)
    assert np.array_equal(und.view(np.uint32), g[cal + "_und"].view(np.uint32))          # cv2 golden, bit-exact
    assert fr.keys_un.tobytes() == oracle.undistort_keypoints(k, K, D).tobytes()
    assert np.array_equal(fr.bounds, oracle.image_bounds(640, 480, K, D))
    ptr, idx = oracle.assign_grid(fr.keys_un, fr.bounds)
    assert np.array_equal(fr.cell_ptr, ptr) and np.array_equal(fr.cell_idx[:ptr[-1]], idx[:ptr[-1]]) and fr.n_assigned == ptr[-1]


@pytest.mark.gpu
def test_gpu_undistort_identity_and_empty(orb, oracle, golden_dir):
    g = _golden(golden_dir)
    k = _keys(oracle, 777, 9)
    fr = orb.FrameFeatures(k, g["tum1_K"], np.zeros(5, np.float32), 752, 480)
    assert fr.keys_un.tobytes() == k.tobytes() and np.array_equal(fr.bounds, np.array([0, 752, 0, 480], np.float32))
    ptr, idx = oracle.assign_grid(k, fr.bounds)
    assert np.array_equal(fr.cell_ptr, ptr) and np.array_equal(fr.cell_idx[:ptr[-1]], idx[:ptr[-1]])
    empty = orb.FrameFeatures(np.zeros(0, oracle.KP_DTYPE), g["tum1_K"], g["tum1_D"], 640, 480)
    assert empty.n_assigned == 0 and not empty.cell_ptr.any()
    assert len(empty.GetFeaturesInArea(100.0, 100.0, 50.0)) == 0


@pytest.mark.gpu
def test_gpu_features_in_area(orb, oracle, golden_dir):
    g = _golden(golden_dir)
    K, D = g["tum2_K"], g["tum2_D"]
    k = _keys(oracle, 2500, 21)
    fr = orb.FrameFeatures(k, K, D, 640, 480)
    rng = np.random.default_rng(2)
    nq = 700
    x = rng.uniform(-30, 670, nq).astype(np.float32); y = rng.uniform(-30, 510, nq).astype(np.float32)
    r = rng.uniform(0.5, 60, nq).astype(np.float32)
    lo = rng.integers(-1, 7, nq).astype(np.int32); hi = rng.integers(-1, 8, nq).astype(np.int32)
    x[:3] = [-500, 5000, 320]; y[:3] = [240, 240, -900]                 # windows off the grid (early returns :338-351)
    got = fr.GetFeaturesInArea(x, y, r, lo, hi)
    total = 0
    for q in range(nq):
        ref = oracle.features_in_area(fr.keys_un, fr.cell_ptr, fr.cell_idx, fr.bounds, x[q], y[q], r[q], int(lo[q]), int(hi[q]))
        assert np.array_equal(got[q], ref), q
        total += len(ref)
    assert total > 1000
    one = fr.GetFeaturesInArea(320.0, 240.0, 25.0)
    assert np.array_equal(one, oracle.features_in_area(fr.keys_un, fr.cell_ptr, fr.cell_idx, fr.bounds, 320.0, 240.0, 25.0, -1, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("nf,nmp,crowded,th", [(1200, 2500, False, 1.0), (1500, 4000, False, 3.0), (300, 3000, True, 1.0),
                                              (64, 900, True, 5.0), (1, 5, False, 1.0), (9000, 5000, False, 2.0)])
def test_gpu_search_by_projection_frame(orb, oracle, synth, golden_dir, nf, nmp, crowded, th):
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, nmp, 100 + nf, K, D, crowded)
    fr = orb.FrameFeatures(keys, K, D, 640, 480)
    mp, dm = _map_points(oracle, fr.keys_un, desc_f, nmp, rng)
    ur = np.where(rng.random(nf) < 0.5, fr.keys_un["x"] - rng.uniform(0, 30, nf), -1).astype(np.float32)      # stereo for half
    occ = (rng.random(nf) < 0.15).astype(np.uint8)
    for nnratio in (0.8, 0.6):
        ref_fp, ref_pf, ref_n = oracle.search_by_projection_frame(fr.keys_un, desc_f, ur, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf,
                                                                  mp, dm, th, nnratio)
        fp, pf, n = orb.search_by_projection_frame(fr, desc_f, ur, occ, sf, mp, dm, th=th, nnratio=nnratio)
        assert n == ref_n and np.array_equal(pf, ref_pf) and np.array_equal(fp, ref_fp)
    if nf >= 300:
        assert ref_n > 50


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nf,nmp,ci", [(907405640, 641, 4287, 0), (907405640, 641, 4287, 1), (5, 400, 4500, 0), (6, 900, 3000, 1)])
def test_gpu_search_by_projection_frame_wide_windows(orb, oracle, synth, golden_dir, seed, nf, nmp, ci):
    """th = 15 in a crowded scene: every window holds a large part of the frame, points wait for each other over ~15 rounds, and a
    point that becomes final early (a LATER one in the reference's order) takes a feature that is the second best of an earlier,
    still unresolved point -- the earlier point must go on seeing it (found by tools/soak_frame.py, seed 7)."""
    g = _golden(golden_dir)
    K, D = (g["tum1_K"], g["tum1_D"]) if ci == 0 else (np.array([500, 500, 320, 240], np.float32), np.zeros(4, np.float32))
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, 0, seed, K, D, True)
    fr = orb.FrameFeatures(keys, K, D, 640, 480)
    ur = np.where(rng.random(nf) < 0.5, fr.keys_un["x"] - rng.uniform(0, 30, nf), -1).astype(np.float32)
    occ = (rng.random(nf) < 0.1).astype(np.uint8)
    mp, dm = _map_points(oracle, fr.keys_un, desc_f, nmp, rng)
    for th, nnratio in ((15.0, 0.6), (15.0, 0.9), (7.0, 0.6)):
        ref_fp, ref_pf, ref_n = oracle.search_by_projection_frame(fr.keys_un, desc_f, ur, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf,
                                                                  mp, dm, th, nnratio)
        fp, pf, n = orb.search_by_projection_frame(fr, desc_f, ur, occ, sf, mp, dm, th=th, nnratio=nnratio)
        assert n == ref_n and np.array_equal(pf, ref_pf) and np.array_equal(fp, ref_fp), (th, nnratio)
    assert ref_n > 50


def _proj_points(oracle, keys_un, desc_f, npts, rng, rot_deg=12.0):
    """Projected points of a previous frame: most land near a current feature and carry a noisy copy of its descriptor;
    their angles differ from the current ones by a common rotation plus outliers (exercises the rotation histogram)."""
    nf = len(keys_un)
    p = np.zeros(npts, oracle.PROJ_DTYPE)
    src = rng.integers(0, nf, npts)
    p["u"] = keys_un["x"][src] + rng.normal(0, 3.0, npts).astype(np.float32)
    p["v"] = keys_un["y"][src] + rng.normal(0, 3.0, npts).astype(np.float32)
    p["ur"] = p["u"] - rng.uniform(0, 30, npts).astype(np.float32)
    ang = keys_un["angle"][src] + np.float32(rot_deg) + rng.normal(0, 3.0, npts).astype(np.float32)
    out = rng.random(npts) < 0.15
    ang[out] = rng.uniform(0, 360, int(out.sum()))
    p["angle"] = np.mod(ang, 360).astype(np.float32)
    p["octave"] = np.clip(keys_un["octave"][src] + rng.integers(-1, 2, npts), 0, 7)
    p["valid"] = rng.random(npts) < 0.9
    p["obs_positive"] = rng.random(npts) < 0.85
    dp = desc_f[src].copy()
    for i in range(npts):
        for b in rng.integers(0, 256, rng.integers(0, 25)):
            dp[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return p, dp


def test_oracle_last_frame_search_rotation_check(oracle, synth, golden_dir):
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, 800, 0, 31, K, D)
    k = oracle.undistort_keypoints(keys, K, D); b = oracle.image_bounds(640, 480, K, D)
    ptr, idx = oracle.assign_grid(k, b)
    pts, dp = _proj_points(oracle, k, desc_f, 700, rng)
    ur = np.full(len(k), -1, np.float32); occ = np.zeros(len(k), np.uint8)
    fp0, pf0, n0 = oracle.search_by_projection_last_frame(k, desc_f, ur, occ, ptr, idx, b, sf, pts, dp, 15.0, 0, False)
    fp1, pf1, n1 = oracle.search_by_projection_last_frame(k, desc_f, ur, occ, ptr, idx, b, sf, pts, dp, 15.0, 0, True)
    assert np.array_equal(pf0, pf1) and n0 == (pf0 >= 0).sum() > 200
    removed = (fp1 == -2)
    assert 0 < removed.sum() < n0 and n1 < n0                     # the angle outliers go, the bulk stays
    assert np.array_equal(fp0[~removed], fp1[~removed])


@pytest.mark.gpu
@pytest.mark.parametrize("nf,npts,crowded,th,direction", [(1200, 1100, False, 15.0, 0), (1200, 1100, False, 7.0, 1), (1500, 2000, False, 30.0, 2),
                                                         (200, 1500, True, 15.0, 0), (1, 3, False, 7.0, 0)])
def test_gpu_search_by_projection_last_frame(orb, oracle, synth, golden_dir, nf, npts, crowded, th, direction):
    g = _golden(golden_dir)
    K, D = g["tum2_K"], g["tum2_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, 0, 300 + nf + direction, K, D, crowded)
    fr = orb.FrameFeatures(keys, K, D, 640, 480)
    pts, dp = _proj_points(oracle, fr.keys_un, desc_f, npts, rng)
    ur = np.where(rng.random(nf) < 0.5, fr.keys_un["x"] - rng.uniform(0, 30, nf), -1).astype(np.float32)
    occ = (rng.random(nf) < 0.1).astype(np.uint8)
    for check in (True, False):
        ref = oracle.search_by_projection_last_frame(fr.keys_un, desc_f, ur, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th,
                                                     direction, check)
        got = orb.search_by_projection_last_frame(fr, desc_f, ur, occ, sf, pts, dp, th, direction=direction, check_orientation=check)
        assert got[2] == ref[2] and np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
    if nf >= 200:
        assert ref[2] > 30


@pytest.mark.gpu
@pytest.mark.parametrize("nf,npts,crowded,th,orb_dist", [(1300, 900, False, 10.0, 100), (1300, 1500, False, 3.0, 64), (150, 1200, True, 10.0, 100)])
def test_gpu_search_by_projection_keyframe(orb, oracle, synth, golden_dir, nf, npts, crowded, th, orb_dist):
    g = _golden(golden_dir)
    K, D = g["four_K"], g["four_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, 0, 500 + nf + orb_dist, K, D, crowded)
    fr = orb.FrameFeatures(keys, K, D, 752, 480)
    pts, dp = _proj_points(oracle, fr.keys_un, desc_f, npts, rng)
    occ = (rng.random(nf) < 0.2).astype(np.uint8)
    for check in (True, False):
        ref = oracle.search_by_projection_keyframe(fr.keys_un, desc_f, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th, orb_dist, check)
        got = orb.search_by_projection_keyframe(fr, desc_f, occ, sf, pts, dp, th, orb_dist, check_orientation=check)
        assert got[2] == ref[2] and np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
    assert ref[2] > 20


@pytest.mark.gpu
@pytest.mark.parametrize("nf,npts,crowded,th", [(1400, 2500, False, 10.0), (180, 1500, True, 4.0)])
def test_gpu_search_by_projection_sim3(orb, oracle, synth, golden_dir, nf, npts, crowded, th):
    g = _golden(golden_dir)
    K, D = g["barrel_K"], g["barrel_D"]
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, 0, 700 + nf, K, D, crowded)
    fr = orb.FrameFeatures(keys, K, D, 640, 480)
    pts, dp = _proj_points(oracle, fr.keys_un, desc_f, npts, rng)
    occ = (rng.random(nf) < 0.3).astype(np.uint8)
    ref = oracle.search_by_projection_sim3(fr.keys_un, desc_f, occ, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, pts, dp, th)
    got = orb.search_by_projection_sim3(fr, desc_f, occ, sf, pts, dp, th)
    assert got[2] == ref[2] > 10 and np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


@pytest.mark.gpu
@pytest.mark.parametrize("gated", [True, False])
def test_gpu_window_best_match_fuse_and_sim3(orb, oracle, synth, golden_dir, gated):
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    nf, npts = 1600, 3000
    keys, desc_f, sf, rng = _scene(oracle, synth, nf, 0, 900 + gated, K, D)
    fr = orb.FrameFeatures(keys, K, D, 640, 480)
    pts, dp = _proj_points(oracle, fr.keys_un, desc_f, npts, rng)
    ur = np.where(rng.random(nf) < 0.5, fr.keys_un["x"] - rng.uniform(0, 30, nf), -1).astype(np.float32)
    pts["ur"] = np.where(rng.random(npts) < 0.7, pts["u"] - 15, pts["ur"])
    sig = (1.0 / (sf * sf)).astype(np.float32) if gated else None
    for th in (3.0, 2.5, 8.0):
        ref = oracle.window_best_match(fr.keys_un, desc_f, ur if gated else None, fr.cell_ptr, fr.cell_idx, fr.bounds, sf, sig, pts, dp, th)
        got = orb.window_best_match(fr, desc_f, sf, pts, dp, th, u_right=ur if gated else None, inv_level_sigma2=sig)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        assert (ref[0] >= 0).sum() > 100
    assert np.all(ref[0][pts["valid"] == 0] == -1) and np.all(ref[1][pts["valid"] == 0] == 256)


@pytest.mark.gpu
def test_gpu_build_frames_on_extractor_device_output(orb, oracle, synth, golden_dir):
    """extract_batch_device -> orbf_build_frames_device on the same stream: undistorted key points and grids of a whole
    batch without leaving the device == oracle per frame."""
    import ctypes as C
    import torch
    g = _golden(golden_dir)
    K, D = g["tum1_K"], g["tum1_D"]
    B, W, H, cap = 6, 640, 480, 1400
    frames = np.stack([synth.frame(40 + i, W, H) for i in range(B)])
    frames[4] = 128                                   # a frame without key points
    ext = orb.ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    dev = torch.device("cuda", 0)
    d_img = torch.from_numpy(frames).to(dev)
    d_k = torch.zeros((B, cap, 7), dtype=torch.int32, device=dev); d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_c = torch.zeros((B,), dtype=torch.int32, device=dev)
    ext.extract_batch_device(d_img.data_ptr(), B, W, H, W, W * H, d_k.data_ptr(), d_d.data_ptr(), cap, d_c.data_ptr())
    d_un = torch.zeros_like(d_k); d_ptr = torch.zeros((B, 64 * 48 + 1), dtype=torch.int32, device=dev)
    d_idx = torch.full((B, cap), -7, dtype=torch.int32, device=dev)
    bounds = oracle.image_bounds(W, H, K, D)
    rc = orb.lib().orbf_build_frames_device(C.c_void_p(d_k.data_ptr()), C.c_void_p(d_c.data_ptr()), B, cap, K.ctypes.data_as(C.c_void_p),
                                            D.ctypes.data_as(C.c_void_p), len(D), bounds.ctypes.data_as(C.c_void_p), C.c_void_p(d_un.data_ptr()),
                                            C.c_void_p(d_ptr.data_ptr()), C.c_void_p(d_idx.data_ptr()), C.c_void_p(ext.stream()))
    assert rc == 0, orb.lib().orb_last_error()
    ext.wait()
    torch.cuda.synchronize()
    cnt = d_c.cpu().numpy()
    keys = d_k.cpu().numpy().reshape(B, -1).view(orb.KP_DTYPE).reshape(B, cap)
    un = d_un.cpu().numpy().reshape(B, -1).view(orb.KP_DTYPE).reshape(B, cap)
    ptr = d_ptr.cpu().numpy(); idx = d_idx.cpu().numpy()
    assert cnt[4] == 0 and cnt[0] > 500
    for f in range(B):
        n = cnt[f]
        ref_un = oracle.undistort_keypoints(keys[f, :n], K, D)
        assert un[f, :n].tobytes() == ref_un.tobytes()
        rptr, ridx = oracle.assign_grid(ref_un, bounds)
        assert np.array_equal(ptr[f], rptr) and np.array_equal(idx[f, :rptr[-1]], ridx[:rptr[-1]])


@pytest.mark.gpu
def test_gpu_frame_api_rejects_bad_arguments(orb, oracle, golden_dir):
    """Error behaviour of the C ABI: status codes, not crashes."""
    import ctypes as C
    g = _golden(golden_dir)
    L = orb.lib()
    k = _keys(oracle, 10, 1)
    out = np.zeros_like(k)
    K = g["tum1_K"]; D = g["tum1_D"]
    assert L.orbf_undistort_keypoints(k.ctypes.data_as(C.c_void_p), 10, K.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), 3,
                                      out.ctypes.data_as(C.c_void_p), 0) == 1          # ORB_ERR_ARG: 3 distortion coefficients
    assert L.orbf_undistort_keypoints(None, 10, K.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), 5, out.ctypes.data_as(C.c_void_p), 0) == 1
    assert b"orbf_undistort_keypoints" in L.orb_last_error()
    fr = orb.FrameFeatures(k, K, D, 640, 480)
    mp = np.zeros(3, orb.MPV_DTYPE); mp["in_view"] = 1; mp["level"] = 9                  # level outside mvScaleFactors
    with pytest.raises(orb.OrbCudaError):
        orb.search_by_projection_frame(fr, np.zeros((10, 32), np.uint8), np.full(10, -1, np.float32), np.zeros(10, np.uint8),
                                       (1.2 ** np.arange(8)).astype(np.float32), mp, np.zeros((3, 32), np.uint8))
    big = np.zeros(40000, oracle.KP_DTYPE)
    ptr = np.zeros(64 * 48 + 1, np.int32); idx = np.zeros(40000, np.int32); n = C.c_int(0)
    assert L.orbf_assign_grid(big.ctypes.data_as(C.c_void_p), 40000, fr.bounds.ctypes.data_as(C.c_void_p), ptr.ctypes.data_as(C.c_void_p),
                              idx.ctypes.data_as(C.c_void_p), C.byref(n), 0) == 1       # more than 32768 key points
    # capacity: the total is reported, the caller retries
    qx = np.full(4, 320, np.float32); qr = np.full(4, 1000, np.float32); lv = np.full(4, -1, np.int32)
    optr = np.zeros(5, np.int32); oidx = np.zeros(3, np.int32)
    rc = L.orbf_features_in_area(fr.keys_un.ctypes.data_as(C.c_void_p), 10, fr.cell_ptr.ctypes.data_as(C.c_void_p), fr.cell_idx.ctypes.data_as(C.c_void_p),
                                 fr.bounds.ctypes.data_as(C.c_void_p), qx.ctypes.data_as(C.c_void_p), qx.ctypes.data_as(C.c_void_p),
                                 qr.ctypes.data_as(C.c_void_p), lv.ctypes.data_as(C.c_void_p), lv.ctypes.data_as(C.c_void_p), 4,
                                 optr.ctypes.data_as(C.c_void_p), oidx.ctypes.data_as(C.c_void_p), 3, 0)
    assert rc == 3 and optr[4] == 4 * fr.n_assigned                                      # ORB_ERR_CAPACITY, out_ptr complete


def _init_pair(oracle, synth, n, seed, crowded=False):
    """Two frames for the monocular initialiser: frame 2 = frame 1 moved by a few pixels, descriptors with a little noise."""
    rng = np.random.default_rng(seed)
    k1 = _keys(oracle, n, seed)
    k1["octave"] = np.where(rng.random(n) < 0.6, 0, k1["octave"])        # only level-0 points take part (:421-423)
    if crowded:
        k1["x"] = (300 + rng.integers(0, 40, n)).astype(np.float32); k1["y"] = (200 + rng.integers(0, 30, n)).astype(np.float32)
    d1 = synth.descriptors(n, seed=seed + 1)
    perm = rng.permutation(n)
    k2 = k1[perm].copy()
    k2["x"] += rng.normal(3, 4, n).astype(np.float32); k2["y"] += rng.normal(-2, 4, n).astype(np.float32)
    k2["angle"] = np.mod(k2["angle"] + 8 + rng.normal(0, 2, n), 360).astype(np.float32)
    wild = rng.random(n) < 0.1
    k2["angle"][wild] = rng.uniform(0, 360, int(wild.sum()))
    d2 = d1[perm].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, 20)):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    if crowded:      # near-duplicate descriptors: many points want the same feature, evictions happen
        d1[: n // 2] = d1[0]; d2[: n // 2] = d1[0]
        for i in range(n // 2):
            d1[i, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8)); d2[i, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8))
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    return k1, d1, k2, d2, prev


def test_oracle_search_for_initialization_properties(oracle, synth):
    k1, d1, k2, d2, prev = _init_pair(oracle, synth, 1500, 61)
    b = np.array([0, 640, 0, 480], np.float32)
    ptr, idx = oracle.assign_grid(k2, b)
    m12, xy, n = oracle.search_for_initialization(k1, d1, k2, d2, ptr, idx, b, prev, 100, 0.9, True)
    good = m12 >= 0
    assert n == good.sum() > 300
    assert len(set(m12[good])) == good.sum()                                  # a frame-2 feature has one owner
    assert np.all(k1["octave"][good] == 0) and np.all(k2["octave"][m12[good]] == 0)
    assert np.array_equal(xy[good], np.stack([k2["x"], k2["y"]], 1)[m12[good]]) and np.array_equal(xy[~good], prev[~good])


@pytest.mark.gpu
@pytest.mark.parametrize("n,crowded,window,ratio", [(1800, False, 100, 0.9), (2500, False, 30, 0.7), (400, True, 100, 0.9), (1, False, 100, 0.9)])
def test_gpu_search_for_initialization(orb, oracle, synth, n, crowded, window, ratio):
    k1, d1, k2, d2, prev = _init_pair(oracle, synth, n, 70 + n, crowded)
    K = np.array([500, 500, 320, 240], np.float32); D = np.zeros(4, np.float32)
    f2 = orb.FrameFeatures(k2, K, D, 640, 480)
    for check in (True, False):
        rm, rxy, rn = oracle.search_for_initialization(k1, d1, f2.keys_un, d2, f2.cell_ptr, f2.cell_idx, f2.bounds, prev, window, ratio, check)
        gm, gxy, gn = orb.search_for_initialization(k1, d1, f2, d2, prev, window, ratio, check)
        assert gn == rn and np.array_equal(gm, rm) and np.array_equal(gxy, rxy)
    if n >= 400:
        assert rn > 20
