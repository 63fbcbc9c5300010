"""GPU parity tests of the matching path against the CPU oracle (bit-exact: integer work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orb():
    import orbcuda
    if orbcuda.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return orbcuda


def _oracle_knn2_full(oracle, q, m, base=0):
    import ctypes as C
    nq = len(q)
    i1, d1, i2, d2 = (np.zeros(nq, np.int32) for _ in range(4))
    oracle.lib().orc_knn2_full(q.ctypes.data, nq, m.ctypes.data, len(m), base, i1.ctypes.data, d1.ctypes.data,
                               i2.ctypes.data, d2.ctypes.data, 8)
    return i1, d1, i2, d2


def test_hamming_kat(orb, oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8); b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        ref = int(np.unpackbits(a[i] ^ b[i]).sum())
        assert orb.ORBmatcher.DescriptorDistance(a[i], b[i]) == ref == oracle.descriptor_distance(a[i], b[i])
    z = np.zeros(32, np.uint8); f = np.full(32, 255, np.uint8)
    assert orb.ORBmatcher.DescriptorDistance(z, f) == 256 and orb.ORBmatcher.DescriptorDistance(f, f) == 0


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("nq,nm", [(1, 1), (7, 300), (257, 5000), (2000, 60000), (100, 0)])
def test_knn2_matches_oracle(orb, oracle, synth, nq, nm, variant):
    m = synth.descriptors(max(nm, 1), seed=5)[:nm]
    if nm > nq:
        q, m, _ = synth.query_set(m, nq=nq, seed=6)
    else:
        q = synth.descriptors(nq, seed=7)
    # duplicates in the map exercise the tie rule (first index wins, second distance counts duplicates)
    if nm > 10:
        m[nm // 2] = m[3]; m[nm - 1] = m[3]
        q[0] = m[3]
    bi, bd, sd, si = orb.ORBmatcher().knn2(q, m, index_base=1000, variant=variant)
    i1, d1, i2, d2 = _oracle_knn2_full(oracle, q, np.ascontiguousarray(m).reshape(-1, 32), base=1000)
    assert np.array_equal(bd, d1) and np.array_equal(sd, d2)
    assert np.array_equal(bi, i1) and np.array_equal(si, i2)


def test_sharded_merge_is_exact(orb, oracle, synth):
    """Map split into G contiguous shards, per-shard records merged == single search, G in {2,4,8}."""
    m = synth.descriptors(40000, seed=8)
    q, m, _ = synth.query_set(m, nq=500, seed=9)
    mt = orb.ORBmatcher()
    full = np.stack(mt.knn2(q, m), 1)[:, [1, 0, 2, 3]]   # d1,i1,d2,i2
    import ctypes as C
    for G in (2, 4, 8):
        parts = []
        bounds = np.linspace(0, len(m), G + 1).astype(int)
        for g in range(G):
            bi, bd, sd, si = mt.knn2(q, m[bounds[g]:bounds[g + 1]], index_base=int(bounds[g]))
            parts.append(np.stack([bd, bi, sd, si], 1))
        parts = np.ascontiguousarray(np.stack(parts), np.int32)
        out = np.zeros((len(q), 4), np.int32)
        assert orb.lib().orbm_merge_top2_host(parts.ctypes.data, G, len(q), out.ctypes.data) == 0
        assert np.array_equal(out, full.astype(np.int32)), G


# ---------------------------------------------------------------------------------- candidate loops
import matchdata


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_bow_kf_f(orb, oracle, seed, check_ori):
    d1, d2, a1, a2, src, dst, rng = matchdata.two_views(1000, 1100, seed)
    valid = (rng.random(len(d1)) < 0.8).astype(np.uint8)
    fv1, fv2 = matchdata.featvec(d1), matchdata.featvec(d2)
    for ratio in (0.7, 0.75, 0.9):
        n_ref, m_ref = oracle.search_by_bow_kf_f(d1, a1, valid, fv1, d2, a2, fv2, ratio, check_ori)
        n_gpu, m_gpu = orb.ORBmatcher(ratio, check_ori).SearchByBoW(d1, a1, valid, fv1, d2, a2, fv2)
        assert n_ref == n_gpu and np.array_equal(m_ref, m_gpu), (ratio, n_ref, n_gpu)
        assert n_ref > 100


@pytest.mark.parametrize("seed", [3, 4])
def test_search_by_bow_kf_kf(orb, oracle, seed):
    d1, d2, a1, a2, src, dst, rng = matchdata.two_views(900, 1000, seed)
    v1 = (rng.random(len(d1)) < 0.7).astype(np.uint8); v2 = (rng.random(len(d2)) < 0.7).astype(np.uint8)
    fv1, fv2 = matchdata.featvec(d1, 5), matchdata.featvec(d2, 5)
    for check_ori in (True, False):
        n_ref, m_ref = oracle.search_by_bow_kf_kf(d1, a1, v1, fv1, d2, a2, v2, fv2, 0.75, check_ori)
        n_gpu, m_gpu = orb.ORBmatcher(0.75, check_ori).SearchByBoW_KF(d1, a1, v1, fv1, d2, a2, v2, fv2)
        assert n_ref == n_gpu and np.array_equal(m_ref, m_gpu)
        assert n_ref > 50


@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("only_stereo", [False, True])
def test_search_for_triangulation(orb, oracle, seed, only_stereo):
    d1, d2, a1, a2, src, dst, rng = matchdata.two_views(1000, 1000, seed)
    f1, f2 = matchdata.tri_features(len(d1), len(d2), src, dst, rng)
    f1["angle"] = a1; f2["angle"] = a2
    fv1, fv2 = matchdata.featvec(d1), matchdata.featvec(d2)
    t = oracle.OracleExtractor(1000).tables()
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)      # sideways translation: epipolar lines y2 = y1
    for (ex, ey) in ((1e6, 240.0), (300.0, 240.0)):                       # far epipole / epipole inside the image
        n_ref, p_ref = oracle.search_for_triangulation(d1, f1, fv1, d2, f2, fv2, F12, ex, ey, t["sf"], t["s2"], only_stereo, False)
        n_gpu, p_gpu = orb.ORBmatcher(0.6, False).SearchForTriangulation(d1, f1, fv1, d2, f2, fv2, F12, (ex, ey), t["sf"],
                                                                         t["s2"], only_stereo)
        assert n_ref == n_gpu and np.array_equal(p_ref, p_gpu), (n_ref, n_gpu)
    assert only_stereo or n_ref > 30
    n_ref, p_ref = oracle.search_for_triangulation(d1, f1, fv1, d2, f2, fv2, F12, 1e6, 0, t["sf"], t["s2"], only_stereo, True)
    n_gpu, p_gpu = orb.ORBmatcher(0.6, True).SearchForTriangulation(d1, f1, fv1, d2, f2, fv2, F12, (1e6, 0), t["sf"], t["s2"],
                                                                    only_stereo)
    assert n_ref == n_gpu and np.array_equal(p_ref, p_gpu)


@pytest.mark.parametrize("seed", [0, 1])
def test_stereo_matches(orb, oracle, synth, seed):
    left, right = synth.stereo_pair(seed, 752, 480)
    el = orb.ORBextractor(1200, 1.2, 8, 20, 7); er = orb.ORBextractor(1200, 1.2, 8, 20, 7)
    kl, dl = el(left); kr, dr = er(right)
    ol = oracle.OracleExtractor(1200, trig_mode=1); orr = oracle.OracleExtractor(1200, trig_mode=1)
    okl, odl = ol.extract(left); okr, odr = orr.extract(right)
    assert kl.tobytes() == okl.tobytes() and kr.tobytes() == okr.tobytes()
    mbf, fx = 40.0, 458.0
    ur, dep, n = orb.compute_stereo_matches(el, er, kl, dl, kr, dr, mbf, mbf / fx)
    our, odep, on = oracle.stereo_matches(okl, odl, okr, odr, ol, orr, mbf, mbf / fx)
    assert n == on and n > 100, (n, on)
    assert np.array_equal(ur, our) and np.array_equal(dep, odep)
    # empty sides are legal
    ur0, dep0, n0 = orb.compute_stereo_matches(el, er, kl, dl, kr[:0], dr[:0], mbf, mbf / fx)
    assert n0 == 0 and (ur0 == -1).all()


@pytest.mark.parametrize("w,h,nfeat,max_disp_px", [(752, 480, 7000, 400.0), (800, 1300, 5000, 30.0)])
def test_stereo_matches_many_keypoints(orb, oracle, synth, w, h, nfeat, max_disp_px):
    """More right key points than one staging chunk of the stereo kernel holds (2048), row bands that hold more candidates than its
    per-warp list (64: a wide disparity range keeps them all), and an image taller than 1024 rows (row bins of more than 8 rows)."""
    left, right = synth.stereo_pair(3, w, h)
    el = orb.ORBextractor(nfeat, 1.2, 8, 20, 7); er = orb.ORBextractor(nfeat, 1.2, 8, 20, 7)
    kl, dl = el(left); kr, dr = er(right)
    ol = oracle.OracleExtractor(nfeat, trig_mode=1); orr = oracle.OracleExtractor(nfeat, trig_mode=1)
    okl, odl = ol.extract(left); okr, odr = orr.extract(right)
    assert kl.tobytes() == okl.tobytes() and kr.tobytes() == okr.tobytes() and len(kr) > 2048
    mbf = 40.0; mb = mbf / max_disp_px      # maxD = mbf / mb = max_disp_px
    ur, dep, n = orb.compute_stereo_matches(el, er, kl, dl, kr, dr, mbf, mb)
    our, odep, on = oracle.stereo_matches(okl, odl, okr, odr, ol, orr, mbf, mb)
    assert n == on and n > 100, (n, on)
    assert np.array_equal(ur, our) and np.array_equal(dep, odep)


def test_knn2_with_device_ratio_test(orb, oracle, synth):
    """2-NN + the reference's acceptance test (ORBmatcher.cc:228-230 / :598-600) on the device in one call chain, and the test
    alone on device-resident records: both equal the host formula on the oracle's records."""
    import ctypes as C
    import torch
    m = synth.descriptors(30000, seed=21)
    q, m, _ = synth.query_set(m, nq=700, seed=22)
    L = oracle.lib()
    i1, d1, d2 = (np.zeros(len(q), np.int32) for _ in range(3))
    L.orc_knn2(q.ctypes.data, len(q), m.ctypes.data, len(m), 0, i1.ctypes.data, d1.ctypes.data, d2.ctypes.data, 4)
    dq = torch.from_numpy(q).cuda(); dm = torch.from_numpy(m).cuda()
    rec = torch.zeros((len(q), 4), dtype=torch.int32, device="cuda"); match = torch.zeros(len(q), dtype=torch.int32, device="cuda")
    lib = orb.lib()
    for ratio, th, strict in ((0.6, 50, 0), (0.75, 50, 1), (0.9, 100, 0)):
        want = np.where(((d1 < th) if strict else (d1 <= th)) & (d1.astype(np.float32) < np.float32(ratio) * d2.astype(np.float32)), i1, -1)
        for variant in (0, 5):
            rc = lib.orbm_knn2_ratio_device(C.c_void_p(dq.data_ptr()), len(q), C.c_void_p(dm.data_ptr()), len(m), 0, ratio, th, strict,
                                            C.c_void_p(rec.data_ptr()), C.c_void_p(match.data_ptr()), variant, None)
            assert rc == 0, lib.orb_last_error()
            torch.cuda.synchronize()
            assert np.array_equal(match.cpu().numpy(), want) and np.array_equal(rec.cpu().numpy()[:, 0], d1)
        match.zero_()
        assert lib.orbm_ratio_test_device(C.c_void_p(rec.data_ptr()), len(q), ratio, th, strict, C.c_void_p(match.data_ptr()), None) == 0
        torch.cuda.synchronize()
        assert np.array_equal(match.cpu().numpy(), want)
    assert (want >= 0).sum() > 100
    # the host-array entry point runs on the calling thread's workspace: repeated calls, growing sizes
    for n in (100, 5000, 400000):
        mm = synth.descriptors(n, seed=n)
        bi, bd, sd, si = orb.ORBmatcher().knn2(q[:64], mm)
        j1, e1, e2 = (np.zeros(64, np.int32) for _ in range(3))
        L.orc_knn2(q.ctypes.data, 64, mm.ctypes.data, n, 0, j1.ctypes.data, e1.ctypes.data, e2.ctypes.data, 4)
        assert np.array_equal(bi, j1) and np.array_equal(bd, e1) and np.array_equal(sd, e2)


@pytest.mark.parametrize("B,nfeat", [(5, 1200), (2, 7000)])
def test_stereo_matches_batch_device(orb, oracle, synth, B, nfeat):
    """Batch form (two extractors per stereo frame, Frame.cc:80-83, then ComputeStereoMatches with its outlier cut, all
    on the device): every pair must equal the oracle's per-pair result.  7000 features: more left key points than the median
    cut keeps in shared memory (4096) and more right key points than one staging chunk of the search kernel."""
    import torch
    W, H = 752, 480
    pairs = [synth.stereo_pair(s, W, H) for s in range(B)]
    if B > 3:
        pairs[3] = (pairs[3][0], np.full((H, W), 90, np.uint8))     # a right image without key points: no matches for this pair
    left = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda(); right = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    el = orb.ORBextractor(nfeat, 1.2, 8, 20, 7); er = orb.ORBextractor(nfeat, 1.2, 8, 20, 7)
    cap = el.max_keypoints(W, H)
    mk = lambda: (torch.zeros((B, cap, 7), dtype=torch.int32, device="cuda"), torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda"),
                  torch.zeros((B,), dtype=torch.int32, device="cuda"))
    kl, dl, cl = mk(); kr, dr, cr = mk()
    el.extract_batch_device(left.data_ptr(), B, W, H, W, W * H, kl.data_ptr(), dl.data_ptr(), cap, cl.data_ptr())
    er.extract_batch_device(right.data_ptr(), B, W, H, W, W * H, kr.data_ptr(), dr.data_ptr(), cap, cr.data_ptr())
    scratch = torch.empty(orb.stereo_scratch_bytes(B, cap), dtype=torch.uint8, device="cuda")
    ur = torch.zeros((B, cap), dtype=torch.float32, device="cuda"); dep = torch.zeros_like(ur)
    nm = torch.zeros((B,), dtype=torch.int32, device="cuda")
    mbf, fx = 40.0, 458.0
    st = torch.cuda.Stream()
    for _ in range(2):      # twice: the call is stream-ordered and must be repeatable
        orb.compute_stereo_matches_batch_device(el, er, kl.data_ptr(), dl.data_ptr(), cl.data_ptr(), kr.data_ptr(), dr.data_ptr(),
                                                cr.data_ptr(), B, cap, mbf, mbf / fx, scratch.data_ptr(), ur.data_ptr(), dep.data_ptr(),
                                                nm.data_ptr(), st.cuda_stream)
    st.synchronize()
    ur = ur.cpu().numpy(); dep = dep.cpu().numpy(); nm = nm.cpu().numpy(); cl = cl.cpu().numpy()
    total = 0
    for f in range(B):
        ol = oracle.OracleExtractor(nfeat, trig_mode=1); orr = oracle.OracleExtractor(nfeat, trig_mode=1)
        okl, odl = ol.extract(pairs[f][0]); okr, odr = orr.extract(pairs[f][1])
        assert cl[f] == len(okl)
        our, odep, on = oracle.stereo_matches(okl, odl, okr, odr, ol, orr, mbf, mbf / fx)
        assert nm[f] == on, (f, nm[f], on)
        assert np.array_equal(ur[f, :len(okl)], our) and np.array_equal(dep[f, :len(okl)], odep)
        assert (ur[f, len(okl):] == -1).all()
        total += on
    assert (B <= 3 or nm[3] == 0) and total > 400
    assert nfeat < 5000 or cl.max() > 4096


def test_distinctive_descriptors(orb, oracle):
    """MapPoint::ComputeDistinctiveDescriptors batched: least-median descriptor per map point."""
    rng = np.random.default_rng(42)
    sizes = [0, 1, 2, 3, 5, 8, 31, 32, 33, 70] + list(rng.integers(1, 40, 500))
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.zeros((ptr[-1], 32), np.uint8)
    for p, n in enumerate(sizes):
        bits = np.unpackbits(np.repeat(base[p][None], n, 0), axis=1) if n else np.zeros((0, 256), np.uint8)
        flip = rng.random(bits.shape) < 0.1
        desc[ptr[p]:ptr[p + 1]] = np.packbits(bits ^ flip, axis=1) if n else desc[ptr[p]:ptr[p + 1]]
    if sizes[4] >= 3:
        desc[ptr[4] + 1] = desc[ptr[4]]          # duplicates -> tied medians, the first index must win
    ref = oracle.distinctive_descriptors(desc, ptr)
    got = orb.compute_distinctive_descriptors(desc, ptr)
    assert np.array_equal(ref, got)
