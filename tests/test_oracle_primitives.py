"""Pin the oracle's OpenCV primitive models: (1) against the committed cv2 4.13.0 golden vectors,
(2) live against cv2 when it is importable (same wheel on the GPU box).  CPU only."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir + "/primitives.npz")


def test_resize_golden(oracle, gold):
    assert np.array_equal(oracle.resize_linear(gold["img"], 167, 133), gold["resize_167x133"])
    assert np.array_equal(oracle.resize_linear(gold["noise"], 67, 53), gold["resize_noise_67x53"])


def test_border_golden(oracle, gold):
    assert np.array_equal(oracle.copy_make_border(gold["img"], 19), gold["border19"])


def test_blur_golden(oracle, gold):
    assert np.array_equal(oracle.gaussian_blur(gold["img"]), gold["blur"])
    assert np.array_equal(oracle.gaussian_blur(gold["noise"]), gold["blur_noise"])


@pytest.mark.parametrize("name", ["img", "noise"])
@pytest.mark.parametrize("th", [20, 7])
def test_fast_golden(oracle, gold, name, th):
    k = oracle.fast(gold[name], th)
    got = np.stack([k["x"], k["y"], k["response"]], 1).astype(np.float32).reshape(-1, 3)
    assert np.array_equal(got, gold["fast_%s_%d" % (name, th)])
    assert (k["size"] == 7).all() and (k["angle"] == -1).all() and (k["class_id"] == -1).all()


def test_fast_atan2_golden(oracle, gold):
    got = np.array([oracle.fast_atan2(y, x) for y, x in gold["atan2_yx"]], np.float32)
    assert np.array_equal(got, gold["atan2"])


def test_live_cv2_chain(oracle, synth):
    cv2 = pytest.importorskip("cv2")
    if cv2.__version__ != "4.13.0":
        pytest.skip("oracle is pinned to cv2 4.13.0")
    e = oracle.OracleExtractor(1000)
    isf = e.tables()["isf"]
    for (w, h) in [(640, 480), (1241, 376), (752, 480)]:
        prev = synth.frame(5, w, h)
        for l in range(1, 8):
            dw = int(np.rint(np.float32(w) * isf[l])); dh = int(np.rint(np.float32(h) * isf[l]))
            a = oracle.resize_linear(prev, dw, dh)
            assert np.array_equal(a, cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)), (w, h, l)
            assert np.array_equal(oracle.gaussian_blur(a),
                                  cv2.GaussianBlur(a, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
            assert np.array_equal(oracle.copy_make_border(a, 19),
                                  cv2.copyMakeBorder(a, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
            prev = a


def test_live_cv2_fast_cells(oracle, synth):
    cv2 = pytest.importorskip("cv2")
    if cv2.__version__ != "4.13.0":
        pytest.skip("oracle is pinned to cv2 4.13.0")
    img = synth.frame(2)
    rng = np.random.default_rng(0)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for _ in range(40):
            x0 = int(rng.integers(0, 600)); y0 = int(rng.integers(0, 440))
            cw = int(rng.integers(7, 40)); ch = int(rng.integers(7, 40))
            view = img[y0:y0 + ch, x0:x0 + cw]
            ref = [(k.pt[0], k.pt[1], k.response) for k in det.detect(np.ascontiguousarray(view))]
            got = oracle.fast(view, th)
            assert ref == [(float(q["x"]), float(q["y"]), float(q["response"])) for q in got]


def test_fast_threshold_filter_equivalence(oracle, synth):
    """FAST(cell, 20) == {k in FAST(cell, 7) : response >= 20}; the CUDA kernel relies on it."""
    img = synth.frame(4)
    rng = np.random.default_rng(1)
    for _ in range(60):
        x0 = int(rng.integers(0, 600)); y0 = int(rng.integers(0, 440))
        view = img[y0:y0 + 36, x0:x0 + 36]
        hi = oracle.fast(view, 20); lo = oracle.fast(view, 7)
        flt = lo[lo["response"] >= 20]
        assert hi.tobytes() == flt.tobytes()


def test_pattern_and_tables(oracle):
    import hashlib, os, re, struct
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "cooperative-orb-slam_b200", "csrc", "orb_pattern.inc")).read()
    vals = [int(v) for v in re.findall(r"-?\d+", re.sub(r"//.*", "", txt))]
    assert len(vals) == 1024
    assert hashlib.sha256(struct.pack("<1024i", *vals)).hexdigest().startswith("7e645581387b8278")
    t = oracle.OracleExtractor(1000).tables()
    assert list(t["nfeat"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(oracle.OracleExtractor(2000).tables()["nfeat"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(oracle.OracleExtractor(1200).tables()["nfeat"]) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert t["sf"][1] == np.float32(1.2) and abs(float(t["sf"][7]) - 3.5831816196) < 1e-6
