"""ctypes bindings to the CPU oracle (oracle/liborb_oracle.so) and to the reference's own
ORBextractor compiled over cvshim (oracle/_ref/liborbref*.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class KeyPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28 == C.sizeof(KeyPoint)


class FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("ptr", C.c_void_p), ("idx", C.c_void_p)]


TRI_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("u_right", "<f4"),
                      ("has_mp", "<i4")])


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liborb_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_ic_angle.restype = C.c_float
        L.orc_ic_angle.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extractor_tables.argtypes = [C.c_void_p] * 7
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]
        L.orc_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_get_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_get_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int,
                                           C.c_size_t]
        L.orc_copy_make_border_reflect101.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                                      C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_gaussian_blur7_sigma2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_fast9_16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_score.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.orc_orb_descriptor.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orc_distribute_octtree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int]
        L.orc_knn2_full.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
        L.orc_search_by_bow_kf_f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_search_by_bow_kf_kf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int,
                                              C.c_void_p]
        L.orc_search_for_triangulation.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def copy_make_border(src, b):
    src = np.ascontiguousarray(src)
    h, w = src.shape
    dst = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib().orc_copy_make_border_reflect101(_p(src), w, h, src.strides[0], _p(dst), dst.strides[0], b, b, b, b)
    return dst


def gaussian_blur(src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_gaussian_blur7_sigma2(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast(img, th, nms=True):
    """img may be a non-contiguous 2-D view (row stride honoured)."""
    assert img.strides[1] == 1
    h, w = img.shape
    cap = w * h
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_16(C.c_void_p(img.ctypes.data), w, h, img.strides[0], th, int(nms), _p(out), cap)
    return out[:n]


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return lib().orc_descriptor_distance(_p(a), _p(b))


# ----------------------------------------------------------------------------- extractor
class OracleExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, trig_mode=0, fma_mode=0):
        self.L = lib()
        self.nlevels = nlevels
        self.h = self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th, trig_mode, fma_mode)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        nf = np.zeros(n, np.int32); um = np.zeros(16, np.int32)
        self.L.orc_extractor_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2), _p(nf), _p(um))
        return dict(sf=sf, isf=isf, s2=s2, is2=is2, nfeat=nf, umax=um)

    def extract(self, img, cap=20000):
        img = np.ascontiguousarray(img)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = self.L.orc_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap,
                                C.byref(n))
        if rc != 0:
            raise RuntimeError("orc_extract rc=%d" % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, l):
        w = C.c_int(); h = C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def pyramid(self, l, with_border=False):
        w, h = self.level_size(l)
        b = 19 if with_border else 0
        out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
        self.L.orc_get_pyramid(self.h, l, int(with_border), _p(out), out.strides[0])
        return out

    def blurred(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        rc = self.L.orc_get_blurred(self.h, l, _p(out), out.strides[0])
        return out if rc == 0 else None

    def candidates(self, l, cap=200000):
        x = np.zeros(cap, np.int16); y = np.zeros(cap, np.int16); s = np.zeros(cap, np.uint8)
        n = self.L.orc_get_candidates(self.h, l, _p(x), _p(y), _p(s), cap)
        return x[:n].copy(), y[:n].copy(), s[:n].copy()

    def level_keypoints(self, l, cap=20000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.orc_get_level_keypoints(self.h, l, _p(out), cap)
        return out[:n].copy()


def distribute_octtree(x, y, score, min_x, max_x, min_y, max_y, nfeat):
    x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    score = np.ascontiguousarray(score, np.uint8)
    cap = len(x) + 8
    out = np.zeros(cap, np.int32)
    n = lib().orc_distribute_octtree(_p(x), _p(y), _p(score), len(x), min_x, max_x, min_y, max_y, nfeat, _p(out), cap)
    if n < 0:
        raise RuntimeError("orc_distribute_octtree rc=%d" % n)
    return out[:n].copy()


# ----------------------------------------------------------------------------- compiled reference
_ref = {}


def ref_available(fma=False):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "liborbref_fma.so" if fma else "liborbref.so"))


def ref_lib(fma=False):
    if fma not in _ref:
        path = os.path.join(ORACLE_DIR, "_ref", "liborbref_fma.so" if fma else "liborbref.so")
        L = C.CDLL(path)
        L.orbref_create.restype = C.c_void_p
        L.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_destroy.argtypes = [C.c_void_p]
        L.orbref_tables.argtypes = [C.c_void_p] * 5
        L.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p]
        L.orbref_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbref_get_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _ref[fma] = L
    return _ref[fma]


class RefExtractor:
    """The reference's ORB_SLAM2::ORBextractor (compiled from /root/reference) behind ctypes."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, fma=False):
        self.L = ref_lib(fma)
        self.nlevels = nlevels
        self.h = self.L.orbref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbref_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        self.L.orbref_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2))
        return dict(sf=sf, isf=isf, s2=s2, is2=is2)

    def extract(self, img, cap=20000):
        img = np.ascontiguousarray(img)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        self.L.orbref_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap,
                              C.byref(n))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def pyramid(self, l, with_border=False):
        w = C.c_int(); h = C.c_int()
        self.L.orbref_level_size(self.h, l, C.byref(w), C.byref(h))
        b = 19 if with_border else 0
        out = np.empty((h.value + 2 * b, w.value + 2 * b), np.uint8)
        self.L.orbref_get_pyramid(self.h, l, int(with_border), _p(out), out.strides[0])
        return out


# ----------------------------------------------------------------------------- matcher loops
def _fv_struct(fv):
    node_ids, ptr, idx = (np.ascontiguousarray(a, np.int32) for a in fv)
    s = FeatVec(len(node_ids), node_ids.ctypes.data, ptr.ctypes.data, idx.ctypes.data)
    s._keep = (node_ids, ptr, idx)
    return s


def search_by_bow_kf_f(dk, ak, vk, fvk, df, af, fvf, ratio, check_ori):
    dk = np.ascontiguousarray(dk, np.uint8); df = np.ascontiguousarray(df, np.uint8)
    ak = np.ascontiguousarray(ak, np.float32); af = np.ascontiguousarray(af, np.float32)
    vk = np.ascontiguousarray(vk, np.uint8)
    a, b = _fv_struct(fvk), _fv_struct(fvf)
    out = np.zeros(len(df), np.int32)
    n = lib().orc_search_by_bow_kf_f(_p(dk), _p(ak), _p(vk), len(dk), C.addressof(a), _p(df), _p(af), len(df), C.addressof(b),
                                     ratio, int(check_ori), _p(out))
    return n, out


def search_by_bow_kf_kf(d1, a1, v1, fv1, d2, a2, v2, fv2, ratio, check_ori):
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
    v1 = np.ascontiguousarray(v1, np.uint8); v2 = np.ascontiguousarray(v2, np.uint8)
    a, b = _fv_struct(fv1), _fv_struct(fv2)
    out = np.zeros(len(d1), np.int32)
    n = lib().orc_search_by_bow_kf_kf(_p(d1), _p(a1), _p(v1), len(d1), C.addressof(a), _p(d2), _p(a2), _p(v2), len(d2),
                                      C.addressof(b), ratio, int(check_ori), _p(out))
    return n, out


def search_for_triangulation(d1, f1, fv1, d2, f2, fv2, F12, ex, ey, sf2, sig2, only_stereo, check_ori):
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    f1 = np.ascontiguousarray(f1, TRI_DTYPE); f2 = np.ascontiguousarray(f2, TRI_DTYPE)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf2 = np.ascontiguousarray(sf2, np.float32); sig2 = np.ascontiguousarray(sig2, np.float32)
    a, b = _fv_struct(fv1), _fv_struct(fv2)
    cap = max(len(d1), 1)
    out = np.zeros((cap, 2), np.int32)
    n = lib().orc_search_for_triangulation(_p(d1), _p(f1), len(d1), C.addressof(a), _p(d2), _p(f2), len(d2), C.addressof(b),
                                           _p(F), float(ex), float(ey), _p(sf2), _p(sig2), int(only_stereo), int(check_ori),
                                           _p(out), cap)
    return n, out[:n].copy()


def stereo_matches(kl, dl, kr, dr, ext_l, ext_r, mbf, mb):
    """ext_l / ext_r: OracleExtractor objects that produced the key points (their pyramids are read)."""
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    L = ext_l.nlevels
    t = ext_l.tables()
    pl = [ext_l.pyramid(l, True) for l in range(L)]
    pr = [ext_r.pyramid(l, True) for l in range(L)]
    PL = (C.c_void_p * L)(*[p.ctypes.data for p in pl]); PR = (C.c_void_p * L)(*[p.ctypes.data for p in pr])
    lw = np.array([p.shape[1] - 38 for p in pl], np.int32); lh = np.array([p.shape[0] - 38 for p in pl], np.int32)
    st = np.array([p.strides[0] for p in pl], np.uint64)
    ur = np.zeros(len(kl), np.float32); dep = np.zeros(len(kl), np.float32)
    n = lib().orc_stereo_matches(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), L, _p(t["sf"]), _p(t["isf"]),
                                 C.cast(PL, C.c_void_p), C.cast(PR, C.c_void_p), _p(lw), _p(lh), _p(st), float(mbf), float(mb),
                                 _p(ur), _p(dep))
    return ur, dep, n


def distinctive_descriptors(desc, ptr):
    d = np.ascontiguousarray(desc, np.uint8); p = np.ascontiguousarray(ptr, np.int32)
    best = np.zeros(len(p) - 1, np.int32)
    L = lib()
    L.orc_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_distinctive_descriptors(_p(d), _p(p), len(p) - 1, _p(best))
    return best


# ----------------------------------------------------------------------------- frame side (oracle/frame_oracle.cc)
MPV_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                      ("in_view", "<i4"), ("obs_positive", "<i4")])
GRID_CELLS = 64 * 48


def undistort_points(xy, K, dist):
    xy = np.ascontiguousarray(xy, np.float32); K = np.ascontiguousarray(K, np.float32); d = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(xy)
    lib().orc_undistort_points(_p(xy), C.c_int(len(xy)), _p(K), _p(d), C.c_int(len(d)), _p(out))
    return out


def undistort_keypoints(kps, K, dist):
    kps = np.ascontiguousarray(kps, KP_DTYPE); K = np.ascontiguousarray(K, np.float32); d = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(kps)
    lib().orc_undistort_keypoints(_p(kps), C.c_int(len(kps)), _p(K), _p(d), C.c_int(len(d)), _p(out))
    return out


def image_bounds(cols, rows, K, dist):
    K = np.ascontiguousarray(K, np.float32); d = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    lib().orc_image_bounds(C.c_int(cols), C.c_int(rows), _p(K), _p(d), C.c_int(len(d)), _p(b))
    return b


def assign_grid(kps_un, bounds):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); b = np.ascontiguousarray(bounds, np.float32)
    ptr = np.zeros(GRID_CELLS + 1, np.int32); idx = np.zeros(max(len(kps_un), 1), np.int32)
    lib().orc_assign_grid(_p(kps_un), C.c_int(len(kps_un)), _p(b), _p(ptr), _p(idx))
    return ptr, idx


def features_in_area(kps_un, ptr, idx, bounds, x, y, r, min_level, max_level):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); b = np.ascontiguousarray(bounds, np.float32)
    out = np.zeros(max(len(kps_un), 1), np.int32)
    L = lib()
    L.orc_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int,
                                       C.c_int, C.c_void_p, C.c_int]
    n = L.orc_features_in_area(_p(kps_un), _p(ptr), _p(idx), _p(b), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n].copy()


def search_by_projection_frame(kps_un, desc_f, u_right, occupied, ptr, idx, bounds, scale_factors, mps, desc_mp, th, nnratio,
                               th_high=100):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); df = np.ascontiguousarray(desc_f, np.uint8)
    ur = np.ascontiguousarray(u_right, np.float32); occ = np.ascontiguousarray(occupied, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    mps = np.ascontiguousarray(mps, MPV_DTYPE); dm = np.ascontiguousarray(desc_mp, np.uint8)
    fp = np.zeros(max(len(kps_un), 1), np.int32); pf = np.zeros(max(len(mps), 1), np.int32)
    L = lib()
    L.orc_search_by_projection_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                                 C.c_int, C.c_void_p, C.c_void_p]
    n = L.orc_search_by_projection_frame(_p(kps_un), _p(df), _p(ur), _p(occ), len(kps_un), _p(ptr), _p(idx), _p(b), _p(sf), _p(mps),
                                         _p(dm), len(mps), th, nnratio, th_high, _p(fp), _p(pf))
    return fp[:len(kps_un)], pf[:len(mps)], n


PROJ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                       ("obs_positive", "<i4")])


def search_by_projection_last_frame(kps_un, desc_f, u_right, occupied, ptr, idx, bounds, scale_factors, pts, desc_pts, th, direction,
                                    check_orientation, th_high=100):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); df = np.ascontiguousarray(desc_f, np.uint8)
    ur = np.ascontiguousarray(u_right, np.float32); occ = np.ascontiguousarray(occupied, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(pts, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    fp = np.zeros(max(len(kps_un), 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32)
    L = lib()
    L.orc_search_by_projection_last_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int,
                                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = L.orc_search_by_projection_last_frame(_p(kps_un), _p(df), _p(ur), _p(occ), len(kps_un), _p(ptr), _p(idx), _p(b), _p(sf), _p(pts),
                                              _p(dp), len(pts), th, direction, int(check_orientation), th_high, _p(fp), _p(pf))
    return fp[:len(kps_un)], pf[:len(pts)], n


def search_by_projection_keyframe(kps_un, desc_f, occupied, ptr, idx, bounds, scale_factors, pts, desc_pts, th, orb_dist,
                                  check_orientation):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); df = np.ascontiguousarray(desc_f, np.uint8)
    occ = np.ascontiguousarray(occupied, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(pts, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    fp = np.zeros(max(len(kps_un), 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32)
    L = lib()
    L.orc_search_by_projection_keyframe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_void_p]
    n = L.orc_search_by_projection_keyframe(_p(kps_un), _p(df), _p(occ), len(kps_un), _p(ptr), _p(idx), _p(b), _p(sf), _p(pts), _p(dp),
                                            len(pts), th, orb_dist, int(check_orientation), _p(fp), _p(pf))
    return fp[:len(kps_un)], pf[:len(pts)], n


def search_by_projection_sim3(kps_un, desc_f, occupied, ptr, idx, bounds, scale_factors, pts, desc_pts, th, th_low=50, grid_origin=None):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); df = np.ascontiguousarray(desc_f, np.uint8)
    occ = np.ascontiguousarray(occupied, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(pts, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    fp = np.zeros(max(len(kps_un), 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32)
    L = lib()
    L.orc_search_by_projection_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    go = None if grid_origin is None else np.ascontiguousarray(grid_origin, np.float32)
    n = L.orc_search_by_projection_sim3(_p(kps_un), _p(df), _p(occ), len(kps_un), _p(ptr), _p(idx), _p(b), _p(sf), _p(pts), _p(dp),
                                        len(pts), th, th_low, _p(fp), _p(pf), None if go is None else _p(go))
    return fp[:len(kps_un)], pf[:len(pts)], n


def window_best_match(kps_un, desc_f, u_right, ptr, idx, bounds, scale_factors, inv_level_sigma2, pts, desc_pts, th, grid_origin=None):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE); df = np.ascontiguousarray(desc_f, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(pts, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    sg = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
    bi = np.zeros(max(len(pts), 1), np.int32); bd = np.zeros(max(len(pts), 1), np.int32)
    L = lib()
    L.orc_window_best_match.restype = None
    L.orc_window_best_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    go = None if grid_origin is None else np.ascontiguousarray(grid_origin, np.float32)
    L.orc_window_best_match(_p(kps_un), _p(df), None if ur is None else _p(ur), len(kps_un), _p(ptr), _p(idx), _p(b), _p(sf),
                            None if sg is None else _p(sg), _p(pts), _p(dp), len(pts), th, _p(bi), _p(bd), None if go is None else _p(go))
    return bi[:len(pts)], bd[:len(pts)]


# ----------------------------------------------------------------------------- BoW transform (oracle/bow_oracle.cc)
def bow_transform(desc, voc, levelsup=4):
    """voc = dict(child_ptr, child_idx, node_desc, word_id, weight, L) -> per-feature (word, node, weight)."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    w = np.zeros(len(d), np.int32); nd = np.zeros(len(d), np.int32); wt = np.zeros(len(d), np.float64)
    L = lib()
    L.orc_bow_transform.restype = None
    L.orc_bow_transform.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_bow_transform(_p(d), len(d), _p(voc["child_ptr"]), _p(voc["child_idx"]), _p(voc["node_desc"]), _p(voc["word_id"]),
                        _p(voc["weight"]), voc["L"], levelsup, _p(w), _p(nd), _p(wt))
    return w, nd, wt


def bow_vectors(word, node, weight, normalize=True):
    n = len(word)
    word = np.ascontiguousarray(word, np.int32); node = np.ascontiguousarray(node, np.int32); weight = np.ascontiguousarray(weight, np.float64)
    bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
    fn = np.zeros(max(n, 1), np.int32); fp = np.zeros(n + 1, np.int32); fi = np.zeros(max(n, 1), np.int32)
    nn = C.c_int(0)
    L = lib()
    L.orc_bow_vectors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    nw = L.orc_bow_vectors(_p(word), _p(node), _p(weight), n, int(normalize), _p(bw), _p(bv), _p(fn), _p(fp), _p(fi), C.byref(nn))
    return (bw[:nw], bv[:nw]), (fn[:nn.value], fp[:nn.value + 1], fi[:fp[nn.value]])


def search_for_initialization(k1, d1, k2, d2, ptr, idx, bounds, prev_xy, window, nnratio, check_orientation, th_low=50):
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); xy = np.ascontiguousarray(prev_xy, np.float32).reshape(-1, 2).copy()
    m12 = np.zeros(max(len(k1), 1), np.int32)
    L = lib()
    L.orc_search_for_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
    n = L.orc_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(ptr), _p(idx), _p(b), _p(xy), window, nnratio,
                                        int(check_orientation), th_low, _p(m12))
    return m12[:len(k1)], xy, n
