"""ctypes driver of the matcher / frame parity harness (oracle/match_harness.cc): the REFERENCE's own Frame, KeyFrame,
MapPoint, Map (compiled verbatim from /root/reference into oracle/_ref/) driven from POD arrays, with one of three
ORBmatcher / ORBextractor implementations behind the same exports:
  "ref"        the reference's own ORBmatcher.cc + ORBextractor.cc                    (ground truth)
  "shim_cpu"   the product's drop-in ORBmatcher class over the CPU restatement (orc_*) (pins the restatement)
  "shim_cuda"  the product's drop-in ORBmatcher + ORBextractor classes over liborbcuda (the product, needs a GPU)
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from oracle_lib import KP_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = {"ref": "libmatchref.so", "shim_cpu": "libmatchshim_cpu.so", "shim_cuda": "libmatchshim_cuda.so"}
_loaded = {}


def available(flavour):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", LIBS[flavour]))


def lib(flavour):
    if flavour not in _loaded:
        L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", LIBS[flavour]))     # RTLD_LOCAL: the flavours define the same symbols
        L.mh_world_create.restype = C.c_void_p
        L.mh_world_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float,
                                      C.c_int, C.c_int, C.c_int]
        _loaded[flavour] = L
    return _loaded[flavour]


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


class World:
    """One calibration + extractor parameters; frames, key frames and map points are addressed by index."""

    def __init__(self, flavour, K, dist, bf, th_depth, cols, rows, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib(flavour)
        self.flavour = flavour
        K = _f32(K); dist = _f32(dist)
        self.h = C.c_void_p(self.L.mh_world_create(_p(K), _p(dist), len(dist), float(bf), float(th_depth), int(cols), int(rows),
                                                   int(nfeatures), float(scale), int(nlevels), int(ini_th), int(min_th)))
        self.cols, self.rows = cols, rows

    def close(self):
        if self.h:
            self.L.mh_world_destroy.argtypes = [C.c_void_p]
            self.L.mh_world_destroy(self.h)
            self.h = None

    # ---- frames
    def frame_from_image(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        self.L.mh_frame_from_image.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        return self.L.mh_frame_from_image(self.h, _p(img), img.strides[0])

    def frame_from_stereo(self, left, right, mb):
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        self.L.mh_frame_from_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
        return self.L.mh_frame_from_stereo(self.h, _p(left), _p(right), left.strides[0], float(mb))

    def frame_from_features(self, kps, desc, u_right=None, depth=None):
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        ur = None if u_right is None else _f32(u_right); dp = None if depth is None else _f32(depth)
        self.L.mh_frame_from_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        return self.L.mh_frame_from_features(self.h, _p(kps), _p(desc), len(kps), _p(ur), _p(dp))

    def frame_n(self, f):
        self.L.mh_frame_n.argtypes = [C.c_void_p, C.c_int]
        return self.L.mh_frame_n(self.h, f)

    def frame_get(self, f):
        n = self.frame_n(f)
        self.L.mh_frame_n_right.argtypes = [C.c_void_p, C.c_int]
        nr = self.L.mh_frame_n_right(self.h, f)
        out = dict(kps=np.zeros(n, KP_DTYPE), kps_un=np.zeros(n, KP_DTYPE), desc=np.zeros((n, 32), np.uint8), u_right=np.zeros(n, np.float32),
                   depth=np.zeros(n, np.float32), bounds=np.zeros(4, np.float32), cell_ptr=np.zeros(64 * 48 + 1, np.int32),
                   cell_idx=np.zeros(max(n, 1), np.int32), kps_right=np.zeros(nr, KP_DTYPE), desc_right=np.zeros((nr, 32), np.uint8))
        self.L.mh_frame_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
        self.L.mh_frame_get(self.h, f, _p(out["kps"]), _p(out["kps_un"]), _p(out["desc"]), _p(out["u_right"]), _p(out["depth"]),
                            _p(out["bounds"]), _p(out["cell_ptr"]), _p(out["cell_idx"]), _p(out["kps_right"]), _p(out["desc_right"]))
        out["cell_idx"] = out["cell_idx"][:out["cell_ptr"][-1]]
        return out

    def frame_set_pose(self, f, Tcw):
        T = _f32(Tcw).reshape(16)
        self.L.mh_frame_set_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.mh_frame_set_pose(self.h, f, _p(T))

    def frame_set_featvec(self, f, fv):
        ids, ptr, idx = (_i32(a) for a in fv)
        self.L.mh_frame_set_featvec.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.mh_frame_set_featvec(self.h, f, len(ids), _p(ids), _p(ptr), _p(idx))

    def frame_set_mappoint(self, f, idx, mp, outlier=False):
        self.L.mh_frame_set_mappoint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        self.L.mh_frame_set_mappoint(self.h, f, int(idx), int(mp), int(outlier))

    def frame_mappoints(self, f):
        out = np.zeros(self.frame_n(f), np.int32)
        self.L.mh_frame_get_mappoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.mh_frame_get_mappoints(self.h, f, _p(out))
        return out

    def frame_features_in_area(self, f, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(self.frame_n(f) + 1, np.int32)
        self.L.mh_frame_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        n = self.L.mh_frame_features_in_area(self.h, f, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    # ---- key frames / map points
    def keyframe(self, f):
        self.L.mh_keyframe.argtypes = [C.c_void_p, C.c_int]
        return self.L.mh_keyframe(self.h, f)

    def keyframe_n(self, k):
        self.L.mh_keyframe_n.argtypes = [C.c_void_p, C.c_int]
        return self.L.mh_keyframe_n(self.h, k)

    def keyframe_mappoints(self, k):
        out = np.zeros(self.keyframe_n(k), np.int32)
        self.L.mh_keyframe_get_mappoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.mh_keyframe_get_mappoints(self.h, k, _p(out))
        return out

    def keyframe_features_in_area(self, k, x, y, r):
        out = np.zeros(self.keyframe_n(k) + 1, np.int32)
        self.L.mh_keyframe_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
        n = self.L.mh_keyframe_features_in_area(self.h, k, x, y, r, _p(out), len(out))
        return out[:n].copy()

    def keyframe_bounds(self, k):
        out = np.zeros(4, np.int32)
        self.L.mh_keyframe_bounds.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.mh_keyframe_bounds(self.h, k, _p(out))
        return out

    def mappoint(self, pos, ref_kf):
        p = _f32(pos)
        self.L.mh_mappoint.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        return self.L.mh_mappoint(self.h, _p(p), int(ref_kf))

    def observe(self, mp, kf, idx):
        self.L.mh_observe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.L.mh_observe(self.h, int(mp), int(kf), int(idx))

    def mappoint_update(self, mp):
        self.L.mh_mappoint_update.argtypes = [C.c_void_p, C.c_int]
        self.L.mh_mappoint_update(self.h, int(mp))

    def mappoint_get(self, mp):
        desc = np.zeros(32, np.uint8); normal = np.zeros(3, np.float32); dist = np.zeros(2, np.float32); state = np.zeros(3, np.int32)
        self.L.mh_mappoint_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.mh_mappoint_get(self.h, int(mp), _p(desc), _p(normal), _p(dist), _p(state))
        return dict(desc=desc, normal=normal, dist=dist, observations=int(state[0]), bad=bool(state[1]), replaced=int(state[2]))

    def mappoint_set_bad(self, mp):
        self.L.mh_mappoint_set_bad.argtypes = [C.c_void_p, C.c_int]
        self.L.mh_mappoint_set_bad(self.h, int(mp))

    # ---- ORBmatcher
    def search_by_bow_kf_f(self, k, f, nnratio, check_ori):
        out = np.zeros(self.frame_n(f), np.int32)
        self.L.mh_search_by_bow_kf_f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        n = self.L.mh_search_by_bow_kf_f(self.h, k, f, nnratio, int(check_ori), _p(out))
        return n, out

    def search_by_bow_kf_kf(self, k1, k2, nnratio, check_ori):
        out = np.zeros(self.keyframe_n(k1), np.int32)
        self.L.mh_search_by_bow_kf_kf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        n = self.L.mh_search_by_bow_kf_kf(self.h, k1, k2, nnratio, int(check_ori), _p(out))
        return n, out

    def search_for_triangulation(self, k1, k2, F12, only_stereo, nnratio=0.6, check_ori=False):
        F = _f32(F12).reshape(9)
        cap = self.keyframe_n(k1) + 1
        pairs = np.zeros((cap, 2), np.int32)
        self.L.mh_search_for_triangulation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
        n = self.L.mh_search_for_triangulation(self.h, k1, k2, _p(F), int(only_stereo), nnratio, int(check_ori), _p(pairs), cap)
        return n, pairs[:n].copy()

    def search_by_projection_local(self, f, mps, th, nnratio):
        mps = _i32(mps)
        out = np.zeros(self.frame_n(f), np.int32); iv = np.zeros(len(mps), np.int32)
        self.L.mh_search_by_projection_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        n = self.L.mh_search_by_projection_local(self.h, f, _p(mps), len(mps), th, nnratio, _p(out), _p(iv))
        return n, out, iv

    def search_by_projection_last(self, cur, last, th, mono, nnratio=0.9, check_ori=True):
        out = np.zeros(self.frame_n(cur), np.int32)
        self.L.mh_search_by_projection_last.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
        n = self.L.mh_search_by_projection_last(self.h, cur, last, th, int(mono), nnratio, int(check_ori), _p(out))
        return n, out

    def search_by_projection_kf(self, cur, k, already, th, orb_dist, nnratio=0.9, check_ori=True):
        already = _i32(already)
        out = np.zeros(self.frame_n(cur), np.int32)
        self.L.mh_search_by_projection_kf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int,
                                                      C.c_void_p]
        n = self.L.mh_search_by_projection_kf(self.h, cur, k, _p(already), len(already), th, int(orb_dist), nnratio, int(check_ori), _p(out))
        return n, out

    def search_by_projection_sim3(self, k, Scw, points, matched, th):
        S = _f32(Scw).reshape(16); points = _i32(points); m = _i32(matched).copy()
        self.L.mh_search_by_projection_sim3.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        n = self.L.mh_search_by_projection_sim3(self.h, k, _p(S), _p(points), len(points), _p(m), int(th))
        return n, m

    def search_for_initialization(self, f1, f2, prev_xy, window, nnratio=0.9, check_ori=True):
        xy = _f32(prev_xy).reshape(-1, 2).copy()
        m12 = np.zeros(self.frame_n(f1), np.int32)
        self.L.mh_search_for_initialization.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        n = self.L.mh_search_for_initialization(self.h, f1, f2, _p(xy), int(window), nnratio, int(check_ori), _p(m12))
        return n, m12, xy

    def fuse(self, k, mps, th):
        mps = _i32(mps)
        self.L.mh_fuse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float]
        return self.L.mh_fuse(self.h, k, _p(mps), len(mps), th)

    def fuse_sim3(self, k, Scw, mps, th):
        S = _f32(Scw).reshape(16); mps = _i32(mps)
        rep = np.zeros(len(mps), np.int32)
        self.L.mh_fuse_sim3.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        n = self.L.mh_fuse_sim3(self.h, k, _p(S), _p(mps), len(mps), th, _p(rep))
        return n, rep

    def search_by_sim3(self, k1, k2, matches12, s12, R12, t12, th):
        m = _i32(matches12).copy(); R = _f32(R12).reshape(9); t = _f32(t12).reshape(3)
        self.L.mh_search_by_sim3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_float]
        n = self.L.mh_search_by_sim3(self.h, k1, k2, _p(m), s12, _p(R), _p(t), th)
        return n, m

    def descriptor_distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        self.L.mh_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        return self.L.mh_descriptor_distance(_p(a), _p(b))

    def extract_latency_ms(self, img, iters=100, mirror=True):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        self.L.mh_extract_latency_ms.restype = C.c_double
        self.L.mh_extract_latency_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        ms = self.L.mh_extract_latency_ms(self.h, _p(img), img.strides[0], int(iters), int(mirror), C.byref(n))
        return ms, n.value

    def frame_stereo_accel(self, f):
        n = self.frame_n(f)
        ur = np.zeros(n, np.float32); dep = np.zeros(n, np.float32)
        self.L.mh_frame_stereo_accel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        m = self.L.mh_frame_stereo_accel(self.h, f, _p(ur), _p(dep))
        return m, ur, dep


# ----------------------------------------------------------------------------------------------------------------------
# A synthetic two-view scene: 3-D points seen from two poses, key points consistent with the calibration (distorted so that
# the reference's own undistortion maps them back), descriptors = per-point base pattern with a few flipped bits.
# ----------------------------------------------------------------------------------------------------------------------
def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def pose(R, t):
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R; T[:3, 3] = t
    return T


def distort(u, v, K, D):
    """pinhole pixel -> distorted pixel (forward Brown model), so that undistortPoints maps it back"""
    fx, fy, cx, cy = K
    k1, k2, p1, p2 = D[:4]
    k3 = D[4] if len(D) > 4 else 0.0
    x = (u - cx) / fx; y = (v - cy) / fy
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return xd * fx + cx, yd * fy + cy


class Scene:
    def __init__(self, seed, distortion=True, n_points=1400, n_clutter=300, cols=752, rows=480, stereo_frac=0.0):
        rng = np.random.Generator(np.random.PCG64(seed))
        self.rng = rng
        self.cols, self.rows = cols, rows
        self.K = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
        self.D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05] if distortion else [0, 0, 0, 0], np.float32)
        self.bf = 47.9; self.th_depth = 35.0
        self.sf = np.array([1.2 ** i for i in range(8)], np.float64)
        # world points in front of the first camera
        z = rng.uniform(2.0, 9.0, n_points)
        x = rng.uniform(-0.8, 0.8, n_points) * z; y = rng.uniform(-0.5, 0.5, n_points) * z
        self.P = np.stack([x, y, z], 1).astype(np.float32)
        self.base = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
        self.level = rng.choice(8, n_points, p=[.3, .2, .15, .12, .09, .07, .04, .03])
        self.T = [pose(np.eye(3), np.zeros(3)),
                  pose(rot(0.01, -0.03, 0.02), np.array([-0.25, 0.02, 0.05])),
                  pose(rot(-0.02, 0.05, -0.01), np.array([0.3, -0.04, -0.1]))]
        self.n_clutter = n_clutter
        self.stereo_frac = stereo_frac

    def observe(self, view, noise_px=0.4, flip=0.045, keep=0.85, angle_offset=0.0):
        """-> (kps, desc, point_of_feature [-1: clutter], u_right, depth)"""
        rng = self.rng
        T = self.T[view]
        Pc = self.P @ T[:3, :3].T + T[:3, 3]
        fx, fy, cx, cy = [float(v) for v in self.K]
        u = fx * Pc[:, 0] / Pc[:, 2] + cx; v = fy * Pc[:, 1] / Pc[:, 2] + cy
        vis = (Pc[:, 2] > 0.1) & (u > 20) & (u < self.cols - 20) & (v > 20) & (v < self.rows - 20) & (rng.random(len(u)) < keep)
        ids = np.nonzero(vis)[0]
        rng.shuffle(ids)
        ud, vd = distort(u[ids] + rng.normal(0, noise_px, len(ids)), v[ids] + rng.normal(0, noise_px, len(ids)), self.K.astype(np.float64),
                         self.D.astype(np.float64))
        n = len(ids) + self.n_clutter
        kps = np.zeros(n, KP_DTYPE)
        kps["x"][:len(ids)] = ud; kps["y"][:len(ids)] = vd
        kps["x"][len(ids):] = rng.uniform(20, self.cols - 20, self.n_clutter); kps["y"][len(ids):] = rng.uniform(20, self.rows - 20, self.n_clutter)
        lv = np.concatenate([np.clip(self.level[ids] + rng.integers(-1, 2, len(ids)), 0, 7), rng.integers(0, 8, self.n_clutter)])
        kps["octave"] = lv; kps["size"] = (31 * self.sf[lv]).astype(np.int32); kps["class_id"] = -1
        kps["response"] = rng.integers(7, 120, n)
        base_angle = (np.arange(len(self.P)) * 37.0) % 360.0
        kps["angle"][:len(ids)] = (base_angle[ids] + angle_offset + rng.normal(0, 2.5, len(ids))) % 360.0
        kps["angle"][len(ids):] = rng.uniform(0, 360, self.n_clutter)
        bits = np.unpackbits(self.base[ids], axis=1)
        desc = np.concatenate([np.packbits(bits ^ (rng.random(bits.shape) < flip), axis=1),
                               rng.integers(0, 256, (self.n_clutter, 32), dtype=np.uint8)])
        owner = np.concatenate([ids, np.full(self.n_clutter, -1)]).astype(np.int64)
        perm = rng.permutation(n)
        kps, desc, owner = kps[perm], desc[perm], owner[perm]
        ur = np.full(n, -1, np.float32); dep = np.full(n, -1, np.float32)
        if self.stereo_frac > 0:
            zc = np.where(owner >= 0, Pc[np.maximum(owner, 0), 2], 5.0)
            st = rng.random(n) < self.stereo_frac
            ur[st] = (kps["x"][st] - self.bf / zc[st]).astype(np.float32); dep[st] = zc[st].astype(np.float32)
        return kps, np.ascontiguousarray(desc), owner, ur, dep
