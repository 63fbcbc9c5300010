"""orbcuda -- Python host mirror of the reference's ORBextractor / ORBmatcher interface on top of the
C ABI of liborbcuda.so (include/orbcuda.h).

The product is the CUDA library; this module only binds it with ctypes (plain pointers and sizes,
no torch types).  There is no CPU fallback: if the library is missing, or no CUDA device is
usable, construction fails loudly.

Reference interface mirrored (R21 = ORB_SLAM2.1 of 530300865/Cooperative-ORB-SLAM):
  ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)   R21/include/ORBextractor.h:45-111
  ORBextractor.__call__(image, mask) -> keypoints, descriptors        R21/src/ORBextractor.cc:1043
  GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
  GetInverseScaleSigmaSquares, mvImagePyramid                          R21/include/ORBextractor.h:63-85
  ORBmatcher(nnratio, checkOri), DescriptorDistance, SearchByBoW, SearchForTriangulation,
  TH_LOW / TH_HIGH / HISTO_LENGTH                                      R21/include/ORBmatcher.h:37-102
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbcuda.so")

ORB_OK, ORB_ERR_ARG, ORB_ERR_CUDA, ORB_ERR_CAPACITY = 0, 1, 2, 3

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
TRI_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("u_right", "<f4"),
                      ("has_mp", "<i4")])


class OrbCudaError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class _FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("ptr", C.c_void_p), ("idx", C.c_void_p)]


_lib = None

# name -> (restype, argtypes); every symbol include/orbcuda.h declares
_VP, _I, _F, _SZ, _I64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
ABI = {
    "orb_last_error": (C.c_char_p, []),
    "orb_device_count": (_I, [_VP]),
    "orb_host_alloc": (_I, [_VP, _SZ]),
    "orb_host_free": (_I, [_VP]),
    "orbx_create": (_I, [_VP, _I, _I, _I, _I, _VP]),
    "orbx_destroy": (_I, [_VP]),
    "orbx_tables": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "orbx_max_keypoints": (_I, [_VP, _I, _I, _VP]),
    "orbx_extract": (_I, [_VP, _VP, _I, _I, _SZ, _VP, _VP, _I, _VP]),
    "orbx_extract_batch": (_I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _VP, _VP, _I, _VP]),
    "orbx_extract_batch_async": (_I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _VP, _VP, _I, _VP]),
    "orbx_wait": (_I, [_VP]),
    "orbx_extract_batch_device": (_I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _VP, _VP, _I, _VP]),
    "orbx_level_size": (_I, [_VP, _I, _VP, _VP]),
    "orbx_download_level": (_I, [_VP, _I, _I, _I, _VP, _SZ]),
    "orbx_download_pyramid": (_I, [_VP, _I, _I, _VP, _VP]),
    "orbx_download_blurred": (_I, [_VP, _I, _I, _VP, _SZ]),
    "orbx_download_candidates": (_I, [_VP, _I, _I, _VP, _VP, _VP, _I, _VP]),
    "orbx_download_scores": (_I, [_VP, _I, _I, _VP, _SZ]),
    "orbx_set_profiling": (_I, [_VP, _I]),
    "orbx_stage_times": (_I, [_VP, _VP]),
    "orbx_launch_count": (_I, [_VP, _VP]),
    "orbx_stream": (_I, [_VP, _VP]),
    "orbx_distribute_octtree": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP, _I, _VP, _I]),
    "orb_hamming256": (_I, [_VP, _VP]),
    "orbm_knn2": (_I, [_VP, _I, _VP, _I64, _I64, _VP, _VP, _VP, _VP, _I, _I]),
    "orbm_knn2_device": (_I, [_VP, _I, _VP, _I64, _I64, _VP, _I, _VP]),
    "orbm_merge_top2_device": (_I, [_VP, _I, _I, _VP, _VP]),
    "orbm_merge_top2_host": (_I, [_VP, _I, _I, _VP]),
    "orbm_ratio_test_host": (_I, [_VP, _I, _F, _I, _I, _VP]),
    "orbm_knn2_ratio_device": (_I, [_VP, _I, _VP, _I64, _I64, _F, _I, _I, _VP, _VP, _I, _VP]),
    "orbm_ratio_test_device": (_I, [_VP, _I, _F, _I, _I, _VP, _VP]),
    "orbm_search_by_bow_kf_f": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _F, _I, _VP, _VP, _I]),
    "orbm_search_by_bow_kf_kf": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _F, _I, _VP, _VP, _I]),
    "orbm_search_for_triangulation": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _F, _F, _VP, _VP, _I, _I, _VP,
                                          _I, _VP, _I]),
    "orbm_stereo_matches": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _F, _VP, _VP, _VP]),
    "orbm_stereo_matches_batch_device": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _F, _F, _VP, _VP, _VP, _VP, _VP]),
    "orbm_stereo_scratch_bytes": (_SZ, [_I, _I]),
    "orbm_distinctive_descriptors": (_I, [_VP, _VP, _I, _VP, _I]),
    "orbf_undistort_keypoints": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I]),
    "orbf_image_bounds": (_I, [_I, _I, _VP, _VP, _I, _VP, _I]),
    "orbf_assign_grid": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I]),
    "orbf_build_frame": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I]),
    "orbf_build_frames_device": (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP]),
    "orbf_features_in_area": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _I]),
    "orbm_search_by_projection_frame": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _F, _I,
                                            _VP, _VP, _VP, _I]),
    "orbm_search_by_projection_last_frame": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _I, _I, _I,
                                                 _VP, _VP, _VP, _I]),
    "orbm_search_by_projection_keyframe": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _I, _I,
                                               _VP, _VP, _VP, _I]),
    "orbm_search_by_projection_sim3": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _I, _VP, _VP, _VP, _VP, _I]),
    "orbm_search_for_initialization": (_I, [_VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _F, _I, _I, _VP, _VP, _I]),
    "orbm_window_best_match": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _F, _VP, _VP, _VP, _I]),
    "orbv_create": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "orbv_destroy": (_I, [_VP]),
    "orbv_transform": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP]),
    "orbv_bow_vectors": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "orbm_peer_create": (_I, [_I, _I, _I, _I, _VP, _VP]),
    "orbm_peer_connect": (_I, [_VP, _VP]),
    "orbm_knn2_exchange_device": (_I, [_VP, _VP, _I, _VP, C.c_int64, C.c_int64, _VP, _I, _VP]),
    "orbm_peer_error": (_I, [_VP, _VP]),
    "orbm_peer_destroy": (_I, [_VP]),
    "orbw_quantize_lcm_host": (_I, [_VP, _I]),
    "orbw_quantize_lcm_device": (_I, [_VP, _VP, _I, _I, _VP]),
    "orbw_message_bytes": (_SZ, [_I, _I, _I]),
    "orbw_pack_keyframes_device": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP]),
    "orbw_unpack_keyframes_device": (_I, [_VP, _I, _SZ, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "orbw_exchange_messages_device": (_I, [_VP, _VP, _SZ, _VP, _SZ, _VP]),
}


def lib():
    """Load liborbcuda.so (in-tree).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbCudaError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(make -C cooperative-orb-slam_b200/csrc); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        missing = [name for name in ABI if not hasattr(L, name)]
        if missing:
            raise OrbCudaError("%s does not export %s (stale build?)" % (LIB_PATH, ", ".join(missing)))
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc, what):
    if rc != ORB_OK:
        msg = lib().orb_last_error()
        raise OrbCudaError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else ""))


def device_count():
    n = C.c_int(0)
    lib().orb_device_count(C.byref(n))
    return n.value


class PinnedArray:
    """numpy view of page-locked host memory from orb_host_alloc (truly asynchronous DMA)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self._ptr = C.c_void_p()
        _check(lib().orb_host_alloc(C.byref(self._ptr), max(self.nbytes, 1)), "orb_host_alloc")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().orb_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ORBextractor:
    """ORB_SLAM2::ORBextractor on a B200 (R21/include/ORBextractor.h:45-111)."""

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, max_width=0, max_height=0,
                 max_batch=0):
        self._L = lib()
        self.nlevels = int(nlevels)
        self._scaleFactor = float(np.float32(scaleFactor))
        self._h = C.c_void_p()
        prm = _Params(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        _check(self._L.orbx_create(C.byref(prm), max_width, max_height, max_batch, device, C.byref(self._h)),
               "orbx_create")
        n = self.nlevels
        self._sf, self._isf, self._s2, self._is2 = (np.zeros(n, np.float32) for _ in range(4))
        self._nfeat = np.zeros(n, np.int32)
        _check(self._L.orbx_tables(self._h, _p(self._sf), _p(self._isf), _p(self._s2), _p(self._is2), _p(self._nfeat)),
               "orbx_tables")
        self._last = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (R21/include/ORBextractor.h:63-83), returned by value like the reference
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self._scaleFactor

    def GetScaleFactors(self):
        return self._sf.copy()

    def GetInverseScaleFactors(self):
        return self._isf.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    @property
    def mnFeaturesPerLevel(self):
        return self._nfeat.copy()

    def max_keypoints(self, width, height):
        cap = C.c_int(0)
        _check(self._L.orbx_max_keypoints(self._h, width, height, C.byref(cap)), "orbx_max_keypoints")
        return cap.value

    # ---- operator() (R21/src/ORBextractor.cc:1043-1105)
    def __call__(self, image, mask=None):
        """Returns (keypoints: structured array with cv::KeyPoint layout, descriptors: N x 32 uint8 or None).
        `mask` is ignored, as in the reference."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), None   # :1046-1047 / :1064-1065
        assert image.dtype == np.uint8 and image.ndim == 2, "image.type() == CV_8UC1"   # :1050
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(self._L.orbx_extract(self._h, C.c_void_p(image.ctypes.data), w, h, image.strides[0], _p(kps), _p(desc),
                                    cap, C.byref(n)), "orbx_extract")
        self._last = (1, w, h)
        if n.value == 0:
            return kps[:0], None
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, cap=None, out=None):
        """images: [B,H,W] uint8 (numpy, C-contiguous rows).  Returns (kps [B,cap], desc [B,cap,32], counts [B])."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.strides[2] == 1
        b, h, w = images.shape
        cap = cap or self.max_keypoints(w, h)
        if out is None:
            out = (np.zeros((b, cap), KP_DTYPE), np.zeros((b, cap, 32), np.uint8), np.zeros(b, np.int32))
        kps, desc, cnt = out
        _check(self._L.orbx_extract_batch(self._h, C.c_void_p(images.ctypes.data), b, w, h, images.strides[1],
                                          images.strides[0], _p(kps), _p(desc), cap, _p(cnt)), "orbx_extract_batch")
        self._last = (b, w, h)
        return kps, desc, cnt

    def extract_batch_async(self, images, kps, desc, cnt):
        """Queue one batch; `kps` [b][cap] KP_DTYPE, `desc` [b][cap][32] uint8 and `cnt` [b] int32 are written by wait().  The
        arrays must be C-contiguous (rows of `images` may be strided); references are kept until wait()."""
        b, h, w = images.shape
        cap = kps.shape[1]
        if images.dtype != np.uint8 or images.strides[2] != 1:
            raise ValueError("images must be uint8 with contiguous rows")
        if not (kps.flags.c_contiguous and desc.flags.c_contiguous and cnt.flags.c_contiguous):
            raise ValueError("output arrays must be C-contiguous")
        if kps.dtype.itemsize != 28 or desc.dtype != np.uint8 or cnt.dtype != np.int32 or kps.shape[0] != b or desc.shape != (b, cap, 32) or cnt.shape != (b,):
            raise ValueError("output arrays: kps [b][cap] of 28-byte records, desc [b][cap][32] uint8, cnt [b] int32")
        self._pending_refs = (images, kps, desc, cnt)
        _check(self._L.orbx_extract_batch_async(self._h, C.c_void_p(images.ctypes.data), b, w, h, images.strides[1],
                                                images.strides[0], _p(kps), _p(desc), cap, _p(cnt)),
               "orbx_extract_batch_async")
        self._last = (b, w, h)

    def extract_batch_device(self, d_images_ptr, b, w, h, row_stride, frame_stride, d_kps_ptr, d_desc_ptr, cap,
                             d_counts_ptr):
        """Device-resident form: all pointers are raw CUDA device addresses (ints)."""
        _check(self._L.orbx_extract_batch_device(self._h, C.c_void_p(d_images_ptr), b, w, h, row_stride, frame_stride,
                                                 C.c_void_p(d_kps_ptr), C.c_void_p(d_desc_ptr), cap,
                                                 C.c_void_p(d_counts_ptr)), "orbx_extract_batch_device")
        self._last = (b, w, h)

    def wait(self):
        try:
            _check(self._L.orbx_wait(self._h), "orbx_wait")
        finally:
            self._pending_refs = None

    # ---- mvImagePyramid and the other stage views
    def level_size(self, level):
        w = C.c_int(); h = C.c_int()
        _check(self._L.orbx_level_size(self._h, level, C.byref(w), C.byref(h)), "orbx_level_size")
        return w.value, h.value

    def pyramid(self, level, with_border=False, frame=0):
        w, h = self.level_size(level)
        b = 19 if with_border else 0
        out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
        _check(self._L.orbx_download_level(self._h, frame, level, int(with_border), _p(out), out.strides[0]),
               "orbx_download_level")
        return out

    @property
    def mvImagePyramid(self):
        """std::vector<cv::Mat> mvImagePyramid (R21/include/ORBextractor.h:85) of the last frame, downloaded lazily."""
        return [self.pyramid(l) for l in range(self.nlevels)]

    def blurred(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(self._L.orbx_download_blurred(self._h, frame, level, _p(out), out.strides[0]), "orbx_download_blurred")
        return out

    def scores(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(self._L.orbx_download_scores(self._h, frame, level, _p(out), out.strides[0]), "orbx_download_scores")
        return out

    def candidates(self, level, frame=0, cap=1 << 20):
        x = np.zeros(cap, np.int16); y = np.zeros(cap, np.int16); s = np.zeros(cap, np.uint8)
        n = C.c_int(0)
        _check(self._L.orbx_download_candidates(self._h, frame, level, _p(x), _p(y), _p(s), cap, C.byref(n)),
               "orbx_download_candidates")
        return x[:n.value].copy(), y[:n.value].copy(), s[:n.value].copy()

    def set_profiling(self, on=True):
        _check(self._L.orbx_set_profiling(self._h, int(on)), "orbx_set_profiling")

    def stage_times(self):
        ms = np.zeros(9, np.float32)
        _check(self._L.orbx_stage_times(self._h, _p(ms)), "orbx_stage_times")
        return dict(zip(["upload", "pyramid", "fast_score", "blur", "cell_nms", "quadtree", "describe", "download",
                         "total"], [float(v) for v in ms]))

    def launch_count(self):
        n = C.c_int64(0)
        _check(self._L.orbx_launch_count(self._h, C.byref(n)), "orbx_launch_count")
        return n.value

    def stream(self):
        s = C.c_void_p()
        _check(self._L.orbx_stream(self._h, C.byref(s)), "orbx_stream")
        return s.value or 0


def distribute_octtree(x, y, score, min_x, max_x, min_y, max_y, n_features, device=0):
    """ORBextractor::DistributeOctTree (R21/src/ORBextractor.cc:539-763) on the GPU; returns kept input indices."""
    x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    score = np.ascontiguousarray(score, np.uint8)
    cap = len(x) + 16
    out = np.zeros(cap, np.int32)
    n = C.c_int(0)
    _check(lib().orbx_distribute_octtree(_p(x), _p(y), _p(score), len(x), min_x, max_x, min_y, max_y, n_features,
                                         _p(out), cap, C.byref(n), device), "orbx_distribute_octtree")
    return out[:n.value].copy()


def _featvec(fv):
    """fv: (node_ids, ptr, idx) int32 arrays -> ctypes struct (keeps references alive)."""
    node_ids, ptr, idx = (np.ascontiguousarray(a, np.int32) for a in fv)
    s = _FeatVec(len(node_ids), node_ids.ctypes.data, ptr.ctypes.data, idx.ctypes.data)
    s._keep = (node_ids, ptr, idx)
    return s


class ORBmatcher:
    """ORB_SLAM2::ORBmatcher hot loops (R21/include/ORBmatcher.h:37-102) on POD arrays."""

    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # R21/src/ORBmatcher.cc:37-39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._L = lib()
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return lib().orb_hamming256(_p(a), _p(b))

    def knn2(self, queries, map_desc, index_base=0, variant=5):
        """Brute-force 2-NN: returns (best_idx, best_dist, second_dist, second_idx).
        variant 0 = LOP3+POPC kernel, 1/2 = mma.sync integer tensor-core AND-popc contraction, 3 = tcgen05 (UMMA + TMEM)
        contraction, 4 = the same with the query operand in TMEM, 5 = the same on CTA pairs (cta_group::2; default); all variants
        return identical results."""
        q = np.ascontiguousarray(queries, np.uint8); m = np.ascontiguousarray(map_desc, np.uint8)
        nq = q.shape[0]
        bi, bd, sd, si = (np.zeros(nq, np.int32) for _ in range(4))
        _check(self._L.orbm_knn2(_p(q), nq, _p(m), m.shape[0], index_base, _p(bi), _p(bd), _p(sd), _p(si), variant,
                                 self.device), "orbm_knn2")
        return bi, bd, sd, si

    def match_ratio(self, queries, map_desc, th=None, strict=False):
        bi, bd, sd, si = self.knn2(queries, map_desc)
        rec = np.stack([bd, bi, sd, si], 1).astype(np.int32)
        out = np.zeros(len(bi), np.int32)
        _check(self._L.orbm_ratio_test_host(_p(rec), len(bi), self.mfNNratio, self.TH_LOW if th is None else th,
                                            int(strict), _p(out)), "orbm_ratio_test_host")
        return out

    def SearchByBoW(self, desc_kf, angle_kf, kf_valid, fv_kf, desc_f, angle_f, fv_f):
        """SearchByBoW(KeyFrame*, Frame&, ...) R21/src/ORBmatcher.cc:159-288 -> (nmatches, match_f)."""
        dk = np.ascontiguousarray(desc_kf, np.uint8); df = np.ascontiguousarray(desc_f, np.uint8)
        ak = np.ascontiguousarray(angle_kf, np.float32); af = np.ascontiguousarray(angle_f, np.float32)
        vk = np.ascontiguousarray(kf_valid, np.uint8)
        a, b = _featvec(fv_kf), _featvec(fv_f)
        out = np.zeros(len(df), np.int32); n = C.c_int(0)
        _check(self._L.orbm_search_by_bow_kf_f(_p(dk), _p(ak), _p(vk), len(dk), C.byref(a), _p(df), _p(af), len(df),
                                               C.byref(b), self.mfNNratio, int(self.mbCheckOrientation), _p(out),
                                               C.byref(n), self.device), "orbm_search_by_bow_kf_f")
        return n.value, out

    def SearchByBoW_KF(self, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2):
        """SearchByBoW(KeyFrame*, KeyFrame*, ...) R21/src/ORBmatcher.cc:522-655 -> (nmatches, match12)."""
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
        v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
        a, b = _featvec(fv1), _featvec(fv2)
        out = np.zeros(len(d1), np.int32); n = C.c_int(0)
        _check(self._L.orbm_search_by_bow_kf_kf(_p(d1), _p(a1), _p(v1), len(d1), C.byref(a), _p(d2), _p(a2), _p(v2),
                                                len(d2), C.byref(b), self.mfNNratio, int(self.mbCheckOrientation),
                                                _p(out), C.byref(n), self.device), "orbm_search_by_bow_kf_kf")
        return n.value, out

    def SearchForTriangulation(self, desc1, feat1, fv1, desc2, feat2, fv2, F12, epipole, scale_factors2,
                               level_sigma2_2, bOnlyStereo=False):
        """R21/src/ORBmatcher.cc:657-823 -> (nmatches, pairs[n,2])."""
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
        f1 = np.ascontiguousarray(feat1, TRI_DTYPE); f2 = np.ascontiguousarray(feat2, TRI_DTYPE)
        F = np.ascontiguousarray(F12, np.float32).reshape(9)
        sf = np.ascontiguousarray(scale_factors2, np.float32); s2 = np.ascontiguousarray(level_sigma2_2, np.float32)
        a, b = _featvec(fv1), _featvec(fv2)
        cap = len(d1)
        out = np.zeros((max(cap, 1), 2), np.int32); n = C.c_int(0)
        _check(self._L.orbm_search_for_triangulation(_p(d1), _p(f1), len(d1), C.byref(a), _p(d2), _p(f2), len(d2),
                                                     C.byref(b), _p(F), float(epipole[0]), float(epipole[1]), _p(sf),
                                                     _p(s2), int(bOnlyStereo), int(self.mbCheckOrientation), _p(out),
                                                     cap, C.byref(n), self.device), "orbm_search_for_triangulation")
        return n.value, out[:n.value].copy()


def compute_distinctive_descriptors(desc, ptr, device=0):
    """MapPoint::ComputeDistinctiveDescriptors (R21/src/MapPoint.cc:242-307) for a batch of map points (CSR)."""
    d = np.ascontiguousarray(desc, np.uint8); p = np.ascontiguousarray(ptr, np.int32)
    best = np.zeros(len(p) - 1, np.int32)
    _check(lib().orbm_distinctive_descriptors(_p(d), _p(p), len(p) - 1, _p(best), device), "orbm_distinctive_descriptors")
    return best


GRID_COLS, GRID_ROWS = 64, 48
MPV_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                      ("in_view", "<i4"), ("obs_positive", "<i4")])


class FrameFeatures:
    """The Frame-side steps right after extraction (R21/src/Frame.cc): UndistortKeyPoints (:409-439),
    ComputeImageBounds (:441-470), AssignFeaturesToGrid (:235-250), GetFeaturesInArea (:332-385)."""

    _bounds_cache = {}

    def __init__(self, keys, K, dist, cols, rows, device=0):
        self.device = device
        self.K = np.ascontiguousarray(K, np.float32); self.dist = np.ascontiguousarray(dist, np.float32).ravel()
        assert self.K.shape == (4,), "K = (fx, fy, cx, cy)"
        L = lib()
        k = np.ascontiguousarray(keys, KP_DTYPE)
        # the image bounds belong to the camera, not to the frame: computed once per (K, dist, size) like the reference's
        # mbInitialComputations block (R21/src/Frame.cc:96-110)
        key = (self.K.tobytes(), self.dist.tobytes(), int(cols), int(rows), int(device))
        b = FrameFeatures._bounds_cache.get(key)
        if b is None:
            b = np.zeros(4, np.float32)
            _check(L.orbf_image_bounds(cols, rows, _p(self.K), _p(self.dist), len(self.dist), _p(b), device), "orbf_image_bounds")
            FrameFeatures._bounds_cache[key] = b
        self.bounds = b.copy()
        self.keys_un = np.zeros(len(k), KP_DTYPE)
        self.cell_ptr = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32); self.cell_idx = np.zeros(max(len(k), 1), np.int32)
        n = C.c_int(0)
        # UndistortKeyPoints + AssignFeaturesToGrid in one pass over the device
        _check(L.orbf_build_frame(_p(k), len(k), _p(self.K), _p(self.dist), len(self.dist), _p(self.bounds), _p(self.keys_un), _p(self.cell_ptr),
                                  _p(self.cell_idx), C.byref(n), device), "orbf_build_frame")
        self.n_assigned = n.value

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        """One window or arrays of windows -> list of index arrays (reference order)."""
        qx = np.atleast_1d(np.asarray(x, np.float32)); nq = len(qx)
        qy = np.broadcast_to(np.asarray(y, np.float32), (nq,)).copy(); qr = np.broadcast_to(np.asarray(r, np.float32), (nq,)).copy()
        mn = np.broadcast_to(np.asarray(minLevel, np.int32), (nq,)).copy(); mx = np.broadcast_to(np.asarray(maxLevel, np.int32), (nq,)).copy()
        ptr = np.zeros(nq + 1, np.int32)
        cap = max(64 * nq, 1024)
        while True:
            out = np.zeros(cap, np.int32)
            rc = lib().orbf_features_in_area(_p(self.keys_un), len(self.keys_un), _p(self.cell_ptr), _p(self.cell_idx), _p(self.bounds),
                                             _p(qx), _p(qy), _p(qr), _p(mn), _p(mx), nq, _p(ptr), _p(out), cap, self.device)
            if rc == 3 and ptr[nq] > cap:      # ORB_ERR_CAPACITY: out_ptr is complete, retry with the exact size
                cap = int(ptr[nq]); continue
            _check(rc, "orbf_features_in_area")
            break
        res = [out[ptr[q]:ptr[q + 1]].copy() for q in range(nq)]
        return res if np.ndim(x) else res[0]


def search_by_projection_frame(frame, desc_f, u_right, occupied, scale_factors, map_points, desc_mp, th=1.0, nnratio=0.6,
                               th_high=100):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (R21/src/ORBmatcher.cc:45-130).
    frame: FrameFeatures; map_points: MPV_DTYPE array.  -> (feature -> point, point -> feature, nmatches)."""
    df = np.ascontiguousarray(desc_f, np.uint8); ur = np.ascontiguousarray(u_right, np.float32)
    occ = np.ascontiguousarray(occupied, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    mp = np.ascontiguousarray(map_points, MPV_DTYPE); dm = np.ascontiguousarray(desc_mp, np.uint8)
    nf = len(frame.keys_un)
    fp = np.zeros(max(nf, 1), np.int32); pf = np.zeros(max(len(mp), 1), np.int32); n = C.c_int(0)
    _check(lib().orbm_search_by_projection_frame(_p(frame.keys_un), _p(df), _p(ur), _p(occ), nf, _p(frame.cell_ptr), _p(frame.cell_idx),
                                                 _p(frame.bounds), _p(sf), len(sf), _p(mp), _p(dm), len(mp), float(th), float(nnratio),
                                                 int(th_high), _p(fp), _p(pf), C.byref(n), frame.device),
           "orbm_search_by_projection_frame")
    return fp[:nf], pf[:len(mp)], n.value


PROJ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                       ("obs_positive", "<i4")])


def search_by_projection_last_frame(frame, desc_f, u_right, occupied, scale_factors, points, desc_pts, th, direction=0,
                                    check_orientation=True, th_high=100):
    """ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (R21/src/ORBmatcher.cc:1328-1470).
    points: PROJ_DTYPE array (one per last-frame feature).  -> (feature -> point | -1 | -2, point -> feature, nmatches)."""
    df = np.ascontiguousarray(desc_f, np.uint8); ur = np.ascontiguousarray(u_right, np.float32)
    occ = np.ascontiguousarray(occupied, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    nf = len(frame.keys_un)
    fp = np.zeros(max(nf, 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32); n = C.c_int(0)
    _check(lib().orbm_search_by_projection_last_frame(_p(frame.keys_un), _p(df), _p(ur), _p(occ), nf, _p(frame.cell_ptr), _p(frame.cell_idx),
                                                      _p(frame.bounds), _p(sf), len(sf), _p(pts), _p(dp), len(pts), float(th), int(direction),
                                                      int(bool(check_orientation)), int(th_high), _p(fp), _p(pf), C.byref(n), frame.device),
           "orbm_search_by_projection_last_frame")
    return fp[:nf], pf[:len(pts)], n.value


def search_by_projection_keyframe(frame, desc_f, occupied, scale_factors, points, desc_pts, th, orb_dist, check_orientation=True):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (R21/src/ORBmatcher.cc:1472-1599)."""
    df = np.ascontiguousarray(desc_f, np.uint8); occ = np.ascontiguousarray(occupied, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    nf = len(frame.keys_un)
    fp = np.zeros(max(nf, 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32); n = C.c_int(0)
    _check(lib().orbm_search_by_projection_keyframe(_p(frame.keys_un), _p(df), _p(occ), nf, _p(frame.cell_ptr), _p(frame.cell_idx),
                                                    _p(frame.bounds), _p(sf), len(sf), _p(pts), _p(dp), len(pts), float(th), int(orb_dist),
                                                    int(bool(check_orientation)), _p(fp), _p(pf), C.byref(n), frame.device),
           "orbm_search_by_projection_keyframe")
    return fp[:nf], pf[:len(pts)], n.value


def search_by_projection_sim3(frame, desc_f, occupied, scale_factors, points, desc_pts, th, th_low=50, grid_origin=None):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (R21/src/ORBmatcher.cc:290-403).  grid_origin =
    (float(pKF.mnMinX), float(pKF.mnMinY)): the key frame's integer-truncated bounds (None: the frame's float bounds)."""
    go = None if grid_origin is None else np.ascontiguousarray(grid_origin, np.float32)
    df = np.ascontiguousarray(desc_f, np.uint8); occ = np.ascontiguousarray(occupied, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    nf = len(frame.keys_un)
    fp = np.zeros(max(nf, 1), np.int32); pf = np.zeros(max(len(pts), 1), np.int32); n = C.c_int(0)
    _check(lib().orbm_search_by_projection_sim3(_p(frame.keys_un), _p(df), _p(occ), nf, _p(frame.cell_ptr), _p(frame.cell_idx),
                                                _p(frame.bounds), _p(sf), len(sf), _p(pts), _p(dp), len(pts), float(th), int(th_low),
                                                _p(fp), _p(pf), C.byref(n), None if go is None else _p(go), frame.device),
           "orbm_search_by_projection_sim3")
    return fp[:nf], pf[:len(pts)], n.value


def search_for_initialization(keys1_un, desc1, frame2, desc2, prev_matched, window_size=100, nnratio=0.9, check_orientation=True,
                             th_low=50):
    """ORBmatcher::SearchForInitialization (R21/src/ORBmatcher.cc:405-520).  frame2: FrameFeatures of F2; prev_matched [n1][2]
    (vbPrevMatched).  -> (vnMatches12, updated vbPrevMatched, nmatches)."""
    k1 = np.ascontiguousarray(keys1_un, KP_DTYPE); d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    xy = np.ascontiguousarray(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.zeros(max(len(k1), 1), np.int32); n = C.c_int(0)
    _check(lib().orbm_search_for_initialization(_p(k1), _p(d1), len(k1), _p(frame2.keys_un), _p(d2), len(frame2.keys_un), _p(frame2.cell_ptr),
                                                _p(frame2.cell_idx), _p(frame2.bounds), _p(xy), int(window_size), float(nnratio),
                                                int(bool(check_orientation)), int(th_low), _p(m12), C.byref(n), frame2.device),
           "orbm_search_for_initialization")
    return m12[:len(k1)], xy, n.value


def window_best_match(frame, desc_f, scale_factors, points, desc_pts, th, u_right=None, inv_level_sigma2=None, grid_origin=None):
    """The window search of ORBmatcher::Fuse x2 and SearchBySim3 (R21/src/ORBmatcher.cc:825-1326): per projected point the best
    feature of levels [l-1, l]; inv_level_sigma2 (+ u_right) enables the chi-square gates of Fuse :905-931.
    -> (best_idx, best_dist)."""
    df = np.ascontiguousarray(desc_f, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    pts = np.ascontiguousarray(points, PROJ_DTYPE); dp = np.ascontiguousarray(desc_pts, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    sg = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
    go = None if grid_origin is None else np.ascontiguousarray(grid_origin, np.float32)
    nf = len(frame.keys_un)
    bi = np.zeros(max(len(pts), 1), np.int32); bd = np.zeros(max(len(pts), 1), np.int32)
    _check(lib().orbm_window_best_match(_p(frame.keys_un), _p(df), None if ur is None else _p(ur), nf, _p(frame.cell_ptr),
                                        _p(frame.cell_idx), _p(frame.bounds), _p(sf), None if sg is None else _p(sg), len(sf), _p(pts),
                                        _p(dp), len(pts), float(th), _p(bi), _p(bd), None if go is None else _p(go), frame.device),
           "orbm_window_best_match")
    return bi[:len(pts)], bd[:len(pts)]


class PeerExchange:
    """Symmetric record buffers of all ranks (one process per GPU) for the fused merge + exchange of the sharded map
    search (csrc/peer.cu).  `all_gather_bytes(b) -> [bytes of rank 0, ..., bytes of rank world-1]` is any host-side
    exchange (bench.py passes a torch.distributed one); this package itself stays free of torch."""

    def __init__(self, nq_cap, rank, world, device, all_gather_bytes):
        self._h = C.c_void_p()
        self.rank, self.world = rank, world
        hnd = (C.c_ubyte * 64)()
        _check(lib().orbm_peer_create(int(nq_cap), int(rank), int(world), int(device), C.byref(self._h), hnd), "orbm_peer_create")
        if world > 1:
            handles = all_gather_bytes(bytes(hnd))
            assert len(handles) == world and all(len(h) == 64 for h in handles)
            blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
            _check(lib().orbm_peer_connect(self._h, blob), "orbm_peer_connect")

    def knn2(self, d_q_ptr, nq, d_m_ptr, nm, index_base, d_out_ptr, variant=5, stream=0):
        """Raw device pointers (ints).  Every rank must call it; d_out = records of the WHOLE map on every rank."""
        _check(lib().orbm_knn2_exchange_device(self._h, C.c_void_p(d_q_ptr), int(nq), C.c_void_p(d_m_ptr), int(nm), int(index_base),
                                               C.c_void_p(d_out_ptr), int(variant), C.c_void_p(stream)), "orbm_knn2_exchange_device")

    def exchange_messages(self, d_msg_ptr, nbytes, d_all_ptr, slot_bytes, stream=0):
        """All-gather of one key-frame message per rank over the peer buffers (orbw_exchange_messages_device):
        d_all = [world][slot_bytes]."""
        _check(lib().orbw_exchange_messages_device(self._h, C.c_void_p(d_msg_ptr), int(nbytes), C.c_void_p(d_all_ptr), int(slot_bytes),
                                                   C.c_void_p(stream)), "orbw_exchange_messages_device")

    def error(self):
        e = C.c_int(0)
        _check(lib().orbm_peer_error(self._h, C.byref(e)), "orbm_peer_error")
        return e.value

    def close(self):
        if self._h:
            lib().orbm_peer_destroy(self._h); self._h = C.c_void_p()


def message_bytes(n_frames, cap, flags=0):
    return int(lib().orbw_message_bytes(int(n_frames), int(cap), int(flags)))


def pack_keyframes_device(d_kps, d_desc, d_counts, n_frames, cap, d_msg, d_kps_un=0, d_u_right=0, d_depth=0, d_mappoints=0, stream=0):
    """The agent -> server key-frame message (csrc/wire.cu) from device-resident extractor output; raw device addresses.
    Returns the message's flags (1 stereo | 2 map points | 4 kps_un)."""
    vp = lambda a: C.c_void_p(a) if a else None
    _check(lib().orbw_pack_keyframes_device(vp(d_kps), vp(d_kps_un), vp(d_desc), vp(d_counts), vp(d_u_right), vp(d_depth), vp(d_mappoints),
                                            int(n_frames), int(cap), vp(d_msg), vp(stream)), "orbw_pack_keyframes_device")
    return (1 if d_u_right else 0) | (2 if d_mappoints else 0) | (4 if d_kps_un else 0)


def unpack_keyframes_device(d_msg, n_frames, cap, flags, d_kps, d_desc, d_counts, d_kps_un=0, d_u_right=0, d_depth=0, d_mappoints=0, stream=0,
                            n_msgs=1, msg_stride=0):
    """n_msgs messages msg_stride bytes apart (the [world][slot_bytes] result of PeerExchange.exchange_messages) in one launch; the
    outputs are then [n_msgs][n_frames][cap]."""
    vp = lambda a: C.c_void_p(a) if a else None
    _check(lib().orbw_unpack_keyframes_device(vp(d_msg), int(n_msgs), int(msg_stride), int(n_frames), int(cap), int(flags), vp(d_kps), vp(d_kps_un), vp(d_desc), vp(d_counts),
                                              vp(d_u_right), vp(d_depth), vp(d_mappoints), vp(stream)), "orbw_unpack_keyframes_device")


def quantize_lcm(keys):
    """Key points as the reference's server receives them (int16 truncation of the LCM message,
    R21/include/lcmKeyFrame/lcmKeyPoint.hpp:19-31)."""
    k = np.ascontiguousarray(keys, KP_DTYPE).copy()
    _check(lib().orbw_quantize_lcm_host(_p(k), len(k)), "orbw_quantize_lcm_host")
    return k


class ORBVocabulary:
    """The device-resident vocabulary tree behind Frame::ComputeBoW (R21/src/Frame.cc:400-407):
    transform(descriptors, levelsup=4) -> (BowVector as (words, values), FeatureVector as (nodes, ptr, idx))."""

    def __init__(self, child_ptr, child_idx, node_desc, word_id, weight, depth_L, device=0):
        cp = np.ascontiguousarray(child_ptr, np.int32); ci = np.ascontiguousarray(child_idx, np.int32)
        nd = np.ascontiguousarray(node_desc, np.uint8); wi = np.ascontiguousarray(word_id, np.int32)
        wt = np.ascontiguousarray(weight, np.float64)
        self._h = C.c_void_p()
        _check(lib().orbv_create(_p(cp), _p(ci), _p(nd), _p(wi), _p(wt), len(cp) - 1, int(depth_L), device, C.byref(self._h)), "orbv_create")

    def close(self):
        if self._h:
            lib().orbv_destroy(self._h); self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        """Per-feature (word id, node id, weight)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        w = np.zeros(len(d), np.int32); nd = np.zeros(len(d), np.int32); wt = np.zeros(len(d), np.float64)
        _check(lib().orbv_transform(self._h, _p(d), len(d), int(levelsup), _p(w), _p(nd), _p(wt)), "orbv_transform")
        return w, nd, wt

    def transform(self, desc, levelsup=4, normalize=True):
        w, nd, wt = self.transform_features(desc, levelsup)
        n = len(w)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.int32); fp = np.zeros(n + 1, np.int32); fi = np.zeros(max(n, 1), np.int32)
        nw = C.c_int(0); nn = C.c_int(0)
        _check(lib().orbv_bow_vectors(_p(w), _p(nd), _p(wt), n, int(bool(normalize)), _p(bw), _p(bv), C.byref(nw), _p(fn), _p(fp), _p(fi),
                                      C.byref(nn)), "orbv_bow_vectors")
        return (bw[:nw.value], bv[:nw.value]), (fn[:nn.value], fp[:nn.value + 1], fi[:fp[nn.value]])


def compute_stereo_matches(ext_left, ext_right, keys_left, desc_left, keys_right, desc_right, mbf, mb):
    """Frame::ComputeStereoMatches (R21/src/Frame.cc:471-645) -> (mvuRight, mvDepth, n)."""
    kl = np.ascontiguousarray(keys_left, KP_DTYPE); kr = np.ascontiguousarray(keys_right, KP_DTYPE)
    dl = np.ascontiguousarray(desc_left, np.uint8); dr = np.ascontiguousarray(desc_right, np.uint8)
    ur = np.zeros(len(kl), np.float32); dep = np.zeros(len(kl), np.float32); n = C.c_int(0)
    _check(lib().orbm_stereo_matches(ext_left._h, ext_right._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                     float(mbf), float(mb), _p(ur), _p(dep), C.byref(n)), "orbm_stereo_matches")
    return ur, dep, n.value


def stereo_scratch_bytes(n_pairs, cap):
    return int(lib().orbm_stereo_scratch_bytes(int(n_pairs), int(cap)))


def compute_stereo_matches_batch_device(ext_left, ext_right, d_kl, d_dl, d_cl, d_kr, d_dr, d_cr, n_pairs, cap, mbf, mb, d_scratch,
                                        d_u_right, d_depth, d_n_matches, stream=0):
    """Frame::ComputeStereoMatches for a batch of stereo pairs, device-resident (all d_* are raw device addresses): the
    outputs of two extract_batch_device/_async calls in, mvuRight / mvDepth [n_pairs][cap] and the match counts out."""
    _check(lib().orbm_stereo_matches_batch_device(ext_left._h, ext_right._h, C.c_void_p(d_kl), C.c_void_p(d_dl), C.c_void_p(d_cl),
                                                  C.c_void_p(d_kr), C.c_void_p(d_dr), C.c_void_p(d_cr), int(n_pairs), int(cap),
                                                  float(mbf), float(mb), C.c_void_p(d_scratch), C.c_void_p(d_u_right),
                                                  C.c_void_p(d_depth), C.c_void_p(d_n_matches), C.c_void_p(stream)),
           "orbm_stereo_matches_batch_device")
