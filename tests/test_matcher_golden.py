"""Golden vectors of the matcher loops: tests/golden/matcher_ref.npz holds the POD inputs of every C-ABI call the drop-in
ORBmatcher class made on two scenes and the outputs -- which tools/gen_golden_matcher.py kept only after checking the
scenes' results against the REFERENCE's own ORBmatcher.cc (oracle/_ref/libmatchref.so).  Replayed here through the CPU
oracle (pins the restatement anywhere, without /root/reference) and, with -m gpu, through liborbcuda's C ABI."""
import ctypes as C
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls():
    z = np.load(os.path.join(ROOT, "tests", "golden", "matcher_ref.npz"))
    index = json.loads(bytes(z["index"]).decode())
    calls = {}
    for call, fn, arg, h in index:
        calls.setdefault((call, fn), {})[arg] = z[h]
    return [(k[0], k[1], v) for k, v in sorted(calls.items())]


def _fv(a, name):
    return a[name + "_ids"], a[name + "_ptr"], a[name + "_idx"]


def _oracle(fn, a, oracle):
    """-> dict of outputs by the recorded output names"""
    sc = a.get("scalars")
    if fn == "assign_grid":
        ptr, idx = oracle.assign_grid(a["kps_un"], a["bounds"])
        return {"out_cell_ptr": ptr, "out_cell_idx": idx[:ptr[-1]]}
    if fn == "bow_kf_f":
        n, m = oracle.search_by_bow_kf_f(a["desc_kf"], a["angle_kf"], a["kf_valid"], _fv(a, "fv_kf"), a["desc_f"], a["angle_f"], _fv(a, "fv_f"),
                                         float(sc[0]), int(sc[1]))
        return {"out_match_f": m, "n": n}
    if fn == "bow_kf_kf":
        n, m = oracle.search_by_bow_kf_kf(a["desc1"], a["angle1"], a["valid1"], _fv(a, "fv1"), a["desc2"], a["angle2"], a["valid2"], _fv(a, "fv2"),
                                          float(sc[0]), int(sc[1]))
        return {"out_match12": m, "n": n}
    if fn == "triangulation":
        n, p = oracle.search_for_triangulation(a["desc1"], a["f1"], _fv(a, "fv1"), a["desc2"], a["f2"], _fv(a, "fv2"), a["F12"], float(sc[0]),
                                               float(sc[1]), a["scale_factors2"], a["level_sigma2_2"], int(sc[2]), int(sc[3]))
        return {"out_pairs": p.reshape(-1, 2), "n": n}
    if fn == "proj_frame":
        fp, pf, n = oracle.search_by_projection_frame(a["kps_un"], a["desc_f"], a["u_right"], a["occupied"], a["cell_ptr"], a["cell_idx"], a["bounds"],
                                                      a["scale_factors"], a["mps"], a["desc_mp"], float(sc[0]), float(sc[1]), int(sc[2]))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_last":
        fp, pf, n = oracle.search_by_projection_last_frame(a["kps_un"], a["desc_f"], a["u_right"], a["occupied"], a["cell_ptr"], a["cell_idx"],
                                                           a["bounds"], a["scale_factors"], a["pts"], a["desc_pts"], float(sc[0]), int(sc[1]),
                                                           int(sc[2]), int(sc[3]))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_keyframe":
        fp, pf, n = oracle.search_by_projection_keyframe(a["kps_un"], a["desc_f"], a["occupied"], a["cell_ptr"], a["cell_idx"], a["bounds"],
                                                         a["scale_factors"], a["pts"], a["desc_pts"], float(sc[0]), int(sc[1]), int(sc[2]))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_sim3":
        fp, pf, n = oracle.search_by_projection_sim3(a["kps_un"], a["desc_f"], a["occupied"], a["cell_ptr"], a["cell_idx"], a["bounds"],
                                                     a["scale_factors"], a["pts"], a["desc_pts"], float(sc[0]), int(sc[1]), a.get("grid_origin"))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "window_best":
        bi, bd = oracle.window_best_match(a["kps_un"], a["desc_f"], a.get("u_right"), a["cell_ptr"], a["cell_idx"], a["bounds"], a["scale_factors"],
                                          a.get("inv_level_sigma2"), a["pts"], a["desc_pts"], float(sc[0]), a.get("grid_origin"))
        return {"out_best_idx": bi, "out_best_dist": bd}
    if fn == "init":
        m12, xy, n = oracle.search_for_initialization(a["kps1_un"], a["desc1"], a["kps2_un"], a["desc2"], a["cell_ptr"], a["cell_idx"], a["bounds"],
                                                      a["prev_xy"], int(sc[0]), float(sc[1]), int(sc[2]), int(sc[3]))
        return {"out_matches12": m12[:len(a["kps1_un"])], "out_prev_xy": xy, "n": n}
    raise AssertionError(fn)


class _GridFrame:
    """what the orbcuda wrappers read from a FrameFeatures object"""

    def __init__(self, a, kps="kps_un", device=0):
        self.keys_un = np.ascontiguousarray(a[kps]); self.cell_ptr = np.ascontiguousarray(a["cell_ptr"])
        self.cell_idx = np.ascontiguousarray(np.concatenate([a["cell_idx"], np.zeros(1, np.int32)]))
        self.bounds = np.ascontiguousarray(a["bounds"]); self.device = device


def _cuda(fn, a, orb):
    sc = a.get("scalars")
    if fn == "assign_grid":
        L = orb.lib()
        n = len(a["kps_un"])
        k = np.ascontiguousarray(a["kps_un"]); b = np.ascontiguousarray(a["bounds"])
        ptr = np.zeros(64 * 48 + 1, np.int32); idx = np.zeros(max(n, 1), np.int32); na = C.c_int(0)
        assert L.orbf_assign_grid(k.ctypes.data, n, b.ctypes.data, ptr.ctypes.data, idx.ctypes.data, C.byref(na), 0) == 0
        return {"out_cell_ptr": ptr, "out_cell_idx": idx[:ptr[-1]]}
    m = orb.ORBmatcher(float(sc[0]) if fn.startswith("bow") else 0.6, bool(int(sc[1])) if fn.startswith("bow") else True)
    if fn == "bow_kf_f":
        n, out = m.SearchByBoW(a["desc_kf"], a["angle_kf"], a["kf_valid"], _fv(a, "fv_kf"), a["desc_f"], a["angle_f"], _fv(a, "fv_f"))
        return {"out_match_f": out, "n": n}
    if fn == "bow_kf_kf":
        n, out = m.SearchByBoW_KF(a["desc1"], a["angle1"], a["valid1"], _fv(a, "fv1"), a["desc2"], a["angle2"], a["valid2"], _fv(a, "fv2"))
        return {"out_match12": out, "n": n}
    if fn == "triangulation":
        mt = orb.ORBmatcher(0.6, bool(int(sc[3])))
        n, p = mt.SearchForTriangulation(a["desc1"], a["f1"], _fv(a, "fv1"), a["desc2"], a["f2"], _fv(a, "fv2"), a["F12"].reshape(3, 3),
                                         (float(sc[0]), float(sc[1])), a["scale_factors2"], a["level_sigma2_2"], bool(int(sc[2])))
        return {"out_pairs": np.asarray(p).reshape(-1, 2), "n": n}
    if fn == "proj_frame":
        fp, pf, n = orb.search_by_projection_frame(_GridFrame(a), a["desc_f"], a["u_right"], a["occupied"], a["scale_factors"], a["mps"], a["desc_mp"],
                                                   th=float(sc[0]), nnratio=float(sc[1]), th_high=int(sc[2]))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_last":
        fp, pf, n = orb.search_by_projection_last_frame(_GridFrame(a), a["desc_f"], a["u_right"], a["occupied"], a["scale_factors"], a["pts"],
                                                        a["desc_pts"], float(sc[0]), int(sc[1]), bool(int(sc[2])), int(sc[3]))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_keyframe":
        fp, pf, n = orb.search_by_projection_keyframe(_GridFrame(a), a["desc_f"], a["occupied"], a["scale_factors"], a["pts"], a["desc_pts"],
                                                      float(sc[0]), int(sc[1]), bool(int(sc[2])))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "proj_sim3":
        fp, pf, n = orb.search_by_projection_sim3(_GridFrame(a), a["desc_f"], a["occupied"], a["scale_factors"], a["pts"], a["desc_pts"], float(sc[0]),
                                                  int(sc[1]), a.get("grid_origin"))
        return {"out_feature_point": fp, "out_point_feature": pf, "n": n}
    if fn == "window_best":
        bi, bd = orb.window_best_match(_GridFrame(a), a["desc_f"], a["scale_factors"], a["pts"], a["desc_pts"], float(sc[0]), a.get("u_right"),
                                       a.get("inv_level_sigma2"), a.get("grid_origin"))
        return {"out_best_idx": bi, "out_best_dist": bd}
    if fn == "init":
        m12, xy, n = orb.search_for_initialization(a["kps1_un"], a["desc1"], _GridFrame(a, "kps2_un"), a["desc2"], a["prev_xy"], int(sc[0]), float(sc[1]),
                                                   bool(int(sc[2])), int(sc[3]))
        return {"out_matches12": m12, "out_prev_xy": xy, "n": n}
    raise AssertionError(fn)


def _check(call, fn, a, got):
    n_rec = {"bow_kf_f": 2, "bow_kf_kf": 2, "triangulation": 4, "proj_frame": 3, "proj_last": 4, "proj_keyframe": 3, "proj_sim3": 2, "init": 4}
    for k, v in got.items():
        if k == "n":
            assert int(v) == int(a["scalars"][n_rec[fn]]), (call, fn, "return value")
        else:
            assert np.array_equal(np.asarray(v), a[k]), (call, fn, k)


def test_oracle_replays_reference_checked_calls(oracle):
    calls = _calls()
    assert len(calls) >= 70 and len(set(fn for _, fn, _ in calls)) == 10
    for call, fn, a in calls:
        _check(call, fn, a, _oracle(fn, a, oracle))


@pytest.mark.gpu
def test_cuda_replays_reference_checked_calls():
    import orbcuda
    for call, fn, a in _calls():
        _check(call, fn, a, _cuda(fn, a, orbcuda))
