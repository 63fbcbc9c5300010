"""Key-frame message semantics: the reference's LCM message truncates the key point floats to int16
(include/lcmKeyFrame/lcmKeyPoint.hpp:19-31, filled ros_mono.cc:2071-2077) and ships descriptors one float per byte."""
import ctypes as C
import numpy as np
import pytest


def _lcm_roundtrip(keys):
    """What the reference's encode/decode does to a key point, written with numpy casts."""
    k = keys.copy()
    for f in ("x", "y", "size", "response"):
        k[f] = k[f].astype(np.int32).astype(np.int16).astype(np.float32)      # float -> int16_t truncates toward zero
    for f in ("octave", "class_id"):
        k[f] = k[f].astype(np.int16).astype(np.int32)
    return k


def _keys(n, seed):
    import orbcuda
    rng = np.random.default_rng(seed)
    k = np.zeros(n, orbcuda.KP_DTYPE)
    k["x"] = rng.uniform(0, 752, n); k["y"] = rng.uniform(0, 480, n); k["size"] = 31 * 1.2 ** rng.integers(0, 8, n)
    k["angle"] = rng.uniform(0, 360, n); k["response"] = rng.uniform(0, 255, n); k["octave"] = rng.integers(0, 8, n); k["class_id"] = -1
    k["x"][:4] = [0.0, 0.99, 751.5, -0.5]      # truncation toward zero, also for (hypothetical) negatives
    return k


def test_host_quantisation_matches_lcm_roundtrip():
    import orbcuda
    k = _keys(5000, 1)
    q = orbcuda.quantize_lcm(k)
    assert q.tobytes() == _lcm_roundtrip(k).tobytes()
    assert np.array_equal(q["angle"], k["angle"])                      # the angle stays a float on the wire
    d = np.random.default_rng(2).integers(0, 256, (100, 32), dtype=np.uint8)
    assert np.array_equal(d.astype(np.float32).astype(np.uint8), d)    # float-per-byte descriptors are lossless


@pytest.mark.gpu
def test_device_quantisation_matches_host():
    import torch
    import orbcuda
    if orbcuda.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    B, cap = 5, 1300
    k = _keys(B * cap, 3).reshape(B, cap)
    counts = np.array([1300, 0, 17, 999, 1], np.int32)
    d_k = torch.from_numpy(k.view(np.uint8).reshape(B, cap * 28).copy()).cuda()
    d_c = torch.from_numpy(counts).cuda()
    rc = orbcuda.lib().orbw_quantize_lcm_device(C.c_void_p(d_k.data_ptr()), C.c_void_p(d_c.data_ptr()), B, cap,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    got = d_k.cpu().numpy().reshape(-1).view(orbcuda.KP_DTYPE).reshape(B, cap)
    for b in range(B):
        n = counts[b]
        assert got[b, :n].tobytes() == orbcuda.quantize_lcm(k[b, :n]).tobytes()
        assert got[b, n:].tobytes() == k[b, n:].tobytes()              # entries past the count are not touched
