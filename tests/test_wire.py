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


def _message_case(B, cap, seed, orbcuda):
    rng = np.random.default_rng(seed)
    k = _keys(B * cap, seed).reshape(B, cap)
    ku = _keys(B * cap, seed + 100).reshape(B, cap)
    desc = rng.integers(0, 256, (B, cap, 32), dtype=np.uint8)
    counts = rng.integers(0, cap + 1, B).astype(np.int32); counts[0] = cap; counts[-1] = 0
    ur = rng.uniform(-1, 700, (B, cap)).astype(np.float32); dep = rng.uniform(-1, 40, (B, cap)).astype(np.float32)
    mp = rng.normal(0, 5, (B, cap, 4)).astype(np.float32); mp[..., 3] = (rng.random((B, cap)) < 0.6).astype(np.float32)
    return k, ku, desc, counts, ur, dep, mp


@pytest.mark.gpu
def test_keyframe_message_pack_unpack_roundtrip():
    """pack -> (the message is one contiguous buffer) -> unpack returns, per key frame, the int16-truncated key points and every
    other field bit for bit; a message read with the wrong shape is refused."""
    import torch
    import orbcuda
    B, cap = 7, 1100
    k, ku, desc, counts, ur, dep, mp = _message_case(B, cap, 5, orbcuda)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).cuda()
    d = dict(k=t(k), ku=t(ku), desc=t(desc), c=t(counts), ur=t(ur), dep=t(dep), mp=t(mp))
    for with_un, with_stereo, with_mp in ((False, False, False), (True, True, True), (False, True, False)):
        flags = (1 if with_stereo else 0) | (2 if with_mp else 0) | (4 if with_un else 0)
        msg = torch.zeros(orbcuda.message_bytes(B, cap, flags), dtype=torch.uint8, device="cuda")
        got_flags = orbcuda.pack_keyframes_device(d["k"].data_ptr(), d["desc"].data_ptr(), d["c"].data_ptr(), B, cap, msg.data_ptr(),
                                                  d["ku"].data_ptr() if with_un else 0, d["ur"].data_ptr() if with_stereo else 0,
                                                  d["dep"].data_ptr() if with_stereo else 0, d["mp"].data_ptr() if with_mp else 0)
        assert got_flags == flags
        o = dict(k=torch.zeros_like(d["k"]), ku=torch.zeros_like(d["ku"]), desc=torch.zeros_like(d["desc"]), c=torch.zeros_like(d["c"]),
                 ur=torch.zeros_like(d["ur"]), dep=torch.zeros_like(d["dep"]), mp=torch.zeros_like(d["mp"]))
        orbcuda.unpack_keyframes_device(msg.data_ptr(), B, cap, flags, o["k"].data_ptr(), o["desc"].data_ptr(), o["c"].data_ptr(),
                                        o["ku"].data_ptr() if with_un else 0, o["ur"].data_ptr() if with_stereo else 0,
                                        o["dep"].data_ptr() if with_stereo else 0, o["mp"].data_ptr() if with_mp else 0)
        torch.cuda.synchronize()
        hdr = msg[:32].cpu().numpy().view(np.int32)
        assert hdr[1] == B and hdr[2] == counts.sum() and hdr[3] == flags and hdr[4] == cap
        gk = o["k"].cpu().numpy().view(orbcuda.KP_DTYPE).reshape(B, cap); gku = o["ku"].cpu().numpy().view(orbcuda.KP_DTYPE).reshape(B, cap)
        gd = o["desc"].cpu().numpy().reshape(B, cap, 32); gc = o["c"].cpu().numpy().view(np.int32)
        gur = o["ur"].cpu().numpy().view(np.float32).reshape(B, cap); gdep = o["dep"].cpu().numpy().view(np.float32).reshape(B, cap)
        gmp = o["mp"].cpu().numpy().view(np.float32).reshape(B, cap, 4)
        assert np.array_equal(gc, counts)
        for b in range(B):
            n = counts[b]
            assert gk[b, :n].tobytes() == orbcuda.quantize_lcm(k[b, :n]).tobytes() and np.array_equal(gd[b, :n], desc[b, :n])
            if with_un:
                assert gku[b, :n].tobytes() == orbcuda.quantize_lcm(ku[b, :n]).tobytes()
            if with_stereo:
                assert np.array_equal(gur[b, :n], ur[b, :n]) and np.array_equal(gdep[b, :n], dep[b, :n])
            if with_mp:
                assert np.array_equal(gmp[b, :n], mp[b, :n])
        # a receiver expecting another shape gets counts = -1, not garbage
        orbcuda.unpack_keyframes_device(msg.data_ptr(), B, cap, flags ^ 4, o["k"].data_ptr(), o["desc"].data_ptr(), o["c"].data_ptr(), 0, 0, 0, 0)
        torch.cuda.synchronize()
        assert (o["c"].cpu().numpy().view(np.int32) == -1).all()
