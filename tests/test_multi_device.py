"""Entry points on a device other than 0 (kernel attributes such as the dynamic shared-memory opt-in are per device).
Skipped on a single-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_second_device_runs_every_large_smem_kernel(oracle, synth):
    import orbcuda
    if orbcuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    img = synth.frame(5, 640, 480)
    ref_k, ref_d = oracle.OracleExtractor(1000, trig_mode=1).extract(img)
    for dev in (0, 1, 0):
        k, d = orbcuda.ORBextractor(1000, 1.2, 8, 20, 7, device=dev)(img)
        assert k.tobytes() == ref_k.tobytes() and np.array_equal(d, ref_d)
    m = synth.descriptors(30000, seed=3)
    q, m, _ = synth.query_set(m, nq=300, seed=4)
    base = orbcuda.ORBmatcher(device=0).knn2(q, m, variant=0)
    for dev in (1, 0):
        for variant in (1, 3, 4, 5):
            got = orbcuda.ORBmatcher(device=dev).knn2(q, m, variant=variant)
            assert all(np.array_equal(a, b) for a, b in zip(got, base)), (dev, variant)
    keys = np.zeros(500, orbcuda.KP_DTYPE)
    rng = np.random.default_rng(1)
    keys["x"] = rng.uniform(20, 620, 500); keys["y"] = rng.uniform(20, 460, 500)
    K = np.array([500, 500, 320, 240], np.float32); D = np.array([0.1, -0.2, 0.001, 0.001], np.float32)
    f0 = orbcuda.FrameFeatures(keys, K, D, 640, 480, device=0); f1 = orbcuda.FrameFeatures(keys, K, D, 640, 480, device=1)
    assert f0.keys_un.tobytes() == f1.keys_un.tobytes() and np.array_equal(f0.cell_idx, f1.cell_idx) and np.array_equal(f0.cell_ptr, f1.cell_ptr)
