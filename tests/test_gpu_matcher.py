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


@pytest.mark.parametrize("nq,nm", [(1, 1), (7, 300), (257, 5000), (2000, 60000), (100, 0)])
def test_knn2_matches_oracle(orb, oracle, synth, nq, nm):
    m = synth.descriptors(max(nm, 1), seed=5)[:nm]
    if nm > nq:
        q, m, _ = synth.query_set(m, nq=nq, seed=6)
    else:
        q = synth.descriptors(nq, seed=7)
    # duplicates in the map exercise the tie rule (first index wins, second distance counts duplicates)
    if nm > 10:
        m[nm // 2] = m[3]; m[nm - 1] = m[3]
        q[0] = m[3]
    bi, bd, sd, si = orb.ORBmatcher().knn2(q, m, index_base=1000)
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
