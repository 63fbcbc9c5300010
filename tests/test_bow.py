"""BoW transform: Frame::ComputeBoW = DBoW2 TemplatedVocabulary::transform(desc, BowVec, FeatVec, 4).
DBoW2 is not vendored by the reference and no vocabulary file is shipped: the oracle restates the published algorithm
("parity unpinned"), the CPU tests check it against an independent numpy walk, the GPU tests check CUDA == oracle."""
import numpy as np
import pytest


def make_vocabulary(k=10, L=4, seed=0, early_leaf=0.05, zero_weight=0.05, near=True):
    """Random tree in creation (BFS) order like DBoW2's: node 0 = root; children get their parent's descriptor with a few
    bits flipped (near=True), which makes equal distances -- the tie rule -- frequent."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]
    children = [[]]
    level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in frontier:
            if lv > 1 and rng.random() < early_leaf:
                continue                                  # an early leaf (empty cluster in k-means)
            for _ in range(k if lv > 1 or True else k):
                d = desc[p].copy() if near and lv > 1 else rng.integers(0, 256, 32, dtype=np.uint8)
                if near and lv > 1:
                    for b in rng.integers(0, 256, 12):
                        d[b >> 3] ^= np.uint8(1 << (b & 7))
                desc.append(d); children.append([]); level.append(lv)
                children[p].append(len(desc) - 1); nxt.append(len(desc) - 1)
        frontier = nxt
    n = len(desc)
    child_ptr = np.zeros(n + 1, np.int32)
    child_idx = []
    for i in range(n):
        child_idx += children[i]
        child_ptr[i + 1] = len(child_idx)
    word_id = np.full(n, -1, np.int32)
    leaves = [i for i in range(n) if not children[i]]
    word_id[leaves] = np.arange(len(leaves), dtype=np.int32)
    weight = np.zeros(n, np.float64)
    weight[leaves] = rng.uniform(0.01, 9.0, len(leaves))
    weight[[l for l in leaves if rng.random() < zero_weight]] = 0.0      # stop words
    return {"child_ptr": child_ptr, "child_idx": np.array(child_idx, np.int32), "node_desc": np.array(desc, np.uint8),
            "word_id": word_id, "weight": weight, "L": L, "level": np.array(level)}


def _queries(voc, n, seed):
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(voc["word_id"] >= 0)
    q = voc["node_desc"][rng.choice(leaves, n)].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, 40)):
            q[i, b >> 3] ^= np.uint8(1 << (b & 7))
    q[: n // 10] = rng.integers(0, 256, (n // 10, 32), dtype=np.uint8)
    return q


def _numpy_walk(q, voc, levelsup):
    pop = np.array([bin(i).count("1") for i in range(256)])
    cp, ci, nd = voc["child_ptr"], voc["child_idx"], voc["node_desc"]
    nid_level = voc["L"] - levelsup
    out = []
    for f in range(len(q)):
        node, nid, lv = 0, 0, 0
        while cp[node + 1] > cp[node]:
            lv += 1
            ch = ci[cp[node]:cp[node + 1]]
            d = pop[nd[ch] ^ q[f]].sum(1)
            node = ch[int(np.argmin(d))]                 # argmin returns the first minimum = DBoW2's strict '<'
            if lv == nid_level:
                nid = node
        out.append((voc["word_id"][node], nid, voc["weight"][node]))
    return out


@pytest.mark.parametrize("k,L,levelsup", [(10, 4, 2), (10, 3, 4), (4, 6, 4), (37, 2, 1)])
def test_oracle_transform_matches_numpy_walk(oracle, k, L, levelsup):
    voc = make_vocabulary(k, L, seed=k + L)
    q = _queries(voc, 300, 1)
    w, nd, wt = oracle.bow_transform(q, voc, levelsup)
    ref = _numpy_walk(q, voc, levelsup)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(w, nd, wt)] == [(int(a), int(b), float(c)) for a, b, c in ref]
    (bw, bv), (fn, fp, fi) = oracle.bow_vectors(w, nd, wt)
    keep = wt > 0
    assert np.array_equal(bw, np.unique(w[keep])) and abs(bv.sum() - 1.0) < 1e-12
    for j, node in enumerate(fn):
        assert np.array_equal(fi[fp[j]:fp[j + 1]], np.flatnonzero(keep & (nd == node)))      # feature order inside a node
    assert np.all(np.diff(fn) > 0) and fp[-1] == keep.sum()


@pytest.fixture(scope="module")
def orb():
    import orbcuda
    if orbcuda.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return orbcuda


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,levelsup,n", [(10, 5, 4, 2000), (10, 4, 2, 1500), (10, 3, 4, 700), (4, 6, 4, 1000), (37, 2, 1, 500), (10, 4, 4, 0)])
def test_gpu_bow_transform(orb, oracle, k, L, levelsup, n):
    voc = make_vocabulary(k, L, seed=100 + k + L)
    q = _queries(voc, max(n, 10), 2)[:n]
    v = orb.ORBVocabulary(voc["child_ptr"], voc["child_idx"], voc["node_desc"], voc["word_id"], voc["weight"], L)
    w, nd, wt = v.transform_features(q, levelsup)
    rw, rn, rwt = oracle.bow_transform(q, voc, levelsup)
    assert np.array_equal(w, rw) and np.array_equal(nd, rn) and np.array_equal(wt.view(np.uint64), rwt.view(np.uint64))
    (bw, bv), (fn, fp, fi) = v.transform(q, levelsup)
    (obw, obv), (ofn, ofp, ofi) = oracle.bow_vectors(rw, rn, rwt)
    assert np.array_equal(bw, obw) and np.array_equal(bv.view(np.uint64), obv.view(np.uint64))      # doubles, bit for bit
    assert np.array_equal(fn, ofn) and np.array_equal(fp, ofp) and np.array_equal(fi, ofi)
    v.close()


@pytest.mark.gpu
def test_gpu_bow_feeds_search_by_bow(orb, oracle):
    """The FeatureVector CSR produced here is the layout SearchByBoW consumes."""
    voc = make_vocabulary(10, 4, seed=7)
    v = orb.ORBVocabulary(voc["child_ptr"], voc["child_idx"], voc["node_desc"], voc["word_id"], voc["weight"], 4)
    d1 = _queries(voc, 800, 3)
    rng = np.random.default_rng(4)
    d2 = d1[rng.permutation(800)].copy()
    for i in range(800):
        for b in rng.integers(0, 256, 6):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    fv1 = v.transform(d1, 2)[1]; fv2 = v.transform(d2, 2)[1]
    a1 = rng.uniform(0, 360, 800).astype(np.float32); a2 = rng.uniform(0, 360, 800).astype(np.float32)
    valid = np.ones(800, np.uint8)
    got = orb.ORBmatcher(0.7, False).SearchByBoW(d1, a1, valid, fv1, d2, a2, fv2)
    ref = oracle.search_by_bow_kf_f(d1, a1, valid, fv1, d2, a2, fv2, 0.7, False)
    assert got[0] == ref[0] > 100 and np.array_equal(got[1], ref[1])
