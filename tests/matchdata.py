"""Synthetic inputs for the matcher loops (the reference ships no vocabulary / sequences):
two correlated descriptor sets, a synthetic vocabulary partition (FeatureVector as CSR), key point
geometry consistent with a sideways camera motion."""
import numpy as np


def featvec(desc, nbits=6):
    """node id = top bits of descriptor byte 0 (nearby descriptors mostly share a node); CSR with ascending
    node ids and ascending feature indices inside a node, like DBoW2 fills its FeatureVector."""
    node = (desc[:, 0] >> (8 - nbits)).astype(np.int32) * 7 + 3     # sparse, non-contiguous ids
    ids = np.unique(node)
    ptr = [0]; idx = []
    for n in ids:
        members = np.nonzero(node == n)[0]
        idx.extend(members.tolist()); ptr.append(len(idx))
    return ids.astype(np.int32), np.array(ptr, np.int32), np.array(idx, np.int32)


def two_views(n1=1000, n2=1100, seed=0, noise=0.06, overlap=0.7):
    rng = np.random.Generator(np.random.PCG64(seed))
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    nshared = int(min(n1, n2) * overlap)
    src = rng.choice(n1, nshared, replace=False); dst = rng.choice(n2, nshared, replace=False)
    bits = np.unpackbits(d1[src], axis=1)
    flip = rng.random(bits.shape) < noise
    d2[dst] = np.packbits(bits ^ flip, axis=1)
    # a few exact duplicates / near ties to exercise best == second-best
    d2[dst[:5]] = d1[src[:5]]
    if n2 > 20:
        d2[(dst[0] + 1) % n2] = d1[src[0]]
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = rng.uniform(0, 360, n2).astype(np.float32)
    a2[dst] = (a1[src] - np.float32(20.0) + rng.normal(0, 3, nshared).astype(np.float32)) % np.float32(360.0)
    return d1, d2, a1, a2, src, dst, rng


def tri_features(n1, n2, src, dst, rng):
    from oracle_lib import TRI_DTYPE
    f1 = np.zeros(n1, TRI_DTYPE); f2 = np.zeros(n2, TRI_DTYPE)
    f1["x"] = rng.uniform(20, 620, n1); f1["y"] = rng.uniform(20, 460, n1)
    f2["x"] = rng.uniform(20, 620, n2); f2["y"] = rng.uniform(20, 460, n2)
    f2["x"][dst] = f1["x"][src] - rng.uniform(2, 40, len(src)).astype(np.float32)
    f2["y"][dst] = f1["y"][src] + rng.normal(0, 0.8, len(src)).astype(np.float32)
    f1["octave"] = rng.integers(0, 8, n1); f2["octave"] = rng.integers(0, 8, n2)
    f1["u_right"] = np.where(rng.random(n1) < 0.3, f1["x"] - 5, -1).astype(np.float32)
    f2["u_right"] = np.where(rng.random(n2) < 0.3, f2["x"] - 5, -1).astype(np.float32)
    f1["has_mp"] = rng.random(n1) < 0.3; f2["has_mp"] = rng.random(n2) < 0.3
    return f1, f2
