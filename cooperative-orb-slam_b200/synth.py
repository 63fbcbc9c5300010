"""Procedural synthetic inputs.  Identical bytes feed the CUDA path, the
oracle and the reference arm.  No reference code or data involved; cv2/numpy only."""
import numpy as np
import cv2


def frame(seed: int, width: int = 640, height: int = 480, low_texture: bool = False) -> np.ndarray:
    """Textured grayscale frame: multi-octave bicubic noise + random rectangles + pixel noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.zeros((height, width), np.float32)
    amps = {8: 90.0, 16: 60.0, 32: 40.0, 64: 30.0}
    for s, amp in amps.items():
        if low_texture:
            amp /= 4.0
        small = rng.random((height // s + 2, width // s + 2), dtype=np.float32)
        up = cv2.resize(small, (width, height), interpolation=cv2.INTER_CUBIC)
        img += amp * up
    nrect = 150
    for _ in range(nrect):
        w = int(rng.integers(8, 41)); h = int(rng.integers(8, 41))
        x = int(rng.integers(0, max(1, width - w))); y = int(rng.integers(0, max(1, height - h)))
        d = float(rng.uniform(-60, 60)) / (4.0 if low_texture else 1.0)
        img[y:y + h, x:x + w] += d
    img += rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    lo, hi = float(img.min()), float(img.max())
    if low_texture:
        # keep the contrast low: centre around mid-grey instead of stretching
        out = np.clip(img - img.mean() + 128.0, 0, 255)
    else:
        out = (img - lo) * (255.0 / max(hi - lo, 1e-6))
    return np.ascontiguousarray(np.rint(out).astype(np.uint8))


def stereo_pair(seed: int, width: int = 752, height: int = 480):
    """Left frame + right frame = left shifted by a per-row-band disparity (4..40 px) + fresh noise."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    wide = frame(seed, width + 64, height)
    left = np.ascontiguousarray(wide[:, 32:32 + width])
    right = np.empty_like(left)
    band = 48
    for y0 in range(0, height, band):
        d = int(rng.integers(4, 41))
        # a point at x in the left image appears at x-d in the right image
        x0 = 32 + d
        x0 = min(max(x0, 0), 64)
        right[y0:y0 + band] = wide[y0:y0 + band, x0:x0 + width]
    noise = rng.normal(0.0, 2.0, right.shape)
    right = np.clip(np.rint(right.astype(np.float32) + noise), 0, 255).astype(np.uint8)
    return left, np.ascontiguousarray(right)


def descriptors(n: int, seed: int = 1234) -> np.ndarray:
    """n x 32 uint8 uniform random 256-bit descriptors."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def _flip_bits(rows: np.ndarray, nflips: np.ndarray, rng) -> np.ndarray:
    out = rows.copy()
    bits = np.unpackbits(out, axis=1)
    for i in range(out.shape[0]):
        pos = rng.choice(256, size=int(nflips[i]), replace=False)
        bits[i, pos] ^= 1
    return np.packbits(bits, axis=1)


def query_set(map_desc: np.ndarray, nq: int = 2000, seed: int = 4321, plant_decoys: bool = True):
    """Queries = noisy copies of random map rows (Binomial(256,0.08) bit flips); for half of them a
    decoy at ~1.4x that distance is planted into the map (ratio-test boundary).  Returns
    (queries, map_with_decoys, source_rows)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = map_desc.copy()
    src = rng.choice(m.shape[0], size=nq, replace=False)
    flips = rng.binomial(256, 0.08, size=nq)
    q = _flip_bits(m[src], flips, rng)
    if plant_decoys:
        free = np.setdiff1d(np.arange(m.shape[0]), src, assume_unique=False)
        spots = rng.choice(free, size=nq // 2, replace=False)
        dflips = np.minimum(255, np.rint(flips[: nq // 2] * 1.4).astype(np.int64) + 1)
        m[spots] = _flip_bits(q[: nq // 2], dflips, rng)
    return np.ascontiguousarray(q), np.ascontiguousarray(m), src
