"""Synthetic frame generators (numpy), byte-identical to adder_codec_rs_b200/csrc/synth.cuh.

h(seed, f, i) = splitmix64(seed ^ (f << 40) ^ i) & 0xFF with i the flat raster index (y*W + x)*C + c.
kinds: 0 gradient, 1 uniform noise, 2 base +-10 jitter, 3 static base with one-frame blips (p = 2/256).
"""
import numpy as np

GRADIENT, NOISE, JITTER, STATIC_BLIPS = 0, 1, 2, 3
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _hash(seed: int, f: int, i: np.ndarray) -> np.ndarray:
    key = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.uint64(f) << np.uint64(40)) ^ i
    return (splitmix64(key) & np.uint64(0xFF)).astype(np.int32)


def frame(kind: int, seed: int, f: int, w: int, h: int, c: int) -> np.ndarray:
    """One (H, W, C) u8 frame."""
    n = w * h * c
    i = np.arange(n, dtype=np.uint64)
    if kind == GRADIENT:
        p = np.arange(n, dtype=np.int64) // c
        v = ((p % w) + 2 * (p // w) + 3 * f) & 255
    elif kind == NOISE:
        v = _hash(seed, f, i)
    elif kind == JITTER:
        v = np.clip(_hash(seed ^ 1, 0, i) + (_hash(seed ^ 2, f, i) % 21) - 10, 0, 255)
    elif kind == STATIC_BLIPS:
        v = np.where(_hash(seed ^ 3, f, i) < 2, _hash(seed ^ 4, f, i), _hash(seed ^ 1, 0, i))
    else:
        raise ValueError(kind)
    return v.astype(np.uint8).reshape(h, w, c)


def frames(kind: int, seed: int, f0: int, n: int, w: int, h: int, c: int) -> np.ndarray:
    return np.stack([frame(kind, seed, f0 + k, w, h, c) for k in range(n)])


def moving_blocks(seed: int, n: int, w: int, h: int, c: int) -> np.ndarray:
    """(n, H, W, C) u8: a dim textured background with bright and dark rectangles that drift a pixel per frame —
    corners come and go, which is what the feature pass of the transcoder reacts to."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(60, 90, (h, w, c)).astype(np.int32)
    rects = [(int(rng.integers(0, w - 12)), int(rng.integers(0, h - 10)), int(rng.integers(6, 12)), int(rng.integers(5, 10)),
              int(rng.choice([-55, 70, 120])), int(rng.choice([-1, 1])), int(rng.choice([-1, 0, 1]))) for _ in range(max(3, w * h // 900))]
    out = np.empty((n, h, w, c), dtype=np.uint8)
    for f in range(n):
        img = bg.copy()
        for (x0, y0, rw, rh, dv, vx, vy) in rects:
            x = (x0 + vx * f) % (w - rw)
            y = (y0 + vy * f) % (h - rh)
            img[y:y + rh, x:x + rw] += dv
        out[f] = np.clip(img + rng.integers(-2, 3, (h, w, c)), 0, 255).astype(np.uint8)
    return out
