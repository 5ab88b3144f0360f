"""Deterministic synthetic networks, frames and lazy-output masks.

The reference ships no network file (its tests name data/dnn.extended.tv.model, absent —
/root/reference/test/java/suskun/nn/FuncTest.java:168), so every network here is synthetic.
Values come from a counter-based integer hash (splitmix64) turned into an Irwin–Hall(8)
approximate normal with exact integer arithmetic and one IEEE division, so the same bytes are
produced on every machine and numpy version (the golden vectors in tests/golden depend on it).
"""
from __future__ import annotations

import os
import numpy as np

from . import formats

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(idx: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def uniform01(count: int, seed: int, start: int = 0) -> np.ndarray:
    """float64 in [0,1) with 32 bits of resolution."""
    h = _splitmix(np.arange(start, start + count, dtype=np.uint64), seed)
    return (h >> np.uint64(32)).astype(np.float64) / 4294967296.0


def normal(count: int, seed: int, start: int = 0) -> np.ndarray:
    """≈N(0,1) float64 (Irwin–Hall of eight 16-bit uniforms; |z| ≤ 4.9)."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    total = np.zeros(count, dtype=np.int64)
    for s in (seed * 2 + 1, seed * 2 + 2):
        h = _splitmix(idx, s)
        for sh in (0, 16, 32, 48):
            total += ((h >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64)
    # mean 8*32767.5 = 262140, var 8*(65536^2-1)/12
    return (total.astype(np.float64) - 262140.0) / 53509.91992145008


# name → (input dim, hidden width, number of hidden layers, outputs)
SHAPES = {
    "tiny": (12, 32, 3, 20),       # smallest legal network: ≥2 int8 layers, hidden % 16 == 0
    "ragged": (24, 48, 3, 37),     # O not a multiple of anything
    "P": (432, 512, 4, 2000),      # config 0: shipped 432-dim features
    "S": (440, 512, 4, 2000),      # config 1
    "L": (440, 2048, 7, 8000),     # config 2/3/4 (headline)
}


def make_network(shape, seed: int = 1234, stress: bool = False):
    """→ (layers, shift, scale).  shape = name in SHAPES or (I, H, n_hidden, O).

    stress=True gives weights of std 0.9 (many beyond the ±3 cutoff → exercises the missing upper
    clip / int8 wrap-around of dnn.cc:493-499 and makes pmaddubsw saturation fire densely).
    """
    I, H, nh, O = SHAPES[shape] if isinstance(shape, str) else shape
    dims = [I] + [H] * nh + [O]
    layers = []
    for j in range(len(dims) - 1):
        fan_in, fan_out = dims[j], dims[j + 1]
        if stress and j > 0:
            sigma = 0.9
        else:
            sigma = 0.05 if j == 0 else 1.5 / np.sqrt(fan_in)
        w = (normal(fan_in * fan_out, seed * 1000 + 10 * j) * sigma).astype(np.float32).reshape(fan_out, fan_in)
        b = (normal(fan_out, seed * 1000 + 10 * j + 1) * 0.1).astype(np.float32)
        layers.append((w, b))
    shift = (normal(I, seed * 1000 + 901) * 0.1).astype(np.float32)
    scale = (0.05 + 0.05 * uniform01(I, seed * 1000 + 902)).astype(np.float32)
    return layers, shift, scale


def make_frames(n: int, dim: int, seed: int = 7, start: int = 0) -> np.ndarray:
    """N(0, 15²) fp32 frames (matches the shipped features' spread); rows [start, start+n) of an
    unbounded seeded stream, so a 1M-frame stream can be produced chunk by chunk."""
    return (normal(n * dim, seed, start * dim) * 15.0).astype(np.float32).reshape(n, dim)


def make_masks(count: int, dimension: int, ratio: float = 0.40, drift: float = 0.03, seed: int = 11) -> np.ndarray:
    """Lazy-output masks following FuncTest.generateMasks (FuncTest.java:121-154): frame 0 has
    ⌊ratio·O⌋ ones; frame t = frame t−1 with ⌊drift·O⌋ entries set to 1 then ⌊drift·O⌋ set to 0,
    each pick changing an entry that does not already hold the value.  Seeded (the reference's
    generator is not)."""
    active, fresh = int(dimension * ratio), int(dimension * drift)
    rng_state = [0]

    def set_random(row, k, val):
        cnt = 0
        while cnt < k:
            draw = _splitmix(np.arange(rng_state[0], rng_state[0] + 4 * k + 16, dtype=np.uint64), seed)
            rng_state[0] += len(draw)
            for pos in (draw % np.uint64(dimension)).astype(np.int64):
                if row[pos] != val:
                    row[pos] = val
                    cnt += 1
                    if cnt == k:
                        break

    masks = np.zeros((count, dimension), dtype=np.int8)
    set_random(masks[0], active, 1)
    for t in range(1, count):
        masks[t] = masks[t - 1]
        set_random(masks[t], fresh, 1)
        set_random(masks[t], fresh, 0)
    return masks


def network_file(shape, seed: int = 1234, stress: bool = False, cache_dir: str | None = None) -> str:
    """Writes (once) the synthetic network as dnn.bin under cache_dir and returns the path."""
    cache_dir = cache_dir or os.environ.get("FDNN_CACHE", "/tmp/fdnn_cache")
    os.makedirs(cache_dir, exist_ok=True)
    tag = shape if isinstance(shape, str) else "x".join(str(v) for v in shape)
    path = os.path.join(cache_dir, f"net_{tag}_s{seed}{'_stress' if stress else ''}.dnn.bin")
    if not os.path.exists(path):
        layers, shift, scale = make_network(shape, seed, stress)
        tmp = f"{path}.{os.getpid()}.tmp"
        formats.write_dnn_bin(tmp, layers, shift, scale)
        os.replace(tmp, path)
    return path


def make_hostile_frames(n: int, dim: int, seed: int = 31) -> np.ndarray:
    """Ordinary frames with the rows the reference was never written for mixed in: NaN, ±inf, huge, tiny, all-zero and constant rows,
    single hostile elements (what happens to them is defined by x86 arithmetic, dnn.h:35-42: NaN and out-of-range products convert
    to INT_MIN and land in bucket 0)."""
    x = make_frames(n, dim, seed=seed)
    if n >= 24:
        x[3, 7 % dim] = np.nan
        x[5, :] = 0.0
        x[9, 100 % dim] = np.inf
        x[11, 5 % dim] = -np.inf
        x[13, :] = 1e30
        x[15, :] = -1e30
        x[17, :] = 1e-30
        x[19, :] = 3.0
        x[21, 0] = 3e38
        x[23, :] = np.nan
    return x
