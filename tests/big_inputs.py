"""Seeded inputs at the north-star geometry (80x128x128, K=100) for the reference-generated goldens
tests/golden/big_*.npz.  The inputs (tens of MB) are NOT stored: they are rebuilt here, identically
on every machine, from numpy's PCG64 stream with IEEE-exact arithmetic only (products, sums, max,
division, comparisons -- no exp/pow/sigmoid, whose last bit may depend on the SIMD path), and the
golden files keep only what the unmodified reference returned for them plus a checksum of the
inputs (tests/golden/make_golden_big.py)."""
import zlib

import numpy as np

F32 = np.float32


def checksum(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8), c)
    return np.uint32(c)


def skewed_heat(rng, shape):
    """Head-like score map: u^4 * (0.25 + 0.75 v) of two uniforms (mean ~0.12, long tail to 1), exact
    float32 products; the second factor spreads the top scores over many mantissas (no exact ties)."""
    u = rng.random(shape, dtype=F32)
    v = rng.random(shape, dtype=F32)
    u2 = u * u
    return ((u2 * u2) * (F32(0.25) + F32(0.75) * v)).astype(F32)


def blob_heat(rng, B, C, H, W, n_obj):
    """Class-sparse map of <=n_obj flat-topped pyramids per image (plateaus, < K peaks per class) x 0.9
    + 0.01 noise -- the second input set of SURVEY 8(d), without transcendental functions."""
    hm = np.zeros((B, C, H, W), F32)
    ys = np.arange(H, dtype=F32).reshape(H, 1); xs = np.arange(W, dtype=F32).reshape(1, W)
    for b in range(B):
        for _ in range(n_obj):
            c = int(rng.integers(0, C)); r = F32(rng.integers(2, 24))
            cy = F32(rng.integers(0, H)); cx = F32(rng.integers(0, W))
            d = np.maximum(np.abs(ys - cy), np.abs(xs - cx))
            v = np.minimum(F32(1), F32(1.25) * np.maximum(F32(0), F32(1) - d / r)).astype(F32)
            np.maximum(hm[b, c], v, out=hm[b, c])
    return (hm * F32(0.9) + F32(0.01) * rng.random((B, C, H, W), dtype=F32)).astype(F32)


def ctdet_inputs(kind, seed=317, B=2, C=80, H=128, W=128):
    rng = np.random.default_rng(seed)
    heat = skewed_heat(rng, (B, C, H, W)) if kind == "noise" else blob_heat(rng, B, C, H, W, 40)
    wh = rng.random((B, 2, H, W), dtype=F32) * F32(32)
    reg = rng.random((B, 2, H, W), dtype=F32)
    return heat, wh, reg


def multi_pose_inputs(seed=318, B=2, H=128, W=128, J=17):
    rng = np.random.default_rng(seed)
    heat = skewed_heat(rng, (B, 1, H, W))
    wh = rng.random((B, 2, H, W), dtype=F32) * F32(40)
    kps = (rng.random((B, 2 * J, H, W), dtype=F32) - F32(0.5)) * F32(24)
    reg = rng.random((B, 2, H, W), dtype=F32)
    hm_hp = skewed_heat(rng, (B, J, H, W))
    hp_offset = rng.random((B, 2, H, W), dtype=F32)
    return heat, wh, kps, reg, hm_hp, hp_offset


def exct_inputs(seed=319, B=1, C=80, H=128, W=128, n_obj=24):
    """t/l/b/r/centre maps with boxes whose extreme points and centre carry peaks of one class, over a
    low noise floor (so that positive-score tuples exist), + the four regression maps."""
    rng = np.random.default_rng(seed)
    base = (F32(0.03) * rng.random((5, B, C, H, W), dtype=F32)).astype(F32)
    for b in range(B):
        for _ in range(n_obj):
            c = int(rng.integers(0, C))
            x0, x1 = sorted(int(v) for v in rng.integers(2, W - 2, 2))
            y0, y1 = sorted(int(v) for v in rng.integers(2, H - 2, 2))
            if x1 - x0 < 4 or y1 - y0 < 4:
                continue
            tx, bx = [int(v) for v in rng.integers(x0, x1 + 1, 2)]
            ly, ry = [int(v) for v in rng.integers(y0, y1 + 1, 2)]
            sc = (F32(0.3) + F32(0.65) * rng.random(5, dtype=F32)).astype(F32)
            base[0, b, c, y0, tx] = sc[0]; base[1, b, c, ly, x0] = sc[1]
            base[2, b, c, y1, bx] = sc[2]; base[3, b, c, ry, x1] = sc[3]
            base[4, b, c, (y0 + y1) // 2, (x0 + x1) // 2] = sc[4]
    regs = [rng.random((B, 2, H, W), dtype=F32) for _ in range(4)]
    return [base[i] for i in range(5)], regs


PRE_CASES = {   # name: (height, width, fix_res, flip_test, input_h, input_w, test scale)
    "a": (375, 500, True, False, 512, 512, 1), "b": (640, 427, True, True, 512, 512, 1),
    "c": (213, 317, False, False, 0, 0, 1), "d": (480, 640, True, False, 384, 640, 1),
    "e": (375, 500, True, False, 512, 512, 0.5), "f": (213, 317, False, True, 0, 0, 1.5),
    "g": (300, 400, True, False, 512, 512, 1.25),
}


def pre_image(name):
    """Seeded uint8 BGR test image (smooth ramps + noise: every interpolation phase occurs), integer arithmetic only."""
    h, w = PRE_CASES[name][:2]
    rng = np.random.default_rng(1000 + ord(name))
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 3 + yy) % 256, (xx + yy * 2) % 256, (xx * yy // 7) % 256], 2)
    return ((base * 3) // 5 + rng.integers(0, 100, (h, w, 3))).astype(np.uint8)
