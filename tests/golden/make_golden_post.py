"""Generates tests/golden/post.npz by running the UNMODIFIED reference post-processing (build container only):
utils/post_process.py (ctdet_post_process, multi_pose_post_process) imported from /root/reference/src/lib and
external/nms.pyx (soft_nms, soft_nms_39) compiled by oracle/build_ref.py into oracle/_ref/.
    python tests/golden/make_golden_post.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/src/lib"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from utils import post_process as P  # noqa: E402
from oracle import build_ref  # noqa: E402

nms = build_ref.load_nms()
assert nms is not None


def boxes(rng, n, extra=0, size=512.0):
    x = rng.uniform(0, size - 40, n); y = rng.uniform(0, size - 40, n)
    # clusters of overlapping boxes so that the Gaussian decay and the discard path both trigger
    k = n // 4
    x[:k * 3] = np.repeat(x[:k], 3) + rng.uniform(-6, 6, k * 3)
    y[:k * 3] = np.repeat(y[:k], 3) + rng.uniform(-6, 6, k * 3)
    w = rng.uniform(20, 120, n); h = rng.uniform(20, 120, n)
    s = rng.uniform(0.0005, 1.0, n)
    cols = [x, y, x + w, y + h, s] + [rng.uniform(0, size, n) for _ in range(extra)]
    return np.stack(cols, 1).astype(np.float32)


def main():
    rng = np.random.default_rng(2718)
    out = {}
    # ---- ctdet_post_process / multi_pose_post_process: two images, non-square originals, scalar and vector scale
    B, K, C = 2, 100, 80
    dets = np.zeros((B, K, 6), np.float32)
    dets[..., 0] = rng.uniform(-4, 120, (B, K)); dets[..., 1] = rng.uniform(-4, 120, (B, K))
    dets[..., 2] = dets[..., 0] + rng.uniform(1, 60, (B, K)); dets[..., 3] = dets[..., 1] + rng.uniform(1, 60, (B, K))
    dets[..., 4] = np.sort(rng.uniform(0, 1, (B, K)), axis=1)[:, ::-1]
    dets[..., 5] = rng.integers(0, C, (B, K))
    c = np.array([[320.0, 213.5], [250.0, 187.0]], np.float32)
    s = np.array([672.0, 512.0], np.float32)             # max(h, w) style scalar scales (base_detector.py:52-54)
    res = P.ctdet_post_process(dets.copy(), [c[0], c[1]], [s[0], s[1]], 128, 128, C)
    flat = []
    for i in range(B):
        for j in range(1, C + 1):
            a = np.array(res[i][j], np.float32).reshape(-1, 5)
            flat.append(np.concatenate([np.full((len(a), 1), i, np.float32), np.full((len(a), 1), j, np.float32), a], 1))
    out.update(ct_dets=dets, ct_c=c, ct_s=s, ct_rows=np.concatenate(flat, 0))
    s2 = np.array([[640.0, 480.0], [512.0, 384.0]], np.float32)   # --keep_res style vector scale (:44-47)
    res2 = P.ctdet_post_process(dets.copy(), [c[0], c[1]], [s2[0], s2[1]], 120, 160, C)
    flat = []
    for i in range(B):
        for j in range(1, C + 1):
            a = np.array(res2[i][j], np.float32).reshape(-1, 5)
            flat.append(np.concatenate([np.full((len(a), 1), i, np.float32), np.full((len(a), 1), j, np.float32), a], 1))
    out.update(ct_s2=s2, ct_rows2=np.concatenate(flat, 0))
    mp = np.zeros((B, K, 40), np.float32)
    mp[..., :4] = dets[..., :4]; mp[..., 4] = dets[..., 4]
    mp[..., 5:39] = rng.uniform(-10, 130, (B, K, 34))
    resm = P.multi_pose_post_process(mp.copy(), [c[0], c[1]], [s[0], s[1]], 128, 128)
    out.update(mp_dets=mp, mp_rows=np.stack([np.array(resm[i][1], np.float32) for i in range(B)], 0))

    # ---- soft_nms / soft_nms_39, all three methods, in-place arrays + returned N
    for name, n, method, Nt in (("g", 60, 2, 0.5), ("lin", 45, 1, 0.3), ("hard", 50, 0, 0.3), ("g1", 1, 2, 0.5),
                                ("gbig", 300, 2, 0.5)):
        b = boxes(rng, n)
        a = b.copy()
        keep = nms.soft_nms(a, Nt=Nt, method=method)
        out["nms_%s_in" % name] = b; out["nms_%s_out" % name] = a; out["nms_%s_n" % name] = len(keep)
        out["nms_%s_cfg" % name] = np.array([0.5, Nt, 0.001, method], np.float32)
    b = boxes(rng, 80, extra=34)
    a = b.copy()
    keep = nms.soft_nms_39(a, Nt=0.5, method=2)
    out.update(nms39_in=b, nms39_out=a, nms39_n=len(keep))

    # ---- merge_outputs (detectors/ctdet.py:76-92) over two scales, 5 classes, max_per_image 100
    NC = 5
    scales = [{j: boxes(rng, int(rng.integers(0, 60))) for j in range(1, NC + 1)} for _ in range(2)]
    results = {}
    for j in range(1, NC + 1):
        results[j] = np.concatenate([d[j] for d in scales], axis=0).astype(np.float32)
        nms.soft_nms(results[j], Nt=0.5, method=2)
    scores = np.hstack([results[j][:, 4] for j in range(1, NC + 1)])
    kth = len(scores) - 100
    assert kth > 0
    thresh = np.partition(scores, kth)[kth]
    for j in range(1, NC + 1):
        results[j] = results[j][results[j][:, 4] >= thresh]
    for si, d in enumerate(scales):
        for j in range(1, NC + 1):
            out["mg_s%d_c%d" % (si, j)] = d[j]
    for j in range(1, NC + 1):
        out["mg_out_c%d" % j] = results[j]
    np.savez_compressed(os.path.join(HERE, "post.npz"), **out)
    print("wrote post.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
