"""Shared test helpers: golden loading and tie-aware detection comparison."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def assert_dets_equal(ref, got, score_col=4, atol=0.0, positive_only=False, what=""):
    """Detections [B, K, D] sorted by score desc.  Rows are compared as multisets inside
    every run of equal scores (torch.topk leaves that order open).  With positive_only
    rows whose reference score is <= 0 (non-peak fillers) are only required to have the
    same score."""
    ref = np.asarray(ref); got = np.asarray(got)
    assert ref.shape == got.shape, (what, ref.shape, got.shape)
    for b in range(ref.shape[0]):
        rs, gs = ref[b, :, score_col], got[b, :, score_col]
        np.testing.assert_array_equal(rs, gs, err_msg="%s: scores differ (image %d)" % (what, b))
        k = 0
        K = ref.shape[1]
        while k < K:
            e = k
            while e + 1 < K and rs[e + 1] == rs[k]:
                e += 1
            if positive_only and rs[k] <= 0:
                k = e + 1
                continue
            a = ref[b, k:e + 1]; c = got[b, k:e + 1]
            if e > k:  # tie group: order-free
                a = a[np.lexsort(a.T[::-1])]; c = c[np.lexsort(c.T[::-1])]
            if atol == 0.0:
                np.testing.assert_array_equal(a, c, err_msg="%s: rows %d..%d image %d" % (what, k, e, b))
            else:
                np.testing.assert_allclose(a, c, rtol=0, atol=atol, err_msg="%s: rows %d..%d image %d" % (what, k, e, b))
            k = e + 1
