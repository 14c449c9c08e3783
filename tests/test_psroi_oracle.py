"""CPU: the PSROI-pooling oracle against the reference's own known-answer test (DCNv2/test.py:117-146) and
against finite differences of itself (the reference kernels are CUDA-only: parity otherwise unpinned)."""
import numpy as np
from oracle import psroi_np as P


def _kat_inputs():
    inp = np.zeros((2, 16, 64, 64), np.float32)
    inp[0, :, 16:26, 16:26] = 1.0
    inp[1, :, 10:20, 20:30] = 2.0
    rois = np.array([[0, 65, 65, 103, 103], [1, 81, 41, 119, 79]], np.float32)
    return inp, rois


def test_zero_offset_equals_no_trans_and_block_means():
    inp, rois = _kat_inputs()
    cfg = dict(spatial_scale=0.25, output_dim=16, group_size=1, pooled_size=7, part_size=7, sample_per_part=4, trans_std=0.1)
    o, c = P.psroi_forward(inp, rois, None, True, **cfg)
    o2, c2 = P.psroi_forward(inp, rois, np.zeros((2, 2, 7, 7), np.float32), False, **cfg)
    assert np.array_equal(o, o2) and np.array_equal(c, c2)           # test.py:136-146: same numbers printed twice
    assert c.min() == 16 and c.max() == 16                          # every sample lands inside the map
    # roi 0 covers [15.75, 25.5) of a block of ones on [16, 26): the mean is just below 1; roi 1 sees the 2-block
    assert 0.95 < o[0].mean() < 1.0 and abs(o[1].mean() - 2 * o[0].mean()) < 1e-6
    assert np.allclose(o[0, 0], o[0, 7]) and o.shape == (2, 16, 7, 7)


def test_backward_matches_finite_differences():
    rng = np.random.default_rng(0)
    data = rng.standard_normal((2, 8, 9, 9)).astype(np.float32)
    rois = np.array([[0, 3.2, 4.1, 22.7, 25.3], [1, 8, 2, 30, 18]], np.float32)
    trans = (rng.standard_normal((2, 4, 3, 3)) * 0.5).astype(np.float32)
    cfg = dict(no_trans=False, spatial_scale=0.25, output_dim=2, group_size=2, pooled_size=3, part_size=3,
               sample_per_part=2, trans_std=0.1)
    o, c = P.psroi_forward(data, rois, trans, **cfg)
    go = rng.standard_normal(o.shape).astype(np.float32)
    gd, gt = P.psroi_backward(go, data, rois, trans, c, **cfg)
    eps = 1e-2
    for idx in [(0, 1, 3, 4), (1, 5, 2, 6), (0, 7, 5, 5)]:
        d2 = data.copy(); d2[idx] += eps
        o2, _ = P.psroi_forward(d2, rois, trans, **cfg)
        assert abs(((o2.astype(np.float64) - o) * go).sum() / eps - gd[idx]) < 2e-4
    for idx in [(0, 0, 1, 1), (0, 1, 2, 0), (1, 2, 0, 2)]:
        tp, tm = trans.copy(), trans.copy()
        tp[idx] += eps; tm[idx] -= eps
        num = ((P.psroi_forward(data, rois, tp, **cfg)[0].astype(np.float64) - P.psroi_forward(data, rois, tm, **cfg)[0]) * go).sum() / (2 * eps)
        assert abs(num - gt[idx]) < 2e-4
