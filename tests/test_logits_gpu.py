"""N1 (SURVEY 8f / 8d "fused-sigmoid variant"): ctdet decode straight from the head's logits.

The reference computes `hm.sigmoid_()` (detectors/ctdet.py:31) and then `ctdet_decode(hm, ...)`.  The fused
path must give the same bits, so it is compared with this library's own heat-map path (pinned against the
oracle in test_decode_gpu.py) fed with torch's sigmoid of the same logits.  Two facts the kernel relies on
are proven here over EVERY float in the range where the sigmoid is not constant:
  * 1/(1+exp(-x)) as compiled into the library == torch's CUDA sigmoid, bit for bit;
  * it is monotone non-decreasing (3x3 max of the heat map == sigmoid of the 3x3 max of the logits)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _float_chunks():
    """Yield sorted-ascending CUDA float32 tensors covering all floats in [-105, 18]."""
    hi_pos = int(np.float32(18.0).view(np.int32))
    hi_neg = int(np.float32(105.0).view(np.int32))
    step = 1 << 26
    # negative side: bits 0x80000000 | m, m descending gives ascending values
    for start in range(hi_neg, -1, -step):
        lo = max(start - step + 1, 0)
        m = torch.arange(start, lo - 1, -1, dtype=torch.int64, device="cuda")
        yield ((m | 0x80000000) - (1 << 32)).to(torch.int32).view(torch.float32)   # -x, ascending
    for start in range(0, hi_pos + 1, step):
        m = torch.arange(start, min(start + step, hi_pos + 1), dtype=torch.int64, device="cuda")
        yield m.to(torch.int32).view(torch.float32)


def test_sigmoid_matches_torch_and_is_monotone_over_all_floats():
    from centernet_b200 import decode as D
    prev_last = None
    total = 0
    for x in _float_chunks():
        s = D.sigmoid(x)
        assert torch.equal(s, torch.sigmoid(x)), "library sigmoid differs from torch's CUDA sigmoid"
        assert bool((s[1:] >= s[:-1]).all()), "sigmoid not monotone inside a chunk"
        if prev_last is not None:
            assert float(s[0]) >= prev_last
        prev_last = float(s[-1])
        total += x.numel()
    assert total > 2_000_000_000
    edge = torch.tensor([-float("inf"), -200.0, -104.0, -88.8, -0.0, 0.0, 17.5, 40.0, float("inf")], device="cuda")
    assert torch.equal(D.sigmoid(edge), torch.sigmoid(edge))


def _check(logits, wh, reg, K=100, cat=False):
    from centernet_b200 import decode as D
    keep = logits.clone()
    got = D.ctdet_decode_from_logits(logits, wh, reg=reg, cat_spec_wh=cat, K=K)
    want = D.ctdet_decode(torch.sigmoid(logits), wh, reg=reg, cat_spec_wh=cat, K=K)
    assert torch.equal(logits, keep), "input modified"
    assert torch.equal(got, want)
    return got


@pytest.mark.parametrize("B,C,H,W,K", [(4, 80, 128, 128, 100), (64, 80, 128, 128, 100), (1, 1, 128, 128, 40),
                                       (2, 5, 40, 60, 33), (3, 7, 128, 128, 256)])
def test_logits_equals_sigmoid_then_decode(B, C, H, W, K):
    g = torch.Generator(device="cuda").manual_seed(317 + B)
    logits = torch.randn(B, C, H, W, device="cuda", generator=g) - 2.19      # head-bias prior, pose_dla_dcn.py:456
    wh = torch.rand(B, 2, H, W, device="cuda", generator=g) * 32
    reg = torch.rand(B, 2, H, W, device="cuda", generator=g)
    _check(logits, wh, reg, K)


def test_logits_saturated_and_colliding():
    """Logits whose sigmoids collide: saturation (everything above ~17 is exactly 1.0, huge plateaus),
    neighbours one ulp apart, and a map so negative that the heat underflows to 0 (zero fillers)."""
    B, C, H, W, K = 2, 80, 128, 128, 100
    g = torch.Generator(device="cuda").manual_seed(5)
    wh = torch.rand(B, 2, H, W, device="cuda", generator=g) * 8
    sat = torch.randn(B, C, H, W, device="cuda", generator=g) * 6 + 9          # many values >= 17 -> heat == 1.0
    _check(sat, wh, None, K)
    for centre in (-3.0, 0.5, 3.0, 6.0, 9.0, 11.5, 14.0):
        base = torch.full((B, C, H, W), centre, device="cuda")
        ulps = torch.randint(0, 6, (B, C, H, W), device="cuda", generator=g, dtype=torch.int32)
        near = (base.view(torch.int32) + ulps).view(torch.float32)           # neighbours 0..5 ulp apart
        _check(near.contiguous(), wh, None, K)
    dead = torch.full((B, C, H, W), -200.0, device="cuda")                   # heat == 0 everywhere: fillers only
    dead[0, 3, 5, 7] = 2.0
    _check(dead, wh, None, K)
    tiny = torch.randn(B, C, H, W, device="cuda", generator=g) * 0.5 - 88.0  # denormal heat values
    _check(tiny, wh, None, K)


def test_logits_overflow_rescan():
    """Planes that keep getting brighter: stale threshold, key-buffer overflow, exact rebuild from the logits."""
    B, C, H, W, K = 4, 80, 128, 128, 100
    g = torch.Generator(device="cuda").manual_seed(11)
    base = torch.rand(B, C, H, W, device="cuda", generator=g)
    logits = (base * 0.05 + torch.linspace(-6, 4, C, device="cuda").view(1, C, 1, 1)).contiguous()
    wh = torch.rand(B, 2, H, W, device="cuda", generator=g) * 8
    _check(logits, wh, None, K)


def test_logits_cat_spec_and_misaligned():
    B, C, H, W, K = 2, 4, 128, 128, 50
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(B, C, H, W, device="cuda", generator=g)
    wh = torch.rand(B, 2 * C, H, W, device="cuda", generator=g) * 8
    reg = torch.rand(B, 2, H, W, device="cuda", generator=g)
    _check(logits, wh, reg, K, cat=True)
    # misaligned base pointer: generic path (materialised heat map)
    buf = torch.randn(B * C * H * W + 1, device="cuda", generator=g)
    mis = buf[1:].view(B, C, H, W)
    assert mis.data_ptr() % 16 != 0
    _check(mis, wh[:, :2].contiguous(), reg, K)


def test_flip_merge_matches_reference_ops():
    """N1: flip-test merges (detectors/ctdet.py:34-37, multi_pose.py:43-51) against the reference's op sequence
    (sigmoid_, flip_tensor / flip_lr / flip_lr_off of models/utils.py, add, divide), bit for bit."""
    from centernet_b200 import decode as D, utils as U
    g = torch.Generator().manual_seed(5)
    flip_idx = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    hm = (torch.randn(4, 6, 12, 20, generator=g) * 3).cuda()
    hps = torch.randn(4, 34, 12, 20, generator=g).cuda()
    hm_hp = (torch.randn(4, 17, 12, 20, generator=g) * 2).cuda()
    s = torch.sigmoid(hm)
    assert torch.equal(D.flip_merge(hm, sigmoid=True), (s[0:2] + U.flip_tensor(s[2:4])) / 2)
    assert torch.equal(D.flip_merge(hps), (hps[0:2] + U.flip_tensor(hps[2:4])) / 2)
    assert torch.equal(D.flip_merge(hps, flip_idx=flip_idx, offsets=True), (hps[0:2] + U.flip_lr_off(hps[2:4], flip_idx)) / 2)
    sp = torch.sigmoid(hm_hp)
    assert torch.equal(D.flip_merge(hm_hp, sigmoid=True, flip_idx=flip_idx), (sp[0:2] + U.flip_lr(sp[2:4], flip_idx)) / 2)
    # whole flip-test decode == reference composition
    wh = (torch.rand(4, 2, 12, 20, generator=g) * 9).cuda(); reg = torch.rand(4, 2, 12, 20, generator=g).cuda()
    want = D.ctdet_decode((s[0:2] + U.flip_tensor(s[2:4])) / 2, (wh[0:2] + U.flip_tensor(wh[2:4])) / 2, reg=reg[0:2], K=15)
    assert torch.equal(D.ctdet_decode_flip_from_logits(hm, wh, reg=reg, K=15), want)
