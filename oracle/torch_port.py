"""CPU port of the reference decode path using the SAME ATen operator sequence the
reference executes (max_pool2d, ==, *, topk x2, gathers, cat) -- TEST INFRASTRUCTURE,
used only as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`)
and cross-checked against the numpy oracle in tests.  /root/reference cannot travel to
the GPU box, so this restatement stands in for "the reference's own PyTorch CPU path";
follows models/decode.py:9-15 (_nms), :103-119 (_topk), :464-495 (ctdet_decode) and
models/utils.py:12-26."""
import torch
import torch.nn.functional as F


def peak_mask_mul(heat):                                   # decode.py:9-15
    pooled = F.max_pool2d(heat, (3, 3), stride=1, padding=1)
    return heat * (pooled == heat).float()


def two_stage_topk(scores, K):                             # decode.py:103-119
    B, C, H, W = scores.shape
    s1, i1 = torch.topk(scores.view(B, C, -1), K)
    i1 = i1 % (H * W)
    y1 = torch.div(i1, W, rounding_mode="floor").float()
    x1 = (i1 % W).float()
    s2, i2 = torch.topk(s1.view(B, -1), K)
    cls = torch.div(i2, K, rounding_mode="floor").int()
    take = lambda a: a.view(B, -1).gather(1, i2)
    return s2, take(i1), cls, take(y1), take(x1)


def gather_nchw(feat, ind):                                # utils.py:12-26
    B, D = feat.shape[:2]
    f = feat.permute(0, 2, 3, 1).contiguous().view(B, -1, D)
    return f.gather(1, ind.unsqueeze(2).expand(B, ind.size(1), D))


def ctdet_decode(heat, wh, reg=None, cat_spec_wh=False, K=100):   # decode.py:464-495
    B, C, H, W = heat.shape
    scores, inds, clses, ys, xs = two_stage_topk(peak_mask_mul(heat), K)
    if reg is not None:
        r = gather_nchw(reg, inds)
        xs = xs.view(B, K, 1) + r[:, :, 0:1]
        ys = ys.view(B, K, 1) + r[:, :, 1:2]
    else:
        xs = xs.view(B, K, 1) + 0.5
        ys = ys.view(B, K, 1) + 0.5
    w = gather_nchw(wh, inds)
    if cat_spec_wh:
        w = w.view(B, K, C, 2).gather(2, clses.view(B, K, 1, 1).expand(B, K, 1, 2).long()).view(B, K, 2)
    boxes = torch.cat([xs - w[..., 0:1] / 2, ys - w[..., 1:2] / 2, xs + w[..., 0:1] / 2, ys + w[..., 1:2] / 2], dim=2)
    return torch.cat([boxes, scores.view(B, K, 1), clses.view(B, K, 1).float()], dim=2)
