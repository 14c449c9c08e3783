"""Generates tests/golden/*.npz by running the UNMODIFIED reference functions
(imported from /root/reference/src/lib, CPU, torch) on seeded synthetic inputs.

Run in the build container only:  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed .npz files.

Every case stores its inputs and the reference outputs.  Cases whose result depends
on torch.topk's order of equal scores are made tie-free (checked here) except the
ones named *_ties, which tests compare tie-aware.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src/lib"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from models import decode as R  # noqa: E402
from models import losses as L  # noqa: E402
from models import utils as U  # noqa: E402
from utils import image as I  # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def noise_heat(g, B, C, H, W, bias=2.19):
    return torch.sigmoid(torch.randn(B, C, H, W, generator=g) - bias)


def blob_heat(rng, B, C, H, W, n_obj):
    """Gaussian-blob maps built with the reference's own splat (utils/image.py) * 0.9 + noise."""
    hm = np.zeros((B, C, H, W), np.float32)
    for b in range(B):
        for _ in range(n_obj):
            c = rng.integers(0, C)
            h, w = rng.uniform(2, H / 2), rng.uniform(2, W / 2)
            r = max(0, int(I.gaussian_radius((np.ceil(h), np.ceil(w)))))
            I.draw_umich_gaussian(hm[b, c], (rng.integers(0, W), rng.integers(0, H)), r)
    return (hm * 0.9 + 0.01 * rng.random((B, C, H, W), dtype=np.float32)).astype(np.float32)


def assert_tie_free(scores, what):
    s = np.asarray(scores)
    for row in s.reshape(-1, s.shape[-1]):
        pos = row[row > 0]
        assert len(np.unique(pos)) == len(pos), "ties in " + what


def main():
    g = torch.Generator().manual_seed(317)  # reference default seed, opts.py:43
    rng = np.random.default_rng(317)

    # ---- A1/A2/A3 raw ops
    heat = noise_heat(g, 2, 6, 24, 40)
    nms = R._nms(heat)
    s, inds, clses, ys, xs = R._topk(nms, K=15)
    assert_tie_free(s, "topk")
    cs, cinds, cys, cxs = R._topk_channel(nms, K=15)
    save("raw_ops", heat=heat, nms=nms, topk_scores=s, topk_inds=inds, topk_clses=clses, topk_ys=ys, topk_xs=xs,
         ch_scores=cs, ch_inds=cinds, ch_ys=cys, ch_xs=cxs, K=15)

    # ---- A5 ctdet
    B, C, H, W, K = 2, 16, 32, 32, 20
    heat = noise_heat(g, B, C, H, W)
    wh = torch.rand(B, 2, H, W, generator=g) * 32
    reg = torch.rand(B, 2, H, W, generator=g)
    out = R.ctdet_decode(heat, wh, reg=reg, K=K)
    assert_tie_free(out[..., 4], "ctdet_noise")
    out_noreg = R.ctdet_decode(heat, wh, reg=None, K=K)
    save("ctdet_noise", heat=heat, wh=wh, reg=reg, dets=out, dets_noreg=out_noreg, K=K)

    heat = t(blob_heat(rng, 2, 8, 32, 48, 12))
    whc = torch.rand(2, 16, 32, 48, generator=g) * 20
    reg = torch.rand(2, 2, 32, 48, generator=g)
    out = R.ctdet_decode(heat, whc, reg=reg, cat_spec_wh=True, K=30)
    assert_tie_free(out[..., 4], "ctdet_blobs")
    save("ctdet_blobs_catspec", heat=heat, wh=whc, reg=reg, dets=out, K=30)

    # few peaks (< K positive): 3 isolated peaks incl. an exact tie
    heat = torch.zeros(1, 2, 8, 8)
    heat[0, 0, 2, 2] = 0.5; heat[0, 0, 2, 4] = 0.5; heat[0, 1, 4, 4] = 0.5; heat[0, 1, 6, 1] = 0.7
    wh = torch.rand(1, 2, 8, 8, generator=g); reg = torch.rand(1, 2, 8, 8, generator=g)
    out = R.ctdet_decode(heat, wh, reg=reg, K=6)
    save("ctdet_ties", heat=heat, wh=wh, reg=reg, dets=out, K=6)

    # ---- A6 multi_pose
    B, H, W, K, J = 2, 32, 32, 20, 17
    heat = noise_heat(g, B, 1, H, W)
    wh = torch.rand(B, 2, H, W, generator=g) * 24
    kps = torch.randn(B, 2 * J, H, W, generator=g) * 6
    reg = torch.rand(B, 2, H, W, generator=g)
    hm_hp = noise_heat(g, B, J, H, W, bias=1.0)
    hp_off = torch.rand(B, 2, H, W, generator=g)
    full = R.multi_pose_decode(heat, wh, kps, reg=reg, hm_hp=hm_hp, hp_offset=hp_off, K=K)
    assert_tie_free(full[..., 4], "multi_pose")
    nooff = R.multi_pose_decode(heat, wh, kps, reg=None, hm_hp=hm_hp, hp_offset=None, K=K)
    nohp = R.multi_pose_decode(heat, wh, kps, reg=reg, K=K)
    save("multi_pose", heat=heat, wh=wh, kps=kps, reg=reg, hm_hp=hm_hp, hp_offset=hp_off, dets=full,
         dets_nooff=nooff, dets_nohp=nohp, K=K)

    # ---- A9 ddd
    B, C, H, W, K = 1, 3, 24, 80, 10
    heat = noise_heat(g, B, C, H, W)
    rot = torch.randn(B, 8, H, W, generator=g); depth = torch.randn(B, 1, H, W, generator=g)
    dim = torch.randn(B, 3, H, W, generator=g); wh = torch.rand(B, 2, H, W, generator=g)
    reg = torch.rand(B, 2, H, W, generator=g)
    save("ddd", heat=heat, rot=rot, depth=depth, dim=dim, wh=wh, reg=reg,
         dets=R.ddd_decode(heat, rot, depth, dim, wh=wh, reg=reg, K=K),
         dets_min=R.ddd_decode(heat, rot, depth, dim, K=K), K=K)

    # ---- A7 edge aggregation
    heat = noise_heat(g, 2, 3, 16, 24, bias=0.5)
    save("aggregate", heat=heat, left=R._left_aggregate(heat), right=R._right_aggregate(heat),
         top=R._top_aggregate(heat), bottom=R._bottom_aggregate(heat),
         h=R._h_aggregate(heat, 0.1).contiguous(), v=R._v_aggregate(heat, 0.1).contiguous())

    # ---- A8 exct / agnex: structured maps so that valid (positive-score) boxes exist
    B, C, H, W, K, ND = 1, 4, 32, 32, 8, 60
    base = 0.02 * rng.random((5, B, C, H, W), dtype=np.float32)
    for _ in range(6):
        c = rng.integers(0, C)
        x0, x1 = sorted(rng.integers(2, W - 2, 2)); y0, y1 = sorted(rng.integers(2, H - 2, 2))
        if x1 - x0 < 4 or y1 - y0 < 4:
            continue
        tx, bx = rng.integers(x0, x1 + 1, 2); ly, ry = rng.integers(y0, y1 + 1, 2)
        sc = rng.uniform(0.3, 0.95, 5).astype(np.float32)
        base[0, 0, c, y0, tx] = sc[0]; base[1, 0, c, ly, x0] = sc[1]
        base[2, 0, c, y1, bx] = sc[2]; base[3, 0, c, ry, x1] = sc[3]
        base[4, 0, c, (y0 + y1) // 2, (x0 + x1) // 2] = sc[4]
        base[4, 0, c, (y0 + y1 + 1) // 2, (x0 + x1 + 1) // 2] = sc[4] * 0.9
    th, lh, bh, rh, ch = [t(base[i]) for i in range(5)]
    regs = [torch.rand(B, 2, H, W, generator=g) for _ in range(4)]
    ex = R.exct_decode(th, lh, bh, rh, ch, *regs, K=K, num_dets=ND)
    ex_noreg = R.exct_decode(th, lh, bh, rh, ch, K=K, num_dets=ND)
    assert (ex[..., 4] > 0).sum() > 0, "exct golden has no positive boxes"
    orig_topk = R._topk
    R._topk = lambda s_, K=40: orig_topk(s_.contiguous(), K=K)   # aggr path needs .contiguous()
    ex_aggr = R.exct_decode(th, lh, bh, rh, ch, *regs, K=K, num_dets=ND, aggr_weight=0.1)
    R._topk = orig_topk
    t1, l1, b1, r1 = [x.max(dim=1, keepdim=True)[0] for x in (th, lh, bh, rh)]
    ag = R.agnex_ct_decode(t1, l1, b1, r1, ch, *regs, K=K, num_dets=ND)
    save("exct", t=th, l=lh, b=bh, r=rh, ct=ch, t_regr=regs[0], l_regr=regs[1], b_regr=regs[2], r_regr=regs[3],
         dets=ex, dets_noreg=ex_noreg, dets_aggr=ex_aggr, t1=t1, l1=l1, b1=b1, r1=r1, dets_agn=ag, K=K,
         num_dets=ND)

    # ---- A13/A14/A16 losses
    B, C, H, W, M = 2, 5, 16, 16, 12
    logits = torch.randn(B, C, H, W, generator=g) - 2
    gt = t(blob_heat(rng, B, C, H, W, 5) / 0.9)
    gt = torch.clamp(gt - 0.011, 0, 1)
    gt[gt > 0.985] = 1.0
    pred = U._sigmoid(logits.clone())
    pr = pred.clone().requires_grad_(True)
    loss = L._neg_loss(pr, gt); loss.backward()
    loss_nopos = L._neg_loss(pred, torch.clamp(gt, max=0.5))
    output = torch.randn(B, 2, H, W, generator=g)
    ind = torch.randint(0, H * W, (B, M), generator=g)
    mask = (torch.rand(B, M, generator=g) > 0.3).to(torch.uint8)
    target = torch.randn(B, M, 2, generator=g)
    wmask = torch.rand(B, M, 2, generator=g)
    import warnings
    warnings.simplefilter("ignore")
    save("losses", logits=logits, gt=gt, pred=pred, neg_loss=loss.detach(), neg_loss_grad=pr.grad,
         neg_loss_nopos=loss_nopos, num_pos=float((gt == 1).sum()),
         output=output, ind=ind, mask=mask, target=target, wmask=wmask,
         reg_l1=L.RegL1Loss()(output, mask, ind, target), reg_sl1=L.RegLoss()(output, mask, ind, target),
         norm_l1=L.NormRegL1Loss()(output, mask, ind, target.abs() + 0.5),
         weighted_l1=L.RegWeightedL1Loss()(output, wmask, ind, target))

    # ---- A15 target splat (datasets/sample/ctdet.py:99-127 loop on random boxes)
    B, C, H, W, M = 2, 4, 32, 40, 10
    cls = rng.integers(0, C, (B, M)).astype(np.int32)
    bw = rng.uniform(1, 30, (B, M)); bh_ = rng.uniform(1, 30, (B, M))
    cx = rng.uniform(0, W, (B, M)); cy = rng.uniform(0, H, (B, M))
    valid = (rng.random((B, M)) > 0.2).astype(np.uint8)
    radius = np.zeros((B, M), np.int32); cxi = cx.astype(np.int32); cyi = cy.astype(np.int32)
    hm = np.zeros((B, C, H, W), np.float32)
    for b in range(B):
        for m in range(M):
            radius[b, m] = max(0, int(I.gaussian_radius((np.ceil(bh_[b, m]), np.ceil(bw[b, m])))))
            if valid[b, m]:
                I.draw_umich_gaussian(hm[b, cls[b, m]], np.array([cx[b, m], cy[b, m]], np.float32).astype(np.int32),
                                      int(radius[b, m]))
    save("splat", obj_cls=cls, obj_cx=cxi, obj_cy=cyi, obj_radius=radius, obj_valid=valid, box_h=bh_, box_w=bw,
         hm=hm, C=C, H=H, W=W)


if __name__ == "__main__":
    main()
