"""Generates tests/golden/big_*.npz, utils.npz and losses_grads.npz by running the UNMODIFIED reference
(imported from /root/reference/src/lib, CPU) -- build container only:
    python tests/golden/make_golden_big.py

big_*: the north-star geometry (80x128x128, K=100; exct K=40 AND K=100).  Inputs come from
tests/big_inputs.py (seeded, IEEE-exact, rebuilt by the tests); only the reference outputs and an
input checksum are stored.  utils.npz: models/utils.py (A4/A13 helpers).  losses_grads.npz: autograd
gradients of the four Reg*Loss classes on the inputs of losses.npz.
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference/src/lib"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import big_inputs as BI  # noqa: E402
from models import decode as R  # noqa: E402
from models import losses as L  # noqa: E402
from models import utils as U  # noqa: E402

warnings.simplefilter("ignore")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def assert_tie_free(scores, what):
    s = np.asarray(scores)
    for row in s.reshape(-1, s.shape[-1]):
        pos = row[row > 0]
        assert len(np.unique(pos)) == len(pos), "ties in " + what


def main():
    torch.set_num_threads(8)
    # ---- ctdet at 2x80x128x128, K=100 (A1+A2+A4+A5)
    for kind in ("noise", "blobs"):
        heat, wh, reg = BI.ctdet_inputs(kind)
        dets = R.ctdet_decode(t(heat), t(wh), reg=t(reg), K=100)
        assert_tie_free(dets[..., 4], "big_ctdet_" + kind)
        dets_noreg = R.ctdet_decode(t(heat), t(wh), reg=None, K=100)
        save("big_ctdet_" + kind, dets=dets, dets_noreg=dets_noreg, crc=BI.checksum(heat, wh, reg), K=100)

    # ---- multi_pose at 2x128x128, K=100 (A3+A6)
    heat, wh, kps, reg, hm_hp, hp_off = BI.multi_pose_inputs()
    dets = R.multi_pose_decode(t(heat), t(wh), t(kps), reg=t(reg), hm_hp=t(hm_hp), hp_offset=t(hp_off), K=100)
    assert_tie_free(dets[..., 4], "big_multi_pose")
    save("big_multi_pose", dets=dets, crc=BI.checksum(heat, wh, kps, reg, hm_hp, hp_off), K=100)

    # ---- exct at 1x80x128x128: K=40 (signature default) and K=100 (detectors/exdet.py:39-51), aggr 0 / 0.1
    maps, regs = BI.exct_inputs()
    tm = [t(m) for m in maps]; tr = [t(r) for r in regs]
    crc = BI.checksum(*maps, *regs)
    orig_topk = R._topk
    out = {}
    for K in (40, 100):
        out["dets_k%d" % K] = R.exct_decode(*tm, *tr, K=K, num_dets=1000)
        assert (out["dets_k%d" % K][..., 4] > 0).sum() > 0
        print("exct K=%d done" % K, flush=True)
    R._topk = lambda s_, K=40: orig_topk(s_.contiguous(), K=K)   # the aggr path needs .contiguous()
    out["dets_k40_aggr"] = R.exct_decode(*tm, *tr, K=40, num_dets=1000, aggr_weight=0.1)
    R._topk = orig_topk
    t1, l1, b1, r1 = [x.max(dim=1, keepdim=True)[0] for x in tm[:4]]
    out["dets_agn_k40"] = R.agnex_ct_decode(t1, l1, b1, r1, tm[4], *tr, K=40, num_dets=1000)
    save("big_exct", crc=crc, **out)

    # ---- models/utils.py helpers (A4, A13)
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(2, 34, 12, 20, generator=g)
    ind = torch.randint(0, 240, (2, 9), generator=g)
    feat3 = torch.randn(2, 50, 7, generator=g)
    ind3 = torch.randint(0, 50, (2, 13), generator=g)
    flip_idx = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]   # coco_hp.py:32-33
    hm17 = torch.randn(2, 17, 6, 10, generator=g)
    logits = torch.randn(2, 3, 6, 10, generator=g) * 6
    save("utils", feat=feat, ind=ind, tg=U._transpose_and_gather_feat(feat, ind), feat3=feat3, ind3=ind3,
         g=U._gather_feat(feat3, ind3), flip_tensor=U.flip_tensor(feat), flip_idx=np.asarray(flip_idx),
         hm17=hm17, flip_lr=U.flip_lr(hm17, flip_idx), flip_lr_off=U.flip_lr_off(feat, flip_idx),
         logits=logits, sigmoid=U._sigmoid(logits.clone()))

    # ---- gradients of the Reg losses (A16) on the inputs of losses.npz
    z = np.load(os.path.join(HERE, "losses.npz"))
    output, ind, mask, target, wmask = [t(z[k]) for k in ("output", "ind", "mask", "target", "wmask")]
    grads = {}
    for name, crit, args in (("reg_l1", L.RegL1Loss(), (mask, ind, target)),
                             ("reg_sl1", L.RegLoss(), (mask, ind, target)),
                             ("norm_l1", L.NormRegL1Loss(), (mask, ind, target.abs() + 0.5)),
                             ("weighted_l1", L.RegWeightedL1Loss(), (wmask, ind, target))):
        o = output.clone().requires_grad_(True)
        loss = crit(o, *args)
        loss.backward()
        assert abs(float(loss) - float(z[name])) < 1e-6
        grads[name + "_grad"] = o.grad
    save("losses_grads", **grads)


if __name__ == "__main__":
    main()
