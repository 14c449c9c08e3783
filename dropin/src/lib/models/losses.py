"""Drop-in overlay for src/lib/models/losses.py: every class trains/{ctdet,multi_pose,exdet,ddd}.py import.
Optional -- the reference's own losses.py also keeps working on top of the overlaid models/utils.py, because
centernet_b200.utils._transpose_and_gather_feat is differentiable."""
from centernet_b200.losses import (  # noqa: F401
    _neg_loss, FocalLoss, RegLoss, RegL1Loss, NormRegL1Loss, RegWeightedL1Loss, L1Loss, BinRotLoss,
    compute_rot_loss,
    FocalSplatLoss, splat_gaussian)   # extra: focal loss against on-the-fly Gaussian targets
