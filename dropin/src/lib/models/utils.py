"""Drop-in overlay for src/lib/models/utils.py."""
from centernet_b200.utils import (  # noqa: F401
    _sigmoid, _gather_feat, _transpose_and_gather_feat, flip_tensor, flip_lr, flip_lr_off)
