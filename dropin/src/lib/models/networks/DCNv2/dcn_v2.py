"""Drop-in overlay for src/lib/models/networks/DCNv2/dcn_v2.py (pose_dla_dcn.py:16 and
resnet_dcn.py:18 do `from .DCNv2.dcn_v2 import DCN`)."""
from centernet_b200.dcn_v2 import DCN, DCNv2, DCNv2Pooling, DCNPooling  # noqa: F401
