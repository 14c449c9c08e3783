"""Drop-in overlay for src/lib/models/networks/DCNv2/dcn_v2_func.py: replaces the cffi module
`_ext.dcn_v2` (dcn_v2_func.py:9) -- no `make.sh` / torch.utils.ffi build step is needed any more."""
from centernet_b200.dcn_v2_func import DCNv2Function, DCNv2PoolingFunction  # noqa: F401
