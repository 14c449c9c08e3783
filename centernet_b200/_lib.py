"""Loads the in-tree pybind11 binding of the C ABI.  No fallback: a missing or stale
extension is an ImportError that says how to build it."""
import torch

try:
    from . import _C as C
except ImportError as e:  # pragma: no cover - exercised only on a broken install
    raise ImportError(
        "centernet_b200: the CUDA extension is not built (%s). Run `python -m centernet_b200.build` "
        "(needs nvcc, targets sm_100a). There is no CPU fallback." % (e,)) from e


def version():
    v = C.version()
    return "%d.%d.%d" % (v // 10000, (v // 100) % 100, v % 100)


def stream_ptr(t):
    """Raw cudaStream_t of the current stream on t's device (decode must be stream-ordered
    because callers immediately do dets.detach().cpu(), detectors/ctdet.py:48)."""
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*tensors, what="centernet_b200"):
    for t in tensors:
        if t is not None and not t.is_cuda:
            # same contract as DCNv2Function.forward (dcn_v2_func.py:23-24)
            raise NotImplementedError("%s: only implemented for CUDA tensors (no CPU fallback)" % what)


def f32c(t):
    """fp32 + contiguous view of t (no copy when it already is)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ptr(t):
    return 0 if t is None else t.data_ptr()


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
