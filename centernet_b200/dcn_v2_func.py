"""Mirror of src/lib/models/networks/DCNv2/dcn_v2_func.py.

``DCNv2Function(stride, padding, dilation=1, deformable_groups=1)`` keeps the reference's
construct-then-call shape (dcn_v2_func.py:13-73, used directly by DCNv2/test.py:88,113) while
delegating to a new-style static autograd Function; the work is done by cnb_dcnv2_forward /
cnb_dcnv2_backward (im2col-free, no `ones`/`columns` scratch).  Non-CUDA tensors raise
NotImplementedError exactly like the reference (:23-24, :41-42).
"""
import os

import torch
from torch.autograd import Function

from ._lib import C, f32c, ptr, stream_ptr, workspace

# CNB_DCN_FP32=1 selects the fp32 CUDA-core contraction instead of the tcgen05 (3xTF32) one
_FORCE_FP32 = os.environ.get("CNB_DCN_FP32", "0") == "1"


def _out_hw(h, w, kh, kw, stride, padding, dilation):
    ho = (h + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1   # dcn_v2_func.py:64-73
    wo = (w + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    return ho, wo


class _DCNv2(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
        if not input.is_cuda:
            raise NotImplementedError
        x, off, msk, w, b = f32c(input), f32c(offset), f32c(mask), f32c(weight), f32c(bias)
        n, cin, h, wd = [int(v) for v in x.shape]
        cout, cin_w, kh, kw = [int(v) for v in w.shape]
        if cin_w != cin:   # dcn_v2_cuda.c:36-38
            raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (cin, cin_w))
        ho, wo = _out_hw(h, wd, kh, kw, stride, padding, dilation)
        kt = kh * kw
        if tuple(off.shape) != (n, 2 * kt * deformable_groups, ho, wo) or \
                tuple(msk.shape) != (n, kt * deformable_groups, ho, wo):
            raise RuntimeError("DCNv2: offset/mask shape does not match the output grid")
        out = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device)
        ws_bytes = 0 if _FORCE_FP32 else C.dcnv2_workspace_bytes(n, cin, cout, h, wd, kh, kw, stride, padding,
                                                                 dilation, deformable_groups)
        ws = workspace(ws_bytes, x.device) if ws_bytes else None
        C.dcnv2_forward(ptr(x), ptr(off), ptr(msk), ptr(w), ptr(b), ptr(out), n, cin, h, wd, cout, kh, kw,
                        stride, stride, padding, padding, dilation, dilation, deformable_groups, ptr(ws),
                        ws.numel() if ws is not None else 0, stream_ptr(x))
        ctx.save_for_backward(x, off, msk, w, b)
        ctx.cfg = (stride, padding, dilation, deformable_groups)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        x, off, msk, w, b = ctx.saved_tensors
        stride, padding, dilation, dg = ctx.cfg
        go = f32c(grad_output)
        n, cin, h, wd = [int(v) for v in x.shape]
        cout, _, kh, kw = [int(v) for v in w.shape]
        # zero-filled, accumulated into -- dcn_v2_func.py:44-48
        gx = torch.zeros_like(x)
        goff = torch.zeros_like(off)
        gmsk = torch.zeros_like(msk)
        gw = torch.zeros_like(w)
        gb = torch.zeros_like(b) if b is not None else None
        C.dcnv2_backward(ptr(x), ptr(off), ptr(msk), ptr(w), ptr(go), ptr(gx), ptr(goff), ptr(gmsk), ptr(gw), ptr(gb),
                         n, cin, h, wd, cout, kh, kw, stride, stride, padding, padding, dilation, dilation, dg,
                         0, 0, stream_ptr(x))
        return gx, goff, gmsk, gw, gb, None, None, None, None


class DCNv2Function(object):
    """Callable instance, same constructor as the reference's old-style Function."""

    def __init__(self, stride, padding, dilation=1, deformable_groups=1):
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.deformable_groups = deformable_groups

    def __call__(self, input, offset, mask, weight, bias):
        return _DCNv2.apply(input, offset, mask, weight, bias, self.stride, self.padding, self.dilation,
                            self.deformable_groups)

    forward = __call__


class DCNv2PoolingFunction(object):
    """Deformable PSROI pooling (dcn_v2_func.py:76-146).  No CenterNet network instantiates it
    (SURVEY.md section 2b); it is a "next" row (section 8f N4) and not implemented yet."""

    def __init__(self, *args, **kwargs):
        self.args = (args, kwargs)

    def __call__(self, *a, **k):
        raise NotImplementedError("DCNv2PoolingFunction: deformable PSROI pooling is not implemented in centernet_b200")
