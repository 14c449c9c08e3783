"""ctypes binding of oracle/_ref/libdcnv2_ref.so (TEST INFRASTRUCTURE): the UNMODIFIED reference
DCNv2 / deformable-PSROI CUDA kernels driven by the cuBLAS restatement of their host loop
(oracle/ref_host.cu, recipe oracle/build_ref.py).  This is "the reference on the GPU": GPU tests
pin oracle/dcn_ref.py, oracle/psroi_np.py and our kernels to it, and bench.py times it as the
`reference_gpu` bar.  Argument meaning follows DCNv2Function / DCNv2PoolingFunction
(DCNv2/dcn_v2_func.py:22-62, 104-146): the caller allocates outputs, gradients arrive zeroed.
"""
import ctypes
import os

import torch

from . import build_ref

_LIB = None


def available():
    return build_ref.available()


def lib():
    global _LIB
    if _LIB is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libdcnv2_ref.so is missing (build it where /root/reference exists: "
                               "python -m oracle.build_ref)")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _out_hw(H, W, kh, kw, stride, padding, dilation):
    return ((H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1,
            (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1)


def dcn_v2_forward(x, offset, mask, weight, bias, stride=1, padding=1, dilation=1, deformable_groups=1):
    x, offset, mask, weight, bias = [t.contiguous().float() for t in (x, offset, mask, weight, bias)]
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho, Wo = _out_hw(H, W, kh, kw, stride, padding, dilation)
    out = torch.empty(B, Co, Ho, Wo, device=x.device)
    columns = torch.empty(C * kh * kw, Ho * Wo, device=x.device)
    ones = torch.ones(Ho, Wo, device=x.device)
    rc = lib().ref_dcn_v2_forward(_p(x), _p(weight), _p(bias), _p(ones), _p(offset), _p(mask), _p(out), _p(columns),
                                  B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation,
                                  deformable_groups, _stream())
    if rc:
        raise RuntimeError("ref_dcn_v2_forward failed: %d" % rc)
    return out


def dcn_v2_backward(x, offset, mask, weight, bias, grad_output, stride=1, padding=1, dilation=1,
                    deformable_groups=1):
    """-> (grad_input, grad_offset, grad_mask, grad_weight, grad_bias), dcn_v2_func.py:40-62."""
    x, offset, mask, weight, grad_output = [t.contiguous().float() for t in (x, offset, mask, weight, grad_output)]
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho, Wo = _out_hw(H, W, kh, kw, stride, padding, dilation)
    gi = torch.zeros_like(x); go = torch.zeros_like(offset); gm = torch.zeros_like(mask)
    gw = torch.zeros_like(weight); gb = torch.zeros(Co, device=x.device)
    columns = torch.empty(C * kh * kw, Ho * Wo, device=x.device)
    ones = torch.ones(Ho, Wo, device=x.device)
    rc = lib().ref_dcn_v2_backward(_p(x), _p(weight), _p(ones), _p(offset), _p(mask), _p(columns), _p(gi), _p(gw),
                                   _p(gb), _p(go), _p(gm), _p(grad_output), B, C, H, W, Co, kh, kw, stride, stride,
                                   padding, padding, dilation, dilation, deformable_groups, _stream())
    if rc:
        raise RuntimeError("ref_dcn_v2_backward failed: %d" % rc)
    return gi, go, gm, gw, gb


def psroi_forward(data, rois, trans, spatial_scale, pooled_size, output_dim, no_trans, group_size=1, part_size=None,
                  sample_per_part=4, trans_std=0.0):
    """-> (out, top_count), dcn_v2_func.py:104-120."""
    part_size = pooled_size if part_size is None else part_size
    data, rois, trans = [t.contiguous().float() for t in (data, rois, trans)]
    B, C, H, W = data.shape
    n = rois.shape[0]
    out = torch.zeros(n, output_dim, pooled_size, pooled_size, device=data.device)
    cnt = torch.zeros_like(out)
    ct = 2 if no_trans else trans.shape[1]
    rc = lib().ref_psroi_forward(_p(data), _p(rois), _p(trans), _p(out), _p(cnt), B, C, H, W, n, ct, int(no_trans),
                                 ctypes.c_float(spatial_scale), output_dim, group_size, pooled_size, part_size,
                                 sample_per_part, ctypes.c_float(trans_std), _stream())
    if rc:
        raise RuntimeError("ref_psroi_forward failed: %d" % rc)
    return out, cnt


def psroi_backward(out_grad, data, rois, trans, top_count, spatial_scale, pooled_size, output_dim, no_trans,
                   group_size=1, part_size=None, sample_per_part=4, trans_std=0.0):
    """-> (grad_input, grad_trans), dcn_v2_func.py:122-146."""
    part_size = pooled_size if part_size is None else part_size
    out_grad, data, rois, trans, top_count = [t.contiguous().float() for t in (out_grad, data, rois, trans, top_count)]
    B, C, H, W = data.shape
    n = rois.shape[0]
    gi = torch.zeros_like(data); gt = torch.zeros_like(trans)
    ct = 2 if no_trans else trans.shape[1]
    rc = lib().ref_psroi_backward(_p(out_grad), _p(data), _p(rois), _p(trans), _p(top_count), _p(gi), _p(gt), B, C, H,
                                  W, n, ct, int(no_trans), ctypes.c_float(spatial_scale), output_dim, group_size,
                                  pooled_size, part_size, sample_per_part, ctypes.c_float(trans_std), _stream())
    if rc:
        raise RuntimeError("ref_psroi_backward failed: %d" % rc)
    return gi, gt
