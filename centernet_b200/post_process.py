"""Mirror of src/lib/utils/post_process.py (ctdet_post_process, multi_pose_post_process), of
src/lib/external/nms.pyx (soft_nms, soft_nms_39) and of CtdetDetector.merge_outputs
(detectors/ctdet.py:76-92), computed on the device.

Two ways in:
  * the drop-in functions keep the reference's signatures and return types (numpy in, python lists /
    in-place arrays out) so `detectors/*.py` work unchanged -- but they also accept the CUDA tensor that
    `ctdet_decode` returned, which removes the `dets.detach().cpu().numpy()` of detectors/ctdet.py:48;
  * `ctdet_merge_device` is the whole tail of a detector run (back-projection of every scale, per-class
    concat, soft-NMS, top max_per_image) in four launches with ONE device->host copy at the end.
"""
import numpy as np
import torch

from ._lib import C, ptr, stream_ptr


# ------------------------------------------------------------------ utils/image.py:27-60 (host, float64, B matrices)
def _inv_affine(center, scale, output_size):
    """2x3 float64 matrix of get_affine_transform(center, scale, 0, output_size, inv=1): output-grid -> image.
    rot = 0, so the three point pairs of the reference define a uniform scale + translation; it is solved
    from the same float32 points in float64 (cv2.getAffineTransform there)."""
    f32 = np.float32
    if not isinstance(scale, (np.ndarray, list, tuple)):
        scale = np.array([scale, scale], dtype=f32)
    scale = np.asarray(scale, dtype=f32).reshape(-1)
    center = np.asarray(center, dtype=f32).reshape(2)
    src_w = scale[0]
    dst_w, dst_h = output_size
    src = np.zeros((3, 2), f32); dst = np.zeros((3, 2), f32)
    src[0] = center
    src[1] = center + np.array([0.0, src_w * f32(-0.5)], f32)
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], f32) + np.array([0, dst_w * -0.5], f32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], f32)
    a = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], axis=1)
    return np.linalg.solve(a, src.astype(np.float64)).T


def _trans_tensor(c, s, w, h, device):
    mats = np.stack([_inv_affine(c[i], s[i], (w, h)) for i in range(len(c))], 0).reshape(len(c), 6)
    return torch.from_numpy(np.ascontiguousarray(mats)).to(device)


def _as_cuda(dets):
    if isinstance(dets, torch.Tensor):
        if not dets.is_cuda:
            raise NotImplementedError("post_process: CUDA tensors or numpy arrays only (no CPU path)")
        return dets.detach().float().contiguous()
    return torch.from_numpy(np.ascontiguousarray(dets, dtype=np.float32)).cuda()


def transform_dets(dets, c, s, h, w, pairs):
    """Back-projects the (x, y) column pairs `pairs` = ((col0, n), (col0, n)) of dets [B, N, D]; returns a new
    CUDA tensor (the decode output is left untouched)."""
    d = _as_cuda(dets)
    b, n, dim = [int(v) for v in d.shape]
    out = torch.empty_like(d)
    t = _trans_tensor(c, s, w, h, d.device)
    (a0, na), (b0, nb) = pairs
    C.post_transform(ptr(d), ptr(out), ptr(t), b, n, dim, a0, na, b0, nb, stream_ptr(d))
    return out


# ------------------------------------------------------------------ utils/post_process.py:83-99
def ctdet_post_process(dets, c, s, h, w, num_classes):
    """dets: [B, K, 6] numpy array or CUDA tensor -> list (per image) of {1-based class id: [[x1,y1,x2,y2,score]]}
    exactly as the reference (python lists of float32 rows)."""
    out = transform_dets(dets, c, s, h, w, ((0, 2), (0, 0)))
    b, n, _ = [int(v) for v in out.shape]
    rows = torch.empty((b, n, 5), dtype=torch.float32, device=out.device)
    offs = torch.empty((b, num_classes + 1), dtype=torch.int32, device=out.device)
    C.group_by_class(ptr(out), b, n, num_classes, ptr(rows), ptr(offs), stream_ptr(out))
    rows_h, offs_h = rows.cpu().numpy(), offs.cpu().numpy()
    ret = []
    for i in range(b):
        ret.append({j + 1: rows_h[i, offs_h[i, j]:offs_h[i, j + 1]].tolist() for j in range(num_classes)})
    return ret


# ------------------------------------------------------------------ utils/post_process.py:102-114
def multi_pose_post_process(dets, c, s, h, w):
    """dets: [B, K, 40] -> list of {1: [[bbox(4), score, 17 x (x, y)] ...]} in image coordinates."""
    out = transform_dets(dets, c, s, h, w, ((0, 2), (5, 17)))
    host = out[:, :, :39].cpu().numpy()
    return [{np.ones(1, dtype=np.int32)[0]: host[i].tolist()} for i in range(host.shape[0])]


# ------------------------------------------------------------------ external/nms.pyx:77-275
def _soft_nms(boxes, sigma, Nt, threshold, method, d):
    if not (isinstance(boxes, np.ndarray) and boxes.dtype == np.float32 and boxes.ndim == 2 and boxes.shape[1] >= d):
        raise TypeError("soft_nms: boxes must be a float32 [N, >=%d] numpy array (modified in place)" % d)
    n = int(boxes.shape[0])
    if n == 0:
        return []
    dev = torch.device("cuda", torch.cuda.current_device())
    src = torch.from_numpy(np.ascontiguousarray(boxes[:, :d])).to(dev)
    dst = torch.empty_like(src)
    offs = torch.tensor([0, n], dtype=torch.int32, device=dev)
    fin = torch.empty(1, dtype=torch.int32, device=dev)
    C.soft_nms(ptr(src), ptr(dst), ptr(offs), 1, 1, n, d, float(sigma), float(Nt), float(threshold), int(method),
               ptr(fin), stream_ptr(src))
    keep = int(fin.item())
    if keep < 0:
        raise NotImplementedError("soft_nms: lists longer than 1536 boxes are not implemented")
    boxes[:, :d] = dst.cpu().numpy()
    return list(range(keep))


def soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """Drop-in for external.nms.soft_nms: `boxes` [N, 5] float32 is modified IN PLACE exactly as the reference
    leaves it (rows beyond the returned length included); returns the keep list."""
    return _soft_nms(boxes, sigma, Nt, threshold, method, 5)


def soft_nms_39(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """Drop-in for external.nms.soft_nms_39 ([N, 39]: box, score, 17 keypoints)."""
    return _soft_nms(boxes, sigma, Nt, threshold, method, 39)


# ------------------------------------------------------------------ detectors/ctdet.py:47-92, on the device
def ctdet_merge_device(dets_per_scale, metas, num_classes, max_per_image=100, nms=False):
    """The tail of CtdetDetector.run for one image batch: for every test scale `ctdet_post_process` (+ the
    division by the scale, detectors/ctdet.py:47-58), then `merge_outputs` (:76-92): per-class concatenation
    over the scales, soft-NMS (Nt 0.5, Gaussian) when there are several scales or `nms` is set, and the cut to
    the max_per_image best scores.
      dets_per_scale: list of [B, K, 6] CUDA tensors (ctdet_decode outputs), one per test scale;
      metas: one dict per scale with the keys of base_detector.pre_process's `meta` (:58-62): 'c' and 's' as
             per-image sequences, 'out_height', 'out_width', plus 'scale' (the test scale, default 1).
    Returns a list (per image) of {class id: float32 [n, 5] array} -- one device->host copy in total."""
    S = len(dets_per_scale)
    assert len(metas) == S
    parts = []
    for k in range(S):
        m = metas[k]
        t = transform_dets(dets_per_scale[k], m["c"], m["s"], m["out_height"], m["out_width"], ((0, 2), (0, 0)))
        sc = float(m.get("scale", 1.0))
        if sc != 1.0:   # a tensor divisor: torch turns division by a Python scalar into a multiplication by 1/scale
            t[:, :, :4] /= torch.tensor(sc, dtype=torch.float32, device=t.device)
        parts.append(t)
    allp = torch.cat(parts, dim=1).contiguous()          # scale-major, like np.concatenate over `detections`
    b, n, _ = [int(v) for v in allp.shape]
    if n > 4096:
        raise NotImplementedError("ctdet_merge_device: at most 4096 detections per image")
    rows = torch.empty((b, n, 5), dtype=torch.float32, device=allp.device)
    offs = torch.empty((b, num_classes + 1), dtype=torch.int32, device=allp.device)
    C.group_by_class(ptr(allp), b, n, num_classes, ptr(rows), ptr(offs), stream_ptr(allp))
    if S > 1 or nms:
        C.soft_nms(ptr(rows), ptr(rows), ptr(offs), b, num_classes, n, 5, 0.5, 0.5, 0.001, 2, 0, stream_ptr(rows))
    keep = torch.empty((b, n), dtype=torch.uint8, device=allp.device)
    C.topk_keep(ptr(rows), b, n, 5, ptr(offs), num_classes, int(max_per_image), ptr(keep), 0, stream_ptr(rows))
    packed = torch.cat([rows.view(b, -1), keep.float(), offs.float()], dim=1).cpu().numpy()   # one D2H
    rows_h = packed[:, :n * 5].reshape(b, n, 5)
    keep_h = packed[:, n * 5:n * 6] > 0
    offs_h = packed[:, n * 6:].astype(np.int64)
    ret = []
    for i in range(b):
        res = {}
        for j in range(num_classes):
            a, e = offs_h[i, j], offs_h[i, j + 1]
            res[j + 1] = rows_h[i, a:e][keep_h[i, a:e]].astype(np.float32)
        ret.append(res)
    return ret
