"""TEST INFRASTRUCTURE ONLY -- CPU restatement of deformable PSROI pooling (DCNv2 package surface,
SURVEY.md section 8f N4).

Follows src/lib/models/networks/DCNv2/src/cuda/dcn_v2_psroi_pooling_cuda.cu:
  bilinear_interp :24-45, forward kernel :47-139, backward kernel :141-255, class bookkeeping :282-283.
Plain Python loops over (roi, channel, bin, sample): small cases only.

Pinning: the reference kernels are CUDA-only and the reference extension cannot be built with the
current PyTorch, so there is no reference output to compare with here: PARITY UNPINNED
except for the reference's own known-answer test (DCNv2/test.py:117-146: zero offsets == no_trans, and the
pooled means of the two constant blocks), which tests/test_psroi_oracle.py reproduces."""
import math
import numpy as np

f32 = np.float32


def _geometry(roi, spatial_scale, pooled, part_size, sample_per_part, ph, pw):
    # :73-96 -- float arithmetic with double literals, as the CUDA source evaluates it
    start_w = f32(float(f32(round(float(roi[1]))) * f32(spatial_scale)) - 0.5)
    start_h = f32(float(f32(round(float(roi[2]))) * f32(spatial_scale)) - 0.5)
    end_w = f32(float(f32(round(float(roi[3])) + 1.0) * f32(spatial_scale)) - 0.5)
    end_h = f32(float(f32(round(float(roi[4])) + 1.0) * f32(spatial_scale)) - 0.5)
    roi_w = f32(max(float(f32(end_w - start_w)), 0.1))
    roi_h = f32(max(float(f32(end_h - start_h)), 0.1))
    bin_h, bin_w = f32(roi_h / f32(pooled)), f32(roi_w / f32(pooled))
    sub_h, sub_w = f32(bin_h / f32(sample_per_part)), f32(bin_w / f32(sample_per_part))
    part_h = int(math.floor(float(f32(f32(ph) / f32(pooled)) * f32(part_size))))
    part_w = int(math.floor(float(f32(f32(pw) / f32(pooled)) * f32(part_size))))
    return start_w, start_h, roi_w, roi_h, bin_w, bin_h, sub_w, sub_h, part_h, part_w


def _round_half_away(x):  # C round()
    return math.floor(abs(x) + 0.5) * (1 if x >= 0 else -1)


def psroi_forward(data, rois, trans, no_trans, spatial_scale, output_dim, group_size, pooled_size, part_size,
                  sample_per_part, trans_std):
    """data [B,C,H,W], rois [N,5] (batch, x1, y1, x2, y2), trans [N, 2*num_classes, part, part] or None.
    Returns (out [N, output_dim, P, P], count [N, output_dim, P, P])."""
    data = np.asarray(data, f32)
    rois = np.asarray(rois, f32)
    B, C, H, W = data.shape
    N, P = rois.shape[0], pooled_size
    num_classes = 1 if no_trans else trans.shape[1] // 2
    ch_each = output_dim if no_trans else output_dim // num_classes
    out = np.zeros((N, output_dim, P, P), f32)
    cnt = np.zeros((N, output_dim, P, P), f32)
    for n in range(N):
        roi = [float(rois[n, 0])] + [_round_half_away(float(v)) for v in rois[n, 1:]]
        bi = int(rois[n, 0])
        for ctop in range(output_dim):
            cls = ctop // ch_each
            for ph in range(P):
                for pw in range(P):
                    sw, sh, rw, rh, bw, bh, subw, subh, part_h, part_w = _geometry(
                        roi, spatial_scale, P, part_size, sample_per_part, ph, pw)
                    tx = f32(0) if no_trans else f32(trans[n, cls * 2, part_h, part_w]) * f32(trans_std)
                    ty = f32(0) if no_trans else f32(trans[n, cls * 2 + 1, part_h, part_w]) * f32(trans_std)
                    wstart = f32(f32(f32(pw) * bw + sw) + f32(tx * rw))
                    hstart = f32(f32(f32(ph) * bh + sh) + f32(ty * rh))
                    gw = min(max(int(math.floor(float(f32(pw) * f32(group_size) / f32(P)))), 0), group_size - 1)
                    gh = min(max(int(math.floor(float(f32(ph) * f32(group_size) / f32(P)))), 0), group_size - 1)
                    c = (ctop * group_size + gh) * group_size + gw
                    s, k = f32(0), 0
                    for ih in range(sample_per_part):
                        for iw in range(sample_per_part):
                            w = f32(wstart + f32(iw) * subw)
                            h = f32(hstart + f32(ih) * subh)
                            if w < -0.5 or w > W - 0.5 or h < -0.5 or h > H - 0.5:
                                continue
                            w = f32(min(max(float(w), 0.0), W - 1.0))
                            h = f32(min(max(float(h), 0.0), H - 1.0))
                            x1, x2 = int(math.floor(w)), int(math.ceil(w))
                            y1, y2 = int(math.floor(h)), int(math.ceil(h))
                            dx, dy = f32(w - f32(x1)), f32(h - f32(y1))
                            pl = data[bi, c]
                            val = f32(f32(f32(f32(1) - dx) * f32(f32(1) - dy)) * pl[y1, x1]) \
                                + f32(f32(f32(f32(1) - dx) * dy) * pl[y2, x1]) \
                                + f32(f32(dx * f32(f32(1) - dy)) * pl[y1, x2]) + f32(f32(dx * dy) * pl[y2, x2])
                            s = f32(s + f32(val))
                            k += 1
                    out[n, ctop, ph, pw] = f32(0) if k == 0 else f32(s / f32(k))
                    cnt[n, ctop, ph, pw] = k
    return out, cnt


def psroi_backward(grad_out, data, rois, trans, count, no_trans, spatial_scale, output_dim, group_size, pooled_size,
                   part_size, sample_per_part, trans_std):
    """Gradients w.r.t. data and trans (float64 accumulation; the CUDA kernels use fp32 atomics)."""
    data = np.asarray(data, f32)
    rois = np.asarray(rois, f32)
    B, C, H, W = data.shape
    N, P = rois.shape[0], pooled_size
    num_classes = 1 if no_trans else trans.shape[1] // 2
    ch_each = output_dim if no_trans else output_dim // num_classes
    gdata = np.zeros(data.shape, np.float64)
    gtrans = None if no_trans else np.zeros(trans.shape, np.float64)
    for n in range(N):
        roi = [float(rois[n, 0])] + [_round_half_away(float(v)) for v in rois[n, 1:]]
        bi = int(rois[n, 0])
        for ctop in range(output_dim):
            cls = ctop // ch_each
            for ph in range(P):
                for pw in range(P):
                    if count[n, ctop, ph, pw] <= 0:
                        continue
                    sw, sh, rw, rh, bw, bh, subw, subh, part_h, part_w = _geometry(
                        roi, spatial_scale, P, part_size, sample_per_part, ph, pw)
                    tx = f32(0) if no_trans else f32(trans[n, cls * 2, part_h, part_w]) * f32(trans_std)
                    ty = f32(0) if no_trans else f32(trans[n, cls * 2 + 1, part_h, part_w]) * f32(trans_std)
                    wstart = f32(f32(f32(pw) * bw + sw) + f32(tx * rw))
                    hstart = f32(f32(f32(ph) * bh + sh) + f32(ty * rh))
                    gw = min(max(int(math.floor(float(f32(pw) * f32(group_size) / f32(P)))), 0), group_size - 1)
                    gh = min(max(int(math.floor(float(f32(ph) * f32(group_size) / f32(P)))), 0), group_size - 1)
                    c = (ctop * group_size + gh) * group_size + gw
                    dv = float(grad_out[n, ctop, ph, pw]) / float(count[n, ctop, ph, pw])
                    for ih in range(sample_per_part):
                        for iw in range(sample_per_part):
                            w = f32(wstart + f32(iw) * subw)
                            h = f32(hstart + f32(ih) * subh)
                            if w < -0.5 or w > W - 0.5 or h < -0.5 or h > H - 0.5:
                                continue
                            w = f32(min(max(float(w), 0.0), W - 1.0))
                            h = f32(min(max(float(h), 0.0), H - 1.0))
                            x0, x1 = int(math.floor(w)), int(math.ceil(w))
                            y0, y1 = int(math.floor(h)), int(math.ceil(h))
                            dx, dy = float(f32(w - f32(x0))), float(f32(h - f32(y0)))
                            gdata[bi, c, y0, x0] += (1 - dx) * (1 - dy) * dv      # :223-231
                            gdata[bi, c, y1, x0] += (1 - dx) * dy * dv
                            gdata[bi, c, y0, x1] += dx * (1 - dy) * dv
                            gdata[bi, c, y1, x1] += dx * dy * dv
                            if no_trans:
                                continue
                            pl = data[bi, c]
                            u00, u01, u10, u11 = float(pl[y0, x0]), float(pl[y1, x0]), float(pl[y0, x1]), float(pl[y1, x1])
                            gx = (u11 * dy + u10 * (1 - dy) - u01 * dy - u00 * (1 - dy)) * trans_std * dv * float(rw)   # :241-246
                            gy = (u11 * dx + u01 * (1 - dx) - u10 * dx - u00 * (1 - dx)) * trans_std * dv * float(rh)
                            gtrans[n, cls * 2, part_h, part_w] += gx
                            gtrans[n, cls * 2 + 1, part_h, part_w] += gy
    return gdata, gtrans
