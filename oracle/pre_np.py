"""CPU restatement of the detector's input pre-processing (TEST INFRASTRUCTURE):
detectors/base_detector.py:37-65 -- get_affine_transform (utils/image.py:27-60), cv2.warpAffine(INTER_LINEAR,
BORDER_CONSTANT 0) restated from OpenCV's fixed-point algorithm (imgwarp.cpp: 10-bit coordinate grid, 5-bit
sub-pixel phase, 15-bit bilinear weights, (sum + 2^14) >> 15), normalisation in float64, HWC -> CHW, flip concat.
Pinned by tests/golden/pre.npz (the unmodified BaseDetector.pre_process run on seeded images)."""
import numpy as np

F32 = np.float32


def forward_affine(center, scale, output_size):
    """get_affine_transform(c, s, 0, output_size) (inv = 0): float32 points as the reference, 3-point solve in
    float64 (cv2.getAffineTransform there)."""
    if not isinstance(scale, (np.ndarray, list, tuple)):
        scale = np.array([scale, scale], dtype=F32)
    scale = np.asarray(scale, dtype=F32)
    center = np.asarray(center, dtype=F32)
    src_w = scale[0]
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), F32); dst = np.zeros((3, 2), F32)
    src[0] = center
    src[1] = center + np.array([0.0, src_w * F32(-0.5)], F32)
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], F32) + np.array([0, dst_w * -0.5], F32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], F32)
    a = np.concatenate([src.astype(np.float64), np.ones((3, 1))], axis=1)
    return np.linalg.solve(a, dst.astype(np.float64)).T          # 2 x 3, src -> dst


def invert_affine(m):
    """The in-place inversion cv::warpAffine performs on its (forward) matrix."""
    m = [float(v) for v in np.asarray(m, np.float64).reshape(6)]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    return m


def warp_affine_linear(img, m_fwd, dsize):
    """cv2.warpAffine(img, m_fwd, dsize, flags=cv2.INTER_LINEAR) for uint8 HxWxC images."""
    img = np.asarray(img, np.uint8)
    H, W = img.shape[:2]
    dw, dh = dsize
    m = invert_affine(m_fwd)
    x = np.arange(dw, dtype=np.float64); y = np.arange(dh, dtype=np.float64)
    adelta = np.rint(m[0] * x * 1024.0).astype(np.int64)
    bdelta = np.rint(m[3] * x * 1024.0).astype(np.int64)
    X0 = np.rint((m[1] * y + m[2]) * 1024.0).astype(np.int64) + 16
    Y0 = np.rint((m[4] * y + m[5]) * 1024.0).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767); sy = np.clip(Y >> 5, -32768, 32767)
    fx = X & 31; fy = Y & 31
    w = [(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32]
    out = np.zeros((dh, dw) + img.shape[2:], np.int64)
    src = img.astype(np.int64)
    for k, (oy, ox) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy = sy + oy; xx = sx + ox
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        v = np.where(ok[..., None] if img.ndim == 3 else ok, v, 0)
        out += v * (w[k][..., None] if img.ndim == 3 else w[k])
    return np.clip((out + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def resize_linear_u8(img, dsize):
    """cv2.resize(img, dsize) (INTER_LINEAR) for uint8 images: OpenCV's 11-bit fixed-point separable filter
    (resize.cpp: coefficients saturate_cast<short>(w * 2048) from a float32 phase; horizontally the source index is
    clamped and the phase zeroed at the borders, vertically only the row indices are clamped; vertical pass
    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2 >> 2)."""
    img = np.asarray(img, np.uint8)
    H, W = img.shape[:2]
    dw, dh = dsize
    d = np.arange(dw)
    f = ((d + 0.5) * (W / dw) - 0.5).astype(F32)
    sx = np.floor(f).astype(np.int64)
    f = (f - sx).astype(F32)
    lo = sx < 0; f[lo] = 0; sx[lo] = 0
    hi = sx >= W - 1; f[hi] = 0; sx[hi] = W - 1
    ax0 = np.rint((F32(1) - f) * F32(2048)).astype(np.int64); ax1 = np.rint(f * F32(2048)).astype(np.int64)
    sx1 = np.minimum(sx + 1, W - 1)
    d = np.arange(dh)
    f = ((d + 0.5) * (H / dh) - 0.5).astype(F32)
    sy = np.floor(f).astype(np.int64)
    f = (f - sy).astype(F32)
    by0 = np.rint((F32(1) - f) * F32(2048)).astype(np.int64); by1 = np.rint(f * F32(2048)).astype(np.int64)
    r0 = np.clip(sy, 0, H - 1); r1 = np.clip(sy + 1, 0, H - 1)
    src = img.astype(np.int64)
    hrow = src[:, sx] * ax0[None, :, None] + src[:, sx1] * ax1[None, :, None]
    out = (((by0[:, None, None] * (hrow[r0] >> 4)) >> 16) + ((by1[:, None, None] * (hrow[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def pre_process(image, scale, mean, std, fix_res=True, input_h=512, input_w=512, pad=31, down_ratio=4, flip_test=False):
    """detectors/base_detector.py:37-65."""
    height, width = image.shape[0:2]
    new_height, new_width = int(height * scale), int(width * scale)
    if fix_res:
        inp_height, inp_width = input_h, input_w
        c = np.array([new_width / 2., new_height / 2.], dtype=F32)
        s = max(height, width) * 1.0
    else:
        inp_height = (new_height | pad) + 1
        inp_width = (new_width | pad) + 1
        c = np.array([new_width // 2, new_height // 2], dtype=F32)
        s = np.array([inp_width, inp_height], dtype=F32)
    trans_input = forward_affine(c, s, [inp_width, inp_height])
    resized = resize_linear_u8(image, (new_width, new_height))          # identity when scale == 1
    inp = warp_affine_linear(resized, trans_input, (inp_width, inp_height))
    mean = np.asarray(mean, F32).reshape(1, 1, 3); std = np.asarray(std, F32).reshape(1, 1, 3)
    inp = ((inp / 255. - mean) / std).astype(F32)
    images = inp.transpose(2, 0, 1).reshape(1, 3, inp_height, inp_width)
    if flip_test:
        images = np.concatenate((images, images[:, :, :, ::-1]), axis=0)
    meta = {'c': c, 's': s, 'out_height': inp_height // down_ratio, 'out_width': inp_width // down_ratio}
    return images, meta
