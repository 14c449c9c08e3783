"""N4: detector input pre-processing (detectors/base_detector.py:37-65).  tests/golden/pre.npz holds CRCs and strided
samples of what the UNMODIFIED BaseDetector.pre_process returned (cv2.warpAffine + normalise + CHW + flip) for
the seeded images of tests/big_inputs.py; the numpy restatement of OpenCV's fixed-point warp (CPU) and the CUDA
kernel (GPU) must reproduce them bit for bit."""
import zlib

import numpy as np
import pytest

import big_inputs as BI
from helpers import golden
from oracle import pre_np

MEAN, STD = [0.408, 0.447, 0.470], [0.289, 0.274, 0.278]


def _check(name, images, meta, g):
    images = np.ascontiguousarray(images)
    assert tuple(images.shape) == tuple(int(v) for v in g["shape_" + name])
    np.testing.assert_array_equal(images[:, :, ::9, ::7], g["sample_" + name])
    assert np.uint32(zlib.crc32(images.view(np.uint8))) == g["crc_" + name]
    np.testing.assert_array_equal(np.asarray(meta["c"], np.float32), g["c_" + name])
    np.testing.assert_array_equal(np.asarray(meta["s"], np.float32).reshape(-1), g["s_" + name])
    assert [meta["out_height"], meta["out_width"]] == [int(v) for v in g["hw_" + name]]


@pytest.mark.parametrize("name", sorted(BI.PRE_CASES))
def test_oracle_pre_process(name):
    g = golden("pre")
    h, w, fix, flip, ih, iw, sc = BI.PRE_CASES[name]
    img = BI.pre_image(name)
    assert np.uint32(zlib.crc32(img.view(np.uint8))) == g["img_crc_" + name]
    images, meta = pre_np.pre_process(img, sc, MEAN, STD, fix_res=fix, input_h=ih or 512, input_w=iw or 512, flip_test=flip)
    _check(name, images, meta, g)


def test_oracle_warp_matches_cv2_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    for dsize in ((65, 48), (131, 97), (200, 150), (262, 194), (17, 301)):
        np.testing.assert_array_equal(pre_np.resize_linear_u8(img, dsize), cv2.resize(img, dsize))
    for c, s, out in (((65.5, 48.5), 131.0, (160, 128)), ((40.0, 60.0), (200.0, 90.0), (96, 64)), ((65.0, 48.0), 64.0, (256, 256))):
        m = pre_np.forward_affine(np.array(c, np.float32), s, out)
        np.testing.assert_array_equal(pre_np.warp_affine_linear(img, m, out), cv2.warpAffine(img, m, out, flags=cv2.INTER_LINEAR))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(BI.PRE_CASES))
def test_gpu_pre_process(name):
    from centernet_b200 import pre_process as P
    g = golden("pre")
    h, w, fix, flip, ih, iw, sc = BI.PRE_CASES[name]
    images, meta = P.pre_process(BI.pre_image(name), sc, MEAN, STD, fix_res=fix, input_h=ih or 512, input_w=iw or 512,
                                 flip_test=flip)
    _check(name, images.cpu().numpy(), meta, g)
