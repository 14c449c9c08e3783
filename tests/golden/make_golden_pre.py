"""Generates tests/golden/pre.npz with the UNMODIFIED BaseDetector.pre_process of the reference
(detectors/base_detector.py:37-65; cv2 from this image) on seeded images -- build container only."""
import os, sys, types
import numpy as np

REF = "/root/reference/src/lib"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
for name, attrs in (("progress", {}), ("progress.bar", {"Bar": object}), ("matplotlib", {}), ("matplotlib.pyplot", {}),
                    ("mpl_toolkits", {}), ("mpl_toolkits.mplot3d", {"Axes3D": object})):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
import torch  # noqa: E402
torch.utils.model_zoo = __import__("torch.utils.model_zoo", fromlist=["x"])
sys.modules.setdefault("models.networks.DCNv2._ext", types.ModuleType("models.networks.DCNv2._ext"))
# models/model.py imports the DCN networks eagerly and the reference's cffi extension cannot be built: give it a
# stand-in module (pre_process itself touches none of this)
_dcn = types.ModuleType("models.networks.DCNv2.dcn_v2"); _dcn.DCN = object
sys.modules["models.networks.DCNv2.dcn_v2"] = _dcn
from detectors.base_detector import BaseDetector  # noqa: E402


class Opt(object):
    pass


def run(image, fix_res, flip, input_h=512, input_w=512, scale=1):
    d = types.SimpleNamespace()
    d.opt = Opt(); d.opt.fix_res = fix_res; d.opt.input_h = input_h; d.opt.input_w = input_w; d.opt.pad = 31
    d.opt.flip_test = flip; d.opt.down_ratio = 4
    d.mean = np.array([0.408, 0.447, 0.470], np.float32).reshape(1, 1, 3)
    d.std = np.array([0.289, 0.274, 0.278], np.float32).reshape(1, 1, 3)
    images, meta = BaseDetector.pre_process(d, image, scale)
    return images.numpy(), meta


def main():
    import zlib
    sys.path.insert(0, os.path.dirname(HERE))
    import big_inputs as BI
    out = {}
    for name, (h, w, fix, flip, ih, iw, sc) in BI.PRE_CASES.items():
        img = BI.pre_image(name)
        images, meta = run(img, fix, flip, ih or 512, iw or 512, sc)
        # inputs are rebuilt by tests/big_inputs.py; of the output (3-6 MB of float32 per case) the CRC and a
        # strided sample are stored
        out["img_crc_" + name] = np.uint32(zlib.crc32(np.ascontiguousarray(img).view(np.uint8)))
        out["crc_" + name] = np.uint32(zlib.crc32(np.ascontiguousarray(images).view(np.uint8)))
        out["shape_" + name] = np.array(images.shape, np.int64)
        out["sample_" + name] = images[:, :, ::9, ::7].copy()
        out["c_" + name] = np.asarray(meta["c"], np.float32); out["s_" + name] = np.asarray(meta["s"], np.float32).reshape(-1)
        out["hw_" + name] = np.array([meta["out_height"], meta["out_width"]], np.int64)
    np.savez_compressed(os.path.join(HERE, "pre.npz"), **out)
    print("wrote pre.npz", {k: tuple(int(x) for x in v) for k, v in out.items() if k.startswith("shape_")})


if __name__ == "__main__":
    main()
