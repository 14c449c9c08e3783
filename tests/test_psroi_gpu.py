"""GPU: deformable PSROI pooling (DCNv2Pooling / DCNPooling / DCNv2PoolingFunction) against the oracle, plus the
reference's own checks (DCNv2/test.py:117-166: zero-offset known answer, gradient check)."""
import numpy as np
import pytest
import torch

from oracle import psroi_np as P

pytestmark = pytest.mark.gpu


def _case(seed, B, C, H, W, N, output_dim, group, pooled, part, spp, classes, trans_std, no_trans):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x = rng.uniform(-6, W * 4 - 8, N); y = rng.uniform(-6, H * 4 - 8, N)
    w = rng.uniform(1, W * 3, N); h = rng.uniform(1, H * 3, N)
    rois = np.stack([rng.integers(0, B, N).astype(np.float64), x, y, x + w, y + h], 1).astype(np.float32)
    trans = None if no_trans else rng.standard_normal((N, 2 * classes, part, part)).astype(np.float32)
    cfg = dict(no_trans=no_trans, spatial_scale=0.25, output_dim=output_dim, group_size=group, pooled_size=pooled,
               part_size=part, sample_per_part=spp, trans_std=trans_std)
    return data, rois, trans, cfg


CASES = [
    (1, 2, 8, 9, 9, 5, 2, 2, 3, 3, 2, 2, 0.1, False),
    (2, 2, 18, 12, 10, 6, 2, 3, 4, 2, 3, 1, 0.3, False),   # part_size != pooled_size
    (3, 1, 4, 7, 7, 4, 4, 1, 3, 3, 4, 2, 0.0, False),      # trans_std 0 as in test.py:148-166
    (4, 3, 16, 16, 16, 7, 16, 1, 7, 7, 4, 1, 0.1, True),
]


@pytest.mark.parametrize("case", CASES)
def test_forward_backward_vs_oracle(case):
    from centernet_b200.dcn_v2_func import DCNv2PoolingFunction
    data, rois, trans, cfg = _case(*case)
    want, cnt = P.psroi_forward(data, rois, trans, **cfg)
    fn = DCNv2PoolingFunction(cfg["spatial_scale"], cfg["pooled_size"], cfg["output_dim"], cfg["no_trans"],
                              cfg["group_size"], cfg["part_size"], cfg["sample_per_part"], cfg["trans_std"])
    d = torch.from_numpy(data).cuda().requires_grad_(True)
    r = torch.from_numpy(rois).cuda()
    t = d.new() if trans is None else torch.from_numpy(trans).cuda().requires_grad_(True)
    out = fn(d, r, t)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=0, atol=1e-5)
    go = np.random.default_rng(9).standard_normal(want.shape).astype(np.float32)
    out.backward(torch.from_numpy(go).cuda())
    gd, gt = P.psroi_backward(go, data, rois, trans, cnt, **cfg)
    np.testing.assert_allclose(d.grad.cpu().numpy(), gd, rtol=1e-4, atol=1e-5)
    if trans is not None:
        np.testing.assert_allclose(t.grad.cpu().numpy(), gt, rtol=1e-4, atol=1e-5)


def test_reference_zero_offset_check():
    """DCNv2/test.py:117-146."""
    from centernet_b200.dcn_v2 import DCNv2Pooling
    inp = torch.zeros(2, 16, 64, 64, device="cuda")
    inp[0, :, 16:26, 16:26] = 1.0
    inp[1, :, 10:20, 20:30] = 2.0
    rois = torch.tensor([[0, 65, 65, 103, 103], [1, 81, 41, 119, 79]], device="cuda").float()
    pooling = DCNv2Pooling(spatial_scale=1.0 / 4, pooled_size=7, output_dim=16, no_trans=True, group_size=1, trans_std=0.1).cuda()
    out = pooling(inp, rois, inp.new())
    dpooling = DCNv2Pooling(spatial_scale=1.0 / 4, pooled_size=7, output_dim=16, no_trans=False, group_size=1, trans_std=0.1).cuda()
    dout = dpooling(inp, rois, torch.zeros(20, 2, 7, 7, device="cuda"))
    assert torch.equal(out, dout)
    want, _ = P.psroi_forward(inp.cpu().numpy(), rois.cpu().numpy(), None, True, 0.25, 16, 1, 7, 7, 4, 0.1)
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=1e-6)
    assert abs(out[0].mean().item() - 0.971507) < 1e-5 and abs(out[1].mean().item() - 1.943014) < 1e-5


def test_dcn_pooling_module_and_errors():
    from centernet_b200.dcn_v2 import DCNPooling
    from centernet_b200.dcn_v2_func import DCNv2PoolingFunction
    torch.manual_seed(0)
    data = torch.randn(2, 8, 20, 20, device="cuda", requires_grad=True)
    rois = torch.tensor([[0, 4, 4, 50, 60], [1, 10, 2, 70, 40], [1, 0, 0, 30, 30]], device="cuda").float()
    m = DCNPooling(0.25, 3, 8, False, group_size=1, trans_std=0.1, deform_fc_dim=32).cuda()
    assert {"offset_fc.0.weight", "offset_fc.4.bias", "mask_fc.2.weight"} <= set(m.state_dict().keys())
    y = m(data, rois)
    # zero-initialised heads: offsets 0, mask sigmoid(0) = 0.5  ->  half of the undeformed pooling (dcn_v2.py:150-169)
    plain = DCNv2PoolingFunction(0.25, 3, 8, True)(data, rois, data.new())
    assert torch.allclose(y, plain * 0.5, atol=1e-6)
    y.sum().backward()
    assert data.grad is not None and torch.isfinite(data.grad).all() and m.offset_fc[4].weight.grad is not None
    with pytest.raises(NotImplementedError):
        DCNv2PoolingFunction(0.25, 3, 8, True)(data.detach().cpu(), rois.cpu(), torch.empty(0))
    with pytest.raises(RuntimeError):
        DCNv2PoolingFunction(0.25, 3, 64, True)(data, rois, data.new())      # needs 64 input channels
