"""CPU: pins the DCNv2 oracle (oracle/dcn_ref.py) with the reference's own known-answer test
(DCNv2/test.py:32-65: zero offsets + identity centre tap + mask 0.5 => 2*DCN(x) == x) and a
cross-check against torchvision.ops.deform_conv2d (third party, same offset layout / bilinear rule)."""
import pytest
import torch

from oracle.dcn_ref import dcn_v2_forward


def test_zero_offset_identity_kat():
    torch.manual_seed(0)
    N, C, H, W = 2, 2, 4, 4
    x = torch.randn(N, C, H, W)
    w = torch.zeros(C, C, 3, 3)
    for i in range(C):
        w[i, i, 1, 1] = 1.0
    out = dcn_v2_forward(x, torch.zeros(N, 18, H, W), torch.full((N, 9, H, W), 0.5), w, torch.zeros(C))
    assert (out * 2 - x.double()).abs().max().item() < 1e-10          # test.py:62 bound


@pytest.mark.parametrize("stride,dg", [(1, 1), (1, 2), (2, 2)])
def test_against_torchvision(stride, dg):
    tv = pytest.importorskip("torchvision")
    from torchvision.ops import deform_conv2d
    torch.manual_seed(1)
    B, Ci, H, W, Co = 2, 6, 9, 11, 5
    Ho = (H + 2 - 3) // stride + 1
    Wo = (W + 2 - 3) // stride + 1
    x = torch.randn(B, Ci, H, W)
    off = torch.randn(B, 18 * dg, Ho, Wo) * 2.5
    m = torch.rand(B, 9 * dg, Ho, Wo)
    w = torch.randn(Co, Ci, 3, 3)
    b = torch.randn(Co)
    a = dcn_v2_forward(x, off, m, w, b, stride, 1, 1, dg)
    t = deform_conv2d(x.double(), off.double(), w.double(), b.double(), stride=stride, padding=1, dilation=1,
                      mask=m.double())
    assert (a - t).abs().max().item() < 1e-10
