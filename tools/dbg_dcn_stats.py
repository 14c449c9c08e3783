"""Cycle breakdown of one CTA of the tcgen05 DCNv2 forward (debug counters in dcnv2_tc.cu)."""
import ctypes, sys
import torch
sys.path.insert(0, '.')
from centernet_b200._lib import C, ptr, stream_ptr, workspace
lib = ctypes.CDLL('centernet_b200/lib/libcenternet_b200.so')
for (B, Cin, H, W, Cout) in [(16, 64, 128, 128, 64), (16, 256, 32, 32, 256)]:
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(B, Cin, H, W, device='cuda', generator=g)
    off = torch.randn(B, 18, H, W, device='cuda', generator=g) * 2
    msk = torch.sigmoid(torch.randn(B, 9, H, W, device='cuda', generator=g))
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) / (3 * Cin ** 0.5)
    bias = torch.randn(Cout, device='cuda', generator=g)
    out = torch.empty(B, Cout, H, W, device='cuda')
    ws = workspace(C.dcnv2_workspace_bytes(B, Cin, Cout, H, W, 3, 3, 1, 1, 1, 1), x.device)
    for _ in range(3):
        C.dcnv2_forward(ptr(x), ptr(off), ptr(msk), ptr(w), ptr(bias), ptr(out), B, Cin, H, W, Cout, 3, 3, 1, 1, 1, 1, 1, 1, 1, ptr(ws), ws.numel(), stream_ptr(x))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    lib.cnb_debug_dcn_stats(buf)
    names = ['total', 'meta', 'wait_mma', 'sample+sync', 'wait_B', 'issue', 'epilogue', 'chunks']
    print((B, Cin, H, W, Cout), {n: int(v) for n, v in zip(names, buf)})
