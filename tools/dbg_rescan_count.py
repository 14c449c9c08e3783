"""Counts, per input family, how many CTAs of the hot decode kernel took the exact-rescan and the
sort-prune paths (debug counters of select.cu), to prove the tests actually reach them."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
from centernet_b200 import decode as D
lib = ctypes.CDLL('centernet_b200/lib/libcenternet_b200.so')
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator(device='cuda').manual_seed(5)
noise = torch.sigmoid(torch.randn(B, C, H, W, device='cuda', generator=g) - 2.19)
ramp = (torch.rand(B, C, H, W, device='cuda', generator=g) * 0.01 + torch.linspace(0.01, 0.98, C, device='cuda').view(1, C, 1, 1)).contiguous()
wh = torch.rand(B, 2, H, W, device='cuda')
for name, heat in (('noise', noise), ('ramp', ramp)):
    buf = torch.zeros(148 * 8 + 8, dtype=torch.int64, device='cuda')
    lib.cnb_debug_set_select_stats(ctypes.c_void_p(buf.data_ptr()))
    D.ctdet_decode(heat, wh, K=K); torch.cuda.synchronize()
    lib.cnb_debug_set_select_stats(ctypes.c_void_p(0))
    a = buf.cpu().numpy()[:148 * 8].reshape(148, 8)
    print(name, 'ctas with rescans:', int(((a[:, 7] >> 8) & 255 > 0).sum()), 'total rescans:', int(((a[:, 7] >> 8) & 255).sum()),
          'prunes:', int(((a[:, 7] >> 16) & 255).sum()), 'max cycles:', int(a[:, 0].max()))
