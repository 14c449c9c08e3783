"""Where does the host->device->decode->host path spend its time?  (GPU box)"""
import os, sys, time
import torch
sys.path.insert(0, '.')
from centernet_b200 import decode as D
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator().manual_seed(1)
heat = torch.sigmoid(torch.randn(B, C, H, W, generator=g) - 2.19).pin_memory()
wh = (torch.rand(B, 2, H, W, generator=g) * 32).pin_memory()
reg = torch.rand(B, 2, H, W, generator=g).pin_memory()
dev = torch.device('cuda', 0)
dh = torch.empty(B, C, H, W, device=dev)

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

print('plain pinned copy 336MB: %.2f ms' % t(lambda: dh.copy_(heat, non_blocking=True)))
def chunked():
    for s in range(0, B, 8):
        dh[s:s + 8].copy_(heat[s:s + 8], non_blocking=True)
print('chunked pinned copy: %.2f ms' % t(chunked))
cs = torch.cuda.Stream()
def chunked_side():
    with torch.cuda.stream(cs):
        for s in range(0, B, 8):
            dh[s:s + 8].copy_(heat[s:s + 8], non_blocking=True)
    torch.cuda.current_stream().wait_stream(cs)
print('chunked copy on side stream: %.2f ms' % t(chunked_side))
print('decode_from_host: %.2f ms' % t(lambda: D.ctdet_decode_from_host(heat, wh, reg=reg, K=K)))
for chunk in (4, 16, 64):
    print('decode_from_host chunk=%d: %.2f ms' % (chunk, t(lambda: D.ctdet_decode_from_host(heat, wh, reg=reg, K=K, chunk=chunk))))
hd = heat.to(dev); wd = wh.to(dev); rd = reg.to(dev)
print('device decode B=64: %.3f ms' % t(lambda: D.ctdet_decode(hd, wd, reg=rd, K=K), 20))
print('device decode B=8: %.3f ms' % t(lambda: D.ctdet_decode(hd[:8], wd[:8], reg=rd[:8], K=K), 20))
