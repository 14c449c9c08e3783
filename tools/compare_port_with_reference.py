"""Build-container only (needs /root/reference): runs the UNMODIFIED reference ctdet_decode and the timed CPU port
(oracle/torch_port.py) on the same synthetic batch, checks the outputs are identical and prints both times, so the
`cpu_baseline` / `--impl reference` numbers of bench.py (which cannot import the reference on the GPU box) are
anchored to the real thing."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src/lib')
from models.decode import ctdet_decode as ref_decode          # the reference, unmodified
from oracle.torch_port import ctdet_decode as port_decode

threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
torch.set_num_threads(threads)
torch.manual_seed(317)
B = 64
heat = torch.sigmoid(torch.randn(B, 80, 128, 128) - 2.19)
wh = torch.rand(B, 2, 128, 128) * 32
reg = torch.rand(B, 2, 128, 128)
a = ref_decode(heat.clone(), wh, reg=reg, K=100)
b = port_decode(heat.clone(), wh, reg=reg, K=100)
print('identical outputs:', torch.equal(a, b))
res = {'reference': [], 'port': []}
for rep in range(8):   # alternate the order: the first call of a pair warms the caches for the second
    pair = (('reference', ref_decode), ('port', port_decode))
    for name, fn in (pair if rep % 2 == 0 else pair[::-1]):
        t0 = time.perf_counter(); fn(heat, wh, reg=reg, K=100); res[name].append(time.perf_counter() - t0)
for name, ts in res.items():
    ts.sort()
    print('%-9s B=%d threads=%d: median %.3f s = %.1f img/s (min %.3f s)' % (name, B, threads, ts[len(ts) // 2], B / ts[len(ts) // 2], ts[0]))
