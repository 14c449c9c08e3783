import ctypes, torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from centernet_b200 import decode as D
lib=ctypes.CDLL('/root/repo/centernet_b200/lib/libcenternet_b200.so')
g=torch.Generator(device='cuda').manual_seed(317)
heat=torch.sigmoid(torch.randn(64,80,128,128,device='cuda',generator=g)-2.19); wh=torch.rand(64,2,128,128,device='cuda'); reg=torch.rand(64,2,128,128,device='cuda')
for _ in range(3): D.ctdet_decode(heat,wh,reg=reg,K=100)
buf=torch.zeros(148*8+8,dtype=torch.int64,device='cuda')
lib.cnb_debug_set_select_stats(ctypes.c_void_p(buf.data_ptr()))
D.ctdet_decode(heat,wh,reg=reg,K=100); torch.cuda.synchronize()
lib.cnb_debug_set_select_stats(ctypes.c_void_p(0))
raw=buf.cpu().numpy(); print('finalize(img5): gather, sort, emit, n =', raw[148*8:148*8+4]); a=raw[:148*8].reshape(148,8)
print('cols: total wait boot flush nflush|tfin<<8 t_cut t_sync nfin|nrescan<<8|nprune<<16')
a=a.copy(); tfin=a[:,4]>>8; a[:,4]&=255; import numpy as _n; a=_n.concatenate([a,tfin[:,None]],1)
order=np.argsort(a[:,0])
for i in list(order[:6])+list(order[-10:]): print(i, a[i])
print('mean', a.mean(0))
for nf in (1,2): 
    m=a[a[:,4]==nf]; print('nflush',nf,len(m), m.mean(0))
