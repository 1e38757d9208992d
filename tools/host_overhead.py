"""Dev tool: host-side enqueue time of one sag_forward (CPU cost of the orchestration + launches) vs its GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spatialaudiogen_b200 import SptAudioGen, weights as Wt
enc = ['audio', 'video']
B = int(os.environ.get('B', '32'))
m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(Wt.init_weights(enc, separation='unet_mask', seed=1))
a = torch.randn(B, 52799, 1, device='cuda') * 0.1
v = torch.rand(B, 1, 224, 448, 3, device='cuda') - 0.5
out = torch.empty(B, 4800, 3, device='cuda')
for _ in range(3):
    m.forward_into(a, v, None, out)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    m.forward_into(a, v, None, out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('B=%d  host enqueue %.3f ms/forward, total %.3f ms/forward' % (B, 1e3 * (t1 - t0) / N, 1e3 * (t2 - t0) / N))
