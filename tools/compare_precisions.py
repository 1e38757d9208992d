"""Dev tool: run one forward with the exact-fp32 FFMA path and with a tensor-core precision and print the relative
difference of every intermediate tensor (first divergence localises a kernel bug).  python tools/compare_precisions.py"""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import sag_oracle as O
from spatialaudiogen_b200 import SptAudioGen, weights as Wt
enc=['audio','video']
W=Wt.init_weights(enc, separation='unet_mask', seed=9, stress=True)
rng=np.random.RandomState(0)
B=2
audio=(0.1*rng.randn(B,52799,1)).astype(np.float32)
video=(rng.randint(0,256,size=(B,1,224,448,3))/255.-0.5).astype(np.float32)
m=SptAudioGen(1, encoders=enc, separation='unet_mask', precision='fp32').load_weights(W)
y0=m.inference_ops(torch.as_tensor(audio).cuda(), video=torch.as_tensor(video).cuda()).clone()
e0={k:v.clone() for k,v in m.ends.items()}
m.set_option('precision','bf16x3')
y1=m.inference_ops(torch.as_tensor(audio).cuda(), video=torch.as_tensor(video).cuda()).clone()
e1=m.ends
def rel(a,b): return float((a.double()-b.double()).abs().max()/b.double().abs().max())
for k in e0:
    if k in e1 and e0[k].shape==e1[k].shape:
        print('%-40s %.3e'%(k, rel(e1[k],e0[k])))
print('out', rel(y1,y0))
