"""Dev tool: compare SptAudioGen.inference_stream with per-batch inference_ops calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spatialaudiogen_b200 import SptAudioGen, weights as Wt
enc = ['audio', 'video']
W = Wt.init_weights(enc, separation='unet_mask', seed=5, stress=True)
m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W)
rng = np.random.RandomState(0)
def mk(i):
    r = np.random.RandomState(i)
    return {'audio': torch.as_tensor((0.1 * r.randn(2, 52799, 1)).astype(np.float32)).pin_memory(),
            'video': torch.as_tensor((r.randint(0, 256, size=(2, 1, 224, 448, 3)) / 255. - 0.5).astype(np.float32)).pin_memory()}
batches = [mk(i) for i in range(6)]
refs = [m.inference_ops(b['audio'], video=b['video']).cpu().clone() for b in batches]
refs2 = [m.inference_ops(b['audio'], video=b['video']).cpu().clone() for b in batches]
print('run-to-run', [float((a - b).abs().max() / b.abs().max()) for a, b in zip(refs, refs2)])
for depth in (1, 2, 3):
    outs = [y.clone() for y in m.inference_stream(iter(batches), depth=depth)]
    print('depth', depth, [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, refs)])
