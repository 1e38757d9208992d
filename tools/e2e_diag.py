"""Dev tool: where does the end-to-end loop (SptAudioGen.inference_stream: pinned host batches in, host waveforms out) spend its time?
Per setting: audio-s/s over N steps after a warm-up that touches every slot, and the host time of one bare forward_into enqueue."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spatialaudiogen_b200 import SptAudioGen, weights as Wt
enc = ['audio', 'video']
B = int(os.environ.get('B', '32'))
N = int(os.environ.get('N', '90'))
m = SptAudioGen(1, encoders=enc, separation='unet_mask', precision='mixed').load_weights(Wt.init_weights(enc, separation='unet_mask', seed=1))
rng = np.random.RandomState(0)
host = [{'audio': torch.as_tensor((rng.randn(B, 52799, 1) * 0.1).astype(np.float32)).pin_memory(),
         'video': torch.as_tensor(rng.randint(0, 256, (B, 1, 224, 448, 3)).astype(np.uint8)).pin_memory()} for _ in range(4)]
a, v = host[0]['audio'].cuda(), host[0]['video'].cuda()
out = torch.empty(B, 4800, 3, device='cuda')
for _ in range(3):
    m.forward_into(a, v, None, out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    m.forward_into(a, v, None, out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('B=%d bare forward_into: host enqueue %.3f ms, device-bound total %.3f ms per forward' % (B, 1e3 * (t1 - t0) / 20, 1e3 * (t2 - t0) / 20))


def batches(n):
    for i in range(n):
        yield host[i % 4]


for lanes, depth, graph in ((1, 3, False), (3, 3, False), (3, 6, False), (3, 9, False), (2, 4, False), (3, 3, True), (1, 3, True)):
    for y in m.inference_stream(batches(3 * max(depth, lanes + 1)), depth=depth, lanes=lanes, use_graph=graph):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = 0.0
    for y in m.inference_stream(batches(N), depth=depth, lanes=lanes, use_graph=graph):
        acc += float(y[0, 0, 0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('lanes=%d depth=%d graph=%d: %.1f audio-s/s (%.3f ms per step)' % (lanes, depth, graph, 0.1 * B * N / dt, 1e3 * dt / N))
