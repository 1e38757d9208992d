"""Dev tool: run-to-run spread of the end-to-end loop at three lanes (same process, repeated)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spatialaudiogen_b200 import SptAudioGen, weights as Wt
enc = ['audio', 'video']
B, N = 32, int(os.environ.get('N', '100'))
m = SptAudioGen(1, encoders=enc, separation='unet_mask', precision='mixed').load_weights(Wt.init_weights(enc, separation='unet_mask', seed=1))
rng = np.random.RandomState(0)
host = [{'audio': torch.as_tensor((rng.randn(B, 52799, 1) * 0.1).astype(np.float32)).pin_memory(),
         'video': torch.as_tensor(rng.randint(0, 256, (B, 1, 224, 448, 3)).astype(np.uint8)).pin_memory()} for _ in range(6)]


def batches(n):
    for i in range(n):
        yield host[i % 6]


def run(lanes, depth, n):
    t0 = time.perf_counter()
    host_t = 0.0
    acc = 0.0
    g = m.inference_stream(batches(n), depth=depth, lanes=lanes)
    while True:
        h0 = time.perf_counter()
        try:
            y = next(g)
        except StopIteration:
            break
        host_t += time.perf_counter() - h0
        acc += float(y[0, 0, 0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 0.1 * B * n / dt, 1e3 * host_t / n


for lanes, depth in ((3, 3), (3, 3), (3, 3), (3, 3), (3, 3), (3, 3), (2, 3), (2, 3), (2, 3), (3, 9), (3, 9), (3, 9)):
    run(lanes, depth, 12)
    r, h = run(lanes, depth, N)
    print('lanes=%d depth=%d: %.1f audio-s/s (time inside the generator %.3f ms per step)' % (lanes, depth, r, h), flush=True)
