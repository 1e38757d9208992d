"""Timing of the frame decode on the GPU box: 32 frames of 224x448 (4:2:0, quality 90), PIL on one host thread vs
readers.JpegDecoder (host entropy decoding on a thread pool + CUDA kernels), and the device part alone (CUDA events)."""
import io
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialaudiogen_b200 import readers as R


def picture(seed):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:224, 0:448]
    img = np.stack([127 + 100 * np.sin(x / 17. + y / 29. + seed), 127 + 90 * np.cos(x / 11. - y / 23.), 127 + 80 * np.sin(x / 7.) * np.cos(y / 13.)], -1)
    return np.clip(img + rng.randn(224, 448, 3) * 10, 0, 255).astype(np.uint8)


files = []
for i in range(32):
    b = io.BytesIO()
    Image.fromarray(picture(i)).save(b, 'JPEG', quality=90, subsampling=2)
    files.append(b.getvalue())
print('32 frames, %.1f KB each on average; host cores: %d' % (sum(map(len, files)) / 32e3, os.cpu_count()))
t = time.perf_counter()
for _ in range(3):
    ref = np.stack([np.asarray(Image.open(io.BytesIO(f)).convert('RGB')) for f in files])
pil_ms = (time.perf_counter() - t) / 3 * 1e3
print('PIL, one thread: %.2f ms per batch (%.3f ms per frame)' % (pil_ms, pil_ms / 32))
for threads in (1, 4, 8, 16, 0):
    dec = R.JpegDecoder(32, 224, 448, threads=threads, device_huffman=False)
    out = torch.empty((32, 224, 448, 3), dtype=torch.uint8, device='cuda')
    for _ in range(3):
        dec.decode(files, out=out)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        dec.decode(files, out=out)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) / 10 * 1e3
    assert np.array_equal(out.cpu().numpy(), ref)
    print('JpegDecoder, host entropy decoding, threads=%2d: %.2f ms per batch end to end (bit-identical to PIL)' % (threads, ms))
for sub in (64, 128, 256, 512):
    for threads in (1, 0):
        dec = R.JpegDecoder(32, 224, 448, threads=threads)
        dec.set_option('sub_bytes', sub)
        for _ in range(3):
            dec.decode(files, out=out)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            dec.decode(files, out=out)
        host_ms = (time.perf_counter() - t) / 10 * 1e3
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t) / 10 * 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec.decode(files, out=out)
        e1.record()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref)
        print('JpegDecoder, device entropy decoding, sub_bytes=%3d host threads=%2d: %.2f ms per batch (host part %.2f ms; device span of one '
              'decode %.3f ms; rounds %s)' % (sub, threads, ms, host_ms, e0.elapsed_time(e1), sorted(set(dec.sync_rounds(32)))))
# device part alone: events around a decode whose host part has already run are not separable through the public call, so time
# two back-to-back decodes and subtract the host time measured with the GPU idle
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dec = R.JpegDecoder(32, 224, 448, device_huffman=False)
torch.cuda.synchronize()
e0.record()
dec.decode(files, out=out)
e1.record()
torch.cuda.synchronize()
print('device span of one decode (H2D of %.1f MB coefficients + idct + upsample/colour kernels): %.3f ms'
      % (32 * 224 * 448 * 1.5 * 2 / 1e6, e0.elapsed_time(e1)))
