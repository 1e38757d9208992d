"""One decode of 32 frames of 224x448 (4:2:0, quality 90) with device and with host entropy decoding: the workload of the ncu pass over
the jpeg_* kernels (tools/gpu_r2_final.sh)."""
import io
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialaudiogen_b200 import readers as R

files = []
for i in range(32):
    rng = np.random.RandomState(i)
    y, x = np.mgrid[0:224, 0:448]
    img = np.stack([127 + 100 * np.sin(x / 17. + y / 29. + i), 127 + 90 * np.cos(x / 11. - y / 23.), 127 + 80 * np.sin(x / 7.) * np.cos(y / 13.)], -1)
    b = io.BytesIO()
    Image.fromarray(np.clip(img + rng.randn(224, 448, 3) * 10, 0, 255).astype(np.uint8)).save(b, 'JPEG', quality=90, subsampling=2)
    files.append(b.getvalue())
print('compressed bytes per batch: %d; coefficient bytes: %d; RGB bytes: %d' % (sum(map(len, files)), 32 * 224 * 448 * 3, 32 * 224 * 448 * 3))
for dh in (True, False):
    dec = R.JpegDecoder(32, 224, 448, device_huffman=dh)
    for _ in range(2):
        out = dec.decode(files)
    torch.cuda.synchronize()
    if dh:
        print('synchronisation rounds per frame:', sorted(dec.sync_rounds(32)))
