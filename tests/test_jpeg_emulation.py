"""Host emulation of the JPEG pixel kernels' per-thread code (no GPU).

`jpeg_idct_kernel` (dequantise + libjpeg's islow inverse DCT) and `jpeg_rgb_kernel` (fancy chroma upsampling + YCbCr -> RGB) are
cut out of the shipped source TEXT of csrc/jpeg.cu, the CUDA keywords are mapped onto plain C++ (threadIdx / blockIdx become
globals that a host loop walks through every thread of the grid), compiled with g++ and run on the coefficients that libsag's
own entropy decoder produces -- the result must equal PIL's decode of the same file, bit for bit.  The inverse DCT has one
block-wide barrier between its column and row passes: the emulation runs every thread of a block up to the barrier (the
`__syncthreads()` becomes an early return in phase 0), then every thread again from the top in phase 1 (the column pass is
recomputed identically and the barrier is a no-op), which is exactly the visibility a barrier gives.  (The entropy decoder's
kernel is emulated through the C ABI instead: sag_jpeg_coefficients_parallel runs its __host__ __device__ phases, test_jpeg.py.)"""
import ctypes as C
import io
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
from PIL import Image

from spatialaudiogen_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, '..', 'spatialaudiogen_b200', 'csrc', 'jpeg.cu')
CUDA_INC = '/usr/local/cuda/include'

PRELUDE = r'''
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector_types.h>
#include <vector_functions.h>
using std::min; using std::max;
struct Dim { unsigned x = 0, y = 0, z = 0; };
static Dim threadIdx, blockIdx, blockDim, gridDim;
static int emu_phase = 0;
'''

HARNESS = r'''
extern "C" void emu_idct(const JpegImage* images, const int16_t* coef, const uint16_t* qt, uint8_t* planes, int grid_x, int n) {
  for (int img = 0; img < n; ++img)
    for (int bx = 0; bx < grid_x; ++bx)
      for (emu_phase = 0; emu_phase < 2; ++emu_phase)
        for (int t = 0; t < 256; ++t) {
          blockIdx.x = bx; blockIdx.y = img; threadIdx.x = t; blockDim.x = 256;
          jpeg_idct_kernel(images, coef, qt, planes);
        }
}
extern "C" void emu_rgb(const JpegImage* images, const uint8_t* planes, int width, int height, uint8_t* frames, int n) {
  for (int img = 0; img < n; ++img)
    for (int y = 0; y < height; ++y)
      for (int bx = 0; bx < (width + 255) / 256; ++bx)
        for (int t = 0; t < 64; ++t) {
          blockIdx.x = bx; blockIdx.y = y; blockIdx.z = img; threadIdx.x = t; blockDim.x = 64;
          jpeg_rgb_kernel(images, planes, width, height, frames);
        }
}
extern "C" int emu_sizeof_image() { return (int)sizeof(JpegImage); }
'''


def _cut(text, start, end):
    a = text.index(start)
    return text[a:text.index(end, a)]


def _host_source():
    t = open(SRC).read()
    body = _cut(t, 'struct JpegImage {', '// ---- parallel entropy decoding on the device')
    body = re.sub(r'__global__\s+void\s+(__launch_bounds__\([^)]*\)\s*)?', 'static void ', body)
    body = body.replace('__device__ __forceinline__', 'static inline').replace('__shared__', 'static')
    body = body.replace('__syncthreads();', 'if (emu_phase == 0) return;')
    assert '<<<' not in body and body.count('if (emu_phase == 0) return;') == 1
    return PRELUDE + body + HARNESS


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    if shutil.which('g++') is None or not os.path.exists(os.path.join(CUDA_INC, 'vector_types.h')):
        pytest.skip('needs g++ and the CUDA headers')
    d = tmp_path_factory.mktemp('jpeg_emu')
    src, lib = str(d / 'emu.cpp'), str(d / 'libemu.so')
    open(src, 'w').write(_host_source())
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-w', '-I', CUDA_INC, src, '-o', lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(lib)


class JpegImage(C.Structure):                      # mirrors csrc/jpeg.cu (checked against sizeof below)
    _fields_ = [('ncomp', C.c_int), ('hmax', C.c_int), ('vmax', C.c_int), ('h', C.c_int * 3), ('v', C.c_int * 3), ('bw', C.c_int * 3),
                ('bh', C.c_int * 3), ('coef_off', C.c_longlong * 3), ('plane_off', C.c_longlong * 3), ('block_base', C.c_longlong * 4)]


def _picture(h, w, seed):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(x / 17. + y / 29. + seed), 127 + 90 * np.cos(x / 11. - y / 23.), 127 + 80 * np.sin(x / 7.) * np.cos(y / 13.)], -1)
    return np.clip(img + rng.randn(h, w, 3) * 12, 0, 255).astype(np.uint8)


CASES = [(64, 80, 2, 90), (37, 53, 2, 60), (41, 67, 1, 95), (33, 49, 0, 75), (16, 16, 2, 30), (48, 64, 2, 100), (40, 56, None, 80),
         (40, 3, 2, 80), (9, 4, 1, 60), (17, 1, 2, 90), (5, 2, 2, 50)]          # narrow: replicated chroma (jinit_upsampler)


@pytest.mark.parametrize('h,w,ss,q', CASES)
def test_pixel_kernels_thread_code_matches_pil(emu, h, w, ss, q):
    img = _picture(h, w, h + w)
    buf = io.BytesIO()
    if ss is None:
        Image.fromarray(img[:, :, 0]).save(buf, 'JPEG', quality=q)          # a grey file
    else:
        Image.fromarray(img).save(buf, 'JPEG', quality=q, subsampling=ss)
    data = buf.getvalue()
    ref = np.asarray(Image.open(io.BytesIO(data)).convert('RGB'))
    # the coefficients and the block grids, from libsag's own (host) entropy decoder
    coef = np.zeros(3 * ((h + 15) // 16 * 16) * ((w + 15) // 16 * 16), np.int16)
    bw, bh, qt = (C.c_int * 3)(), (C.c_int * 3)(), np.zeros(192, np.uint16)
    L.check(L.lib().sag_jpeg_coefficients(data, len(data), coef.ctypes.data, coef.size, bw, bh, qt.ctypes.data))
    v = [C.c_int() for _ in range(5)]
    L.check(L.lib().sag_jpeg_info(data, len(data), *[C.byref(x) for x in v]))
    ncomp, hmax, vmax = v[2].value, v[3].value, v[4].value
    assert emu.emu_sizeof_image() == C.sizeof(JpegImage)
    im = JpegImage()
    im.ncomp, im.hmax, im.vmax = ncomp, hmax, vmax
    mcux, mcuy = -(-w // (8 * hmax)), -(-h // (8 * vmax))
    off = blocks = 0
    for c in range(3):
        im.block_base[c] = blocks
        im.h[c] = im.v[c] = 1
        if c < ncomp:
            im.bw[c], im.bh[c] = bw[c], bh[c]
            im.h[c], im.v[c] = bw[c] // mcux, bh[c] // mcuy
            im.coef_off[c] = im.plane_off[c] = off
            off += bw[c] * bh[c] * 64
            blocks += bw[c] * bh[c]
    im.block_base[3] = blocks
    planes = np.zeros(off, np.uint8)
    frames = np.zeros((h, w, 3), np.uint8)
    emu.emu_idct(C.byref(im), coef.ctypes.data_as(C.c_void_p), qt.ctypes.data_as(C.c_void_p), planes.ctypes.data_as(C.c_void_p),
                 (blocks + 31) // 32, 1)
    emu.emu_rgb(C.byref(im), planes.ctypes.data_as(C.c_void_p), w, h, frames.ctypes.data_as(C.c_void_p), 1)
    assert np.array_equal(frames, ref)
