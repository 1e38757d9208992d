"""Host emulation of the streaming BN kernels' per-thread code (no GPU, no oracle).

The bodies of `bn_apply_stats_kernel`, `bn_relu_maxpool_stats_kernel` and `space_to_depth16_kernel` (csrc/pointwise.cu) -- their grid-stride /
unrolled loops, index arithmetic, clamped taps and the split-bf16 load / store helpers -- are cut out of the shipped
source TEXT, the CUDA keywords are mapped onto plain C++ (threadIdx / blockIdx become globals that a host loop walks
through every thread of the grid), compiled with g++ and compared with numpy formulas of the same operators
(reference core.py:209-210 batch-norm with batch statistics, resnet.py:135 3x3/2 SAME max-pool; the 2x2 space-to-depth
of the zero-bordered frame that turns resnet.py:133's 7x7/2 conv into a 4x4/1 one).  It checks exactly what
a GPU-less change to those loops can break: every element visited once, right channel, right tap set, right planes."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, '..', 'spatialaudiogen_b200', 'csrc', 'pointwise.cu')
CUDA_INC = '/usr/local/cuda/include'

PRELUDE = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <cuda_bf16.h>
#include <vector_types.h>
#include <vector_functions.h>
using std::min; using std::max;
struct Dim { unsigned x = 0, y = 0, z = 0; };
static Dim threadIdx, blockIdx, blockDim, gridDim;
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
#define __ldg(p) (*(p))
#define __ldcs(p) (*(p))
enum { ACT_F32 = 0, ACT_BF2 = 1 };
struct ActView { void* p = nullptr; int fmt = ACT_F32; int64_t plane = 0; };
struct BnStats { const double* sum = nullptr; const double* sqs = nullptr; const float* gamma = nullptr; const float* beta = nullptr;
                 double inv_count = 0.0; float eps = 1e-3f; };
static float s_ss[16384];
enum { FRAMES_F32 = 0, FRAMES_U8_VIDEO = 1, FRAMES_U8_FLOW = 2, FRAMES_U8_VIDEO_INT = 3 };
static void emu_fill_lut(float* lut);
// the block prologue (bn_scale_shift_to_smem) needs a barrier: here every emulated thread fills the whole table itself
static void emu_scale_shift(const BnStats& bn, int c, float* s_scale, float* s_shift) {
  for (int i = 0; i < c; ++i) {
    const double mean = bn.sum[i] * bn.inv_count;
    double var = bn.sqs[i] * bn.inv_count - mean * mean;
    if (var < 0) var = 0;
    const double sc = (double)bn.gamma[i] / std::sqrt(var + (double)bn.eps);
    s_scale[i] = (float)sc;
    s_shift[i] = (float)((double)bn.beta[i] - mean * sc);
  }
}
'''

HARNESS = r'''
template <class F> static void for_each_thread(int grid, int block, F f) {
  gridDim.x = grid; blockDim.x = block;
  for (int b = 0; b < grid; ++b) for (int t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; f(); }
}
static BnStats mk(const double* sum, const double* sqs, const float* g, const float* be, double inv) {
  BnStats s; s.sum = sum; s.sqs = sqs; s.gamma = g; s.beta = be; s.inv_count = inv; return s;
}
static ActView view(void* p, int fmt, int64_t plane) { ActView v; v.p = p; v.fmt = fmt; v.plane = plane; return v; }
extern "C" void emu_bn_apply(const float* x, const double* sum, const double* sqs, const float* g, const float* be, double inv,
                             void* res, int res_fmt, int64_t res_plane, int relu, void* y, int y_fmt, int64_t y_plane, int64_t n4,
                             int c, int grid, int block) {
  for_each_thread(grid, block, [&] {
    bn_apply_stats_kernel<true>(reinterpret_cast<const float4*>(x), mk(sum, sqs, g, be, inv), view(res, res_fmt, res_plane), relu,
                          view(y, y_fmt, y_plane), n4, c);
  });
}
extern "C" void emu_maxpool(const float* x, const double* sum, const double* sqs, const float* g, const float* be, double inv, int n,
                            int h, int w, int c, int oh, int ow, int pt, int pl, void* y, int y_fmt, int64_t y_plane, int grid, int block) {
  for_each_thread(grid, block, [&] {
    bn_relu_maxpool_stats_kernel(x, mk(sum, sqs, g, be, inv), n, h, w, c, oh, ow, pt, pl, view(y, y_fmt, y_plane));
  });
}
// the block-wide table fill of the kernel (one entry per thread, then a barrier): here one emulated thread fills it all
static void emu_fill_lut(float* lut) {
  const Dim t = threadIdx, b = blockDim;
  threadIdx.x = 0; blockDim.x = 1;
  frame_lut_to_smem(lut);
  threadIdx = t; blockDim = b;
}
extern "C" void emu_s2d(const void* x, const double* lims, int kind, int n, int h, int w, int c, int pt, int pl, int h2, int w2, void* y,
                        int y_fmt, int64_t y_plane, int grid, int block) {
  for_each_thread(grid, block, [&] {
    if (kind == FRAMES_U8_VIDEO) space_to_depth16_kernel<3, FRAMES_U8_VIDEO>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else if (kind == FRAMES_U8_VIDEO_INT) space_to_depth16_kernel<3, FRAMES_U8_VIDEO_INT>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else if (kind == FRAMES_U8_FLOW) space_to_depth16_kernel<3, FRAMES_U8_FLOW>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else if (c == 1) space_to_depth16_kernel<1, FRAMES_F32>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else if (c == 2) space_to_depth16_kernel<2, FRAMES_F32>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else if (c == 3) space_to_depth16_kernel<3, FRAMES_F32>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
    else space_to_depth16_kernel<4, FRAMES_F32>(x, lims, n, h, w, pt, pl, h2, w2, view(y, y_fmt, y_plane));
  });
}
'''


def _cut(text, start, end):
    a = text.index(start)
    return text[a:text.index(end, a)]


def _host_source():
    t = open(SRC).read()
    parts = [_cut(t, '__device__ __forceinline__ void store_act4(', 'static int g_num_sms'),
             _cut(t, 'constexpr int BN_UNROLL', 'int launch_bn_apply_stats('),
             _cut(t, '__global__ void __launch_bounds__(256, 4) bn_relu_maxpool_stats_kernel(', 'int launch_bn_relu_maxpool_stats('),
             _cut(t, 'template <int C, int KIND>\n__device__ __forceinline__ void load_frame_pixel(', 'template <int C>\nstatic void launch_s2d_kind(')
             .replace('__syncthreads();', '').replace('__shared__ float s_lut[256];', 'static float s_lut[256];')
             .replace('frame_lut_to_smem(s_lut);', 'emu_fill_lut(s_lut);')]
    body = '\n'.join(parts)
    body = re.sub(r'__global__\s+void\s+(__launch_bounds__\([^)]*\)\s*)?', 'static void ', body)
    body = body.replace('__device__ __forceinline__', 'static inline').replace('extern __shared__ float s_ss[];', '')
    body = body.replace('pdl_prologue();', '').replace('bn_scale_shift_to_smem(', 'emu_scale_shift(')
    assert '__syncthreads' not in body and '<<<' not in body
    return PRELUDE + body + HARNESS


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    if shutil.which('g++') is None or not os.path.exists(os.path.join(CUDA_INC, 'cuda_bf16.h')):
        pytest.skip('needs g++ and the CUDA headers')
    d = tmp_path_factory.mktemp('emu')
    src, lib = str(d / 'emu.cpp'), str(d / 'libemu.so')
    open(src, 'w').write(_host_source())
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-w', '-I', CUDA_INC, src, '-o', lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _bf16_round(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000                       # round to nearest even
    return u.astype(np.uint32).view(np.float32)


def _split(x):
    hi = _bf16_round(x)
    return hi, _bf16_round(x - hi)


def _planes_to_f32(buf, n):
    hi = (buf[:n].astype(np.uint32) << 16).view(np.float32)
    lo = (buf[n:2 * n].astype(np.uint32) << 16).view(np.float32)
    return hi, lo


def _stats(x2d, rng):
    c = x2d.shape[1]
    s, q = x2d.astype(np.float64).sum(0), (x2d.astype(np.float64) ** 2).sum(0)
    g, b = rng.uniform(-1.5, 1.5, c).astype(np.float32), rng.randn(c).astype(np.float32)
    inv = 1.0 / x2d.shape[0]
    mean = s * inv
    var = np.maximum(q * inv - mean * mean, 0)
    sc = (g.astype(np.float64) / np.sqrt(var + np.float64(np.float32(1e-3)))).astype(np.float32)
    sh = (b.astype(np.float64) - mean * (g.astype(np.float64) / np.sqrt(var + np.float64(np.float32(1e-3))))).astype(np.float32)
    return s, q, g, b, inv, sc, sh


@pytest.mark.parametrize('rows,c,grid,block', [(37, 8, 3, 32), (5, 4, 1, 1), (129, 64, 2, 64), (1000, 16, 7, 32), (3, 12, 4, 8)])
@pytest.mark.parametrize('res_fmt', [None, 0, 1])
@pytest.mark.parametrize('y_fmt', [0, 1])
@pytest.mark.parametrize('relu', [0, 1])
def test_bn_apply_thread_code(emu, rows, c, grid, block, res_fmt, y_fmt, relu):
    rng = np.random.RandomState(rows * 131 + c)
    x = rng.randn(rows, c).astype(np.float32)
    s, q, g, b, inv, sc, sh = _stats(x, rng)
    n = rows * c
    res = rng.randn(rows, c).astype(np.float32)
    want = (x.astype(np.float64) * sc + sh).astype(np.float32)
    res_buf, res_plane, rfmt = None, 0, 0
    if res_fmt == 0:
        res_buf, want = res.copy(), want + res
    elif res_fmt == 1:
        hi, lo = _split(res)
        res_buf = np.concatenate([(hi.view(np.uint32) >> 16).astype(np.uint16).ravel(), (lo.view(np.uint32) >> 16).astype(np.uint16).ravel()])
        res_plane, rfmt, want = 2 * n, 1, want + (hi + lo)
    if relu:
        want = np.maximum(want, 0)
    y = np.full(n, np.nan, np.float32) if y_fmt == 0 else np.full(2 * n, 0xffff, np.uint16)
    emu.emu_bn_apply(_p(x), _p(s), _p(q), _p(g), _p(b), C.c_double(inv), _p(res_buf), rfmt, C.c_int64(res_plane), relu, _p(y), y_fmt,
                     C.c_int64(2 * n if y_fmt else 0), C.c_int64(n // 4), c, grid, block)
    if y_fmt == 0:
        assert np.allclose(y.reshape(rows, c), want, rtol=2e-6, atol=2e-7)
    else:
        hi, lo = _planes_to_f32(y, n)
        assert np.abs(lo).max() <= np.abs(hi).max() * 2.0 ** -8                               # lo is the bf16 remainder of hi
        assert np.allclose((hi + lo).reshape(rows, c), want, rtol=2e-5, atol=1e-6)


def _same(nin, k, s):
    out = -(-nin // s)
    tot = max((out - 1) * s + k - nin, 0)
    return tot // 2, out


@pytest.mark.parametrize('n,h,w,c,grid,block', [(2, 7, 9, 8, 3, 32), (1, 8, 6, 4, 1, 7), (3, 1, 1, 4, 2, 2), (1, 2, 5, 12, 5, 16), (2, 112 // 8, 224 // 8, 8, 4, 64)])
@pytest.mark.parametrize('y_fmt', [0, 1])
def test_bn_relu_maxpool_thread_code(emu, n, h, w, c, grid, block, y_fmt):
    rng = np.random.RandomState(h * 17 + w)
    x = rng.randn(n, h, w, c).astype(np.float32)
    s, q, g, b, inv, sc, sh = _stats(x.reshape(-1, c), rng)
    (pt, oh), (pl, ow) = _same(h, 3, 2), _same(w, 3, 2)
    a = np.maximum((x.astype(np.float64) * sc + sh).astype(np.float32), 0)
    padded = np.full((n, h + 3, w + 3, c), -np.inf, np.float32)
    padded[:, pt:pt + h, pl:pl + w] = a
    want = np.full((n, oh, ow, c), -np.inf, np.float32)
    for dy in range(3):
        for dx in range(3):
            want = np.maximum(want, padded[:, dy:dy + 2 * oh:2, dx:dx + 2 * ow:2][:, :oh, :ow])
    m = n * oh * ow * c
    y = np.full(m, np.nan, np.float32) if y_fmt == 0 else np.full(2 * m, 0xffff, np.uint16)
    emu.emu_maxpool(_p(x), _p(s), _p(q), _p(g), _p(b), C.c_double(inv), n, h, w, c, oh, ow, pt, pl, _p(y), y_fmt, C.c_int64(2 * m if y_fmt else 0),
                    grid, block)
    if y_fmt == 0:
        assert np.allclose(y.reshape(want.shape), want, rtol=2e-6, atol=2e-7)
    else:
        hi, lo = _planes_to_f32(y, m)
        assert np.allclose((hi + lo).reshape(want.shape), want, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize('n,h,w,c,pt,pl,grid,block', [(2, 8, 12, 3, 2, 2, 3, 16), (1, 7, 5, 3, 3, 3, 2, 8), (1, 6, 6, 1, 0, 0, 1, 5), (2, 5, 9, 4, 2, 3, 4, 32),
                                                     (1, 224 // 8, 448 // 8, 3, 2, 2, 2, 64), (1, 4, 4, 2, 1, 0, 7, 3)])
@pytest.mark.parametrize('y_fmt', [0, 1])
def test_space_to_depth16_thread_code(emu, n, h, w, c, pt, pl, grid, block, y_fmt):
    rng = np.random.RandomState(h * 31 + w + c)
    x = rng.randn(n, h, w, c).astype(np.float32)
    h2, w2 = (h + pt + 3) // 2 + 1, (w + pl + 3) // 2 + 1                  # a frame with a zero border on every side
    big = np.zeros((n, 2 * h2 + 2, 2 * w2 + 2, c), np.float32)
    big[:, pt:pt + h, pl:pl + w] = x
    want = np.zeros((n, h2, w2, 16), np.float32)
    for sub in range(4):
        want[..., sub * c:(sub + 1) * c] = big[:, (sub >> 1):(sub >> 1) + 2 * h2:2, (sub & 1):(sub & 1) + 2 * w2:2][:, :h2, :w2]
    m = n * h2 * w2 * 16
    y = np.full(m, np.nan, np.float32) if y_fmt == 0 else np.full(2 * m, 0xffff, np.uint16)
    emu.emu_s2d(_p(x), None, 0, n, h, w, c, pt, pl, h2, w2, _p(y), y_fmt, C.c_int64(2 * m if y_fmt else 0), grid, block)
    if y_fmt == 0:
        assert np.array_equal(y.reshape(want.shape), want)
    else:
        hi, lo = _planes_to_f32(y, m)
        assert np.allclose((hi + lo).reshape(want.shape), want, rtol=2e-5, atol=0) and np.array_equal((hi + lo).reshape(want.shape) == 0, want == 0)


@pytest.mark.parametrize('kind', [1, 2, 3])
def test_uint8_frame_ingest_thread_code(emu, kind):
    """The uint8 frame sources of the ingest kernel against the reference's host preparation written out in numpy:
    video = img_prep_fcn (myutils.py:88-89: x/255. - 0.5 in float64, rounded to float32 when fed) -- bit-exact;
    flow = FlowReader.get_by_index (feeder.py:147-161) on a float32 chunk with float64 limits -- to float32 sin/cos accuracy."""
    rng = np.random.RandomState(kind)
    n, h, w, c, pt, pl = 2, 6, 10, 3, 2, 2
    u = rng.randint(0, 256, size=(n, h, w, c)).astype(np.uint8)
    u[0, 0, 0] = (0, 0, 0)
    u[0, 0, 1] = (255, 255, 255)
    lims = np.stack([rng.uniform(0, 2, n), rng.uniform(5, 30, n)], 1).astype(np.float64)
    if kind == 1:
        x = (u / 255. - 0.5).astype(np.float32)
    elif kind == 3:
        x = 2. * u.astype(np.float32) - 255.                 # the integer form: 510 * (u/255 - 0.5), exact in bf16
    else:
        chunk = u.astype(np.float32)
        m_min, m_max = lims[:, 0].reshape((-1, 1, 1)), lims[:, 1].reshape((-1, 1, 1))
        chunk[:, :, :, 2] *= (m_max - m_min) / 255.
        chunk[:, :, :, 2] += m_min
        chunk[:, :, :, 0] *= (2 * np.pi) / 255.
        chunk[:, :, :, 1] = chunk[:, :, :, 2] * np.sin(chunk[:, :, :, 0])
        chunk[:, :, :, 0] = chunk[:, :, :, 2] * np.cos(chunk[:, :, :, 0])
        x = chunk
    h2, w2 = (h + pt + 3) // 2 + 1, (w + pl + 3) // 2 + 1
    big = np.zeros((n, 2 * h2 + 2, 2 * w2 + 2, c), np.float32)
    big[:, pt:pt + h, pl:pl + w] = x
    want = np.zeros((n, h2, w2, 16), np.float32)
    for sub in range(4):
        want[..., sub * c:(sub + 1) * c] = big[:, (sub >> 1):(sub >> 1) + 2 * h2:2, (sub & 1):(sub & 1) + 2 * w2:2][:, :h2, :w2]
    y = np.full(n * h2 * w2 * 16, np.nan, np.float32)
    emu.emu_s2d(_p(u), _p(lims), kind, n, h, w, c, pt, pl, h2, w2, _p(y), 0, C.c_int64(0), 3, 16)
    if kind in (1, 3):
        assert np.array_equal(y.reshape(want.shape), want)
        if kind == 3:
            import torch
            assert torch.equal(torch.from_numpy(want).bfloat16().float(), torch.from_numpy(want))     # one bf16 plane holds it exactly
    else:
        assert np.allclose(y.reshape(want.shape), want, rtol=2e-6, atol=2e-6)
        assert np.array_equal(y.reshape(want.shape)[..., 2::3][..., :4], want[..., 2::3][..., :4])      # magnitudes: exact
