"""GPU parity tests: every CUDA stage and the whole forward, called through the C ABI (libsag.so), against the CPU
oracle on the same seeded inputs.  Tolerances are stated per test; indices are bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import sag_oracle as O
from spatialaudiogen_b200 import weights as Wt

pytestmark = pytest.mark.gpu

RESNET_NPY = '/root/reference/pyutils/tflib/models/image/resnet18.npy'


def _L():
    from spatialaudiogen_b200 import _lib as L
    return L


def _audio(B, seed=0, n=52799):
    rng = np.random.RandomState(seed)
    t = np.arange(n)
    ph = rng.uniform(0, 2 * np.pi, size=(B, 1))
    x = 0.1 * rng.randn(B, n) + 0.3 * np.sin(2 * np.pi * 440 * t / 48000. + ph)
    return np.clip(x, -1, 1).astype(np.float32)[:, :, None]


def _video(B, seed=1, h=224, w=448):
    rng = np.random.RandomState(seed)
    return (rng.randint(0, 256, size=(B, 1, h, w, 3)) / 255. - 0.5).astype(np.float32)


def _flow(B, seed=2, h=224, w=448):
    rng = np.random.RandomState(seed)
    mag = rng.uniform(0, 20, size=(B, 1, h, w))
    th = rng.uniform(0, 2 * np.pi, size=(B, 1, h, w))
    return np.stack([mag * np.cos(th), mag * np.sin(th), mag], -1).astype(np.float32)


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------------------------------------ STFT / iSTFT
def test_stft_matches_oracle_and_frame_indexing_bit_exact():
    from spatialaudiogen_b200 import myutils
    a = _audio(3, 4)[:, :, 0]
    s = myutils.stft(cu(a)[:, None, :], 1024, 4)
    ref = O.stft(torch.as_tensor(a)[:, None, :].double(), 1024, 4)
    assert tuple(s.shape) == (3, 1, 200, 1024)
    assert _rel(torch.view_as_real(s), torch.view_as_real(ref)) < 2e-6
    # bit-exact frame indexing: an impulse at sample n lights exactly the frames with 256t <= n < 256t+1024
    w = O.hann(1024, torch.float32)
    for n in [0, 255, 256, 1023, 1024, 30000, 51967, 51968, 52798]:
        x = torch.zeros(1, 52799)
        x[0, n] = 1.0
        e = myutils.stft(x.cuda(), 1024, 4).abs().sum(-1)[0].cpu()
        hit = set(torch.nonzero(e > 0).flatten().tolist())
        exp = {t for t in range(200) if 256 * t <= n < 256 * t + 1024 and w[n - 256 * t] > 0}
        assert hit == exp, (n, hit, exp)


@pytest.mark.parametrize('wind,ov,n', [(1200, 2, 4800), (2048, 2, 8192), (64, 4, 1000), (1024, 4, 2048)])
def test_stft_other_windows(wind, ov, n):
    from spatialaudiogen_b200 import myutils
    x = np.random.RandomState(0).randn(2, 3, n).astype(np.float32)
    s = myutils.stft(cu(x), wind, ov)
    ref = O.stft(torch.as_tensor(x).double(), wind, ov)
    assert tuple(s.shape) == tuple(ref.shape)
    assert _rel(torch.view_as_real(s), torch.view_as_real(ref)) < 3e-6


def test_istft_matches_oracle_and_gain_half():
    from spatialaudiogen_b200 import myutils
    a = _audio(2, 6)[:, :, 0]
    s = myutils.stft(cu(a), 1024, 4)[:, 89:117]
    y = myutils.istft(s.contiguous(), 4)
    assert tuple(y.shape) == (2, 6400)
    assert _rel(y, 0.5 * torch.as_tensor(a[:, 23552:29952])) < 2e-6
    g = torch.Generator().manual_seed(0)
    z = torch.complex(torch.randn(3, 2, 30, 1024, generator=g), torch.randn(3, 2, 30, 1024, generator=g))
    y = myutils.istft(z.cuda(), 4)
    ref = O.istft(z.to(torch.complex128), 4)
    assert tuple(y.shape) == tuple(ref.shape) == (3, 2, 6400)     # 30 frames -> 28 used (myutils.py:187-188)
    assert _rel(y, ref) < 3e-6


def test_stft_rejects_short_input():
    from spatialaudiogen_b200 import myutils
    with pytest.raises(ValueError):
        myutils.stft(torch.zeros(1, 1500).cuda(), 1024, 4)


# ------------------------------------------------------------------------------------------------ dense stages
def _conv_case(L, n, h, w, cin, kh, kw, cout, sh, sw, same, bias, relu, prec, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(kh, kw, cin, cout, generator=g) / np.sqrt(kh * kw * cin)
    b = torch.randn(cout, generator=g) if bias else None
    ref = O.tf_conv2d(x.double(), wt.double(), (sh, sw), 'SAME' if same else 'VALID')
    if bias:
        ref = ref + b.double()
    if relu:
        ref = torch.relu(ref)
    y = torch.empty(tuple(ref.shape), dtype=torch.float32, device='cuda')
    xc, wc, bc = x.cuda(), wt.cuda(), (b.cuda() if bias else None)
    L.check(L.lib().sag_conv2d(L.ptr(xc), n, h, w, cin, L.ptr(wc), kh, kw, cout, sh, sw, int(same), L.ptr(bc), int(relu),
                               L.ptr(y), prec, L.stream()))
    torch.cuda.synchronize()
    return _rel(y, ref)


CONV_CASES = [
    (2, 127, 1024, 1, 7, 16, 32, 4, 8, 0, 1, 1),      # audio conv1
    (2, 31, 127, 32, 3, 7, 64, 2, 4, 0, 1, 1),        # audio conv2
    (2, 7, 14, 128, 3, 5, 256, 1, 1, 0, 1, 1),        # audio conv4
    (1, 64, 96, 3, 7, 7, 64, 2, 2, 1, 0, 0),          # resnet conv1 (SAME 2,3)
    (2, 14, 28, 64, 3, 3, 128, 2, 2, 1, 0, 0),        # 3x3/2 SAME (0,1)
    (2, 14, 28, 64, 3, 3, 64, 1, 1, 1, 0, 0),         # 3x3/1 SAME
    (2, 14, 28, 64, 1, 1, 128, 2, 2, 1, 0, 0),        # shortcut 1x1/2
    (1, 5, 7, 5, 3, 3, 7, 1, 2, 1, 1, 0),             # ragged channels
    (1, 9, 9, 16, 2, 4, 40, 1, 1, 0, 1, 1),           # even kernel, Cout not multiple of 32
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d_fp32(case):
    assert _conv_case(_L(), *case, prec=0) < 2e-5


def _deconv_case(L, n, h, w, cin, kh, kw, cout, sh, sw, relu, prec, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(kh, kw, cout, cin, generator=g) / np.sqrt(kh * kw * cin / (sh * sw))
    b = torch.randn(cout, generator=g)
    ref = O.tf_conv2d_transpose_valid(x.double(), wt.double(), (sh, sw)) + b.double()
    if relu:
        ref = torch.relu(ref)
    y = torch.full(tuple(ref.shape), float('nan'), dtype=torch.float32, device='cuda')
    xc, wc, bc = x.cuda(), wt.cuda(), b.cuda()
    L.check(L.lib().sag_deconv2d(L.ptr(xc), n, h, w, cin, L.ptr(wc), kh, kw, cout, sh, sw, L.ptr(bc), int(relu), L.ptr(y),
                                 prec, L.stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(y).any()
    return _rel(y, ref)


DECONV_CASES = [
    (2, 3, 6, 1024, 3, 5, 256, 1, 1, 1),      # deconv5
    (2, 7, 14, 256, 3, 5, 64, 2, 2, 1),       # deconv3
    (1, 15, 31, 128, 3, 7, 32, 2, 4, 1),      # deconv2
    (1, 8, 20, 64, 7, 16, 32, 4, 8, 0),       # deconv1 geometry (small)
    (1, 4, 5, 3, 2, 2, 5, 3, 3, 0),           # stride > kernel (holes get only the bias)
]


@pytest.mark.parametrize('case', DECONV_CASES)
def test_deconv2d_fp32(case):
    assert _deconv_case(_L(), *case, prec=0) < 2e-5


# tcgen05 path: bf16x3 (operands split hi+lo, 3 MMAs per K step) must be fp32-grade; plain bf16 is the fast mode
UMMA_TOL = {3: 5e-5, 2: 2e-2}


@pytest.mark.parametrize('prec', [3, 2])
@pytest.mark.parametrize('case', CONV_CASES + [
    (2, 56, 112, 64, 3, 3, 64, 1, 1, 1, 0, 0),        # resnet conv2_x (many M tiles, BN=64)
    (3, 14, 28, 256, 3, 3, 256, 1, 1, 1, 0, 0),       # resnet conv4_x (two N tiles of 128, 36 K chunks)
    (1, 6, 6, 64, 1, 1, 99, 1, 1, 0, 1, 0),           # N = 99 (localization/fc3 width)
])
def test_conv2d_tcgen05(case, prec):
    assert _conv_case(_L(), *case, prec=prec) < UMMA_TOL[prec]


# stream-K: shapes whose tiles fill the last wave of the persistent grid badly are dealt out by K chunk; the pieces of a tile meet
# in the epilogue of the CTA that finishes it (bias / ReLU after the fix-up).  Ragged M, several N tiles, odd piece counts.
STREAM_K_CASES = [
    (32, 7, 14, 512, 3, 3, 512, 1, 1, 1, 0, 0),       # resnet conv5_x at the benchmarked batch: 100 tiles of 72 chunks on 148 CTAs
    (16, 14, 28, 256, 3, 3, 256, 1, 1, 1, 1, 1),      # 98 tiles, bias + ReLU after the fix-up
    (21, 13, 27, 128, 3, 3, 256, 1, 1, 1, 1, 0),      # last M tile partly filled
    (29, 7, 13, 256, 3, 3, 384, 1, 1, 1, 0, 1),       # three N tiles
]


@pytest.mark.parametrize('prec', [3, 2])
@pytest.mark.parametrize('case', STREAM_K_CASES)
def test_conv2d_tcgen05_stream_k(case, prec, monkeypatch):
    monkeypatch.setenv('SAG_UMMA_STREAMK', '1')          # (the default; read at every call)
    L = _L()
    n, h, w, cin, kh, kw, cout = case[:7]
    assert L.lib().sag_plan_stream_k(kh * kw * cin, cout, n * h * w) == 1
    assert _conv_case(L, *case, prec=prec) < UMMA_TOL[prec]
    assert _conv_case(L, *case, prec=prec, seed=1) < UMMA_TOL[prec]        # other data


@pytest.mark.parametrize('prec', [3, 2])
@pytest.mark.parametrize('case', DECONV_CASES)
def test_deconv2d_tcgen05(case, prec):
    assert _deconv_case(_L(), *case, prec=prec) < UMMA_TOL[prec]


def test_fc_fp32():
    L = _L()
    g = torch.Generator().manual_seed(0)
    for rows, cin, cout, relu in [(6, 3072, 1024, 1), (2, 12544, 512, 1), (6, 512, 99, 0), (1, 7, 3, 0)]:
        x = torch.randn(rows, cin, generator=g)
        w = torch.randn(cin, cout, generator=g) / np.sqrt(cin)
        b = torch.randn(cout, generator=g)
        ref = x.double() @ w.double() + b.double()
        if relu:
            ref = torch.relu(ref)
        y = torch.empty(rows, cout, device='cuda')
        xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
        L.check(L.lib().sag_fc(L.ptr(xc), rows, cin, L.ptr(wc), cout, L.ptr(bc), relu, L.ptr(y), 0, L.stream()))
        assert _rel(y, ref) < 2e-5


def test_batchnorm_train_and_maxpool():
    L = _L()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 10, 12, 64, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    res = torch.randn(3, 10, 12, 64, generator=g)
    ref = torch.relu(O.tf_batch_norm_train(x.double(), gamma.double(), beta.double()) + res.double())
    y = torch.empty_like(x, device='cuda')
    scratch = torch.empty(4 * 64, dtype=torch.float64, device='cuda')
    xc, gc, bc, rc = x.cuda(), gamma.cuda(), beta.cuda(), res.cuda()
    L.check(L.lib().sag_batchnorm_train(L.ptr(xc), 3 * 10 * 12, 64, L.ptr(gc), L.ptr(bc), L.ptr(rc), 1, L.ptr(y),
                                        C.c_void_p(scratch.data_ptr()), L.stream()))
    assert _rel(y, ref) < 1e-5
    for (h, w) in [(112, 224), (7, 9), (8, 8)]:
        x = torch.randn(2, h, w, 64, generator=g)
        ref = O.tf_max_pool_same_3x3s2(x)
        y = torch.empty(tuple(ref.shape), device='cuda')
        xc = x.cuda()
        L.check(L.lib().sag_maxpool_3x3s2_same(L.ptr(xc), 2, h, w, 64, L.ptr(y), L.stream()))
        assert torch.equal(y.cpu(), ref)


def test_mix_matches_reference_decode():
    L = _L()
    g = torch.Generator().manual_seed(0)
    B, K, T = 2, 32, 4800
    xs = torch.randn(B, K, T, generator=g)
    loc = torch.randn(B, 3, 3 * (K + 1), generator=g)
    w = loc.reshape(B, 3, 3, 1, K + 1).unsqueeze(2).repeat(1, 1, T // 3, 1, 1, 1).reshape(B, T, 3, 1, K + 1)
    ref = (w[..., :-1].double() * xs.double().permute(0, 2, 1)[:, :, None, None, :]).sum(4).sum(3) + w[..., -1].double()[:, :, :, 0]
    y = torch.empty(B, T, 3, device='cuda')
    xc, lc = xs.cuda(), loc.cuda()
    L.check(L.lib().sag_mix(L.ptr(xc), L.ptr(lc), B, K, T, 3, L.ptr(y), L.stream()))
    assert _rel(y, ref) < 1e-5


# ------------------------------------------------------------------------------------------------ whole forward
def _models(encoders, separation, seed, B, precision='fp32', resnet=None, stress=True):
    from spatialaudiogen_b200 import SptAudioGen
    W = Wt.init_weights(encoders, separation=separation, seed=seed, stress=stress, resnet_npy=resnet)
    ref = O.SptAudioGen(W, encoders=encoders, separation=separation, dtype=torch.float64)
    m = SptAudioGen(1, encoders=encoders, separation=separation, precision=precision).load_weights(W)
    return ref, m


def test_forward_audio_only_fp32_parity():
    """BASELINE config 1 shape (audio-only encoder, random weights) at B=1 and B=3; <= 1e-3 rel required,
    the fp32 path lands around 1e-5."""
    ref, m = _models(['audio'], 'unet_mask', 3, 1)
    for B in (1, 3):
        a = _audio(B, 10 + B)
        y = m.inference_ops(cu(a))
        yr = ref.inference_ops(a)
        assert tuple(y.shape) == (B, 4800, 3)
        assert _rel(y, yr) < 1e-4
        assert _rel(m.sep_channels, ref.sep_channels) < 1e-4
        assert _rel(m.ends['separation/mask_logits'], ref.ends['separation/mask_logits'][:, 0]) < 1e-4
        for l in range(6):
            assert _rel(m.ends['audio_encoder/%d' % l], ref.ends['audio_encoder'][l]) < 1e-4, l
        assert _rel(m.ends['bottleneck'], ref.ends['bottleneck']) < 1e-4


def test_forward_skip_unused_is_bit_identical_to_full():
    ref, m = _models(['audio'], 'unet_mask', 5, 2)
    a = cu(_audio(2, 21))
    y0 = m.inference_ops(a).clone()
    m.set_option('skip_unused', 0)
    y1 = m.inference_ops(a).clone()
    assert torch.equal(y0, y1)
    s = m.ends['stft']
    assert tuple(s.shape) == (2, 1, 200, 1024)
    assert _rel(torch.view_as_real(s), torch.view_as_real(O.stft(torch.as_tensor(_audio(2, 21)).double().permute(0, 2, 1), 1024, 4))) < 2e-6


def test_forward_no_separation():
    ref, m = _models(['audio'], 'none', 7, 2)
    a = _audio(2, 22)
    assert _rel(m.inference_ops(cu(a)), ref.inference_ops(a)) < 1e-4


def test_forward_audio_video_fp32_parity():
    """BASELINE config 2 shape (audio+video, batch-statistics BN) at a small batch."""
    ref, m = _models(['audio', 'video'], 'unet_mask', 9, 2)
    a, v = _audio(2, 23), _video(2, 24)
    y = m.inference_ops(cu(a), video=cu(v))
    yr = ref.inference_ops(a, video=v)
    assert _rel(m.ends['video_encoder/conv2_1'], ref.ends['video_encoder/conv2_1']) < 1e-4
    assert _rel(m.ends['video_encoder/conv5_2'], ref.ends['video_encoder/conv5_2']) < 1e-3
    assert _rel(m.ends['bottleneck'], ref.ends['bottleneck']) < 1e-3
    assert _rel(y, yr) < 1e-3


def test_forward_audio_video_flow_fp32_parity():
    ref, m = _models(['audio', 'video', 'flow'], 'unet_mask', 11, 2)
    a, v, fl = _audio(2, 25), _video(2, 26), _flow(2, 27)
    y = m.inference_ops(cu(a), video=cu(v), flow=cu(fl))
    yr = ref.inference_ops(a, video=v, flow=fl)
    assert _rel(y, yr) < 1e-3


def test_forward_tcgen05_bf16x3_parity_and_bf16_error():
    """The tensor-core path: bf16x3 must meet the same <= 1e-3 waveform tolerance as the fp32 path (north_star);
    plain bf16 is reported against a looser bound (BASELINE config 3)."""
    ref, m = _models(['audio', 'video'], 'unet_mask', 9, 2, precision='bf16x3')
    a, v = _audio(2, 23), _video(2, 24)
    y = m.inference_ops(cu(a), video=cu(v))
    yr = ref.inference_ops(a, video=v)
    assert _rel(m.ends['audio_encoder/5'], ref.ends['audio_encoder'][5]) < 1e-4
    assert _rel(m.ends['video_encoder/conv2_1'], ref.ends['video_encoder/conv2_1']) < 1e-4
    assert _rel(m.ends['video_encoder/conv5_2'], ref.ends['video_encoder/conv5_2']) < 1e-3
    assert _rel(m.ends['separation/mask_logits'], ref.ends['separation/mask_logits'][:, 0]) < 1e-3
    assert _rel(y, yr) < 1e-3
    m.set_option('precision', 'bf16')
    y16 = m.inference_ops(cu(a), video=cu(v))
    assert _rel(y16, yr) < 1e-1


def test_forward_tcgen05_audio_only_and_flow():
    ref, m = _models(['audio'], 'unet_mask', 3, 3, precision='bf16x3')
    a = _audio(3, 13)
    assert _rel(m.inference_ops(cu(a)), ref.inference_ops(a)) < 1e-3
    ref, m = _models(['audio', 'video', 'flow'], 'unet_mask', 11, 2, precision='bf16x3')
    a, v, fl = _audio(2, 25), _video(2, 26), _flow(2, 27)
    assert _rel(m.inference_ops(cu(a), video=cu(v), flow=cu(fl)), ref.inference_ops(a, video=v, flow=fl)) < 1e-3


def test_forward_tma_gather_matches_cp_async_gather():
    """The TMA im2col producer and the cp.async producer put the same bytes into the operand ring: the audio-only
    forward (no batch-norm atomics) is bit-identical, the audio+video forward agrees to the atomics' rounding."""
    _, m = _models(['audio'], 'unet_mask', 5, 3, precision='bf16x3')
    a = cu(_audio(3, 31))
    m.set_option('tma_gather', 1)
    y1 = m.inference_ops(a).clone()
    m.set_option('tma_gather', 0)
    y0 = m.inference_ops(a).clone()
    assert torch.equal(y0, y1)
    _, m = _models(['audio', 'video'], 'unet_mask', 7, 2, precision='bf16x3')
    a, v = cu(_audio(2, 32)), cu(_video(2, 33))
    m.set_option('tma_gather', 1)
    y1 = m.inference_ops(a, video=v).clone()
    m.set_option('tma_gather', 0)
    y0 = m.inference_ops(a, video=v).clone()
    assert _rel(y1, y0) < 1e-4


def test_fused_istft_mix_matches_the_two_kernel_path_and_the_oracle():
    """forward_into (deploy / eval hot loop) folds the 32->3 mixing into the inverse STFT by linearity; inference_ops
    keeps x_sep and mixes afterwards.  Same waveform to fp32 rounding, and within tolerance of the oracle."""
    for enc, B in ((['audio'], 3), (['audio', 'video'], 2)):
        ref, m = _models(enc, 'unet_mask', 13, B, precision='bf16x3')
        a = _audio(B, 61)
        v = _video(B, 62) if 'video' in enc else None
        y_two = m.inference_ops(cu(a), video=None if v is None else cu(v)).clone()
        assert m.sep_channels is not None
        y_fused = torch.empty_like(y_two)
        m.forward_into(cu(a), None if v is None else cu(v), None, y_fused)
        torch.cuda.synchronize()
        assert _rel(y_fused, y_two) < (2e-5 if v is None else 1e-4)
        assert _rel(y_fused, ref.inference_ops(a, video=v)) < 1e-3
    # exact-arithmetic check of the fusion itself on the fp32 path, B = 1 (one window: segments x channels)
    ref, m = _models(['audio'], 'unet_mask', 14, 1, precision='fp32')
    a = _audio(1, 63)
    y = torch.empty((1, 4800, 3), device='cuda')
    m.forward_into(cu(a), None, None, y)
    torch.cuda.synchronize()
    assert _rel(y, ref.inference_ops(a)) < 1e-4


def test_forward_cta_pair_matches_single_cta():
    """cta_group::2 (two M tiles per cluster, each CTA feeding half of every weight tile) computes the same products in
    the same order as the single-CTA kernel: bit-identical without batch-norm atomics, to rounding with them."""
    _, m = _models(['audio'], 'unet_mask', 5, 3, precision='bf16x3')
    a = cu(_audio(3, 41))
    m.set_option('cta_pair', 1)
    y1 = m.inference_ops(a).clone()
    m.set_option('cta_pair', 0)
    y0 = m.inference_ops(a).clone()
    assert torch.equal(y0, y1)
    _, m = _models(['audio', 'video'], 'unet_mask', 7, 3, precision='bf16x3')     # odd M tile counts: padding tiles
    a, v = cu(_audio(3, 42)), cu(_video(3, 43))
    m.set_option('cta_pair', 1)
    y1 = m.inference_ops(a, video=v).clone()
    m.set_option('cta_pair', 0)
    y0 = m.inference_ops(a, video=v).clone()
    assert _rel(y1, y0) < 1e-4


def test_forward_errors():
    from spatialaudiogen_b200 import SptAudioGen
    with pytest.raises(ValueError):
        SptAudioGen(1, encoders=['audio'], separation='bogus')
    m = SptAudioGen(1, encoders=['audio'], separation='unet_mask')
    with pytest.raises(RuntimeError):
        m.inference_ops(torch.zeros(1, 52799, 1).cuda())           # weights not loaded
    W = Wt.init_weights(['audio'])
    bad = dict(W)
    bad['audio_encoder/conv1/weights'] = np.zeros((7, 16, 2, 32), np.float32)
    with pytest.raises(ValueError):
        m.load_weights(bad)
    m.load_weights(W)
    with pytest.raises(ValueError):
        m.inference_ops(torch.zeros(1, 1000, 1).cuda())
    with pytest.raises(TypeError):
        m.forward_into(torch.zeros(1, 52799, 1), None, None, torch.zeros(1, 4800, 3).cuda())   # CPU tensor: no CPU path


# ------------------------------------------------------------------------------------------------ metrics
def test_metrics_match_oracle():
    from spatialaudiogen_b200 import SptAudioGen
    from spatialaudiogen_b200 import metrics as M
    rng = np.random.RandomState(0)
    B = 5
    gt = (rng.randn(B, 4800, 3) * 0.1).astype(np.float32)
    pred = (gt + rng.randn(B, 4800, 3) * 0.03).astype(np.float32)
    mask = np.ones((B, 3), np.float32)
    mask[1, 1] = 0
    ref = O.SptAudioGen({}, encoders=['audio'], separation='unet_mask', dtype=torch.float64)
    mr, stft_r, lsd_r, mse_r, snr_r = ref.evaluation_ops(pred, gt, None, mask)
    m = SptAudioGen(1, encoders=['audio'], separation='unet_mask')
    mm, stft_g, lsd_g, mse_g, snr_g = m.evaluation_ops(cu(pred), cu(gt), None, cu(mask))
    assert _rel(stft_g, stft_r) < 1e-4
    assert _rel(lsd_g, lsd_r) < 1e-4
    assert _rel(mse_g, mse_r) < 1e-5
    assert _rel(snr_g, snr_r) < 1e-5
    for k in mr:
        assert abs(float(mm[k]) - float(mr[k])) <= 1e-4 * max(1.0, abs(float(mr[k]))), k
    env_r = np.stack([O.compute_envelope_dist(pred[b], gt[b]) for b in range(B)])
    assert _rel(m.last_eval['env'], env_r) < 1e-4
    amp = m.last_eval['amp'].cpu().numpy()
    assert np.array_equal(amp[:, 0], np.abs(pred).reshape(B, -1).max(1)) and np.array_equal(amp[:, 1], np.abs(gt).reshape(B, -1).max(1))
    # spherical-harmonic RMS maps (84 directions at 30 degrees; 37x72 at 5 degrees)
    ambi = (rng.randn(B, 4800, 4) * 0.1).astype(np.float32)
    for res, shape in ((30., (7, 12)), (5., (37, 72))):
        g = M.ambix_rms_map(cu(ambi), res)
        assert tuple(g.shape) == (B,) + shape
        r = np.stack([O.ambix_rms_map(ambi[b], res) for b in range(B)])
        assert _rel(g, r) < 1e-5


# ------------------------------------------------------------------------------------------------ drivers
class _P(object):
    """What myutils.load_params returns for a model_dir (reference myutils.py:40-85)."""
    def __init__(self, encoders, separation='unet_mask'):
        self.encoders, self.separation = encoders, separation
        self.ambi_order, self.audio_rate, self.video_rate, self.context = 1, 48000, 10, 1.0
        self.num_sep_tracks, self.fft_window = 32, 0.025
        self.context_units, self.freq_mask_units, self.loc_units = [64, 128, 128], [256], [512, 512]


def test_w2xyz_deploy_matches_reference_loop_with_zero_padded_tail():
    """deploy.py:112-151: 13 windows -> one full batch of 10 and a zero-padded batch of 3 (batch statistics make the
    padding visible in the video tower); rows [W, Y, Z, X], float64."""
    from spatialaudiogen_b200.deploy import W2XYZ
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=21, stress=True)
    N = 13
    amb = np.concatenate([_audio(N, 31), _audio(N, 32), _audio(N, 33), _audio(N, 34)], axis=2)     # (N, 52799, 4)
    vid = _video(N, 35)
    ref = O.deploy_assemble(O.SptAudioGen(W, encoders=enc, separation='unet_mask'), amb, video_windows=vid)
    w = W2XYZ(params=_P(enc), weights=W)
    out = w.deploy_windows(amb, video_windows=vid)
    assert out.dtype == np.float64 and out.shape == (N * 4800, 4) == ref.shape
    assert np.array_equal(out[:, 0], amb[:, 24000:28800, 0].reshape(-1).astype(np.float64))   # W: bit-exact crop
    assert _rel(out[:, 1:], ref[:, 1:]) < 1e-3


def test_w2xyz_restores_a_tf_checkpoint_bundle(tmp_path):
    """deploy.py:79-87: the weights come from `checkpoint` + model.ckpt-N.index/.data (variables, Adam slots and
    global_step side by side); restoring by name gives the same network as handing the arrays over directly."""
    from spatialaudiogen_b200.deploy import W2XYZ
    from spatialaudiogen_b200 import tf_checkpoint as T
    enc = ['audio']
    W = Wt.init_weights(enc, separation='unet_mask', seed=5, stress=True)
    bundle = dict(W)
    for k, v in W.items():
        bundle[k + '/Adam'] = np.zeros_like(v)
    bundle['global_step'] = np.asarray(777, dtype=np.int64)
    T.write_bundle(str(tmp_path / 'model.ckpt-777'), bundle)
    amb = np.concatenate([_audio(4, 51)] * 4, axis=2)
    a = W2XYZ(model_dir=str(tmp_path), params=_P(enc)).deploy_windows(amb)
    b = W2XYZ(params=_P(enc), weights=W).deploy_windows(amb)
    assert np.array_equal(a, b)


def test_w2xyz_deploy_reads_a_video_folder(tmp_path):
    """deploy.py:90-152 end to end: wav / jpg files on disk -> readers.SampleReader -> batches -> [W, Y, Z, X] rows; equal to
    handing the same windows to deploy_windows."""
    from spatialaudiogen_b200.deploy import W2XYZ
    from spatialaudiogen_b200 import readers as R
    from test_host import _make_video_folder
    folder, full = _make_video_folder(str(tmp_path), seconds=4)
    enc = ['audio', 'video']
    w = W2XYZ(params=_P(enc), weights=Wt.init_weights(enc, separation='unet_mask', seed=9, stress=True))
    out = w.deploy(folder, 1.2, 2.0)                                     # windows at 1.2 and 2.2 s (schedule 1.5, 2.5 shifted by 0.3)
    assert out.shape == (2 * 4800, 4) and out.dtype == np.float64
    s0 = int(1.2 * 48000)                                                # W = the mono crop at the window centre (deploy.py:143-147)
    assert np.array_equal(out[:4800, 0], full[s0:s0 + 4800, 0])
    r = R.SampleReader(folder, shuffle=False, random_rotations=False, start_time=1.2, sample_duration=2.0, img_prep=lambda x: x / 255. - 0.5)
    r.chunks_t = [t - 0.3 for t in r.chunks_t]
    chunks = list(r.loop_chunks())
    ref = w.deploy_windows(np.stack([c['ambix'] for c in chunks]), np.stack([c['video'] for c in chunks]).astype(np.float32))
    assert np.array_equal(out, ref)
    # two full batches of 10 + a tail of 5: the frames' jpg files decoded on the GPU (full batches stay on the device, lanes in
    # flight) against the PIL readers
    with open(os.path.join(folder, 'audio_pow.lst'), 'w') as f:
        for k in range(30):
            f.write('%.1f %.3f\n' % (0.5 + 0.1 * k, 0.3))
    gpu = w.deploy(folder, 0.5, 2.5, gpu_jpeg=True)
    pil = w.deploy(folder, 0.5, 2.5, gpu_jpeg=False)
    assert gpu.shape == (25 * 4800, 4) and np.array_equal(gpu, pil)


def test_evaluate_from_video_folders(tmp_path):
    """eval.py end to end on disk data: folder_batches (eval schedule: every 10th audio_pow.lst entry) -> evaluate_batches
    -> eval-detailed.txt rows; equal to running the same windows through the model and metric_rows by hand."""
    from spatialaudiogen_b200 import evaluate as E, readers as R, SptAudioGen
    from test_host import _make_video_folder
    folder, full = _make_video_folder(str(tmp_path), seconds=4)
    with open(os.path.join(folder, 'audio_pow.lst'), 'w') as f:          # preprocessing writes one entry per 0.1 s
        for k in range(30):
            f.write('%.1f %.3f\n' % (0.5 + 0.1 * k, 0.3))
    enc = ['audio', 'video']
    m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(Wt.init_weights(enc, separation='unet_mask', seed=4, stress=True))
    assert list(E.folder_batches([folder], _P(enc), batch_size=16)) == []          # the short batch is dropped like the reference's queue does
    batches = list(E.folder_batches([folder], _P(enc), batch_size=16, channel_masks={'vidA': np.array([1., 1., 0., 1.])}, drop_remainder=False))
    assert len(batches) == 1 and batches[0].get('short_batch') and batches[0]['id'] == ['vidA 0.5', 'vidA 1.5', 'vidA 2.5']
    assert tuple(batches[0]['ambix'].shape) == (3, 52799, 4) and tuple(batches[0]['video'].shape) == (3, 1, 224, 448, 3)
    ids, rows = E.evaluate_batches(m, batches, rms_maps=True)
    assert ids == batches[0]['id'] and tuple(rows.shape) == (3, 28) and not torch.isnan(rows).any()
    # batches in flight on lane twins: same rows in the same order, whatever the number of lanes
    many = [batches[0]] * 5
    ids1, rows1 = E.evaluate_batches(m, many, rms_maps=True, lanes=1)
    ids3, rows3 = E.evaluate_batches(m, many, rms_maps=True, lanes=3)
    assert ids1 == ids3 == batches[0]['id'] * 5 and torch.equal(rows1, rows3) and torch.equal(rows1[:3], rows)
    amb = batches[0]['ambix']
    pred = m.inference_ops(amb[:, :, :1].contiguous(), video=batches[0]['video'])
    ref_rows, _ = E.metric_rows(pred, amb[:, 24000:28800, 1:].contiguous(), mono=amb[:, 24000:28800, :1].contiguous(),
                                layout=batches[0]['mask'], rms_maps=True)
    assert torch.allclose(rows, ref_rows, rtol=2e-3, atol=1e-5)
    E.write_eval_detailed(str(tmp_path / 'eval-detailed.txt'), ids, rows)
    assert open(str(tmp_path / 'eval-detailed.txt')).read().splitlines()[1].startswith('vidA 0.5 | ')


def test_deploy_post_processing_arithmetic():
    """myutils.gen_360video without ffmpeg: stereo down-mix (myutils.py:285-291) and the 5-degree heat-map frames
    (myutils.py:251-275) from the GPU energy maps."""
    from spatialaudiogen_b200 import myutils
    rng = np.random.RandomState(8)
    ambix = (rng.randn(48000, 4) * np.array([0.2, 0.1, 0.05, 0.15])).astype(np.float32)        # 1 s of [W, Y, Z, X]
    st = myutils.ambix_to_stereo(ambix)
    assert st.shape == (48000, 2) and abs(np.abs(st).max() - 0.95) < 1e-12
    a64 = ambix.astype(np.float64)
    assert np.allclose(st[:, 0] / st[:, 1], (a64[:, 0] + a64[:, 1]) / (a64[:, 0] - a64[:, 1]))
    maps = myutils.energy_map_frames(ambix, 48000, 10.)
    ref = O.energy_map_frames(ambix, 48000, 10.)
    assert maps.shape == ref.shape == (5 * (2 - 1), 37, 72)                 # 10 fps: one map per 0.5 s, 5 blended frames between two
    assert np.abs(maps - ref).max() < 2e-4


def test_evaluate_rows_follow_eval_detailed_columns(tmp_path):
    from spatialaudiogen_b200 import evaluate as E
    rng = np.random.RandomState(3)
    B = 4
    gt = (rng.randn(B, 4800, 3) * 0.1).astype(np.float32)
    pred = (gt + rng.randn(B, 4800, 3) * 0.02).astype(np.float32)
    mono = (rng.randn(B, 4800, 1) * 0.1).astype(np.float32)
    layout = np.ones((B, 4), np.float32)
    layout[2, 2] = 0                                     # a WXY clip: no Z (feeder.py:312-314)
    rows, maps = E.metric_rows(cu(pred), cu(gt), mono=cu(mono), layout=cu(layout), rms_maps=True)
    rows = rows.cpu().numpy()
    assert rows.shape == (B, 28) and E.ALL_METRICS[0] == 'amplitude/predicted' and E.ALL_METRICS[-1] == 'emd/dir2'
    ref = O.SptAudioGen({}, encoders=['audio'], separation='unet_mask', dtype=torch.float64)
    _, stft_r, lsd_r, mse_r, snr_r = ref.evaluation_ops(pred, gt, None, np.ones((B, 3), np.float32))
    col = {k: i for i, k in enumerate(E.ALL_METRICS)}
    for name, r in (('stft', stft_r), ('lsd', lsd_r), ('mse', mse_r), ('snr', snr_r)):
        r = r.numpy()
        assert np.allclose(rows[:, col[name + '/avg']], r.mean(1), rtol=2e-4, atol=1e-7)
        for i, ch in enumerate('YZX'):                   # keyed by channel name, channel order of the model is Y,Z,X
            assert np.allclose(rows[:, col[name + '/' + ch]], r[:, i], rtol=2e-4, atol=1e-7)
    env_r = np.stack([O.compute_envelope_dist(pred[b], gt[b]) for b in range(B)])
    assert np.allclose(rows[:, col['env_mse/X']], env_r[:, 2], rtol=2e-4)
    mel_r = np.stack([O.compute_lsd_dist(pred[b], gt[b], 48000) for b in range(B)])      # myutils.py:96-106
    assert np.allclose(rows[:, col['mel_lsd/avg']], mel_r.mean(1), rtol=1e-3)
    for i, ch in enumerate('YZX'):
        assert np.allclose(rows[:, col['mel_lsd/' + ch]], mel_r[:, i], rtol=1e-3)
    for b in range(B):
        rp = O.ambix_rms_map(np.concatenate((mono[b], pred[b]), 1) * layout[b], 30.)
        assert _rel(maps[0][b], np.ascontiguousarray(rp)) < 1e-5
        # emd/dir, emd/dir2 (eval.py:190): GPU maps + host min-cost flow against the oracle's maps + LP
        e1, e2 = O.ambix_emd(np.concatenate((mono[b], pred[b]), 1) * layout[b], np.concatenate((mono[b], gt[b]), 1) * layout[b], 30.)
        assert abs(rows[b, col['emd/dir']] - e1) < 1e-5 * max(1.0, abs(e1)) + 1e-6
        assert abs(rows[b, col['emd/dir2']] - e2) < 1e-4 * max(1.0, abs(e2)) + 1e-5
    rows_nomaps, _ = E.metric_rows(cu(pred), cu(gt))                     # without the maps the EMD columns stay nan
    assert np.isnan(rows_nomaps.cpu().numpy()[:, col['emd/dir']]).all()
    fn = str(tmp_path / 'eval-detailed.txt')
    E.write_eval_detailed(fn, ['vid%d 0.5' % b for b in range(B)], rows)
    lines = open(fn).read().splitlines()
    assert lines[0] == 'SampleID | ' + ' '.join(E.ALL_METRICS) and len(lines) == B + 1
    assert lines[1].startswith('vid0 0.5 | ') and len(lines[1].split(' | ')[1].split(' ')) == 28


def test_inference_stream_matches_per_batch_calls():
    """The pipelined driver loop returns, in order, exactly what one inference_ops call per batch returns."""
    from spatialaudiogen_b200 import SptAudioGen
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=5, stress=True)
    m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W)
    batches = [{'audio': torch.as_tensor(_audio(2, 40 + i)).pin_memory(), 'video': torch.as_tensor(_video(2, 50 + i)).pin_memory()}
               for i in range(5)]
    outs = [y.clone() for y in m.inference_stream(iter(batches))]
    assert len(outs) == 5
    out = torch.empty((2, 4800, 3), device='cuda')
    for b, y in zip(batches, outs):
        # bit-equal to the same hot loop called batch by batch: batch-norm statistics are summed in a fixed order (no
        # floating-point atomics), so repeated forwards reproduce each other exactly
        m.forward_into(b['audio'].cuda(), b['video'].cuda(), None, out)
        assert torch.equal(y, out.cpu())
        ref = m.inference_ops(b['audio'], video=b['video']).cpu()      # two-kernel inverse STFT + mixing: same values to rounding
        assert _rel(y, ref) < 2e-5
    # forwards in flight (default: three lanes): one lane and two lanes, eager launches, return the same bits in the same order;
    # options set on the model reach the lane twins
    for lanes in (1, 2):
        again = [y.clone() for y in m.inference_stream(iter(batches), lanes=lanes, use_graph=False)]
        assert len(again) == 5 and all(torch.equal(x, y) for x, y in zip(outs, again)), lanes
    m.set_option('precision', 'fp32')
    exact = [y.clone() for y in m.inference_stream(iter(batches), use_graph=False)]
    for b, y in zip(batches, exact):
        m.forward_into(b['audio'].cuda(), b['video'].cuda(), None, out)
        assert _rel(y, out.cpu()) < 1e-5                             # (the FFMA path sums its split-K pieces with float atomics)
    assert not torch.equal(exact[1], outs[1])                        # (the fp32 path really ran on the twin)


def test_cuda_graph_replay_equals_eager_forward():
    """capture_graph / inference_stream(use_graph=True): the replayed chain (programmatic dependent launches included) returns
    what the eager calls return, bit for bit, also after the buffers are refilled."""
    from spatialaudiogen_b200 import SptAudioGen
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=8, stress=True)
    m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W)
    for B in (1, 10):
        a, v = cu(_audio(B, 70)), cu(_video(B, 71))
        out_g, out_e = torch.empty((B, 4800, 3), device='cuda'), torch.empty((B, 4800, 3), device='cuda')
        g = m.capture_graph(a, v, None, out_g)
        for seed in (72, 73):
            a.copy_(cu(_audio(B, seed)))
            v.copy_(cu(_video(B, seed + 10)))
            out_g.zero_()
            g.replay()
            m.forward_into(a, v, None, out_e)
            torch.cuda.synchronize()
            assert torch.equal(out_g, out_e) and float(out_e.abs().max()) > 0
    batches = [{'audio': torch.as_tensor(_audio(2, 80 + i)).pin_memory(), 'video': torch.as_tensor(_video(2, 90 + i)).pin_memory()}
               for i in range(5)]
    eager = [y.clone() for y in m.inference_stream(iter(batches), use_graph=False)]
    graph = [y.clone() for y in m.inference_stream(iter(batches), use_graph=True)]
    assert all(torch.equal(x, y) for x, y in zip(eager, graph))


def test_forward_is_bit_reproducible():
    """Run-to-run determinism of the tensor-core forward at a split-K batch (B=2) and at the benchmarked batch (B=32)."""
    from spatialaudiogen_b200 import SptAudioGen
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=6, stress=True)
    m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W)
    for B in (2, 32):
        a, v = cu(_audio(B, 60)), cu(_video(B, 61))
        outs = []
        for _ in range(3):
            o = torch.empty((B, 4800, 3), device='cuda')
            m.forward_into(a, v, None, o)
            outs.append(o)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_forward_stream_k_matches_the_tile_schedule_and_is_bit_reproducible(monkeypatch):
    """Stream-K on conv4_x / conv5_x of the benchmarked batch: the pieces of a tile meet in the finishing CTA's epilogue
    (batch-norm statistics + TMA-store path), the flags are left cleared, so repeated forwards agree bit for bit."""
    from spatialaudiogen_b200 import SptAudioGen
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=6, stress=True)
    B = 32
    a, v = cu(_audio(B, 60)), cu(_video(B, 61))
    ref = torch.empty((B, 4800, 3), device='cuda')
    monkeypatch.setenv('SAG_UMMA_STREAMK', '0')                 # whole tiles per CTA
    SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W).forward_into(a, v, None, ref)
    monkeypatch.setenv('SAG_UMMA_STREAMK', '1')
    assert _L().lib().sag_plan_stream_k(9 * 512, 512, B * 7 * 14) == 1
    m = SptAudioGen(1, encoders=enc, separation='unet_mask').load_weights(W)      # planned and run under the same setting
    outs = []
    for _ in range(3):
        o = torch.empty((B, 4800, 3), device='cuda')
        m.forward_into(a, v, None, o)
        outs.append(o)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert _rel(outs[0], ref.double().cpu()) < 2e-4            # other summation order of the K chunks (both are ~6e-5 from the oracle)


def test_stage_methods_match_oracle():
    """The reference's per-stage methods (stages.StageOps) on the GPU primitives against the oracle's stages."""
    from spatialaudiogen_b200 import myutils
    ref, m = _models(['audio', 'video'], 'unet_mask', 17, 2, precision='fp32')
    a, v = _audio(2, 41), _video(2, 42)
    yr = ref.inference_ops(a, video=v)
    mono = cu(a).permute(0, 2, 1).contiguous()
    s = myutils.stft(mono, m.wind_size, 4)
    a_enc = m.audio_encoder_ops(s)
    for got, want in zip(a_enc, ref.ends['audio_encoder']):
        assert _rel(got, want) < 1e-4
    vis = m.visual_encoding_ops(cu(v), is_training=False, finetune=True, scope='video_encoder')
    assert _rel(vis, ref.ends['video_encoder/conv5_2']) < 1e-3
    feats = m.bottleneck_ops({'audio': a_enc, 'video': vis}, True)
    assert _rel(feats, ref.ends['bottleneck']) < 1e-3
    w, b = m.localization_ops(feats)
    assert _rel(w, ref.loc_channels[0]) < 1e-3 and _rel(b, ref.loc_channels[1]) < 1e-3
    x_sep = m.separation_ops(mono, s, a_enc, feats)
    assert _rel(x_sep, ref.sep_channels) < 1e-3
    y = (w * x_sep.permute(0, 3, 1, 2).unsqueeze(2)).sum(4).sum(3) + b[:, :, :, 0]
    assert _rel(y, yr) < 1e-3


def test_halo_resident_conv_matches_the_im2col_kernel_and_the_oracle():
    """The 3x3 / stride-1 convolutions of the ResNet trunk on the halo-resident kernel (tiled TMA boxes shared by three vertical
    taps, 4-D tensor store, statistics from registers) against the im2col kernel and the fp64 oracle; odd batch and a frame size
    whose tiles overhang the image border (the clipped store / masked statistics path)."""
    from spatialaudiogen_b200 import SptAudioGen
    enc = ['audio', 'video']
    for frame, B in (((224, 448), 3), ((208, 432), 2)):       # 52 x 108 feature maps: 16 x 8 tiles overhang both borders
        W = Wt.init_weights(enc, separation='unet_mask', seed=21, stress=True)
        if frame != (224, 448):                                              # the video-fc input follows the frame size
            fh, fw = -(-frame[0] // 32), -(-frame[1] // 32)
            rng = np.random.RandomState(3)
            W['bottleneck/video-fc/weights'] = (rng.randn(fh * fw * 128, 512) * 0.01).astype(np.float32)
        m = SptAudioGen(1, encoders=enc, separation='unet_mask', frame_size=frame).load_weights(W)
        a, v = cu(_audio(B, 150)), cu(_video(B, 151, frame[0], frame[1]))
        res = {}
        for halo in (1, 0):
            m.set_option('halo_conv', halo)                                      # 1 = forced wherever eligible (default: zero-padding shapes only)
            y = m.inference_ops(a, video=v).clone()
            res[halo] = (y, {k: m.ends[k].clone() for k in ('video_encoder/conv2_1', 'video_encoder/conv2_2', 'video_encoder/conv3_2', 'video_encoder/conv5_2')})
        for k in res[1][1]:
            assert _rel(res[1][1][k], res[0][1][k]) < 1e-4, k           # (different accumulation order over the taps)
        assert _rel(res[1][0], res[0][0]) < 1e-4
        m.set_option('cta_pair', 0)                                              # the single-CTA variant of the halo kernel
        m.set_option('halo_conv', 1)
        y1 = m.inference_ops(a, video=v).clone()
        assert _rel(m.ends['video_encoder/conv2_2'], res[1][1]['video_encoder/conv2_2']) < 1e-4 and _rel(y1, res[1][0]) < 1e-4
        m.set_option('cta_pair', -1)
        if frame == (224, 448):
            ref = O.SptAudioGen(W, encoders=enc, separation='unet_mask', dtype=torch.float64)
            yr = ref.inference_ops(a.cpu().numpy(), video=v.cpu().numpy())
            assert _rel(res[1][1]['video_encoder/conv3_2'], ref.ends['video_encoder/conv3_2']) < 1e-4
            assert _rel(res[1][0], yr) < 1e-3
