"""The DSP helpers and parameter parsing of the reference's myutils.py that sit on the inference path, with the same
names and argument meaning: stft / istft (myutils.py:119-147, 181-211) run as fused shared-memory FFT kernels in
libsag.so; load_params / img_prep_fcn (myutils.py:40-89) are host-side and kept verbatim in behaviour."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .metrics import compute_envelope_dist  # noqa: F401  (myutils.py:109-116)


def stft(inp, wind_size, n_overlap):
    """reference myutils.py:119-147.  inp (..., n_samples) real CUDA tensor -> (..., n_overlap*n_winds, wind_size)
    complex64: frame t = samples [t*hop, t*hop+wind) x periodic Hann, two-sided unnormalised FFT, no padding."""
    inp = L.f32(inp)
    lead = list(inp.shape[:-1])
    n = inp.shape[-1]
    rows = int(np.prod(lead)) if lead else 1
    n_winds = n // wind_size - 1
    if n_winds < 1:
        raise ValueError('stft: %d samples are too few for window %d' % (n, wind_size))
    nf = n_overlap * n_winds
    with torch.cuda.device(inp.device):
        out = torch.empty((rows, nf, wind_size, 2), dtype=torch.float32, device=inp.device)
        L.check(L.lib().sag_stft(L.ptr(inp), rows, n, wind_size, n_overlap, 0, nf, L.ptr(out), 0, 0, None, L.stream()))
    return torch.view_as_complex(out).reshape(lead + [nf, wind_size])


def istft(inp, n_overlap):
    """reference myutils.py:181-211.  inp (..., n_frames, wind) complex -> (..., n_samples): real(ifft), the
    n_overlap interleaved streams trimmed and summed / n_overlap (no synthesis window: stft->istft gain is 0.5)."""
    if not torch.is_complex(inp):
        raise TypeError('istft expects a complex tensor')
    x = torch.view_as_real(inp.to(torch.complex64)).contiguous()
    lead = list(inp.shape[:-2])
    n_frames, wind = inp.shape[-2], inp.shape[-1]
    rows = int(np.prod(lead)) if lead else 1
    nf = (n_frames // n_overlap) * n_overlap
    if nf == 0:
        raise ValueError('istft: needs at least %d frames' % n_overlap)
    n_out = (nf // n_overlap) * wind - (n_overlap - 1) * (wind // n_overlap)
    with torch.cuda.device(x.device):
        out = torch.empty((rows, n_out), dtype=torch.float32, device=x.device)
        L.check(L.lib().sag_istft(L.ptr(x), rows, n_frames, wind, n_overlap, L.ptr(out), L.stream()))
    return out.reshape(lead + [n_out])


def load_params(model_dir):
    """reference myutils.py:40-85: parse `model_dir/train-params.txt` (`key: value` lines) with the reference's
    fall-backs for missing keys (num_sep_tracks 64, fft_window 0.025, context_units [64,128,128], freq_mask_units [],
    loc_units [256,256]) -- they differ from definitions.py on purpose: old checkpoints rely on them."""
    params = {l.split(':')[0]: l.strip().split(':')[1].strip() for l in open(model_dir + '/train-params.txt')}
    for k in ['encoders', 'separation']:
        params[k] = params[k].lower()
    for k in ['ambi_order', 'audio_rate', 'video_rate', 'n_iters', 'batch_size']:
        params[k] = int(params[k])
    for k in ['context', 'sample_dur', 'lr', 'lr_decay', 'lr_iters']:
        params[k] = float(params[k])
    params['encoders'] = [enc.strip()[1:-1] for enc in params['encoders'][1:-1].split(',')]
    params['num_sep_tracks'] = int(params.get('num_sep_tracks', '64'))
    params['fft_window'] = float(params.get('fft_window', '0.025'))
    for k, default in (('context_units', '[64, 128, 128]'), ('freq_mask_units', '[]'), ('loc_units', '[256, 256]')):
        v = params.get(k, default)
        params[k] = [int(l.strip()) for l in v[1:-1].split(',')] if len(v[1:-1]) > 0 else []

    class Struct:
        def __init__(self, **entries):
            self.__dict__.update(entries)

    return Struct(**params)


def save_params(args, model_dir=None):
    """reference myutils.py:29-32 (`key: value` per line)."""
    d = args if isinstance(args, dict) else args.__dict__
    with open((model_dir or d['model_dir']) + '/train-params.txt', 'w') as f:
        for k, v in d.items():
            f.write('{}: {}\n'.format(k, v))


def img_prep_fcn():
    """reference myutils.py:88-89"""
    return lambda x: x / 255. - 0.5
