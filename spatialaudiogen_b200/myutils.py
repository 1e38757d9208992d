"""The DSP helpers and parameter parsing of the reference's myutils.py that sit on the inference path, with the same
names and argument meaning: stft / istft (myutils.py:119-147, 181-211) run as fused shared-memory FFT kernels in
libsag.so; load_params / img_prep_fcn (myutils.py:40-89) are host-side and kept verbatim in behaviour."""

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


def flow_prep_fcn():
    """reference myutils.py:92-93: `imresize(x, (224, 448), 'nearest')` -- scipy.misc.imresize is PIL's NEAREST resize of the
    uint8 frame (rows, cols order); frames that already are 224 x 448 pass through unchanged."""
    def prep(x):
        from PIL import Image
        x = np.asarray(x)
        if x.dtype != np.uint8:
            raise TypeError('flow_prep_fcn resizes the uint8 frames as decoded (scipy.misc.imresize would rescale other types to bytes)')
        if x.shape[:2] == (224, 448):
            return x
        return np.asarray(Image.fromarray(x).resize((448, 224), Image.NEAREST))
    return prep


def compute_lsd_dist(pred, gt, rate):
    """reference myutils.py:96-106 (mel log-spectral distance per channel); runs on the GPU (metrics.compute_lsd_dist)."""
    from . import metrics
    return metrics.compute_lsd_dist(pred, gt, rate)


def compute_envelope_dist(pred, gt):
    """reference myutils.py:109-116 (Hilbert-envelope distance per channel); runs on the GPU (metrics.compute_envelope_dist)."""
    from . import metrics
    return metrics.compute_envelope_dist(pred, gt)


# ---- deploy post-processing (reference myutils.gen_360video, myutils.py:224-311): the parts that are arithmetic.
# Splitting / muxing the streams (ffmpeg) and the spatial-media metadata injection are external tools and stay out.
def ambix_to_stereo(ambix):
    """myutils.py:285-291 ("binauralize"): left = W + Y, right = W - Y, peak-normalised to 0.95.  ambix (N, 4) [W, Y, Z, X]
    -> (N, 2) float64."""
    import numpy as np
    ambix = np.asarray(ambix, np.float64)
    stereo = np.stack([ambix[:, 0] + ambix[:, 1], ambix[:, 0] - ambix[:, 1]], 1)
    return stereo / (np.abs(stereo).max() / 0.95)


def energy_map_frames(ambix, snd_rate, video_fps, device=None):
    """The heat maps gen_360video overlays on the video (myutils.py:251-275): the ambisonics are decimated by 5, decoded on
    the 5-degree mesh (37 x 72 directions) and reduced to one RMS map per 5 video frames (the GPU kernel behind
    metrics.ambix_rms_map); each map is min-max normalised with the +0.005 guard, consecutive maps are blended linearly over
    the 5 video frames between them, then `2*rms - 0.7` clipped at 0.  Returns (n_video_frames, 37, 72) float32 in [0, 1.3):
    the reference indexes its colour map with int(255*rms) (clipped) and blends with alpha = 0.6*rms."""
    import numpy as np
    import torch
    from . import metrics as M
    x = np.asarray(ambix, np.float32)[::5]                                  # SphericalAmbisonicsVisualizer(ambix[::5], snd_rate/5., 5./fps, 5.)
    window_frames = int((5. / video_fps) * (snd_rate / 5.))                 # distance.py:30
    n_frames = x.shape[0] // window_frames
    if n_frames < 2:
        return np.zeros((0, 37, 72), np.float32)
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    chunks = torch.as_tensor(x[:n_frames * window_frames].reshape(n_frames, window_frames, 4)).to(dev)
    rms = M.ambix_rms_map(chunks, 5.).cpu().numpy()
    lo = rms.reshape(n_frames, -1).min(1)[:, None, None]
    hi = rms.reshape(n_frames, -1).max(1)[:, None, None]
    rms = (rms - lo) / (hi - lo + 0.005)                                    # myutils.py:256,262
    beta = (np.arange(5, dtype=np.float32) / 5.)[None, :, None, None]       # myutils.py:269-270
    out = (1 - beta) * rms[:-1, None] + beta * rms[1:, None]
    out = out.reshape((-1,) + rms.shape[1:]) * 2. - 0.7                     # myutils.py:271
    return np.maximum(out, 0.).astype(np.float32)
