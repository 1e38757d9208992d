"""Per-window evaluation metrics on the GPU: the in-graph metrics of reference model.py:110-154 (STFT distance, LSD,
MSE, SNR), the Hilbert-envelope distance of myutils.py:109-116, the amplitudes of eval.py:197-198 and the
spherical-harmonic RMS energy maps of pyutils/ambisonics (decoder.py:24-28, distance.py:41-52).  All arithmetic runs
in libsag.so (metrics.cu); this module only allocates outputs."""
import ctypes as C

import torch

from . import _lib as L


def window_metrics(preds, targets, audio_rate=48000, envelope=True):
    """preds, targets: (B, T, 3) float32 CUDA.  Returns dict of (B,3) tensors 'stft','lsd','mse','snr','env' and
    'amp' (B,2) = max|pred|, max|gt|."""
    preds, targets = L.f32(preds), L.f32(targets, preds.device if isinstance(preds, torch.Tensor) and preds.is_cuda else None)
    if preds.shape != targets.shape or preds.dim() != 3 or preds.shape[2] != 3:
        raise ValueError('preds/targets must both be (B, T, 3), got %s and %s' % (tuple(preds.shape), tuple(targets.shape)))
    B, T, _ = preds.shape
    dev = preds.device
    with torch.cuda.device(dev):
        out = {k: torch.empty((B, 3), dtype=torch.float32, device=dev) for k in ('stft', 'lsd', 'mse', 'snr', 'env')}
        out['amp'] = torch.empty((B, 2), dtype=torch.float32, device=dev)
        scratch = torch.empty(max(int(L.lib().sag_metrics_scratch_bytes(B, T)), 256), dtype=torch.uint8, device=dev)
        L.check(L.lib().sag_metrics(L.ptr(preds), L.ptr(targets), B, T, int(audio_rate), L.ptr(out['stft']), L.ptr(out['lsd']),
                                    L.ptr(out['mse']), L.ptr(out['snr']), L.ptr(out['env']) if envelope else None,
                                    L.ptr(out['amp']), C.c_void_p(scratch.data_ptr()), L.stream()))
    if not envelope:
        del out['env']
    return out


def compute_envelope_dist(pred, gt):
    """reference myutils.py:109-116 for one window: pred, gt (T, 3) -> (3,)."""
    r = window_metrics(L.f32(pred)[None], L.f32(gt)[None])
    return r['env'][0]


def ambix_rms_map(ambi, ang_res=30.):
    """AmbiDecoder 'projection' decode on the spherical mesh + RMS over time, rows flipped (reference decoder.py:24-28,
    distance.py:9-13,41-52).  ambi: (B, T, 4) [W,Y,Z,X] float32 CUDA -> (B, n_nu, n_phi)."""
    ambi = L.f32(ambi)
    if ambi.dim() == 2:
        ambi = ambi[None]
    if ambi.dim() != 3 or ambi.shape[2] != 4:
        raise ValueError('ambi must be (B, T, 4), got %s' % (tuple(ambi.shape),))
    B, T, _ = ambi.shape
    n_nu, n_phi = C.c_int(), C.c_int()
    L.check(L.lib().sag_sh_rms_dims(float(ang_res), C.byref(n_nu), C.byref(n_phi)))
    with torch.cuda.device(ambi.device):
        rms = torch.empty((B, n_nu.value, n_phi.value), dtype=torch.float32, device=ambi.device)
        L.check(L.lib().sag_sh_rms(L.ptr(ambi), B, T, float(ang_res), L.ptr(rms), L.stream()))
    return rms
