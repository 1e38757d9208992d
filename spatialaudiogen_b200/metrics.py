"""Per-window evaluation metrics on the GPU: the in-graph metrics of reference model.py:110-154 (STFT distance, LSD,
MSE, SNR), the Hilbert-envelope distance of myutils.py:109-116, the amplitudes of eval.py:197-198 and the
spherical-harmonic RMS energy maps of pyutils/ambisonics (decoder.py:24-28, distance.py:41-52).  All arithmetic runs
in libsag.so (metrics.cu); this module only allocates outputs."""
import ctypes as C

import numpy as np
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


def mel_lsd(preds, targets, audio_rate=48000):
    """reference myutils.compute_lsd_dist (myutils.py:96-106) for a batch: preds, targets (B, T, 3) float32 CUDA ->
    (B, 3) distances between the 128-band mel power spectrograms in dB (librosa melspectrogram restated in metrics.cu)."""
    preds = L.f32(preds)
    targets = L.f32(targets, preds.device)
    if preds.shape != targets.shape or preds.dim() != 3 or preds.shape[2] != 3:
        raise ValueError('preds/targets must both be (B, T, 3), got %s and %s' % (tuple(preds.shape), tuple(targets.shape)))
    B, T, _ = preds.shape
    with torch.cuda.device(preds.device):
        out = torch.empty((B, 3), dtype=torch.float32, device=preds.device)
        L.check(L.lib().sag_mel_lsd(L.ptr(preds), L.ptr(targets), B, T, int(audio_rate), L.ptr(out), L.stream()))
    return out


def compute_lsd_dist(pred, gt, rate):
    """reference myutils.compute_lsd_dist for one window: pred, gt (T, 3) -> (3,)."""
    return mel_lsd(L.f32(pred)[None], L.f32(gt)[None], rate)[0]


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


def spherical_mesh(ang_res):
    """reference distance.py:9-13: (phi_mesh, nu_mesh), each (n_nu, n_phi), radians."""
    phi_rg = np.flip(np.arange(-180., 180., ang_res) / 180. * np.pi, 0)
    nu_rg = np.arange(-90., 90.1, ang_res) / 180. * np.pi
    return np.meshgrid(phi_rg, nu_rg)


def emd_hat(first, second, dist, extra_mass_penalty=-1.0):
    """pyemd.emd semantics (EMD-hat, exact) for `count` pairs of histograms: first, second (count, n) or (n,), dist (n, n).
    Host solver in libsag.so (csrc/emd.cu), float64.  Returns (count,) float64."""
    first = np.ascontiguousarray(np.atleast_2d(np.asarray(first, np.float64)))
    second = np.ascontiguousarray(np.atleast_2d(np.asarray(second, np.float64)))
    dist = np.ascontiguousarray(np.asarray(dist, np.float64))
    if first.shape != second.shape or dist.shape != (first.shape[1], first.shape[1]):
        raise ValueError('emd_hat: histograms %s / %s do not match the %s distance matrix' % (first.shape, second.shape, dist.shape))
    out = np.zeros(first.shape[0], np.float64)
    L.check(L.lib().sag_emd_hat(first.ctypes.data_as(C.c_void_p), second.ctypes.data_as(C.c_void_p), first.shape[1],
                                dist.ctypes.data_as(C.c_void_p), float(extra_mass_penalty), first.shape[0],
                                out.ctypes.data_as(C.c_void_p)))
    return out


def ambix_emd_from_maps(maps1, maps2, ang_res=30.):
    """The two EMD columns of eval-detailed.txt from the RMS energy maps of one 0.1 s window each (reference
    distance.py:100-143 `emd` inside `ambix_emd`, eval.py:190): ground distance = great-circle angle between mesh
    directions; `dir` compares map / n_nodes, `dir2` compares map / (sum(map) + 0.01).  maps (B, n_nu, n_phi), CUDA or
    host.  Returns (dir, dir2), each (B,) float64."""
    m1 = np.asarray(maps1.detach().cpu() if isinstance(maps1, torch.Tensor) else maps1, np.float64)
    m2 = np.asarray(maps2.detach().cpu() if isinstance(maps2, torch.Tensor) else maps2, np.float64)
    B = m1.shape[0]
    phi_mesh, nu_mesh = spherical_mesh(ang_res)
    if m1.shape[1:] != phi_mesh.shape or m2.shape != m1.shape:
        raise ValueError('maps %s / %s do not match the %s mesh' % (m1.shape, m2.shape, phi_mesh.shape))
    p_mesh = np.stack((np.cos(nu_mesh) * np.cos(phi_mesh), np.cos(nu_mesh) * np.sin(phi_mesh), np.sin(nu_mesh)), 0).reshape((3, -1))
    ang_dist = np.arccos(np.clip(np.dot(p_mesh.T, p_mesh), -1., 1.))       # distance.py:106-109
    m1, m2 = m1.reshape(B, -1), m2.reshape(B, -1)
    n_nodes = m1.shape[1]
    d1 = emd_hat(m1 / n_nodes, m2 / n_nodes, ang_dist)                      # distance.py:124
    d2 = emd_hat(m1 / (m1.sum(1, keepdims=True) + 0.01), m2 / (m2.sum(1, keepdims=True) + 0.01), ang_dist)   # :125
    return d1, d2
