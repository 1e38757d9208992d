"""CPU ORACLE for the spatialaudiogen inference hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch-CPU tensors, fp32 by default, fp64 on request) of the
reference's TF-1.4 graph for the path named in BASELINE.json:north_star.  It is the *checker*:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
The product (spatialaudiogen_b200/) never imports it and has no CPU fallback.

PINNING.  The reference ships no tests, golden vectors or fixtures for this path and its
TF graph cannot be executed here (python2 + tensorflow-gpu==1.4.0rc1, neither installable).  What CAN run here is
the reference's own python around the graph, and that is what pins this file (tests/test_reference_goldens.py against
tests/golden/reference_goldens.npz, generated in the build container by tests/golden/make_reference_goldens.py from
/root/reference through a purely syntactic python-2 shim):
  * pinned by the reference's numpy / scipy code: a1 derived constants (model.py:24-60), a12 envelope distance
    (myutils.py:109-116), a13 mesh / SH matrix / energy maps (distance.py, common.py, decoder.py, position.py),
    f2 reader logic (feeder.py:50-161), load_params (myutils.py:40-85);
  * pinned by the reference's graph-building code (indexing, crops, reshapes, reductions) run eagerly on a numpy stand-in
    for the dozen elementary TF ops it calls: a2 stft, a8 istft (myutils.py:119-211), a11 evaluation_ops
    (model.py:62-154);
  * pinned by the reference's MODEL-BUILDING code -- SptAudioGen.inference_ops with audio_encoder_ops,
    visual_encoding_ops, bottleneck_ops, localization_ops, separation_ops (model.py:161-434), the layer wrappers
    (pyutils/tflib/wrappers/core.py:10-220) and ResNet18 (pyutils/tflib/models/image/resnet.py:110-249) -- executed
    verbatim and eagerly at full size (B=2, audio-only and audio+video+flow) with the weights served by the scoped
    variable names it asks for: a3-a7, a9, a10 wiring (scopes, the 30 / 214 variable names and shapes, layer order,
    strides, paddings, crops, concat / tile / reshape order, mask and mixing arithmetic; restore_pretrained checked
    against the reference's resnet18.npy).  This file agrees with it to 5e-8 (float64) / 2e-5 (float32); likewise the
    deploy loop (deploy.py:90-152 W2XYZ.deploy: batches of 10, zero-padded tail, mono crop, rows [W,Y,Z,X]) run around
    that model code, and the EMD columns' wrapper (distance.py:100-143) run around a pyemd stand-in;
  * STILL RESTATED BY HAND, beneath that: the TF kernels themselves (tf.nn.convolution / conv2d_transpose / max_pool /
    contrib batch_norm / matmul of the un-vendored TensorFlow 1.4.0rc1, requirements.txt:10), evaluated in the generator
    by tap-by-tap float64 numpy loops from their published semantics -- independent of the torch
    calls below but not executed from TensorFlow.  They are anchored by (i) the analytic invariants of SURVEY.md 8c
    (tests/test_oracle.py) and (ii) a semantic known-answer test of the ResNet-18 trunk with the reference's own
    resnet18.npy and test images (correct ImageNet classes; tests/golden/); likewise the restated third-party pieces
    of the widened rows (pyemd's EMD-hat: its defining LP; librosa's mel spectrogram).  For those op kernels alone the
    header says it plainly: parity unpinned against TensorFlow itself.

Every function cites the reference file:line (relative to /root/reference) it follows.
Python-2 integer division of the reference is written `//` here.
"""
from __future__ import annotations

from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

AUDIO, VIDEO, FLOW = 'audio', 'video', 'flow'          # definitions.py:1-4
NO_SEPARATION, FREQ_MASK = 'none', 'unet_mask'         # definitions.py:6-8
FFT_WINDOW = 25 * 0.001                                # definitions.py:10
FFT_OVERLAP_R = 2                                      # definitions.py:11


# --------------------------------------------------------------------------------------
# TF-1.4 op semantics
# --------------------------------------------------------------------------------------

def _same_pads(n, k, s):
    """TF 'SAME': out=ceil(n/s); pad_total=max((out-1)*s+k-n,0); before=total//2, after=rest."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def tf_conv2d(x, w_hwio, stride, padding):
    """tf.nn.convolution NHWC / HWIO cross-correlation (core.py:206).  x: (B,H,W,C)."""
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == 'SAME':
        pt, pb = _same_pads(x.shape[1], kh, sh)
        pl, pr = _same_pads(x.shape[2], kw, sw)
        xn = F.pad(xn, (pl, pr, pt, pb))
    elif padding != 'VALID':
        raise ValueError(padding)
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=(sh, sw))
    return y.permute(0, 2, 3, 1)


def tf_conv2d_transpose_valid(x, w_hwoi, stride):
    """tf.nn.conv2d_transpose VALID (core.py:139-140): y[b,i*sh+p,j*sw+q,co] += x[b,i,j,ci]*w[p,q,co,ci]."""
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w_hwoi.permute(3, 2, 0, 1), stride=(sh, sw))
    return y.permute(0, 2, 3, 1)


def tf_batch_norm_train(x, gamma, beta, eps=1e-3):
    """contrib.layers.batch_norm(decay=.99, scale=True, is_training=True) forward (core.py:209-210):
    batch statistics over (N,H,W), biased variance, eps=0.001."""
    mu = x.mean(dim=(0, 1, 2), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(0, 1, 2), keepdim=True)
    return gamma * (x - mu) / torch.sqrt(var + eps) + beta


def tf_batch_norm_infer(x, gamma, beta, mean, var, eps=1e-3):
    """is_training=False branch (moving statistics) -- only used by the ImageNet known-answer test."""
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta


def tf_max_pool_same_3x3s2(x):
    """tf.nn.max_pool [1,3,3,1]/[1,2,2,1] SAME (resnet.py:135); padded cells ignored (-inf)."""
    pt, pb = _same_pads(x.shape[1], 3, 2)
    pl, pr = _same_pads(x.shape[2], 3, 2)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float('-inf'))
    return F.max_pool2d(xn, 3, 2).permute(0, 2, 3, 1)


# --------------------------------------------------------------------------------------
# wrappers/core.py
# --------------------------------------------------------------------------------------

class Weights(object):
    """name -> torch tensor view of a {tf_var_name: ndarray} dict (checkpoint layout)."""

    def __init__(self, arrays, dtype=torch.float32):
        self.dtype = dtype
        self.t = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in arrays.items()}

    def __getitem__(self, k):
        return self.t[k]

    def __contains__(self, k):
        return k in self.t


def conv_2d(W, scope, x, kernel_stride, padding, relu, use_bias=True, use_bn=False, bn_train=True):
    """core.py:156-220 conv_2d: convolution -> (BN | bias) -> activation."""
    y = tf_conv2d(x, W[scope + '/weights'], kernel_stride, padding)
    if use_bn:
        g, b = W[scope + '/bn/gamma'], W[scope + '/bn/beta']
        if bn_train:
            y = tf_batch_norm_train(y, g, b)
        else:
            y = tf_batch_norm_infer(y, g, b, W[scope + '/bn/moving_mean'], W[scope + '/bn/moving_variance'])
    elif use_bias:
        y = y + W[scope + '/biases']
    return torch.relu(y) if relu else y


def deconv_2d(W, scope, x, stride, relu=False):
    """core.py:96-153 deconv_2d (VALID): conv2d_transpose -> bias -> activation."""
    y = tf_conv2d_transpose_valid(x, W[scope + '/weights'], stride) + W[scope + '/biases']
    return torch.relu(y) if relu else y


def fully_connected(W, scope, x, relu=True):
    """core.py:43-93: acts on the last axis only (reshape-matmul-reshape), bias, activation."""
    y = x.reshape(-1, x.shape[-1]) @ W[scope + '/weights'] + W[scope + '/biases']
    y = y.reshape(tuple(x.shape[:-1]) + (-1,))
    return torch.relu(y) if relu else y


# --------------------------------------------------------------------------------------
# pyutils/tflib/models/image/resnet.py  ResNet18
# --------------------------------------------------------------------------------------

def resnet18(W, scope, x, bn_train=True, truncate_at='conv5_2'):
    """resnet.py:123-198 inference_ops; blocks :200-236.  x: (N,H,W,3) NHWC.  Returns (x, ends)."""
    p = (scope + '/') if scope else ''
    filters = [64, 64, 128, 256, 512]
    ends = OrderedDict()

    def block(x, name):                                    # resnet.py:224-236
        sc = x
        y = conv_2d(W, p + name + '/conv_1', x, 1, 'SAME', True, use_bias=False, use_bn=True, bn_train=bn_train)
        y = conv_2d(W, p + name + '/conv_2', y, 1, 'SAME', False, use_bias=False, use_bn=True, bn_train=bn_train)
        return torch.relu(y + sc)

    def block_first(x, cout, s, name):                     # resnet.py:200-222
        sc = conv_2d(W, p + name + '/shortcut', x, s, 'SAME', False, use_bias=False)   # 1x1/s, no BN, no bias
        y = conv_2d(W, p + name + '/conv_1', x, s, 'SAME', True, use_bias=False, use_bn=True, bn_train=bn_train)
        y = conv_2d(W, p + name + '/conv_2', y, 1, 'SAME', False, use_bias=False, use_bn=True, bn_train=bn_train)
        return torch.relu(y + sc)

    x = conv_2d(W, p + 'conv1/conv', x, 2, 'SAME', True, use_bias=False, use_bn=True, bn_train=bn_train)
    ends['conv'] = x
    x = tf_max_pool_same_3x3s2(x)                          # resnet.py:135
    if truncate_at == 'conv1':
        return x, ends
    for name, first, f in [('conv2_1', False, 0), ('conv2_2', False, 0), ('conv3_1', True, filters[2]),
                           ('conv3_2', False, 0), ('conv4_1', True, filters[3]), ('conv4_2', False, 0),
                           ('conv5_1', True, filters[4]), ('conv5_2', False, 0)]:
        x = block_first(x, f, 2, name) if first else block(x, name)
        ends[name] = x
        if truncate_at == name:
            return x, ends
    x = x.mean(dim=(1, 2))                                 # resnet.py:193-196 (logits; known-answer test only)
    x = fully_connected(W, p + 'logits/fc', x, relu=True)
    ends['fc'] = x
    return x, ends


# --------------------------------------------------------------------------------------
# myutils.py DSP
# --------------------------------------------------------------------------------------

def hann(n, dtype):
    """myutils.py:134 -- periodic Hann, computed in float64 then cast (tf.constant(..., float32))."""
    w = 0.5 - 0.5 * np.cos(2 * np.pi / n * np.arange(n))
    return torch.as_tensor(w).to(dtype)


def stft(inp, wind_size, n_overlap):
    """myutils.py:119-147.  inp (..., n_frames) real -> (..., n_overlap*n_winds, wind_size) complex."""
    inp_sz = list(inp.shape)
    if len(inp_sz) > 2:
        inp = inp.reshape(int(np.prod(inp_sz[:-1])), inp_sz[-1])
    batch_size, n_frames = inp.shape
    n_winds = int(np.floor(n_frames // wind_size)) - 1           # py2 int division inside floor
    x_crops = []
    for ss in range(0, wind_size, wind_size // n_overlap):
        x_crops.append(inp[:, ss:ss + wind_size * n_winds])
    x = torch.stack(x_crops, 1).reshape(batch_size, n_overlap, -1, wind_size)
    x = x * hann(wind_size, inp.dtype)[None, None, :]
    cdt = torch.complex64 if inp.dtype == torch.float32 else torch.complex128
    s = torch.fft.fft(x.to(cdt), dim=-1)                         # tf.fft: unnormalised forward
    s = s.permute(0, 2, 1, 3)
    sz = s.shape
    s = s.reshape(sz[0], sz[1] * sz[2], sz[3])
    if len(inp_sz) > 2:
        s = s.reshape(inp_sz[:-1] + list(s.shape[-2:]))
    return s


def stft_for_loss(signal, window, n_overlap):
    """myutils.py:151-178.  signal (BS,N,nC) -> (BS,nC,nW_total,window_pow2) complex."""
    BS, N, nC = signal.shape
    window = int(2 ** np.ceil(np.log(window) / np.log(2)))
    hw = hann(window, signal.dtype)
    if n_overlap == 1:
        nW = int(float(N) / window)
        if nW > 1:
            if N > window * nW:
                signal = signal[:, :window * nW, :]
            windows = signal.reshape(BS, nW, window, nC)
        else:
            windows = signal
    else:
        windows = []
        stride = int(window / n_overlap)
        for i in range(n_overlap):
            nW = int(float(N - i * stride - 1) / window)
            y = signal[:, (i * stride):(i * stride) + window * nW, :]
            windows.append(y.reshape(BS, nW, window, nC))
        windows = torch.cat(windows, 1)
    windows = windows.permute(0, 3, 1, 2) * hw[None, None, None, :]
    cdt = torch.complex64 if signal.dtype == torch.float32 else torch.complex128
    return torch.fft.fft(windows.to(cdt), dim=-1)


def istft(inp, n_overlap):
    """myutils.py:181-211: real(ifft), de-interleave n_overlap streams, trim, sum / n_overlap (no window)."""
    inp_sz = list(inp.shape)
    if len(inp_sz) > 3:
        inp = inp.reshape(int(np.prod(inp_sz[:-2])), inp_sz[-2], inp_sz[-1])
    batch_size, n_frames, n_freqs = inp.shape
    n_frames = int(int(float(n_frames) / n_overlap) * n_overlap)
    inp = inp[:, :n_frames, :]
    x = torch.fft.ifft(inp, dim=-1).real                          # tf.ifft: 1/N normalised
    x = x.reshape(batch_size, -1, n_overlap, n_freqs).permute(0, 2, 1, 3).reshape(batch_size, n_overlap, -1)
    x_list = list(x.unbind(1))
    skip = n_freqs // n_overlap
    for i in range(n_overlap):
        if i == 0:
            x_list[i] = x_list[i][:, (n_overlap - i - 1) * skip:]
        else:
            x_list[i] = x_list[i][:, (n_overlap - i - 1) * skip:-i * skip]
    x = sum(x_list) / float(n_overlap)
    if len(inp_sz) > 3:
        x = x.reshape(inp_sz[:-2] + [x.shape[-1]])
    return x


def compute_envelope_dist(pred, gt):
    """myutils.py:109-116 (float64 numpy/scipy.signal.hilbert)."""
    from scipy.signal import hilbert
    pred, gt = np.asarray(pred, np.float64), np.asarray(gt, np.float64)
    dist = np.zeros(gt.shape[1])
    for i in range(gt.shape[1]):
        dist[i] = np.sqrt(np.mean((np.abs(hilbert(gt[:, i])) - np.abs(hilbert(pred[:, i]))) ** 2))
    return dist


# --------------------------------------------------------------------------------------
# model.py  SptAudioGen
# --------------------------------------------------------------------------------------

class SptAudioGenParams(object):
    """model.py:10-21"""

    def __init__(self, sep_num_tracks=32, ctx_feats_fc_units=(64, 128, 128), loc_fc_units=(512, 512),
                 sep_freq_mask_fc_units=(256,), sep_fft_window=0.025):
        self.sep_num_tracks = sep_num_tracks
        self.ctx_feats_fc_units = list(ctx_feats_fc_units)
        self.loc_fc_units = list(loc_fc_units)
        self.sep_freq_mask_fc_units = list(sep_freq_mask_fc_units)
        self.sep_fft_window = sep_fft_window


class SptAudioGen(object):
    """model.py:24-434 restated eagerly on CPU tensors.  `weights` = {tf_var_name: ndarray}."""

    def __init__(self, weights, ambi_order=1, audio_rate=48000, video_rate=10, context=1., sample_duration=0.1,
                 encoders=None, separation='none', params=None, dtype=torch.float32):
        assert float(audio_rate) / video_rate == int(audio_rate) // int(video_rate)
        self.dtype = dtype
        self.W = Weights(weights, dtype)
        self.ambi_order = ambi_order
        self.num_ambi_channels = sum([2 * i + 1 for i in range(ambi_order + 1)])
        self.snd_rate, self.vid_rate = audio_rate, video_rate
        self.context, self.duration = context, sample_duration
        self.snd_contx = int(context * audio_rate)                       # model.py:38
        self.snd_dur = int(sample_duration * audio_rate)                 # :39
        self.snd_size = self.snd_contx + self.snd_dur - 1                # :40
        self.encoders = [AUDIO, VIDEO, FLOW] if encoders is None else encoders
        self.separation = separation
        self.params = params or SptAudioGenParams()
        self.ends = OrderedDict()
        self.wind_size = int(self.params.sep_fft_window * self.snd_rate)
        self.wind_size = int(2 ** np.round(np.log2(self.wind_size)))     # :59-60
        self.loc_channels = None
        self.sep_channels = None

    # ---- index helpers (model.py:166-172 and :313-317, 344-346) --------------------------------
    def encoder_crop(self):
        inp_dim = 95.
        ss = (self.snd_contx / 2.) * (4. / self.wind_size)
        ss = int(ss - (inp_dim - 1) / 2.)
        tt = (self.snd_contx / 2. + self.snd_dur) * (4. / self.wind_size)
        tt = int(tt + (inp_dim - 1) / 2.)
        tt = int((np.ceil((tt - ss - inp_dim) / 16.)) * 16 + inp_dim + ss)
        return ss, tt

    def mask_crop(self):
        ss = np.floor((self.snd_contx / 2. - self.wind_size) * (4. / self.wind_size))
        tt = np.ceil((self.snd_contx / 2. + self.snd_dur + self.wind_size) * (4. / self.wind_size))
        inp_dim = 95.
        skip = (self.snd_contx / 2.) * (4. / self.wind_size)
        skip = int(skip - (inp_dim - 1) / 2.)
        return int(ss), int(tt), skip

    def final_crop(self):
        ss = self.snd_contx / 2.
        skip = np.floor((self.snd_contx / 2. - self.wind_size) * (4. / self.wind_size)) * (self.wind_size / 4.)
        skip += 3. * self.wind_size / 4.
        return int(ss - skip)

    # ---- ops --------------------------------------------------------------------------------
    def audio_encoder_ops(self, s):
        """model.py:161-187.  s: (B,1,200,1024) complex -> list of 6 NHWC tensors."""
        n_filters = [32, 64, 128, 256, 512]
        stride = [(4, 8), (2, 4), (2, 2), (1, 1), (1, 1)]
        ss, tt = self.encoder_crop()
        x = s[:, :, ss:tt, :].permute(0, 2, 3, 1).abs().to(self.dtype)
        out = [x]
        for l in range(len(n_filters)):
            x = conv_2d(self.W, 'audio_encoder/conv%d' % (l + 1), x, stride[l], 'VALID', True)
            out.append(x)
        return out

    def visual_encoding_ops(self, inp, scope):
        """model.py:189-201: reshape (B,T,H,W,C)->(B*T,H,W,C); ResNet18 with is_training=finetune=True."""
        x = inp.reshape((inp.shape[0] * inp.shape[1],) + tuple(inp.shape[2:]))
        x, ends = resnet18(self.W, scope, x, bn_train=True, truncate_at='conv5_2')
        for k, v in ends.items():
            self.ends[scope + '/' + k] = v
        return x

    def bottleneck_ops(self, x_enc, use_audio=True):
        """model.py:203-239."""
        bott = []
        audio_sz = x_enc[AUDIO][-1].shape
        for k in [AUDIO, VIDEO, FLOW]:
            if k == AUDIO and not use_audio:
                continue
            if k in x_enc:
                x = x_enc[k][-1] if k == AUDIO else x_enc[k]
                if k != AUDIO:
                    x = fully_connected(self.W, 'bottleneck/%s-fc-red' % k, x)
                sz = x.shape
                x = x.reshape(sz[0], sz[1], sz[2] * sz[3]) if k == AUDIO else x.reshape(sz[0], 1, sz[1] * sz[2] * sz[3])
                x = fully_connected(self.W, 'bottleneck/%s-fc' % k, x)
                if k in [VIDEO, FLOW]:
                    x = x.repeat(1, audio_sz[1], 1)
                bott.append(x)
        return torch.cat(bott, 2)

    def localization_ops(self, x):
        """model.py:241-271."""
        num_out = (self.ambi_order + 1) ** 2 - self.ambi_order ** 2
        num_in = self.ambi_order ** 2
        for i, u in enumerate(self.params.loc_fc_units):
            x = fully_connected(self.W, 'localization/fc%d' % (i + 1), x)
        x = fully_connected(self.W, 'localization/fc%d' % (len(self.params.loc_fc_units) + 1), x, relu=False)
        sz = x.shape
        x = x.reshape(sz[0], sz[1], num_out, num_in, self.params.sep_num_tracks + 1)
        sz = x.shape
        x = x.unsqueeze(2).repeat(1, 1, self.snd_dur // sz[1], 1, 1, 1)
        x = x.reshape(sz[0], self.snd_dur, sz[2], sz[3], sz[4])
        return x[..., :-1], x[..., -1]

    def separation_ops(self, mono, s, audio_enc, feats):
        """model.py:273-354.  mono (B,1,N); s (B,1,200,1024) complex."""
        if self.separation == NO_SEPARATION:
            ss = self.snd_contx // 2
            return mono[:, :, ss:ss + self.snd_dur].unsqueeze(1)
        elif self.separation != FREQ_MASK:
            raise ValueError('Unknown separation mode.')
        n_filters = [32, 64, 128, 256, 512]
        stride = [(4, 8), (2, 4), (2, 2), (1, 1), (1, 1)]
        feats = fully_connected(self.W, 'separation/fc-feats', feats)
        enc_sz = audio_enc[-1].shape
        feats = feats.unsqueeze(2).repeat(1, 1, enc_sz[2], 1)
        x = torch.cat([audio_enc[-1], feats], dim=3)
        n_chann_in = mono.shape[1]
        for l in reversed(range(len(n_filters))):
            x = deconv_2d(self.W, 'separation/deconv%d' % (l + 1), x, stride[l], relu=False)
            if l == 0:
                break
            x = torch.cat((torch.relu(x), audio_enc[l]), 3)          # audio_enc[:-1][l]
        ss, tt, skip = self.mask_crop()
        s_c = s[:, :, ss:tt]
        x = x[:, ss - skip:tt - skip, :]
        x = x.permute(0, 3, 1, 2)
        x_sz = x.shape
        x = x.reshape(x_sz[0], n_chann_in, -1, x_sz[2], x_sz[3])
        self.ends['separation/mask_logits'] = x
        f_mask = torch.sigmoid(x).to(s.dtype)
        stft_sep = s_c.unsqueeze(2) * f_mask
        x_sep = istft(stft_sep, 4)
        c0 = self.final_crop()
        x_sep = x_sep[:, :, :, c0:c0 + self.snd_dur]
        self.ends['separation/all_channels'] = x_sep
        return x_sep

    def inference_ops(self, audio, video=None, flow=None):
        """model.py:356-434.  audio (B,52799,1); video/flow (B,1,224,448,3).  Returns (B,4800,3) [Y,Z,X]."""
        audio = torch.as_tensor(audio).to(self.dtype).permute(0, 2, 1)
        s = stft(audio, self.wind_size, 4)
        self.ends['stft'] = s
        x_enc = {}
        if AUDIO in self.encoders:
            x_enc[AUDIO] = self.audio_encoder_ops(s)
            self.ends['audio_encoder'] = x_enc[AUDIO]
        if VIDEO in self.encoders:
            x_enc[VIDEO] = self.visual_encoding_ops(torch.as_tensor(video).to(self.dtype), 'video_encoder')
        if FLOW in self.encoders:
            x_enc[FLOW] = self.visual_encoding_ops(torch.as_tensor(flow).to(self.dtype), 'flow_encoder')
        feats = self.bottleneck_ops(x_enc, AUDIO in self.encoders)
        self.ends['bottleneck'] = feats
        weights, biases = self.localization_ops(feats)
        self.loc_channels = [weights, biases]
        x_sep = self.separation_ops(audio, s, x_enc[AUDIO] if len(x_enc) else None, feats)
        self.sep_channels = x_sep
        x_sep = x_sep.permute(0, 3, 1, 2)                                 # (B,4800,1,32)
        x_ambi = (weights * x_sep.unsqueeze(2)).sum(4).sum(3) + biases[:, :, :, 0]
        self.ends['decoder/ambix'] = x_ambi
        return x_ambi

    # ---- metrics (model.py:62-154) ----------------------------------------------------------
    @staticmethod
    def _stft_mse_ops(gt, pred, window, overlap):
        d = (stft_for_loss(gt, window, overlap) - stft_for_loss(pred, window, overlap)).abs()
        return (d ** 2).mean(3).mean(2)

    @staticmethod
    def _lsd_ops(gt, pred, window, overlap):
        EPS = 1e-2
        s_gt = stft(gt.permute(0, 2, 1), window, overlap)
        s_pr = stft(pred.permute(0, 2, 1), window, overlap)

        def power_spect(x):
            return 10 * torch.log(x.abs() + EPS) / math.log(10.)
        d = power_spect(s_gt) - power_spect(s_pr)
        return torch.sqrt((d ** 2).mean(3)).mean(2)

    @staticmethod
    def _temporal_mse_ops(gt, pred):
        return ((gt - pred) ** 2).mean(1)

    @staticmethod
    def _temporal_snr_ops(gt, pred):
        EPS = 1e-1
        ps = (gt ** 2).sum(1)
        pn = ((gt - pred) ** 2).sum(1)
        return 10. * torch.log((ps + EPS) / (pn + EPS)) / math.log(10.)

    def evaluation_ops(self, preds, targets, w, mask_channels):
        """model.py:110-154 -> (metrics, stft_ps, lsd_ps, mse_ps, snr_ps)."""
        preds, targets = torch.as_tensor(preds).to(self.dtype), torch.as_tensor(targets).to(self.dtype)
        mask = torch.as_tensor(mask_channels).to(self.dtype)
        num_masked = mask.sum(0).clamp(min=1)
        metrics = OrderedDict()
        window = int(FFT_WINDOW * self.snd_rate)
        overlap = FFT_OVERLAP_R
        stft_ps = self._stft_mse_ops(targets, preds, window, overlap)
        v = (stft_ps * mask).sum(0) / num_masked * 100.
        metrics['stft/avg'] = v.mean()
        for i, ch in zip(range(3), 'YZX'):
            metrics['stft/' + ch] = v[i]
        lsd_ps = self._lsd_ops(targets, preds, window, overlap)
        v = (lsd_ps * mask).sum(0) / num_masked
        metrics['lsd/avg'] = v.mean()
        for i, ch in zip(range(3), 'YZX'):
            metrics['lsd/' + ch] = v[i]
        mse_ps = self._temporal_mse_ops(targets, preds)
        v = (mse_ps * mask).sum(0) / num_masked * 5e3
        metrics['mse/avg'] = v.mean()
        for i, ch in zip(range(3), 'YZX'):
            metrics['mse/' + ch] = v[i]
        snr_ps = self._temporal_snr_ops(targets, preds)
        v = (snr_ps * mask).sum(0) / num_masked
        metrics['snr/avg'] = v.mean()
        for i, ch in zip(range(3), 'YZX'):
            metrics['snr/' + ch] = v[i]
        metrics['pow/pred'] = (preds ** 2).mean(2).mean(0).sum()
        metrics['pow/gt'] = (targets ** 2).mean(2).mean(0).sum()
        return metrics, stft_ps, lsd_ps, mse_ps, snr_ps


# --------------------------------------------------------------------------------------
# deploy.py assembly
# --------------------------------------------------------------------------------------

def deploy_assemble(model, ambix_windows, video_windows=None, flow_windows=None, batch_size=10):
    """deploy.py:112-151: batches of `batch_size` consecutive windows, last batch zero-padded, W from the
    input crop, output rows [W,Y,Z,X] (float64 as numpy promotes).  ambix_windows: (n,52799,C>=1)."""
    n = ambix_windows.shape[0]
    ss = model.snd_contx // 2
    mono, pred = [], []
    for b0 in range(0, n, batch_size):
        a = np.asarray(ambix_windows[b0:b0 + batch_size], np.float64)
        ns = a.shape[0]
        if ns != batch_size:
            a = np.concatenate([a, np.zeros((batch_size - ns,) + a.shape[1:])], 0)
        kw = {}
        for key, src in (('video', video_windows), ('flow', flow_windows)):
            if src is not None:
                v = np.asarray(src[b0:b0 + batch_size], np.float64)
                if ns != batch_size:
                    v = np.concatenate([v, np.zeros((batch_size - ns,) + v.shape[1:])], 0)
                kw[key] = v
        out = model.inference_ops(a[:, :, :1], **kw).to(torch.float32).numpy()
        pred.append(np.copy(out[:ns]).reshape(ns * out.shape[1], out.shape[2]))
        mono.append(np.copy(a[:ns, ss:ss + model.snd_dur, :1]).reshape(-1, 1))
    mono = np.concatenate(mono, 0)
    return np.concatenate((mono, np.concatenate(pred, 0)), 1)


# --------------------------------------------------------------------------------------
# pyutils/ambisonics: position decoder + RMS energy map
# --------------------------------------------------------------------------------------

def _sn3d_norm(n, m):
    """common.py:136-137."""
    return math.sqrt((2. - float(m == 0)) * float(math.factorial(n - abs(m))) / float(math.factorial(n + abs(m))))


def spherical_harmonic_mn(order, degree, phi, nu):
    """common.py:151-157 (ACN/SN3D defaults), scipy.special.lpmv as in the reference."""
    from scipy.special import lpmv
    norm = _sn3d_norm(order, degree)
    return (-1) ** degree * norm * lpmv(abs(degree), order, np.sin(nu)) * \
        (np.cos(abs(degree) * phi) if degree >= 0 else np.sin(abs(degree) * phi))


def spherical_harmonics_matrix(phis, nus, max_order=1):
    """common.py:160-178 with ACN index i -> (order=int(sqrt(i)), degree=i-order^2-order) (common.py:89-94)."""
    nch = (max_order + 1) ** 2
    Y = np.zeros((len(phis), nch))
    for p, (phi, nu) in enumerate(zip(phis, nus)):
        for i in range(nch):
            order = int(math.sqrt(i))
            degree = i - order ** 2 - order
            Y[p, i] = spherical_harmonic_mn(order, degree, phi, nu)
    return Y


def _position_polar_roundtrip(phi, nu):
    """position.py:23-37 set_polar -> calc_cartesian -> calc_polar (atan2 round trip, r=1)."""
    x = math.cos(phi) * math.cos(nu)
    y = math.sin(phi) * math.cos(nu)
    z = math.sin(nu)
    return math.atan2(y, x), math.atan2(z, math.sqrt(x ** 2 + y ** 2))


def spherical_mesh(angular_res):
    """distance.py:9-13."""
    phi_rg = np.flip(np.arange(-180., 180., angular_res) / 180. * np.pi, 0)
    nu_rg = np.arange(-90., 90.1, angular_res) / 180. * np.pi
    return np.meshgrid(phi_rg, nu_rg)


def ambix_rms_map(ambi, angular_res=30.):
    """distance.py:17-52 for one window: decode (T,4) on the mesh (decoder.py:24-26), RMS over T, flipud."""
    phi_mesh, nu_mesh = spherical_mesh(angular_res)
    pts = [_position_polar_roundtrip(p, n) for p, n in zip(phi_mesh.reshape(-1), nu_mesh.reshape(-1))]
    Y = spherical_harmonics_matrix([p[0] for p in pts], [p[1] for p in pts], 1)
    decoded = np.dot(np.asarray(ambi, np.float64), Y.T)
    rms = np.sqrt(np.mean(decoded ** 2, 0)).reshape(phi_mesh.shape)
    return np.flipud(rms)


def emd_hat_lp(first, second, dist, extra_mass_penalty=-1.0):
    """pyemd.emd(first, second, dist) restated as its defining linear program (Pele & Werman's EMD-hat; pyemd==0.5.1 is
    absent): min sum f_ij d_ij + |sum P - sum Q| * penalty over flows f >= 0 with row sums <= P, column sums <= Q and
    total flow min(sum P, sum Q); penalty = max(dist) for -1.  Solved with scipy's HiGHS -- independent of the
    product's min-cost-flow solver."""
    from scipy.optimize import linprog
    p, q, D = np.asarray(first, np.float64), np.asarray(second, np.float64), np.asarray(dist, np.float64)
    n = p.size
    pen = D.max() if extra_mass_penalty < 0 else extra_mass_penalty
    rows = np.zeros((2 * n, n * n))
    for i in range(n):
        rows[i, i * n:(i + 1) * n] = 1.0
        rows[n + i, i::n] = 1.0
    res = linprog(D.reshape(-1), A_ub=rows, b_ub=np.concatenate((p, q)), A_eq=np.ones((1, n * n)), b_eq=[min(p.sum(), q.sum())],
                  bounds=(0, None), method='highs')
    return float(res.fun + abs(p.sum() - q.sum()) * pen)


def ambix_emd(ambi1, ambi2, ang_res=30.):
    """distance.py:129-143 for one 0.1 s window (one visualizer frame): ambi (T, 4) -> (emd_dir, emd_dir2)."""
    phi_mesh, nu_mesh = spherical_mesh(ang_res)
    p_mesh = np.stack((np.cos(nu_mesh) * np.cos(phi_mesh), np.cos(nu_mesh) * np.sin(phi_mesh), np.sin(nu_mesh)), 0).reshape((3, -1))
    ang_dist = np.dot(p_mesh.T, p_mesh)
    ang_dist[ang_dist >= 1] = 1
    ang_dist[ang_dist <= -1] = -1
    ang_dist = np.arccos(ang_dist)                                        # distance.py:106-109
    m1, m2 = ambix_rms_map(ambi1, ang_res).reshape(-1), ambix_rms_map(ambi2, ang_res).reshape(-1)
    n_nodes = m1.size
    return (emd_hat_lp(m1 / n_nodes, m2 / n_nodes, ang_dist),
            emd_hat_lp(m1 / (m1.sum() + 0.01), m2 / (m2.sum() + 0.01), ang_dist))       # distance.py:124-125


def _mel_filter_bank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel of librosa 0.6.0 (htk=False, norm=1): Slaney mel scale (linear below 1 kHz, log above),
    triangles between consecutive mel points, area normalised.  librosa is absent: restated from its algorithm."""
    fmax = float(sr) / 2 if fmax is None else fmax
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return weights * enorm[:, None]


def compute_lsd_dist(pred, gt, rate):
    """myutils.py:96-106: librosa.feature.melspectrogram(y, sr=rate, n_mels=128, fmax=12000) (n_fft 2048, hop 512,
    centred with reflect padding, periodic Hann, power 2) -> 10*log10(|S| + 0.01) -> RMS difference.  (T, 3) x 2 -> (3,)."""
    n_fft, hop = 2048, 512
    basis = _mel_filter_bank(rate, n_fft, 128, 0.0, 12000.0)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)

    def melspec(y):
        y = np.pad(np.asarray(y, np.float64), n_fft // 2, mode='reflect')
        n_frames = 1 + (len(y) - n_fft) // hop
        frames = np.stack([y[f * hop:f * hop + n_fft] * win for f in range(n_frames)], 1)
        return basis.dot(np.abs(np.fft.rfft(frames, axis=0)) ** 2)

    def power_spect(x):
        return 10 * np.log(np.abs(x) + 1e-2) / np.log(10.)

    pred, gt = np.asarray(pred), np.asarray(gt)
    dist = np.zeros(gt.shape[1])
    for i in range(gt.shape[1]):
        dist[i] = np.sqrt(np.mean((power_spect(melspec(gt[:, i])) - power_spect(melspec(pred[:, i]))) ** 2))
    return dist


def energy_map_frames(ambix, snd_rate, video_fps):
    """myutils.py:251-275 (the heat-map arithmetic of gen_360video) with the oracle's own RMS maps."""
    x = np.asarray(ambix, np.float64)[::5]
    window_frames = int((5. / video_fps) * (snd_rate / 5.))
    n_frames = x.shape[0] // window_frames
    maps = []
    for f in range(n_frames):
        r = ambix_rms_map(x[f * window_frames:(f + 1) * window_frames], 5.)
        maps.append((r - r.min()) / (r.max() - r.min() + 0.005))
    out = []
    for f in range(1, n_frames):
        for i in range(5):
            beta = i / 5.
            rms = ((1 - beta) * maps[f - 1] + beta * maps[f]) * 2. - 0.7
            rms[rms < 0] = 0
            out.append(rms)
    return np.stack(out, 0) if out else np.zeros((0, 37, 72))
