"""Checkpoint layout (variable name -> shape, TF layouts) and synthetic initialisers.

The variable names / shapes are the reference's checkpoint contract:
  wrappers create `<name>/weights`, `<name>/biases`, `<name>/bn/{beta,gamma,moving_mean,moving_variance}`
  (reference pyutils/tflib/wrappers/core.py:21,69,127,191,210); conv weights HWIO (core.py:184),
  transposed-conv weights [kh,kw,Cout,Cin] (core.py:118), FC [in,out] (core.py:67); scopes from
  model.py:378-428; ResNet-18 names from pyutils/tflib/models/image/resnet.py:123-236.
Initialisers follow core.py:14 (zero biases), core.py:34 (Xavier uniform) and model.py:255
(localization/fc3 truncated normal, stddev 1e-3).
"""
from collections import OrderedDict

import numpy as np

from .definitions import AUDIO, VIDEO, FLOW, FREQ_MASK

AUDIO_FILTERS = [32, 64, 128, 256, 512]                              # model.py:162 / :283
AUDIO_KERNELS = [(7, 16), (3, 7), (3, 5), (3, 5), (3, 5)]            # model.py:163 / :284
AUDIO_STRIDES = [(4, 8), (2, 4), (2, 2), (1, 1), (1, 1)]             # model.py:164 / :285

RESNET_BLOCKS = [('conv2_1', 64, 64, False), ('conv2_2', 64, 64, False),
                 ('conv3_1', 64, 128, True), ('conv3_2', 128, 128, False),
                 ('conv4_1', 128, 256, True), ('conv4_2', 256, 256, False),
                 ('conv5_1', 256, 512, True), ('conv5_2', 512, 512, False)]


def resnet18_shapes(scope, in_channels=3, with_logits=False):
    sh = OrderedDict()

    def bn(p, c):
        for k in ('beta', 'gamma', 'moving_mean', 'moving_variance'):
            sh['%s/bn/%s' % (p, k)] = (c,)

    p = scope + '/' if scope else ''
    sh[p + 'conv1/conv/weights'] = (7, 7, in_channels, 64)
    bn(p + 'conv1/conv', 64)
    for name, cin, cout, first in RESNET_BLOCKS:
        if first:
            sh['%s%s/shortcut/weights' % (p, name)] = (1, 1, cin, cout)
        sh['%s%s/conv_1/weights' % (p, name)] = (3, 3, cin, cout)
        bn('%s%s/conv_1' % (p, name), cout)
        sh['%s%s/conv_2/weights' % (p, name)] = (3, 3, cout, cout)
        bn('%s%s/conv_2' % (p, name), cout)
    if with_logits:
        sh[p + 'logits/fc/weights'] = (512, 1000)
        sh[p + 'logits/fc/biases'] = (1000,)
    return sh


def variable_shapes(encoders, separation=FREQ_MASK, sep_num_tracks=32, loc_fc_units=(512, 512), ambi_order=1):
    """OrderedDict tf_var_name -> shape for a model configuration (214 vars / 49 005 763 params for A+V+F)."""
    sh = OrderedDict()
    if AUDIO in encoders:
        cin = 1
        for l, (nf, ks) in enumerate(zip(AUDIO_FILTERS, AUDIO_KERNELS)):
            sh['audio_encoder/conv%d/weights' % (l + 1)] = ks + (cin, nf)
            sh['audio_encoder/conv%d/biases' % (l + 1)] = (nf,)
            cin = nf
    for k in (VIDEO, FLOW):
        if k in encoders:
            sh.update(resnet18_shapes(k + '_encoder'))
    D = 0
    if AUDIO in encoders:
        sh['bottleneck/audio-fc/weights'] = (3 * 2 * 512, 1024)
        sh['bottleneck/audio-fc/biases'] = (1024,)
        D += 1024
    for k in (VIDEO, FLOW):
        if k in encoders:
            sh['bottleneck/%s-fc-red/weights' % k] = (512, 128)
            sh['bottleneck/%s-fc-red/biases' % k] = (128,)
            sh['bottleneck/%s-fc/weights' % k] = (7 * 14 * 128, 512)
            sh['bottleneck/%s-fc/biases' % k] = (512,)
            D += 512
    num_out = (ambi_order + 1) ** 2 - ambi_order ** 2
    num_in = ambi_order ** 2
    prev = D
    for i, u in enumerate(loc_fc_units):
        sh['localization/fc%d/weights' % (i + 1)] = (prev, u)
        sh['localization/fc%d/biases' % (i + 1)] = (u,)
        prev = u
    n3 = num_out * num_in * (sep_num_tracks + 1)
    sh['localization/fc%d/weights' % (len(loc_fc_units) + 1)] = (prev, n3)
    sh['localization/fc%d/biases' % (len(loc_fc_units) + 1)] = (n3,)
    if separation == FREQ_MASK:
        sh['separation/fc-feats/weights'] = (D, 512)
        sh['separation/fc-feats/biases'] = (512,)
        outs = [sep_num_tracks] + AUDIO_FILTERS[:-1]
        for l in reversed(range(5)):
            cin = 1024 if l == 4 else 2 * AUDIO_FILTERS[l]
            sh['separation/deconv%d/weights' % (l + 1)] = AUDIO_KERNELS[l] + (outs[l], cin)
            sh['separation/deconv%d/biases' % (l + 1)] = (outs[l],)
    return sh


def _fans(name, shape):
    if len(shape) == 4:
        rf = shape[0] * shape[1]
        return rf * shape[2], rf * shape[3]          # TF xavier: fan_in = k*k*shape[-2], fan_out = k*k*shape[-1]
    return shape[0], shape[1]


def init_weights(encoders, separation=FREQ_MASK, sep_num_tracks=32, loc_fc_units=(512, 512), seed=1234,
                 resnet_npy=None, stress=False):
    """Synthetic weights in the checkpoint layout.

    stress=False: the reference's initialisers (Xavier-uniform, zero biases, fc3 ~ truncN(0, 1e-3), BN
    gamma=1/beta=0); ResNet towers from `resnet_npy` (the reference's resnet18.npy) when given.
    stress=True: additionally random biases / BN affine and a 100x larger fc3, so that every term of the
    forward is exercised by parity tests.
    """
    rng = np.random.RandomState(seed)
    shapes = variable_shapes(encoders, separation, sep_num_tracks, loc_fc_units)
    pre = None
    if resnet_npy is not None:
        pre = np.load(resnet_npy, allow_pickle=True, encoding='latin1').item()
    fc3 = 'localization/fc%d/weights' % (len(loc_fc_units) + 1)
    W = OrderedDict()
    for name, shape in shapes.items():
        scope = name.split('/')[0]
        if pre is not None and scope in ('video_encoder', 'flow_encoder'):
            W[name] = np.ascontiguousarray(pre[name[len(scope) + 1:]], dtype=np.float32)
            assert W[name].shape == tuple(shape), (name, W[name].shape, shape)
            continue
        leaf = name.split('/')[-1]
        if leaf == 'weights':
            if name == fc3:
                std = 1e-1 if stress else 1e-3
                w = rng.randn(*shape)
                bad = np.abs(w) > 2
                while bad.any():                          # truncated normal: resample beyond 2 sigma
                    w[bad] = rng.randn(int(bad.sum()))
                    bad = np.abs(w) > 2
                W[name] = (w * std).astype(np.float32)
            else:
                fi, fo = _fans(name, shape)
                lim = np.sqrt(6.0 / (fi + fo))
                W[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf == 'biases':
            W[name] = (rng.randn(*shape) * 0.1).astype(np.float32) if stress else np.zeros(shape, np.float32)
        elif leaf == 'gamma':
            W[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if stress else np.ones(shape, np.float32)
        elif leaf == 'beta':
            W[name] = (rng.randn(*shape) * 0.1).astype(np.float32) if stress else np.zeros(shape, np.float32)
        elif leaf == 'moving_mean':
            W[name] = np.zeros(shape, np.float32)
        elif leaf == 'moving_variance':
            W[name] = np.ones(shape, np.float32)
        else:
            raise KeyError(name)
    return W


def num_params(W):
    return int(sum(int(np.prod(v.shape)) for v in W.values()))
