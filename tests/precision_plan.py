"""Dev tool (test infrastructure, runs on the CPU): which contractions of the forward need which operand precision to keep
the ambisonic waveform within the north-star tolerance (<= 1e-3 of max |ref|)?

The fp64 oracle is run with the OPERANDS of chosen layer groups rounded the way a tensor-core precision rounds them
(accumulation stays fp64 -- the fp32 TMEM accumulators contribute ~1e-7, far below any operand rounding):
  exact : no rounding                       f16   : x, w -> fp16           (one MMA per K step)
  bf16  : x, w -> bf16 (one MMA)            f16a2 : x -> fp16 hi+lo, w -> fp16   (two MMAs: activations fp32-grade)
  bf16x3: hi+lo bf16 of both, lo*lo dropped (what SAG_PREC_BF16X3 computes)
  f16x3 : hi+lo fp16 of both, lo*lo dropped
Groups: audio_encoder, video_encoder / flow_encoder (ResNet towers), fc (bottleneck / localization / fc-feats),
decoder (deconv5..1).

    python tests/precision_plan.py [--batch 8] [--resnet-npy PATH] [--stress 0|1]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sag_oracle as O                      # noqa: E402
from spatialaudiogen_b200 import weights as Wt           # noqa: E402


def _hi_lo(t, dt):
    hi = t.to(dt).to(t.dtype)
    lo = (t - hi).to(dt).to(t.dtype)
    return hi, lo


def _contract(fn, x, w, mode):
    """fn(x, w) is bilinear; emulate the operand rounding of `mode`."""
    if mode == 'exact':
        return fn(x, w)
    if mode in ('bf16', 'f16'):
        dt = torch.bfloat16 if mode == 'bf16' else torch.float16
        return fn(x.to(dt).to(x.dtype), w.to(dt).to(w.dtype))
    if mode in ('bf16x3', 'f16x3'):
        dt = torch.bfloat16 if mode == 'bf16x3' else torch.float16
        xh, xl = _hi_lo(x, dt)
        wh, wl = _hi_lo(w, dt)
        return fn(xh, wh) + fn(xh, wl) + fn(xl, wh)
    if mode in ('bf16a2', 'f16a2'):                      # activations hi+lo, weights hi only
        dt = torch.bfloat16 if mode == 'bf16a2' else torch.float16
        xh, xl = _hi_lo(x, dt)
        wh = w.to(dt).to(w.dtype)
        return fn(xh + xl, wh)
    if mode in ('bf16w2', 'f16w2'):                      # weights hi+lo, activations hi only
        dt = torch.bfloat16 if mode == 'bf16w2' else torch.float16
        wh, wl = _hi_lo(w, dt)
        xh = x.to(dt).to(x.dtype)
        return fn(xh, wh + wl)
    raise ValueError(mode)


class Plan(object):
    def __init__(self, modes):
        self.modes = modes                                # group -> mode

    def mode_of(self, scope):
        top = scope.split('/')[0]
        if top in ('video_encoder', 'flow_encoder'):
            g = 'tower'
        elif top == 'audio_encoder':
            g = 'audio_encoder'
        elif top == 'separation' and 'deconv' in scope:
            g = 'decoder'
        else:
            g = 'fc'
        return self.modes.get(g, 'exact')


_PLAN = Plan({})
_conv2d, _deconv, _fc = O.conv_2d, O.deconv_2d, O.fully_connected


def conv_2d(W, scope, x, kernel_stride, padding, relu, use_bias=True, use_bn=False, bn_train=True):
    mode = _PLAN.mode_of(scope)
    y = _contract(lambda a, b: O.tf_conv2d(a, b, kernel_stride, padding), x, W[scope + '/weights'], mode)
    if use_bn:
        y = O.tf_batch_norm_train(y, W[scope + '/bn/gamma'], W[scope + '/bn/beta'])
    elif use_bias:
        y = y + W[scope + '/biases']
    return torch.relu(y) if relu else y


def deconv_2d(W, scope, x, stride, relu=False):
    mode = _PLAN.mode_of(scope)
    y = _contract(lambda a, b: O.tf_conv2d_transpose_valid(a, b, stride), x, W[scope + '/weights'], mode) + W[scope + '/biases']
    return torch.relu(y) if relu else y


def fully_connected(W, scope, x, relu=True):
    mode = _PLAN.mode_of(scope)
    y = _contract(lambda a, b: a.reshape(-1, a.shape[-1]) @ b, x, W[scope + '/weights'], mode) + W[scope + '/biases']
    y = y.reshape(tuple(x.shape[:-1]) + (-1,))
    return torch.relu(y) if relu else y


O.conv_2d, O.deconv_2d, O.fully_connected = conv_2d, deconv_2d, fully_connected


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    global _PLAN
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--resnet-npy', default=None)
    ap.add_argument('--stress', type=int, default=1)
    ap.add_argument('--seed', type=int, default=9)
    ap.add_argument('--encoders', default='audio,video')
    args = ap.parse_args()
    enc = args.encoders.split(',')
    B = args.batch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_parity import _audio, _video, _flow
    W = Wt.init_weights(enc, separation='unet_mask', seed=args.seed, stress=bool(args.stress), resnet_npy=args.resnet_npy)
    a, v = _audio(B, 23), _video(B, 24)
    kw = dict(video=v)
    if 'flow' in enc:
        kw['flow'] = _flow(B, 27)
    ref = O.SptAudioGen(W, encoders=enc, separation='unet_mask', dtype=torch.float64)

    def run(modes):
        global _PLAN
        _PLAN = Plan(modes)
        t0 = time.time()
        y = ref.inference_ops(a, **kw).clone()
        ends = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ref.ends.items()}
        return y, ends, time.time() - t0

    y0, e0, dt = run({})
    print('exact: %.1f s, max|y| %.3e' % (dt, float(y0.abs().max())))
    plans = [
        ('all bf16', dict(audio_encoder='bf16', tower='bf16', fc='bf16', decoder='bf16')),
        ('all bf16x3', dict(audio_encoder='bf16x3', tower='bf16x3', fc='bf16x3', decoder='bf16x3')),
        ('tower bf16, rest bf16x3', dict(audio_encoder='bf16x3', tower='bf16', fc='bf16x3', decoder='bf16x3')),
        ('tower bf16a2, rest bf16x3', dict(audio_encoder='bf16x3', tower='bf16a2', fc='bf16x3', decoder='bf16x3')),
        ('tower bf16w2, rest bf16x3', dict(audio_encoder='bf16x3', tower='bf16w2', fc='bf16x3', decoder='bf16x3')),
        ('tower f16, rest bf16x3', dict(audio_encoder='bf16x3', tower='f16', fc='bf16x3', decoder='bf16x3')),
        ('audio_encoder bf16, rest bf16x3', dict(audio_encoder='bf16', tower='bf16x3', fc='bf16x3', decoder='bf16x3')),
        ('decoder bf16, rest bf16x3', dict(audio_encoder='bf16x3', tower='bf16x3', fc='bf16x3', decoder='bf16')),
        ('fc bf16, rest bf16x3', dict(audio_encoder='bf16x3', tower='bf16x3', fc='bf16', decoder='bf16x3')),
        ('enc+dec bf16, rest bf16x3', dict(audio_encoder='bf16', tower='bf16x3', fc='bf16x3', decoder='bf16')),
        ('enc+dec bf16w2, rest bf16x3', dict(audio_encoder='bf16w2', tower='bf16x3', fc='bf16x3', decoder='bf16w2')),
        ('enc+dec bf16a2, rest bf16x3', dict(audio_encoder='bf16a2', tower='bf16x3', fc='bf16x3', decoder='bf16a2')),
    ]
    print('%-34s %10s %10s %10s %10s' % ('plan', 'waveform', 'conv5_2', 'bottleneck', 'mask_logit'))
    for name, modes in plans:
        y, e, dt = run(modes)
        print('%-34s %10.2e %10.2e %10.2e %10.2e' % (
            name, rel(y, y0), rel(e['video_encoder/conv5_2'], e0['video_encoder/conv5_2']) if 'video_encoder/conv5_2' in e else 0,
            rel(e['bottleneck'], e0['bottleneck']), rel(e['separation/mask_logits'], e0['separation/mask_logits'])))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
