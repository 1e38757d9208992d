"""The stage methods of the reference's SptAudioGen (model.py:161-354) as eager compositions of libsag.so's stage entry
points (sag_conv2d / sag_deconv2d / sag_fc / sag_resnet18, myutils.stft / istft).

The reference calls these methods once, while it builds the graph; `inference_ops` here runs the whole graph as ONE
`sag_forward` and never goes through them.  They exist so that code written against the reference's class (feature
extraction from one tower, a decoder fed with other features, ...) finds the same names, argument meaning and tensor
layouts: NHWC activations, `(B, 1, frames, bins)` complex STFTs, `(B, snd_dur, 3, 1, K)` localization weights.  The dense
contractions run in the handle's precision on the same kernels as the forward; the reshapes / tiles / concats / crops
between them -- pure data movement in the reference graph too -- are torch views and copies on the caller's stream, and so
is the sigmoid-mask product of separation_ops (the forward fuses it into the inverse-STFT kernel).

Tests: the glue on the CPU against the oracle with the four primitives stood in (tests/test_host.py); end to end on the GPU
in tests/test_gpu_parity.py::test_stage_methods_match_oracle."""
import ctypes as C

import torch

from . import _lib as L
from . import myutils
from .definitions import AUDIO, VIDEO, FLOW, NO_SEPARATION, FREQ_MASK

AUDIO_FILTERS = [32, 64, 128, 256, 512]                                   # model.py:162-164 / 282-284
AUDIO_KERNELS = [(7, 16), (3, 7), (3, 5), (3, 5), (3, 5)]
AUDIO_STRIDES = [(4, 8), (2, 4), (2, 2), (1, 1), (1, 1)]


class StageOps(object):
    """Mixin of SptAudioGen: needs self._h, self._w (name -> host array), self.dims, self.params, self.precision, self.device."""

    # ---- primitives: one C-ABI call each ---------------------------------------------------------------------------
    def _dev_weight(self, name):
        cache = self.__dict__.setdefault('_wd', {})
        if name not in cache:
            cache[name] = torch.as_tensor(self._w[name]).to(self.device).contiguous()
        return cache[name]

    def _prec(self):
        return L.PRECISIONS[self.precision]

    def _conv(self, scope, x, stride, same, relu):
        """tfw.conv_2d with bias (core.py:156-220).  x (N, H, W, Cin) -> (N, OH, OW, Cout)."""
        x = L.f32(x, self.device)
        w, b = self._dev_weight(scope + '/weights'), self._dev_weight(scope + '/biases')
        n, h, wd, cin = x.shape
        kh, kw, _, cout = w.shape
        if same:
            oh, ow = -(-h // stride[0]), -(-wd // stride[1])
        else:
            oh, ow = (h - kh) // stride[0] + 1, (wd - kw) // stride[1] + 1
        y = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=self.device)
        L.check(L.lib().sag_conv2d(L.ptr(x), n, h, wd, cin, L.ptr(w), kh, kw, cout, stride[0], stride[1], int(same), L.ptr(b), int(relu),
                                   L.ptr(y), self._prec(), L.stream()))
        return y

    def _deconv(self, scope, x, stride, relu):
        """tfw.deconv_2d VALID (core.py:96-153).  x (N, H, W, Cin), weights (kh, kw, Cout, Cin)."""
        x = L.f32(x, self.device)
        w, b = self._dev_weight(scope + '/weights'), self._dev_weight(scope + '/biases')
        n, h, wd, cin = x.shape
        kh, kw, cout, _ = w.shape
        y = torch.empty((n, (h - 1) * stride[0] + kh, (wd - 1) * stride[1] + kw, cout), dtype=torch.float32, device=self.device)
        L.check(L.lib().sag_deconv2d(L.ptr(x), n, h, wd, cin, L.ptr(w), kh, kw, cout, stride[0], stride[1], L.ptr(b), int(relu), L.ptr(y),
                                     self._prec(), L.stream()))
        return y

    def _fc(self, scope, x, relu=True):
        """tfw.fully_connected (core.py:43-93): acts on the last axis."""
        x = L.f32(x, self.device)
        w, b = self._dev_weight(scope + '/weights'), self._dev_weight(scope + '/biases')
        rows = x.numel() // x.shape[-1]
        y = torch.empty(tuple(x.shape[:-1]) + (w.shape[1],), dtype=torch.float32, device=self.device)
        L.check(L.lib().sag_fc(L.ptr(x), rows, x.shape[-1], L.ptr(w), w.shape[1], L.ptr(b), int(relu), L.ptr(y), self._prec(), L.stream()))
        return y

    def _resnet(self, scope, x):
        """ResNet18.inference_ops(truncate_at='conv5_2'), batch-statistics BN (resnet.py:123-190).  x (N, H, W, 3)."""
        x = L.f32(x, self.device)
        if x.dim() != 4 or tuple(x.shape[1:]) != (self._frame[0], self._frame[1], 3):
            # sag_resnet18 takes no spatial arguments: it reads B * frame_h * frame_w * 3 floats of the handle's frame size
            raise ValueError('visual tower input must be (N, %d, %d, 3) -- the frame_size this model was built with -- got %s'
                             % (self._frame[0], self._frame[1], tuple(x.shape)))
        n, h, wd, _ = x.shape
        y = torch.empty((n, -(-h // 32), -(-wd // 32), 512), dtype=torch.float32, device=self.device)
        ws = self._workspace(n)
        L.check(L.lib().sag_resnet18(self._h, scope.encode(), L.ptr(x), n, L.ptr(y), C.c_void_p(ws.data_ptr()), ws.numel(), L.stream()))
        return y

    # ---- model.py:161-187 ------------------------------------------------------------------------------------------
    def audio_encoder_ops(self, stft):
        """stft (B, 1, frames, bins) complex -> [magnitude crop, conv1 .. conv5] NHWC (the U-Net skip list)."""
        d = self.dims
        x = stft[:, :, d.enc_ss:d.enc_tt, :].permute(0, 2, 3, 1).abs().contiguous()
        downsampling_l = [x]
        for l in range(len(AUDIO_FILTERS)):
            x = self._conv('audio_encoder/conv%d' % (l + 1), x, AUDIO_STRIDES[l], same=False, relu=True)
            downsampling_l.append(x)
        return downsampling_l

    # ---- model.py:189-201 ------------------------------------------------------------------------------------------
    def visual_encoding_ops(self, inp, is_training=True, finetune=False, scope=None):
        """inp (B, T, H, W, 3) -> (B*T, H/32, W/32, 512).  The reference passes `finetune` (always True at its call sites) as
        the tower's is_training, i.e. batch statistics; `is_training` is ignored there too."""
        if scope not in ('video_encoder', 'flow_encoder'):
            raise ValueError("scope must be 'video_encoder' or 'flow_encoder'")
        sh = tuple(inp.shape)
        return self._resnet(scope, inp.reshape((sh[0] * sh[1],) + sh[2:]))

    # ---- model.py:203-239 ------------------------------------------------------------------------------------------
    def bottleneck_ops(self, x_enc, use_audio=True):
        if len(x_enc) == 0:
            return None
        bottleneck = []
        audio_sz = tuple(x_enc[AUDIO][-1].shape)
        for k in [AUDIO, VIDEO, FLOW]:
            if k == AUDIO and not use_audio:
                continue
            if k in x_enc:
                x = x_enc[k][-1] if k == AUDIO else x_enc[k]
                if k != AUDIO:
                    x = self._fc('bottleneck/%s-fc-red' % k, x)
                sz = tuple(x.shape)
                x = x.reshape((sz[0], sz[1], sz[2] * sz[3]) if k == AUDIO else (sz[0], 1, sz[1] * sz[2] * sz[3]))
                x = self._fc('bottleneck/%s-fc' % k, x)
                if k in [VIDEO, FLOW]:
                    x = x.repeat(1, audio_sz[1], 1)
                bottleneck.append(x)
        return torch.cat(bottleneck, 2)

    # ---- model.py:241-271 ------------------------------------------------------------------------------------------
    def localization_ops(self, x):
        """x (B, NF, D) -> weights (B, snd_dur, 3, 1, K), biases (B, snd_dur, 3, 1) like the reference (tiled over time)."""
        num_out = (self.ambi_order + 1) ** 2 - self.ambi_order ** 2
        num_in = self.ambi_order ** 2
        units = list(self.params.loc_fc_units)
        for i in range(len(units)):
            x = self._fc('localization/fc%d' % (i + 1), x)
        x = self._fc('localization/fc%d' % (len(units) + 1), x, relu=False)
        sz = tuple(x.shape)
        k1 = self.params.sep_num_tracks + 1
        x = x.reshape(sz[0], sz[1], num_out, num_in, k1)
        x = x.unsqueeze(2).repeat(1, 1, self.snd_dur // sz[1], 1, 1, 1).reshape(sz[0], self.snd_dur, num_out, num_in, k1)
        return x[..., :-1], x[..., -1]

    # ---- model.py:273-354 ------------------------------------------------------------------------------------------
    def separation_ops(self, mono, stft, audio_enc, feats, scope='separation'):
        """mono (B, 1, snd_size); stft (B, 1, frames, bins) complex; audio_enc = audio_encoder_ops(stft); feats =
        bottleneck_ops(...).  Returns x_sep (B, 1, K, snd_dur)."""
        d = self.dims
        if self.separation == NO_SEPARATION:
            ss = self.snd_contx // 2
            return mono[:, :, ss:ss + self.snd_dur].unsqueeze(1)
        if self.separation != FREQ_MASK:
            raise ValueError('Unknown separation mode.')
        feats = self._fc(scope + '/fc-feats', feats)
        enc_sz = tuple(audio_enc[-1].shape)
        feats = feats.unsqueeze(2).repeat(1, 1, enc_sz[2], 1)
        x = torch.cat([audio_enc[-1], feats], dim=3)
        n_chann_in = mono.shape[1]
        for l in reversed(range(len(AUDIO_FILTERS))):
            x = self._deconv(scope + '/deconv%d' % (l + 1), x, AUDIO_STRIDES[l], relu=False)
            if l == 0:
                break
            x = torch.cat((torch.relu(x), audio_enc[l]), 3)                  # zip(..., audio_enc[:-1]) reversed: skip of level l
        stft_c = stft[:, :, d.mask_ss:d.mask_tt]
        x = x[:, d.mask_ss - d.mask_skip:d.mask_tt - d.mask_skip, :]
        x = x.permute(0, 3, 1, 2)
        x_sz = tuple(x.shape)
        x = x.reshape(x_sz[0], n_chann_in, -1, x_sz[2], x_sz[3])
        f_mask = torch.sigmoid(x).to(stft.dtype)
        stft_sep = (stft_c.unsqueeze(2) * f_mask).contiguous()
        x_sep = myutils.istft(stft_sep, 4)
        return x_sep[:, :, :, d.final_crop:d.final_crop + self.snd_dur]
