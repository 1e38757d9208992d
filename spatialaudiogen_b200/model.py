"""SptAudioGen: the reference's model.py operator API (reference model.py:10-434) backed by libsag.so.

Same constructor arguments, attribute names (snd_contx, snd_dur, snd_size, wind_size, num_ambi_channels, encoders,
separation, ends, loc_channels, sep_channels, init_ops) and method names as the reference; the TF graph +
`sess.run` is replaced by eager calls on `torch.cuda` tensors that go straight to hand-written sm_100a kernels
through the C ABI (include/sag.h).  Weights enter by TF variable name like `Saver.restore`.
PyTorch only owns the buffers; there is no CPU / PyTorch compute fallback.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from .stages import StageOps
from .definitions import *          # noqa: F401,F403  (AUDIO, VIDEO, FLOW, ENCODERS, NO_SEPARATION, FREQ_MASK, ...)
from .definitions import (AUDIO, VIDEO, FLOW, ENCODERS, NO_SEPARATION, FREQ_MASK, FFT_WINDOW, FFT_OVERLAP_R,
                          NUM_SEP_TRACKS_DEF, CTX_FEATS_FCUNITS_DEF, LOC_FCUNITS_DEF, SEP_FREQ_MASK_FCUNITS_DEF,
                          SEP_FFT_WINDOW_DEF)


class SptAudioGenParams:
    """reference model.py:10-21 (ctx_feats_fc_units / sep_freq_mask_fc_units are stored but unused there too)."""

    def __init__(self,
                 sep_num_tracks=NUM_SEP_TRACKS_DEF,
                 ctx_feats_fc_units=CTX_FEATS_FCUNITS_DEF,
                 loc_fc_units=LOC_FCUNITS_DEF,
                 sep_freq_mask_fc_units=SEP_FREQ_MASK_FCUNITS_DEF,
                 sep_fft_window=SEP_FFT_WINDOW_DEF):
        self.sep_num_tracks = sep_num_tracks
        self.ctx_feats_fc_units = ctx_feats_fc_units
        self.loc_fc_units = loc_fc_units
        self.sep_freq_mask_fc_units = sep_freq_mask_fc_units
        self.sep_fft_window = sep_fft_window


class SptAudioGen(StageOps):
    """reference model.py:24-434.  (The per-stage methods audio_encoder_ops / visual_encoding_ops / bottleneck_ops /
    localization_ops / separation_ops live in stages.StageOps.)  Extra keyword arguments (not in the reference): `precision`
    ('fp32' | 'bf16' | 'bf16x3': arithmetic of the dense contractions), `device`, `frame_size`."""

    def __init__(self, ambi_order,
                 audio_rate=48000,
                 video_rate=10,
                 context=1.,
                 sample_duration=0.1,
                 encoders=None,
                 separation='none',
                 params=None,
                 precision=None,
                 device=None,
                 frame_size=(224, 448)):
        assert float(audio_rate) / video_rate == int(audio_rate) // int(video_rate)          # model.py:33
        if params is None:
            params = SptAudioGenParams()
        self.ambi_order = ambi_order
        self.num_ambi_channels = sum([2 * i + 1 for i in range(ambi_order + 1)])
        self.snd_rate, self.vid_rate = audio_rate, video_rate
        self.context, self.duration = context, sample_duration
        self.snd_contx = int(context * audio_rate)
        self.snd_dur = int(sample_duration * audio_rate)
        self.snd_size = self.snd_contx + self.snd_dur - 1
        assert self.snd_rate % self.vid_rate == 0

        if encoders is None:
            encoders = [AUDIO, VIDEO, FLOW]
        assert isinstance(encoders, list)
        assert all([e in ENCODERS for e in encoders])
        self.encoders = encoders
        if separation not in (NO_SEPARATION, FREQ_MASK):
            raise ValueError('Unknown separation mode.')                                      # model.py:351
        self.separation = separation
        self.params = params

        self.model = None
        self.deploy = None
        self.solver = None
        self.ends = OrderedDict()
        self.init_ops = []
        self.loc_channels = None
        self.sep_channels = None
        self.wind_size = int(self.params.sep_fft_window * self.snd_rate)
        self.wind_size = int(2 ** np.round(np.log2(self.wind_size)))

        # ---- native handle ----
        if not torch.cuda.is_available():
            raise RuntimeError('spatialaudiogen_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if precision is None:
            precision = L.default_precision()
        self.precision = precision
        self._ctor = dict(ambi_order=ambi_order, audio_rate=audio_rate, video_rate=video_rate, context=context,
                          sample_duration=sample_duration, separation=separation, params=params, precision=precision,
                          frame_size=tuple(frame_size))
        self._opts = OrderedDict()      # sag_set_option calls so far (replayed onto the lane twins of inference_stream)
        self._twins = []
        lib = L.lib()
        cfg = L.sag_config()
        L.check(lib.sag_config_default(C.byref(cfg)))
        cfg.ambi_order, cfg.audio_rate, cfg.video_rate = int(ambi_order), int(audio_rate), int(video_rate)
        cfg.context, cfg.sample_duration = float(context), float(sample_duration)
        cfg.enc_audio, cfg.enc_video, cfg.enc_flow = int(AUDIO in encoders), int(VIDEO in encoders), int(FLOW in encoders)
        cfg.separation = L.SAG_SEP_UNET_MASK if separation == FREQ_MASK else L.SAG_SEP_NONE
        cfg.sep_num_tracks = int(params.sep_num_tracks)
        units = list(params.loc_fc_units)
        if len(units) > 4:
            raise ValueError('at most 4 localization FC layers are supported')
        cfg.n_loc_fc = len(units)
        for i, u in enumerate(units):
            cfg.loc_fc_units[i] = int(u)
        cfg.sep_fft_window = float(params.sep_fft_window)
        cfg.precision = L.PRECISIONS[precision]
        cfg.frame_h, cfg.frame_w = int(frame_size[0]), int(frame_size[1])
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(lib.sag_create(C.byref(self._h), C.byref(cfg)))
        d = L.sag_dims()
        L.check(lib.sag_get_dims(self._h, C.byref(d)))
        self.dims = d
        assert (d.snd_contx, d.snd_dur, d.snd_size, d.wind_size) == (self.snd_contx, self.snd_dur, self.snd_size, self.wind_size)
        self._frame = (int(frame_size[0]), int(frame_size[1]))
        self._ws = None
        self._ws_batch = 0
        self._weights_ready = False
        self._w = {}

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None and self._h.value:
                L.lib().sag_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- checkpoint layout ------------------------------------------------------------------------------------
    def variable_shapes(self):
        """OrderedDict TF variable name -> shape the handle expects."""
        lib = L.lib()
        out = OrderedDict()
        buf = C.create_string_buffer(256)
        shape = (C.c_int64 * 4)()
        rank = C.c_int()
        for i in range(lib.sag_num_weights_expected(self._h)):
            L.check(lib.sag_weight_name(self._h, i, buf, 256, shape, C.byref(rank)))
            out[buf.value.decode()] = tuple(int(shape[k]) for k in range(rank.value))
        return out

    def load_weights(self, arrays, strict=True):
        """tf.train.Saver().restore replacement (deploy.py:79-87, eval.py:98-118): `arrays` maps TF variable
        names to ndarrays in TF layouts.  Unknown names (optimizer slots, 'metrics/...', 'step') are ignored."""
        lib = L.lib()
        expected = self.variable_shapes()
        with torch.cuda.device(self.device):
            for name, shape in expected.items():
                if name not in arrays:
                    if strict and '/bn/moving_' not in name:
                        raise KeyError('checkpoint is missing variable %s' % name)
                    continue
                a = np.ascontiguousarray(np.asarray(arrays[name], dtype=np.float32))
                if tuple(a.shape) != tuple(shape):
                    raise ValueError('variable %s has shape %s, expected %s' % (name, a.shape, shape))
                sh = (C.c_int64 * len(shape))(*shape)
                L.check(lib.sag_load_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), sh, len(shape)))
                self._w[name] = a
            L.check(lib.sag_finalize_weights(self._h, L.stream()))
        self._weights_ready = True
        self._twins = []                # lane twins carry the old weights
        self._ws_batch = 0              # re-plan: the packed tensor-core images of reloaded layers are rebuilt in sag_workspace_bytes
        self.__dict__.pop('_wd', None)
        return self

    def set_option(self, key, value):
        for t in self._twins:
            t.set_option(key, value)
        if key != 'profile':
            self._opts[key] = value
        if key == 'precision' and isinstance(value, str):
            self.precision = value
            value = L.PRECISIONS[value]
        L.check(L.lib().sag_set_option(self._h, key.encode(), int(value)))
        if key not in ('keep_sep_channels', 'profile', 'cta_pair', 'overlap', 'fuse_gains', 'halo_conv'):     # (these do not change the workspace plan)
            self._ws_batch = 0

    # ---- forward ---------------------------------------------------------------------------------------------
    def _workspace(self, B):
        if self._ws is None or self._ws_batch != B:
            need = L.lib().sag_workspace_bytes(self._h, B)
            if need == 0:
                raise L.SagError(L.SAG_EUNSUPPORTED, L.lib().sag_last_error().decode('utf-8', 'replace'))
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_batch = B
        return self._ws

    def forward_into(self, audio, video, flow, out, flow_limits=None):
        """sag_forward on pre-staged contiguous CUDA tensors (no allocation, no copies, no sync).  audio / out are float32;
        video / flow are either the float32 frames the reference's feeder prepares, or the uint8 frames as decoded from disk
        (video: x/255 - 0.5 applied on the device, myutils.py:88-89; flow: de-quantised on the device with `flow_limits`, a
        (B, 2) float64 CUDA tensor of the frames' (min, max) rows of flow_limits.npy, feeder.py:147-161)."""
        B = audio.shape[0]
        ws = self._workspace(B)
        u8 = [t is not None and t.dtype == torch.uint8 for t in (video, flow)]
        for t in (audio, out) + tuple(x for x, q in zip((video, flow), u8) if x is not None and not q):
            if t.dtype != torch.float32:
                raise TypeError('expected float32 (or uint8 frames), got %s' % t.dtype)
        if not any(u8):
            L.check(L.lib().sag_forward(self._h, L.ptr(audio), L.ptr(video), L.ptr(flow), L.ptr(out), C.c_void_p(ws.data_ptr()),
                                        ws.numel(), B, L.stream()))
            return out
        if u8[1]:
            if flow_limits is None or flow_limits.dtype != torch.float64 or tuple(flow_limits.shape) != (B, 2):
                raise ValueError('uint8 flow frames need flow_limits: a (B, 2) float64 CUDA tensor')
        L.check(L.lib().sag_forward_frames(self._h, L.ptr(audio), L.ptr(video), int(u8[0]), L.ptr(flow), int(u8[1]),
                                           L.ptr(flow_limits) if u8[1] else None, L.ptr(out), C.c_void_p(ws.data_ptr()), ws.numel(), B,
                                           L.stream()))
        return out

    def capture_graph(self, audio, video, flow, out, flow_limits=None):
        """One forward on these (static) device buffers as a CUDA graph: at small batches the forward is a chain of ~60 short
        kernels and the host cannot enqueue them as fast as the GPU runs them (B=1: 0.33 ms of launches for 0.1 ms of work); a
        graph replays the whole chain -- programmatic dependent launches included -- with one call.  Returns the
        torch.cuda.CUDAGraph; refill the buffers, then `.replay()` on the stream that orders the refill.  The uncaptured warm-up
        call plans the batch (sag_workspace_bytes) and sets the kernel attributes, so nothing allocates during the capture."""
        with torch.cuda.device(self.device):
            self.forward_into(audio, video, flow, out, flow_limits)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward_into(audio, video, flow, out, flow_limits)
        return g

    def inference_ops(self, audio, video=None, flow=None, is_training=True, flow_limits=None):
        """reference model.py:356-434.  audio (B, snd_size, 1); video / flow (B, 1, H, W, 3); returns the
        (B, snd_dur, 3) first-order channels (Y, Z, X) as a CUDA tensor.  `is_training` is accepted for API
        compatibility: the reference never forwards it to the visual towers (they always use batch statistics,
        model.py:197) and no other layer depends on it.  video / flow may also be the uint8 frames as decoded from disk (see
        forward_into; uint8 flow with `flow_limits`)."""
        if not self._weights_ready:
            raise RuntimeError('load_weights() must be called before inference_ops()')
        with torch.cuda.device(self.device):
            audio = L.f32(audio, self.device)
            if audio.dim() != 3 or audio.shape[1] != self.snd_size or audio.shape[2] != 1:
                raise ValueError('audio must be (B, %d, 1), got %s' % (self.snd_size, tuple(audio.shape)))
            B = audio.shape[0]
            ins = {}
            for key, t in ((VIDEO, video), (FLOW, flow)):
                if key in self.encoders:
                    if t is None:
                        raise ValueError('%s input required by encoders=%s' % (key, self.encoders))
                    t = torch.as_tensor(t)
                    t = t.to(self.device).contiguous() if t.dtype == torch.uint8 else L.f32(t, self.device)
                    want = (B, 1, self.dims_frame()[0], self.dims_frame()[1], 3)
                    if tuple(t.shape) != want:
                        raise ValueError('%s must be %s, got %s' % (key, want, tuple(t.shape)))
                    ins[key] = t
            out = torch.empty((B, self.snd_dur, self.num_ambi_channels - self.ambi_order ** 2), dtype=torch.float32,
                              device=self.device)
            # this call exposes the graph's intermediate tensors (`ends`, `sep_channels`) like the reference's
            # inference_ops: the separated tracks must exist, so the inverse STFT and the mixing run as two kernels;
            # forward_into / inference_stream / W2XYZ (the deploy and eval hot loops) use the fused kernel
            self.set_option('keep_sep_channels', 1)
            try:
                if flow_limits is not None:
                    flow_limits = torch.as_tensor(flow_limits, dtype=torch.float64).to(self.device).contiguous()
                self.forward_into(audio, ins.get(VIDEO), ins.get(FLOW), out, flow_limits)
            finally:
                self.set_option('keep_sep_channels', 0)
            self._collect_ends()
        return out

    def dims_frame(self):
        return self._frame

    def _lanes(self, n):
        """[self, twin, ...]: n models with the same configuration, options and weights, each with its own native handle (side
        streams, events, stream-K flags) and workspace, so that n forwards can be in flight on n streams."""
        while len(self._twins) < n - 1:
            t = SptAudioGen(encoders=list(self.encoders), device=self.device, **self._ctor)
            t.load_weights(self._w)
            for k, v in self._opts.items():
                t.set_option(k, v)
            self._twins.append(t)
        return [self] + self._twins[:n - 1]

    def inference_stream(self, batches, depth=2, use_graph=None, lanes=3):
        """The driver loop around `sess.run` (reference deploy.py:112-148, eval.py:140-201) as a generator: `batches`
        yields dicts of HOST tensors {'audio': (B, snd_size, 1)[, 'video', 'flow': (B, 1, H, W, 3)]} (pinned memory
        makes the copies asynchronous; CUDA tensors -- e.g. frames decoded on the GPU by readers.JpegDecoder on the caller's stream --
        are taken with device copies ordered after that stream); for each one a HOST (B, snd_dur, 3) float32 tensor (pinned, reused every
        `depth` steps) is yielded in order.  'video' / 'flow' may be the uint8 frames as decoded from disk (a quarter of the
        PCIe bytes; they are prepared on the device, see forward_into) -- uint8 flow comes with 'flow_limits' (B, 2) float64.
        use_graph: replay each slot's forward as a CUDA graph (capture_graph); default: batches of at most 16 windows, where
        the host's launch rate, not the GPU, bounds the step (the reference's deploy loop feeds 10, its eval loop 16).  Host->device copies of step i+1 and the device->host copy of step i-1 run
        on their own streams while step i computes, so the PCIe transfers hide behind the forward.
        lanes: forwards in flight.  A forward is a serial chain of ~75 kernels of which the FCs, the decoder and the batch-norm
        passes of the small feature maps fill a fraction of the 148 SMs; consecutive batches are independent (deploy.py:112-148),
        so batch i+1 runs on its own stream, handle and workspace (a twin model with the same weights) and its kernels take the
        SMs batch i leaves idle.  Results are bit-identical to one lane and are yielded in order."""
        if not self._weights_ready:
            raise RuntimeError('load_weights() must be called before inference_stream()')
        dev = self.device
        lanes = max(1, int(lanes))
        depth = max(int(depth), lanes + 1)
        depth = (depth + lanes - 1) // lanes * lanes      # slot k always runs on lane k % lanes (its graph is captured there)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            models = self._lanes(lanes)
            # every lane computes on a stream of its own (one lane: the caller's stream), so that the caller's stream only orders
            compute = [main] if lanes == 1 else [torch.cuda.Stream(device=dev) for _ in range(lanes)]
            start = torch.cuda.Event()
            start.record(main)
            for cs in compute:
                if cs is not main:
                    cs.wait_event(start)                   # the lanes start after what the caller queued before this loop
            side = torch.cuda.Stream(device=dev)           # host -> device copies
            back = torch.cuda.Stream(device=dev)           # device -> host copies (own stream: a D2H waiting for its
                                                           # forward must not block the next step's H2D behind it)
            slots = []
            pending = []                                   # (slot index) in flight, oldest first

            def make_slot(b):
                def _dt(k, v):
                    v = torch.as_tensor(v)
                    return v.dtype if (k == 'flow_limits' or (k in (VIDEO, FLOW) and v.dtype == torch.uint8)) else torch.float32
                sl = {'in': {k: torch.empty(tuple(torch.as_tensor(v).shape), dtype=_dt(k, v), device=dev) for k, v in b.items()
                             if k in (AUDIO, VIDEO, FLOW, 'flow_limits')},
                      'out': torch.empty((b[AUDIO].shape[0], self.snd_dur, 3), dtype=torch.float32, device=dev),
                      'host': torch.empty((b[AUDIO].shape[0], self.snd_dur, 3), dtype=torch.float32).pin_memory(),
                      'in_ready': torch.cuda.Event(), 'done': torch.cuda.Event(), 'out_ready': torch.cuda.Event(),
                      'free': torch.cuda.Event()}
                sl['free'].record(main)
                sl['graph'] = None
                return sl

            def finish(idx):
                sl = slots[idx]
                sl['out_ready'].synchronize()
                return sl['host']

            i = 0
            try:
                for b in batches:
                    if len(slots) < depth:
                        slots.append(make_slot(b))
                    idx = i % depth
                    if len(pending) == depth:                  # the slot we are about to reuse must have been consumed
                        yield finish(pending.pop(0))
                    sl = slots[idx]
                    if sl['in'][AUDIO].shape != b[AUDIO].shape:
                        raise ValueError('all batches of a stream must have the same shape')
                    srcs = {k: torch.as_tensor(b[k]) for k in sl['in']}
                    produced = None
                    if any(t.is_cuda for t in srcs.values()):  # device-resident inputs (frames decoded on the GPU): the copies
                        produced = torch.cuda.Event()          # follow what the caller queued on its stream to produce them
                        produced.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(sl['free'])            # previous forward that read these inputs has finished
                        if produced is not None:
                            side.wait_event(produced)
                        for k, t in sl['in'].items():
                            t.copy_(srcs[k], non_blocking=True)
                            if srcs[k].is_cuda:
                                srcs[k].record_stream(side)
                        sl['in_ready'].record(side)
                    m, cs = models[idx % lanes], compute[idx % lanes]
                    with torch.cuda.stream(cs):
                        cs.wait_event(sl['in_ready'])
                        graph = use_graph if use_graph is not None else b[AUDIO].shape[0] <= 16
                        if graph and sl['graph'] is None:
                            try:                               # (warm-up forward + capture; this first use also produces the result)
                                sl['graph'] = m.capture_graph(sl['in'][AUDIO], sl['in'].get(VIDEO), sl['in'].get(FLOW), sl['out'],
                                                              sl['in'].get('flow_limits'))
                            except RuntimeError as e:          # capture unsupported: stay eager, say so once
                                import warnings
                                warnings.warn('CUDA graph capture of the forward failed (%s); running eagerly' % e)
                                sl['graph'] = False
                        if graph and sl['graph']:
                            sl['graph'].replay()
                        else:
                            m.forward_into(sl['in'][AUDIO], sl['in'].get(VIDEO), sl['in'].get(FLOW), sl['out'], sl['in'].get('flow_limits'))
                        sl['done'].record(cs)
                        sl['free'].record(cs)
                    with torch.cuda.stream(back):
                        back.wait_event(sl['done'])
                        sl['host'].copy_(sl['out'], non_blocking=True)
                        sl['out_ready'].record(back)
                    pending.append(idx)
                    i += 1
                while pending:
                    yield finish(pending.pop(0))
            finally:
                for cs in compute:                             # what the caller queues next follows every lane
                    if cs is not main:
                        ev = torch.cuda.Event()
                        ev.record(cs)
                        main.wait_event(ev)

    def _view(self, name):
        lib = L.lib()
        p = C.c_void_p()
        shape = (C.c_int64 * 5)()
        rank = C.c_int()
        ld = C.c_int64()
        L.check(lib.sag_get_tensor(self._h, name.encode(), C.byref(p), shape, C.byref(rank), C.byref(ld)))
        shp = [int(shape[k]) for k in range(rank.value)]
        fmt, plane = C.c_int(), C.c_int64()
        L.check(lib.sag_get_tensor_format(self._h, name.encode(), C.byref(fmt), C.byref(plane)))
        off = p.value - self._ws.data_ptr()
        strides = [1] * len(shp)
        if len(shp) >= 2:
            strides[-2] = int(ld.value)
            for k in range(len(shp) - 3, -1, -1):
                strides[k] = strides[k + 1] * shp[k + 1]
        if fmt.value == 0:
            assert off >= 0 and off % 4 == 0
            return self._ws[off:].view(torch.float32).as_strided(shp, strides)
        # tensor-core path: the value is hi + lo of two bf16 planes (a copy, not a view)
        assert off >= 0 and off % 2 == 0
        span = sum((n - 1) * s for n, s in zip(shp, strides)) + 1
        out = self._ws[off:off + 2 * span].view(torch.bfloat16).as_strided(shp, strides).float()
        if plane.value:
            o2 = off + plane.value
            out = out + self._ws[o2:o2 + 2 * span].view(torch.bfloat16).as_strided(shp, strides).float()
        return out

    def _collect_ends(self):
        """Views (aliasing the workspace, valid until the next forward) of the taps the reference keeps in
        `self.ends`, `self.sep_channels`, `self.loc_channels` (model.py:371-432)."""
        lib = L.lib()
        buf = C.create_string_buffer(256)
        self.ends = OrderedDict()
        for i in range(lib.sag_num_tensors(self._h)):
            L.check(lib.sag_tensor_name(self._h, i, buf, 256))
            name = buf.value.decode()
            self.ends[name] = self._view(name)
        if 'stft' in self.ends:
            self.ends['stft'] = torch.view_as_complex(self.ends['stft'])
        self.sep_channels = self.ends.get('separation/all_channels')
        loc = self.ends.get('localization')
        if loc is not None:
            self.loc_channels = [loc[..., :-1], loc[..., -1]]       # per segment; the reference tiles x1600 (model.py:262)

    # ---- evaluation (model.py:110-154) ----------------------------------------------------------------------------
    def evaluation_ops(self, preds_t, targets_t, w_t, mask_channels=None):
        """reference model.py:110-154 -> (metrics, stft_dist_ps, lsd_ps, mse_ps, snr_ps); tensors (B,3), channel
        order (Y,Z,X).  Like the reference, the channel mask is mandatory (model.py:114 dereferences it first)."""
        if mask_channels is None:
            raise ValueError('mask_channels is required (reference model.py:114-118)')
        from . import metrics as M
        preds = L.f32(preds_t, self.device)
        targets = L.f32(targets_t, self.device)
        mask = L.f32(mask_channels, self.device)
        if mask.dim() == 1:
            mask = mask[None].expand(preds.shape[0], -1)
        if mask.shape[1] == self.num_ambi_channels:
            mask = mask[:, self.ambi_order ** 2:]
        res = M.window_metrics(preds, targets, self.snd_rate)
        stft_ps, lsd_ps, mse_ps, snr_ps = res['stft'], res['lsd'], res['mse'], res['snr']
        self.last_eval = res
        num_masked = mask.sum(0).clamp(min=1)                                              # model.py:115-116
        metrics = OrderedDict()
        for key, ps, scale in (('stft', stft_ps, 100.), ('lsd', lsd_ps, 1.), ('mse', mse_ps, 5e3), ('snr', snr_ps, 1.)):
            v = (ps * mask).sum(0) / num_masked * scale
            metrics[key + '/avg'] = v.mean()
            for i, ch in zip(range(3), 'YZX'):
                metrics[key + '/' + ch] = v[i]
        metrics['pow/pred'] = (preds ** 2).mean(2).mean(0).sum()
        metrics['pow/gt'] = (targets ** 2).mean(2).mean(0).sum()
        return metrics, stft_ps, lsd_ps, mse_ps, snr_ps

    def loss_ops(self, metrics_t, step_t=None):
        """reference model.py:156-159: the training losses, keyed like the reference's OrderedDict -- the STFT distance
        (training itself is out of scope; this only keeps code that reads `losses['stft/mse']` working)."""
        losses = OrderedDict()
        losses['stft/mse'] = metrics_t['stft/avg']
        return losses
