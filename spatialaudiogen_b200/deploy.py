"""W2XYZ: the reference's deploy driver (reference deploy.py:41-152) on top of libsag.so.

Same construction (model_dir with train-params.txt + weights) and the same hot loop: consecutive 0.1 s windows in
batches of 10, the last batch zero-padded (deploy.py:124-139 -- it matters: the visual towers use batch statistics),
W taken from the input crop, rows [W, Y, Z, X] appended window after window (deploy.py:143-151).  The reference reads
windows from disk through feeder.SampleReader; here the windows are handed in as arrays (`deploy_windows`), the
on-disk reader is a "next" row.  `sess.run` becomes one sag_forward per batch on the caller's stream,
with the host<->device copies on pinned buffers.
"""
import numpy as np
import torch

from . import myutils, readers, tf_checkpoint
from .definitions import AUDIO, VIDEO, FLOW, NO_SEPARATION
from .model import SptAudioGen, SptAudioGenParams


def load_weights_file(path):
    """name -> ndarray dict from an .npz written with np.savez (TF variable names, TF layouts: SURVEY.md App. B)."""
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


class W2XYZ(object):
    def __init__(self, model_dir=None, params=None, weights=None, precision=None, device=None):
        """model_dir: directory holding train-params.txt (reference myutils.load_params) and the weights -- a TensorFlow
        V2 checkpoint as the reference saves it (`checkpoint` + `model.ckpt-N.index/.data-*`, restored like
        deploy.py:79-87 through tf_checkpoint.read_bundle) or weights.npz; or pass `params` (object with the fields
        load_params returns) and `weights` (dict) directly."""
        if params is None:
            params = myutils.load_params(model_dir)
        self.params = params
        self.duration = 0.1                                                        # deploy.py:49
        self.batch_size = 10                                                       # deploy.py:50
        num_sep = params.num_sep_tracks if params.separation != NO_SEPARATION else 1          # deploy.py:54
        net_params = SptAudioGenParams(sep_num_tracks=num_sep, ctx_feats_fc_units=params.context_units,
                                       loc_fc_units=params.loc_units, sep_freq_mask_fc_units=params.freq_mask_units,
                                       sep_fft_window=params.fft_window)
        self.model = SptAudioGen(ambi_order=params.ambi_order, audio_rate=params.audio_rate, video_rate=params.video_rate,
                                 context=params.context, sample_duration=self.duration, encoders=list(params.encoders),
                                 separation=params.separation, params=net_params, precision=precision, device=device)
        self.audio_size = self.model.snd_dur + self.model.snd_contx - 1
        self.video_size = int(self.duration * params.video_rate)
        if weights is None:
            weights = tf_checkpoint.load_model_dir(model_dir, names=set(self.model.variable_shapes()))
        self.model.load_weights(weights)
        B, dev = self.batch_size, self.model.device
        H, W = self.model.dims_frame()
        # pinned staging + device buffers, allocated once (the reference's placeholders, deploy.py:69-76); visual frames have a
        # float32 set (prepared frames) and a uint8 set (frames as decoded from disk, prepared by the ingest kernel)
        self._h = {AUDIO: torch.zeros((B, self.audio_size, 1), dtype=torch.float32).pin_memory()}
        for k in (VIDEO, FLOW):
            if k in params.encoders:
                self._h[k] = torch.zeros((B, self.video_size, H, W, 3), dtype=torch.float32).pin_memory()
                self._h[k + '_u8'] = torch.zeros((B, self.video_size, H, W, 3), dtype=torch.uint8).pin_memory()
        if FLOW in params.encoders:
            self._h['flow_limits'] = torch.zeros((B, 2), dtype=torch.float64).pin_memory()
        self._d = {k: torch.empty_like(v, device=dev) for k, v in self._h.items()}
        self._out = torch.empty((B, self.model.snd_dur, 3), dtype=torch.float32, device=dev)
        self._out_h = torch.empty((B, self.model.snd_dur, 3), dtype=torch.float32).pin_memory()

    def run_batch(self, audio, video=None, flow=None, flow_limits=None):
        """One `sess.run(ambi_pred_t, feed_dict)` (deploy.py:141): host arrays of n <= batch_size windows in, host
        (n, snd_dur, 3) predictions out.  The batch is zero-padded to batch_size like the reference.  video / flow: prepared
        float32 frames, or the uint8 frames as decoded from disk (flow then with `flow_limits` (n, 2), the frames' rows of
        flow_limits.npy); uint8 frames of a FULL batch are uploaded as they are and prepared on the device, a short batch is
        prepared on the host first, because the reference pads with zeros AFTER its preparation (deploy.py:127-139)."""
        n = audio.shape[0]
        srcs = {AUDIO: audio, VIDEO: video, FLOW: flow}
        use = {}
        with torch.cuda.device(self.model.device):
            for k in (AUDIO, VIDEO, FLOW):
                if k not in self._h:
                    continue
                if srcs[k] is None:
                    raise ValueError('%s windows required by encoders=%s' % (k, self.params.encoders))
                x = np.asarray(srcs[k])
                if k != AUDIO and x.dtype == np.uint8:
                    if k == FLOW and flow_limits is None:
                        raise ValueError('uint8 flow frames need flow_limits')
                    if n == self.batch_size:
                        self._h[k + '_u8'].copy_(torch.from_numpy(np.ascontiguousarray(x)))
                        self._d[k + '_u8'].copy_(self._h[k + '_u8'], non_blocking=True)
                        use[k] = self._d[k + '_u8']
                        if k == FLOW:
                            self._h['flow_limits'].copy_(torch.as_tensor(np.asarray(flow_limits, np.float64)))
                            self._d['flow_limits'].copy_(self._h['flow_limits'], non_blocking=True)
                        continue
                    x = myutils.img_prep_fcn()(x) if k == VIDEO else readers.dequantize_flow(x, flow_limits)
                h = self._h[k]
                h[:n].copy_(torch.as_tensor(np.asarray(x, dtype=np.float32)))
                if n != self.batch_size:
                    h[n:].zero_()
                self._d[k].copy_(h, non_blocking=True)
                use[k] = self._d[k]
            lims = self._d['flow_limits'] if (FLOW in use and use[FLOW].dtype == torch.uint8) else None
            self.model.forward_into(use[AUDIO], use.get(VIDEO), use.get(FLOW), self._out, lims)
            self._out_h.copy_(self._out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self._out_h[:n].numpy().copy()

    def deploy_stream(self, windows, lanes=3):
        """The loop of deploy.py:112-151 over an ITERATOR of windows -- dicts {'ambix': (audio_size, C>=1) with W in channel 0
        [, 'video', 'flow': (video_size, H, W, 3) float32 or uint8, 'flow_limits': (video_size, 2)]} -- consumed batch_size at a
        time like the reference does, so host memory holds a few batches of inputs plus the growing output.  Returns
        (N*snd_dur, 4) float64 rows [W, Y, Z, X].  Full batches go through SptAudioGen.inference_stream (copies on their own
        streams, `lanes` forwards in flight, the batch of 10 replayed as a CUDA graph; same bits as one run_batch per batch);
        the short last batch, which the reference zero-pads AFTER preparing its frames, goes through run_batch.  Windows whose
        'video' / 'flow' are lists of jpg FILES (bytes: SampleReader(jpeg_files=True)) are decoded on the GPU, a batch at a time
        (readers.JpegDecoder, bit-identical to the PIL decode)."""
        ss = self.model.snd_contx // 2
        mono, pred = [], []
        enc = self.params.encoders
        tail = []

        decoders = {}

        def frames(batch, k):
            if k not in enc:
                return None
            if not isinstance(batch[0][k], (list, tuple)):
                return np.stack([c[k] for c in batch], 0)
            files = [f for c in batch for f in c[k]]            # undecoded jpg files: a whole batch goes to the GPU
            if k not in decoders:
                h, w = readers.jpeg_info(files[0])[:2]
                decoders[k] = readers.JpegDecoder(self.batch_size * len(batch[0][k]), h, w, device=self.model.device)
            with torch.cuda.device(self.model.device):
                x = decoders[k].decode(files)
            return x.view(len(batch), len(batch[0][k]), x.shape[1], x.shape[2], 3)

        def pack(batch):
            a = np.stack([np.asarray(c['ambix'], np.float64) for c in batch], 0)
            v, f = frames(batch, VIDEO), frames(batch, FLOW)
            fl = np.stack([np.asarray(c['flow_limits']).reshape(-1, 2)[0] for c in batch], 0) if (f is not None and 'flow_limits' in batch[0]) else None
            mono.append(np.copy(a[:, ss:ss + self.model.snd_dur, :1]).reshape(-1, 1))
            return a, v, f, fl

        def full_batches():
            batch = []
            for c in windows:
                batch.append(c)
                if len(batch) == self.batch_size:
                    a, v, f, fl = pack(batch)
                    batch = []
                    b = {AUDIO: torch.from_numpy(np.ascontiguousarray(a[:, :, :1], dtype=np.float32))}
                    for k, x in ((VIDEO, v), (FLOW, f)):
                        if x is None:
                            continue
                        if k == FLOW and x.dtype in (np.uint8, torch.uint8) and fl is None:
                            raise ValueError('uint8 flow frames need flow_limits')
                        if isinstance(x, torch.Tensor):         # decoded on the GPU
                            b[k] = x
                            continue
                        if x.dtype != np.uint8:
                            x = np.asarray(x, dtype=np.float32)
                        b[k] = torch.from_numpy(np.ascontiguousarray(x))
                    if FLOW in b and b[FLOW].dtype == torch.uint8:
                        b['flow_limits'] = torch.from_numpy(np.ascontiguousarray(fl, dtype=np.float64))
                    yield b
            tail.extend(batch)

        for y in self.model.inference_stream(full_batches(), lanes=lanes):
            pred.append(y.numpy().reshape(-1, y.shape[2]).copy())
        if tail:
            a, v, f, fl = pack(tail)
            v, f = [x.cpu().numpy() if isinstance(x, torch.Tensor) else x for x in (v, f)]
            out = self.run_batch(a[:, :, :1], v, f, fl)
            pred.append(out.reshape(-1, out.shape[2]))
        if not pred:
            return np.zeros((0, 4))
        return np.concatenate((np.concatenate(mono, 0), np.concatenate(pred, 0)), 1)   # float64, like numpy promotes in deploy.py:151

    def deploy_windows(self, ambix_windows, video_windows=None, flow_windows=None):
        """deploy_stream over arrays: ambix_windows (N, audio_size, C>=1); video / flow (N, video_size, H, W, 3)."""
        def it():
            for i in range(len(ambix_windows)):
                c = {'ambix': ambix_windows[i]}
                if video_windows is not None:
                    c['video'] = video_windows[i]
                if flow_windows is not None:
                    c['flow'] = flow_windows[i]
                yield c
        return self.deploy_stream(it())

    def deploy(self, input_folder, deploy_start, deploy_duration, gpu_jpeg=True):
        """deploy.py:90-152: read the windows of `input_folder` (the per-video folder layout of readers.SampleReader)
        scheduled by its audio_pow.lst from `deploy_start` for `deploy_duration` seconds -- the schedule is shifted so
        that the first window sits exactly at deploy_start (deploy.py:108-109) -- and generate their ambisonics, batch by
        batch (the reader is consumed lazily; gpu_jpeg: the frames' jpg files are decoded on the GPU -- baseline files only --
        else by PIL and travel as uint8; either way they are prepared on the device).
        Returns (N*snd_dur, 4) float64 rows [W, Y, Z, X]."""
        p = self.params
        reader = readers.SampleReader(input_folder, ambi_order=p.ambi_order, audio_rate=p.audio_rate, video_rate=p.video_rate,
                                      context=p.context, duration=self.duration, return_video=VIDEO in p.encoders,
                                      img_prep=None, return_flow=FLOW in p.encoders, start_time=deploy_start,
                                      sample_duration=deploy_duration, skip_silence_thr=None, shuffle=False,
                                      random_rotations=False, skip_rate=None, raw_flow=True, jpeg_files=gpu_jpeg)
        if not reader.chunks_t:
            raise ValueError('%s has no windows in [%s, %s)' % (input_folder, deploy_start, deploy_start + deploy_duration))
        dt = reader.chunks_t[0] - deploy_start
        reader.chunks_t = [t - dt for t in reader.chunks_t]
        return self.deploy_stream(reader.loop_chunks())


def parse_arguments(argv=None):
    """deploy.py:14-38 -- the same positional and optional arguments (the video / overlay outputs need ffmpeg and the
    spatial-media injector, which stay external: they are accepted and reported as skipped)."""
    import argparse
    import sys
    parser = argparse.ArgumentParser(description='Generate first-order ambisonics for a clip (reference deploy.py).')
    parser.add_argument('model_dir', help='Directory containing model snapshot.')
    parser.add_argument('input_folder', default='', help='Folder with input sample.')
    parser.add_argument('video', default='', nargs='?', help='High resolution video (only used by the video outputs).')
    parser.add_argument('--deploy_start', default=0., type=float)
    parser.add_argument('--deploy_duration', default=10., type=float)
    parser.add_argument('--output_fn', default='output', help='Basename for output files.')
    parser.add_argument('--save_ambix', action='store_true', help='Output ambix video file.')
    parser.add_argument('--save_video', action='store_true', help='Output video file.')
    parser.add_argument('--overlay_map', action='store_true', help='Overlay spherical map.')
    parser.add_argument('--VR', action='store_true', help='360 video.')
    parser.add_argument('--gpu', type=int, default=0, help='GPU id')
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    if args.deploy_duration <= 0:
        args.deploy_duration = None                                                # deploy.py:35-36: the whole clip
    return args


def main(args):
    """deploy.py:155-215 up to the ambisonics: model restored from `model_dir`, windows of `input_folder` generated, the
    (N, 4) [W, Y, Z, X] track written as `<output_fn>.wav` and its stereo down-mix as `<output_fn>-stereo.wav`."""
    import sys
    with torch.cuda.device(args.gpu):
        model = W2XYZ(args.model_dir)
        ambi_pred = model.deploy(args.input_folder, args.deploy_start, args.deploy_duration)
    readers.save_wav(args.output_fn + '.wav', ambi_pred, model.params.audio_rate)
    readers.save_wav(args.output_fn + '-stereo.wav', myutils.ambix_to_stereo(ambi_pred), model.params.audio_rate)
    if args.save_ambix or args.save_video or args.overlay_map:
        sys.stderr.write('video outputs (ffmpeg mux, spatial-media metadata, heat-map overlay) are not produced by this package: '
                         'see myutils.energy_map_frames for the overlay frames\n')
    return ambi_pred


if __name__ == '__main__':                       # python -m spatialaudiogen_b200.deploy MODEL_DIR INPUT_FOLDER [VIDEO] ...
    main(parse_arguments())
