"""On-disk sample readers: the host side of the reference's feeder.py:50-278.

Layout written by the reference's preprocessing (scraping/preprocess.py:98-204), one folder per video:

    <folder>/ambix/%06d.wav          1 s of first-order ambisonics each, channels (W, Y, Z, X) at 48 kHz
    <folder>/video/%06d.jpg          equirectangular frames, 224 x 448, 10 per second
    <folder>/flow/%06d.jpg           optical flow quantised to 8 bit: channel 0 = angle, 2 = magnitude
    <folder>/flow/flow_limits.npy    (n_frames, 2): per-frame (min, max) of the flow magnitude
    <folder>/audio_pow.lst           "<chunk centre time> <power>" per line (the eval schedule: t = 0.5 + k)

`SampleReader.get()` returns the same dict as the reference ({'id', 'ambix', 'video', 'flow'}) with the same edge
behaviour: audio chunks are zero-padded before the start / after the end of the clip (feeder.py:66-90), video chunks
start at max(int(t * rate), 0), flow is de-quantised to (mag cos, mag sin, mag) (feeder.py:147-161).  Decoding runs on
the host like the reference's feeder threads (scipy wav reader, PIL JPEG decoder); the arrays it yields are what
`W2XYZ.deploy` / `evaluate.evaluate_batches` upload.  With `jpeg_files=True` the visual readers hand out the jpg FILES instead
(bytes, undecoded) and `JpegDecoder` decodes a whole batch of them on the GPU -- Huffman decoding by a pool of host threads
in libsag.so, inverse DCT / chroma upsampling / colour conversion as CUDA kernels -- into the uint8 frames the ingest kernel
takes, bit-identical to PIL's decode (tests/test_jpeg.py).  Files at another sample rate are rejected instead of resampled
(the reference resamples with resampy 'kaiser_fast', which is not available; its own preprocessing already writes
the model's rate).
"""
import os
import random

import numpy as np


def load_wav(fname, rate=None):
    """pyutils/iolib/audio.py:11-27: (n, channels) float64 in [-1, 1) and the rate."""
    from scipy.io import wavfile
    _rate, data = wavfile.read(fname)
    if data.ndim == 1:
        data = data.reshape((-1, 1))
    if data.dtype.kind == 'i':                                  # libsndfile's float conversion: / 2^(bits-1)
        data = data.astype(np.float64) / float(1 << (8 * data.dtype.itemsize - 1))
    elif data.dtype.kind == 'u':                                # 8-bit PCM is unsigned
        data = (data.astype(np.float64) - 128.0) / 128.0
    else:
        data = data.astype(np.float64)
    if rate is not None and int(rate) != int(_rate):
        raise ValueError('%s is sampled at %d Hz, the model reads %d Hz (resampling is not built)' % (fname, _rate, rate))
    return data, float(_rate)


def save_wav(fname, signal, rate):
    """pyutils/iolib/audio.py:30-33 (16-bit PCM, what libsndfile's default 'wav' format writes)."""
    from scipy.io import wavfile
    x = np.clip(np.asarray(signal, np.float64), -1.0, 1.0 - 1.0 / 32768)
    wavfile.write(fname, int(rate), np.round(x * 32768.0).astype(np.int16))


def _imread(fn):
    from PIL import Image
    with Image.open(fn) as im:
        return np.asarray(im.convert('RGB'))


def _read_file(fn):
    with open(fn, 'rb') as f:
        return f.read()


def jpeg_info(data):
    """(height, width, components, h_samp, v_samp) of a jpg file held in `data` (bytes), read by libsag's marker parser."""
    import ctypes as C
    from . import _lib as L
    v = [C.c_int() for _ in range(5)]
    L.check(L.lib().sag_jpeg_info(data, len(data), *[C.byref(x) for x in v]))
    w, h, nc, hs, vs = [x.value for x in v]
    return h, w, nc, hs, vs


class JpegDecoder(object):
    """Batches of jpg files -> uint8 RGB frames on the GPU: what `scipy.misc.imread` does per frame in the reference's feeder
    (feeder.py:120-127), bit-identical to PIL / libjpeg (islow inverse DCT, fancy upsampling).  libsag.so decodes the entropy-coded
    segments with a pool of host threads and runs the rest as CUDA kernels on the current stream (include/sag.h sag_jpeg_*).
    Baseline sequential YCbCr / grey files only -- what the reference's preprocessing writes (skimage.io.imsave = PIL's encoder with its
    defaults: baseline, 4:2:0, quality 75; scraping/preprocess.py:141-143, 198); others raise."""

    def __init__(self, max_frames, height, width, device=None, threads=0, device_huffman=True):
        """device_huffman: decode the entropy-coded segments on the GPU too (parallel, self-synchronising subsequences: only the
        compressed bytes cross PCIe); False: on `threads` host threads (0: one per core), the coefficients cross PCIe."""
        import ctypes as C
        import torch
        from . import _lib as L
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_frames, self.height, self.width, self.threads = int(max_frames), int(height), int(width), int(threads)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().sag_jpeg_create(C.byref(self._h), self.max_frames, self.height, self.width))
        self.set_option('device_huffman', int(bool(device_huffman)))

    def set_option(self, key, value):
        from . import _lib as L
        L.check(L.lib().sag_jpeg_set_option(self._h, key.encode(), int(value)))

    def sync_rounds(self, n):
        """Synchronisation rounds the device entropy decoder needed for each of the first n frames of the last decode."""
        import ctypes as C
        from . import _lib as L
        r = (C.c_int * n)()
        L.check(L.lib().sag_jpeg_sync_rounds(self._h, r, n))
        return list(r)

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None and self._h.value:
                from . import _lib as L
                L.lib().sag_jpeg_destroy(self._h)
                self._h.value = None
        except Exception:
            pass

    def decode(self, files, out=None):
        """files: a list of at most max_frames `bytes` objects, each one jpg file.  Returns (len(files), height, width, 3) uint8
        on the device (`out` if given), valid in stream order on the current stream."""
        import ctypes as C
        import torch
        from . import _lib as L
        n = len(files)
        if out is None:
            out = torch.empty((n, self.height, self.width, 3), dtype=torch.uint8, device=self.device)
        if out.dtype != torch.uint8 or tuple(out.shape) != (n, self.height, self.width, 3) or not out.is_contiguous():
            raise ValueError('out must be a contiguous uint8 (%d, %d, %d, 3) CUDA tensor' % (n, self.height, self.width))
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(f), C.c_void_p) for f in files])
        sizes = (C.c_size_t * n)(*[len(f) for f in files])
        with torch.cuda.device(self.device):
            L.check(L.lib().sag_jpeg_decode(self._h, ptrs, sizes, n, L.ptr(out), self.threads, L.stream()))
        return out


class AudioReader(object):
    """The `ambix/` folder of a clip -- one wav file per second -- as one zero-padded timeline (reference feeder.py:50-105).

    `get(t, n)` returns the n samples that start at time t; samples before the clip or past its last file are zeros.  Like the
    reference, positions are truncated, not rounded: the first sample is int(t * rate) of the clip, and the offset inside the
    file that holds it is int((t - second) * rate)."""

    def __init__(self, audio_folder, rate=None, ambi_order=1):
        self.audio_folder = audio_folder
        names = sorted(f for f in os.listdir(audio_folder) if f.endswith('.wav'))
        self.num_files = len(names)
        probe, file_rate = load_wav(os.path.join(audio_folder, names[0]))
        self.rate = float(file_rate) if rate is None else rate
        self.num_channels = min(probe.shape[1], (ambi_order + 1) ** 2)
        self.duration = self.num_files
        self.num_frames = int(self.duration * self.rate)

    def _file(self, second):
        return load_wav('{}/{:06d}.wav'.format(self.audio_folder, second), self.rate)[0]

    def _copy_span(self, out, dst, t0, count):
        """`count` samples starting at time t0 >= 0 (inside the clip) into out[dst:], file by file."""
        second = int(t0)
        skip = int((t0 - second) * self.rate)                     # offset inside the first file
        last = min(int(np.ceil(t0 + count / float(self.rate))), self.num_files)
        while count > 0 and second < last:
            data = self._file(second)[skip:skip + count, :self.num_channels]
            out[dst:dst + data.shape[0]] = data
            dst += data.shape[0]
            count -= data.shape[0]
            second, skip = second + 1, 0

    def get(self, start_time, size, rotation=None):
        out = np.zeros((size, self.num_channels))
        first = int(start_time * self.rate)                        # (truncation toward zero, also for negative times)
        lead = max(-first, 0)                                      # zeros before the start of the clip
        t0, first = (0., 0) if lead > 0 else (start_time, first)
        count = size - lead
        count -= max(first + count - self.num_frames, 0)           # zeros past the end of the clip
        if count > 0:
            self._copy_span(out, lead, t0, count)
        if rotation is not None:                                   # yaw of the sound field about the vertical axis (feeder.py:92-102)
            assert -np.pi <= rotation < np.pi
            c, s = np.cos(rotation), np.sin(rotation)
            out = np.dot(out, np.array([[1, 0, 0, 0], [0, c, 0, s], [0, 0, 1, 0], [0, -s, 0, c]]).T)      # rows act on (W, Y, Z, X)
        return out


class VideoReader(object):
    """The `video/` (or `flow/`) folder of a clip: numbered jpg frames at 10 per second (reference feeder.py:108-135).
    Without `img_prep` the frames come back as decoded -- uint8 -- which is what the device-side ingest takes
    (SptAudioGen.forward_into prepares them in the frame-ingest kernel)."""
    RAW_RATE = 10.

    def __init__(self, video_folder, rate=None, img_prep=None, jpeg_files=False):
        self.video_folder = video_folder
        self.rate = self.RAW_RATE if rate is None else rate
        self.img_prep = img_prep if img_prep is not None else (lambda x: x)
        self.jpeg_files = jpeg_files                              # hand out the files (bytes) for JpegDecoder instead of frames
        names = sorted(f for f in os.listdir(video_folder) if f.endswith('.jpg'))
        self.num_frames = len(names)
        self.duration = self.num_frames / self.RAW_RATE
        if jpeg_files:
            h, w = jpeg_info(_read_file(os.path.join(video_folder, names[0])))[:2]
            self.frame_shape = (h, w, 3)
        else:
            self.frame_shape = self.img_prep(_imread(os.path.join(video_folder, names[0]))).shape

    def frame(self, index):
        return self.img_prep(_imread(os.path.join(self.video_folder, '{:06d}.jpg'.format(index))))

    def get_by_index(self, start_time, size, rotation=None):
        first = max(int(start_time * self.rate), 0)
        if self.jpeg_files:
            if rotation is not None:
                raise ValueError('jpeg_files readers do not rotate (roll the decoded frames on the device instead)')
            return [_read_file(os.path.join(self.video_folder, '{:06d}.jpg'.format(first + k))) for k in range(size)]
        chunk = np.stack([self.frame(first + k) for k in range(size)], 0)
        if rotation is not None:                                   # the same yaw as the audio: a roll along the panorama's width
            chunk = np.roll(chunk, -int(rotation / (2. * np.pi) * self.frame_shape[1]), axis=2)
        return chunk


def dequantize_flow(chunk_u8, limits):
    """feeder.py:147-161 on (..., H, W, 3) uint8 flow frames (channel 0 = angle, 2 = magnitude) with one (min, max) row of
    flow_limits.npy per frame -> float32 (mag cos, mag sin, mag).  (The device does the same inside the frame-ingest kernel.)"""
    chunk = np.asarray(chunk_u8).astype(np.float32)
    shape = chunk.shape
    chunk = chunk.reshape((-1,) + shape[-3:])
    limits = np.asarray(limits).reshape(-1, 2)
    lo, hi = limits[:, 0].reshape((-1, 1, 1)), limits[:, 1].reshape((-1, 1, 1))
    mag, ang = chunk[:, :, :, 2], chunk[:, :, :, 0]                # (views: the in-place steps below round like the reference's)
    mag *= (hi - lo) / 255.
    mag += lo
    ang *= (2 * np.pi) / 255.
    chunk[:, :, :, 1] = mag * np.sin(ang)
    chunk[:, :, :, 0] = mag * np.cos(ang)
    return chunk.reshape(shape)


class FlowReader(object):
    """Optical flow stored as 8-bit frames + per-frame magnitude limits (reference feeder.py:138-161).  raw=True returns the
    quantised frames and their limits (what the device-side ingest takes) instead of de-quantising on the host."""

    def __init__(self, flow_dir, flow_lims_fn, rate=None, flow_prep=None, raw=False, jpeg_files=False):
        self.reader = VideoReader(flow_dir, rate=rate, jpeg_files=jpeg_files)
        self.lims = np.load(flow_lims_fn)
        self.rate = self.reader.rate
        self.duration = self.reader.duration
        self.flow_prep = flow_prep if flow_prep is not None else (lambda x: x)
        self.raw = raw

    def limits(self, start_time, size):
        first = max(int(start_time * self.rate), 0)
        return self.lims[first:first + size]

    def get_by_index(self, start_time, size, rotation=None):
        chunk = self.reader.get_by_index(start_time, size, rotation)
        lims = self.limits(start_time, len(chunk))
        return (chunk, lims) if getattr(self, 'raw', False) else dequantize_flow(chunk, lims)


def sample_folders(directory, subset_fn=None):
    """feeder.py:12-47 (FilenameProvider, one epoch, no shuffling): the per-video folders of `directory` in os.listdir order,
    restricted to the ids listed one per line in `subset_fn` (meta/subsets/*.lst)."""
    sample_ids = os.listdir(directory)
    if len(sample_ids) == 0:
        raise ValueError('Dataset directory is empty.')
    if subset_fn is not None:
        if not os.path.exists(subset_fn):
            raise IOError('subset file %s does not exist' % subset_fn)
        subset = set(open(subset_fn).read().splitlines())
        sample_ids = [y for y in sample_ids if y in subset]
    return [os.path.join(directory, y) for y in sample_ids]


def load_channel_masks(audio_layouts_fn):
    """feeder.py:312-314: meta/audio_layouts.txt (`<video id> <WXYZ|WXY>` per line) -> {video id: (4,) mask over [W, Y, Z, X]};
    a WXY recording has no height channel (mask [1, 1, 0, 1])."""
    masks = {'WXYZ': np.array([1., 1., 1., 1.]), 'WXY': np.array([1., 1., 0., 1.])}
    out = {}
    for line in open(audio_layouts_fn).read().splitlines():
        if line.strip():
            vid, layout = line.split()[:2]
            out[vid] = masks[layout]
    return out


def chunk_schedule(times, powers, skip_rate=None, skip_silence_thr=None, start_time=0.5, sample_duration=None, num_threads=1,
                   thread_id=0):
    """Which entries of a clip's audio_pow.lst a reader visits (reference feeder.py:208-236), as array filters applied in the
    reference's order: every skip_rate-th entry, then entries louder than the silence threshold, then the [start_time,
    start_time + sample_duration) window (a start of 0.5 or less keeps everything), then this thread's contiguous slice."""
    t, p = np.asarray(times, np.float64), np.asarray(powers, np.float64)
    if skip_rate is not None:
        t, p = t[::skip_rate], p[::skip_rate]
    if skip_silence_thr is not None:
        t = t[p > skip_silence_thr]
    if start_time > 0.5:
        t = t[t >= start_time]
    if sample_duration is not None:
        t = t[t < start_time + sample_duration]
    if num_threads > 1:
        cut = np.linspace(0, len(t), num_threads + 1).astype(int)
        t = t[cut[thread_id]:cut[thread_id + 1]]
    return [float(x) for x in t]


class SampleReader(object):
    """One clip's windows in schedule order (reference feeder.py:164-278): `get()` returns {'id', 'ambix'[, 'video'][, 'flow']}
    for the next scheduled time, None at the end; `loop_chunks(n)` iterates.  The audio window starts context/2 before the
    scheduled time; all readers of a window share one random yaw when random_rotations is on.  raw_flow=True yields the
    quantised flow frames plus 'flow_limits' instead of float frames; jpeg_files=True yields the undecoded jpg files (lists of
    bytes) under 'video' / 'flow' for JpegDecoder."""

    def __init__(self, folder, ambi_order=1, audio_rate=48000, video_rate=10, context=1.0, duration=0.1, return_video=True,
                 img_prep=None, return_flow=False, flow_prep=None, skip_silence_thr=None, shuffle=True, start_time=0.5,
                 sample_duration=None, skip_rate=None, random_rotations=True, num_threads=1, thread_id=0, raw_flow=False,
                 jpeg_files=False):
        n_audio, n_video, n_context = duration * audio_rate, duration * video_rate, context * audio_rate
        assert float(audio_rate) / video_rate == int(float(audio_rate) / video_rate)
        for v in (n_audio, n_video, n_context):
            assert abs(v - round(v)) < 1e-6                       # whole numbers of samples / frames
        self.folder, self.video_id = folder, os.path.split(folder)[-1]
        self.duration, self.context = duration, context
        self.audio_rate, self.video_rate = audio_rate, video_rate
        self.return_video, self.return_flow, self.random_rotations, self.raw_flow = return_video, return_flow, random_rotations, raw_flow
        self.audio_reader = AudioReader(os.path.join(folder, 'ambix'), audio_rate, ambi_order)
        if jpeg_files and (random_rotations or (return_flow and not raw_flow)):
            raise ValueError('jpeg_files needs random_rotations=False and raw_flow=True (undecoded files cannot be rolled or de-quantised on the host)')
        self.video_reader = VideoReader(os.path.join(folder, 'video'), video_rate, img_prep, jpeg_files=jpeg_files) if return_video else None
        if return_flow:
            flow_dir = os.path.join(folder, 'flow')
            self.flow_reader = FlowReader(flow_dir, os.path.join(flow_dir, 'flow_limits.npy'), video_rate, flow_prep, jpeg_files=jpeg_files)
            if raw_flow:
                self.flow_reader.raw = True
        self.audio_size = int(round(n_audio)) + int(round(n_context)) - 1
        self.video_size = int(round(n_video))
        self.video_shape = self.video_reader.frame_shape if return_video else None
        table = np.loadtxt(os.path.join(folder, 'audio_pow.lst'), ndmin=2)
        self.chunks_t = chunk_schedule(table[:, 0], table[:, 1], skip_rate, skip_silence_thr, start_time, sample_duration,
                                       num_threads, thread_id)
        if shuffle:
            random.shuffle(self.chunks_t)
        self.head = -1

    def get(self):
        self.head += 1
        if self.head >= len(self.chunks_t):
            return None
        t = self.cur_t = self.chunks_t[self.head]
        yaw = random.random() * 2 * np.pi - np.pi if self.random_rotations else None
        window = {'id': self.video_id + ' ' + str(t),
                  'ambix': self.audio_reader.get(t - self.context / 2, self.audio_size, yaw)}
        if self.return_video:
            window['video'] = self.video_reader.get_by_index(t, self.video_size, yaw)
        if self.return_flow:
            flow = self.flow_reader.get_by_index(t, self.video_size, yaw)
            if self.raw_flow:
                window['flow'], window['flow_limits'] = flow
            else:
                window['flow'] = flow
        return window

    def loop_chunks(self, n=np.inf):
        served = 0
        while served < n:
            window = self.get()
            if window is None:
                return
            served += 1
            yield window
