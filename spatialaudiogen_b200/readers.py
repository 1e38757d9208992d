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
`W2XYZ.deploy` / `evaluate.evaluate_batches` upload.  Files at another sample rate are rejected instead of resampled
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


class AudioReader(object):
    """feeder.py:50-105."""
    def __init__(self, audio_folder, rate=None, ambi_order=1):
        self.audio_folder = audio_folder
        fns = [f for f in os.listdir(audio_folder) if f.endswith('.wav')]
        self.num_files = len(fns)
        data, file_rate = load_wav(os.path.join(audio_folder, sorted(fns)[0]))
        self.rate = float(file_rate) if rate is None else rate
        self.num_channels = min((data.shape[1], (ambi_order + 1) ** 2))
        self.duration = self.num_files
        self.num_frames = int(self.duration * self.rate)

    def get(self, start_time, size, rotation=None):
        """`size` samples starting at `start_time` seconds; whatever falls before 0 or after the last file is zero
        (feeder.py:66-90).  The offset inside the first file is int(frac(start_time) * rate), like the reference."""
        out = np.zeros((size, self.num_channels))
        lead = max(-int(start_time * self.rate), 0)              # samples before the start of the clip
        t0 = 0. if lead > 0 else start_time
        want = size - lead
        first_frame = 0 if lead > 0 else int(start_time * self.rate)
        want -= max(first_frame + want - self.num_frames, 0)      # samples beyond the end of the clip
        if want > 0:
            sec0 = int(t0)
            sec1 = min(int(np.ceil(t0 + want / float(self.rate))), self.num_files)
            parts = [load_wav('{}/{:06d}.wav'.format(self.audio_folder, i), self.rate)[0] for i in range(sec0, sec1)]
            data = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)
            ss = int((t0 - sec0) * self.rate)
            data = data[ss:ss + want, :self.num_channels]
            out[lead:lead + data.shape[0]] = data
        if rotation is not None:                                 # feeder.py:92-102: yaw rotation of (W, Y, Z, X)
            assert -np.pi <= rotation < np.pi
            c, s = np.cos(rotation), np.sin(rotation)
            out = np.dot(out, np.array([[1, 0, 0, 0], [0, c, 0, s], [0, 0, 1, 0], [0, -s, 0, c]]).T)
        return out


class VideoReader(object):
    """feeder.py:108-135."""
    def __init__(self, video_folder, rate=None, img_prep=None):
        raw_rate = 10.
        self.video_folder = video_folder
        self.rate = rate if rate is not None else raw_rate
        self.img_prep = img_prep if img_prep is not None else lambda x: x
        frame_fns = [fn for fn in os.listdir(video_folder) if fn.endswith('.jpg')]
        self.num_frames = len(frame_fns)
        self.duration = self.num_frames / raw_rate
        self.frame_shape = self.img_prep(_imread(os.path.join(video_folder, sorted(frame_fns)[0]))).shape

    def get_by_index(self, start_time, size, rotation=None):
        ss = max(int(start_time * self.rate), 0)
        chunk = [self.img_prep(_imread(os.path.join(self.video_folder, '{:06d}.jpg'.format(fno)))) for fno in range(ss, ss + size)]
        chunk = np.stack(chunk, 0) if len(chunk) > 1 else chunk[0][np.newaxis]
        if rotation is not None:
            roll = -int(rotation / (2. * np.pi) * self.frame_shape[1])
            chunk = np.roll(chunk, roll, axis=2)
        return chunk


def dequantize_flow(chunk_u8, limits):
    """feeder.py:147-161 on (T, H, W, 3) uint8 frames (channel 0 = angle, 2 = magnitude) with their (T, 2) (min, max) rows of
    flow_limits.npy -> float32 (mag cos, mag sin, mag).  (The device does the same inside the frame-ingest kernel.)"""
    chunk = np.asarray(chunk_u8).astype(np.float32)
    lead = chunk.shape[:-3]
    chunk = chunk.reshape((-1,) + chunk.shape[-3:])
    limits = np.asarray(limits).reshape(-1, 2)
    m_min = limits[:, 0].reshape((-1, 1, 1))
    m_max = limits[:, 1].reshape((-1, 1, 1))
    chunk[:, :, :, 2] *= (m_max - m_min) / 255.              # magnitude back to its range
    chunk[:, :, :, 2] += m_min
    chunk[:, :, :, 0] *= (2 * np.pi) / 255.                  # angle
    chunk[:, :, :, 1] = chunk[:, :, :, 2] * np.sin(chunk[:, :, :, 0])
    chunk[:, :, :, 0] = chunk[:, :, :, 2] * np.cos(chunk[:, :, :, 0])
    return chunk.reshape(lead + chunk.shape[-3:])


class FlowReader(object):
    """feeder.py:138-161."""
    def __init__(self, flow_dir, flow_lims_fn, rate=None, flow_prep=None):
        self.reader = VideoReader(flow_dir, rate=rate)
        self.lims = np.load(flow_lims_fn)
        self.rate = self.reader.rate
        self.duration = self.reader.duration
        self.flow_prep = flow_prep if flow_prep is not None else lambda x: x

    def get_by_index(self, start_time, size, rotation=None):
        chunk = self.reader.get_by_index(start_time, size, rotation)
        ss = max(int(start_time * self.rate), 0)
        return dequantize_flow(chunk, self.lims[ss:ss + chunk.shape[0]])


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


class SampleReader(object):
    """feeder.py:164-278: iterates the chunk times of `<folder>/audio_pow.lst`."""
    def __init__(self, folder, ambi_order=1, audio_rate=48000, video_rate=10, context=1.0, duration=0.1, return_video=True,
                 img_prep=None, return_flow=False, flow_prep=None, skip_silence_thr=None, shuffle=True, start_time=0.5,
                 sample_duration=None, skip_rate=None, random_rotations=True, num_threads=1, thread_id=0):
        a2v = float(audio_rate) / video_rate
        snd_dur, vid_dur, snd_ctx = duration * audio_rate, duration * video_rate, context * audio_rate
        self.video_id = os.path.split(folder)[-1]
        assert a2v == int(a2v)
        assert abs(snd_dur - round(snd_dur)) < 1e-6 and abs(vid_dur - round(vid_dur)) < 1e-6 and abs(snd_ctx - round(snd_ctx)) < 1e-6
        self.audio_reader = AudioReader(os.path.join(folder, 'ambix'), audio_rate, ambi_order)
        self.video_reader = VideoReader(os.path.join(folder, 'video'), video_rate, img_prep) if return_video else None
        if return_flow:
            flow_dir = os.path.join(folder, 'flow')
            self.flow_reader = FlowReader(flow_dir, os.path.join(flow_dir, 'flow_limits.npy'), video_rate, flow_prep)
        self.folder, self.duration, self.context = folder, duration, context
        self.audio_rate, self.video_rate = audio_rate, video_rate
        self.audio_size = int(round(snd_dur)) + int(round(snd_ctx)) - 1
        self.video_size = int(round(vid_dur))
        self.video_shape = self.video_reader.frame_shape if return_video else None
        self.return_video, self.return_flow, self.random_rotations = return_video, return_flow, random_rotations
        lines = [l.strip().split() for l in open(os.path.join(folder, 'audio_pow.lst')) if l.strip()]
        chunks_t, chunks_pow = [float(l[0]) for l in lines], [float(l[1]) for l in lines]
        if skip_rate is not None:
            keep = range(0, len(chunks_t), skip_rate)
            chunks_t, chunks_pow = [chunks_t[i] for i in keep], [chunks_pow[i] for i in keep]
        if skip_silence_thr is not None:
            chunks_t = [t for t, p in zip(chunks_t, chunks_pow) if p > skip_silence_thr]
        if start_time > 0.5:
            chunks_t = [t for t in chunks_t if t >= start_time]
        if sample_duration is not None:
            chunks_t = [t for t in chunks_t if t < start_time + sample_duration]
        if num_threads > 1:
            lims = np.linspace(0, len(chunks_t), num_threads + 1).astype(int)
            chunks_t = chunks_t[lims[thread_id]:lims[thread_id + 1]]
        if shuffle:
            random.shuffle(chunks_t)
        self.chunks_t = chunks_t
        self.head = -1

    def get(self):
        self.head += 1
        if self.head >= len(self.chunks_t):
            return None
        self.cur_t = cur_t = self.chunks_t[self.head]
        rotation = random.random() * 2 * np.pi - np.pi if self.random_rotations else None
        chunks = {'id': self.video_id + ' ' + str(cur_t)}
        chunks['ambix'] = self.audio_reader.get(cur_t - self.context / 2, self.audio_size, rotation)
        if self.return_video:
            chunks['video'] = self.video_reader.get_by_index(cur_t, self.video_size, rotation)
        if self.return_flow:
            chunks['flow'] = self.flow_reader.get_by_index(cur_t, self.video_size, rotation)
        return chunks

    def loop_chunks(self, n=np.inf):
        k = 0
        while True:
            k += 1
            if k > n:
                break
            chunks = self.get()
            if chunks is None:
                break
            yield chunks
