"""Evaluation pass of the reference's eval.py (reference eval.py:29-232) on top of libsag.so.

Per batch of windows: forward (eval.py:90), the in-graph per-sample metrics (model.evaluation_ops, model.py:110-154),
the Hilbert-envelope distance (myutils.py:109-116), amplitudes (eval.py:197-198) and the 84-direction RMS energy maps
that feed the EMD (distance.py:41-52, ang_res=30) -- all as GPU kernels.  Rows come out in the reference's
`eval-detailed.txt` format (`SampleID | <28 metric names>`, eval.py:125-133, 212-215).  The EMD columns (pyemd in the
reference) come from the exact host solver in libsag.so when the energy maps are requested (nan otherwise); the mel_lsd
columns (librosa in the reference) from the restated mel spectrogram kernel.  With several ranks
(one process per GPU) whole batches are sharded and the rows meet in one all-gather (dist.gather_rows).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import metrics as M

ALL_METRICS = ['amplitude/predicted', 'amplitude/gt',
               'mse/avg', 'mse/X', 'mse/Y', 'mse/Z',
               'stft/avg', 'stft/X', 'stft/Y', 'stft/Z',
               'lsd/avg', 'lsd/X', 'lsd/Y', 'lsd/Z',
               'mel_lsd/avg', 'mel_lsd/X', 'mel_lsd/Y', 'mel_lsd/Z',
               'snr/avg', 'snr/X', 'snr/Y', 'snr/Z',
               'env_mse/avg', 'env_mse/X', 'env_mse/Y', 'env_mse/Z',
               'emd/dir', 'emd/dir2']                                      # eval.py:125-132
_COL = {k: i for i, k in enumerate(ALL_METRICS)}
N_COLS = len(ALL_METRICS)
# metric_rows gathers its columns from [amp_pred, amp_gt | mse, stft, lsd, mel_lsd, snr, env x (Y,Z,X) | the 6 channel means | nan]
_KEYS = ('mse', 'stft', 'lsd', 'mel', 'snr', 'env')
_NAMES = ('mse', 'stft', 'lsd', 'mel_lsd', 'snr', 'env_mse')
_PERM_CACHE = {}


def _perm(device):
    """Column gather of metric_rows: ALL_METRICS position -> position in the concatenated metric buffer (filled by key,
    not by position, like eval.py:159-171; the emd columns point at the nan column until metric_rows fills them)."""
    key = str(device)
    if key not in _PERM_CACHE:
        src = {'amplitude/predicted': 0, 'amplitude/gt': 1}
        for k, name in enumerate(_NAMES):
            for i, ch in enumerate('YZX'):
                src[name + '/' + ch] = 2 + 3 * k + i
            src[name + '/avg'] = 2 + 3 * len(_NAMES) + k
        nan_col = 2 + 4 * len(_NAMES)
        _PERM_CACHE[key] = torch.tensor([src.get(m, nan_col) for m in ALL_METRICS], dtype=torch.int64, device=device)
    return _PERM_CACHE[key]


def metric_rows(pred, target, mono=None, layout=None, audio_rate=48000, rms_maps=False, mel_lsd=True, emd=True):
    """pred, target (B, T, 3) CUDA, channels (Y, Z, X).  Returns (rows (B, 28) CUDA float32 in ALL_METRICS order,
    maps or None).  maps = (rms_pred, rms_gt), each (B, 7, 12): energy maps of [W | pred or gt] * layout
    (eval.py:147-148, 190), needs `mono` (B, T, 1)."""
    res = M.window_metrics(pred, target, audio_rate)
    # eval.py:173-178 (myutils.compute_lsd_dist); mel_lsd=False leaves those four columns nan
    res['mel'] = M.mel_lsd(pred, target, audio_rate) if mel_lsd else torch.full_like(res['mse'], float('nan'))
    B = pred.shape[0]
    # four small launches instead of one per column: [amp | per-channel metrics] -> channel means -> one column gather
    per_ch = torch.cat([res['amp']] + [res[k] for k in _KEYS], 1)          # (B, 2 + 6*3), channels in (Y, Z, X) order
    avg = per_ch[:, 2:].reshape(B, len(_KEYS), 3).mean(2)                  # np.mean over the 3 channels (eval.py:159-171)
    full = torch.cat((per_ch, avg, torch.full((B, 1), float('nan'), dtype=torch.float32, device=pred.device)), 1)
    rows = full.index_select(1, _perm(pred.device))
    maps = None
    if rms_maps:
        if mono is None:
            raise ValueError('rms_maps needs the W channel (mono)')
        lay = torch.ones((B, 1, 4), device=pred.device) if layout is None else torch.as_tensor(layout, device=pred.device).float()[:, None, :]
        maps = (M.ambix_rms_map(torch.cat((mono, pred), 2) * lay, 30.), M.ambix_rms_map(torch.cat((mono, target), 2) * lay, 30.))
        if emd:
            # emd/dir, emd/dir2 (eval.py:190-193): exact EMD of the two 84-node maps, solved on the host like the reference
            d1, d2 = M.ambix_emd_from_maps(maps[0], maps[1], 30.)
            rows[:, _COL['emd/dir']:_COL['emd/dir2'] + 1] = torch.as_tensor(np.stack((d1, d2), 1), dtype=torch.float32).to(rows.device)
    return rows, maps


def evaluate_batches(model, batches, audio_rate=48000, rms_maps=False, lanes=3):
    """eval.py:140-201 for an iterable of dicts {'id': list, 'ambix': (B, snd_size, 4), 'video'/'flow': ..., 'mask': (B,4)}
    (CUDA float32 tensors; uint8 frames as decoded).  Returns (ids, rows (N, 28) CUDA).  Consecutive batches are independent, so up
    to `lanes` of them are in flight, each on its own stream with a twin of the model (SptAudioGen._lanes: same weights and
    options); rows come back in batch order and do not depend on `lanes`."""
    ids, out = [], []
    ss, t = model.snd_contx // 2, model.snd_dur                            # eval.py:67-68
    lanes = max(1, int(lanes))
    with torch.cuda.device(model.device):
        main = torch.cuda.current_stream()
        models = model._lanes(lanes)
        streams = [main] + [torch.cuda.Stream(device=model.device) for _ in range(lanes - 1)]
        try:
            for i, b in enumerate(batches):
                m, st = models[i % lanes], streams[i % lanes]
                ready = torch.cuda.Event()
                ready.record(main)                                          # the batch was produced on the caller's stream
                with torch.cuda.stream(st):
                    st.wait_event(ready)
                    ambix = b['ambix']
                    audio_input = ambix[:, :, :1].contiguous()              # eval.py:69
                    target = ambix[:, ss:ss + t, 1:].contiguous()           # eval.py:70
                    pred = torch.empty((ambix.shape[0], t, 3), dtype=torch.float32, device=ambix.device)
                    m.forward_into(audio_input, b.get('video'), b.get('flow'), pred, b.get('flow_limits'))   # the hot loop (fused inverse STFT + mixing)
                    rows, _ = metric_rows(pred, target, mono=audio_input[:, ss:ss + t], layout=b.get('mask'), audio_rate=audio_rate,
                                          rms_maps=rms_maps)
                    if st is not main:
                        for v in b.values():                                # memory of the caller's stream read on this one
                            if isinstance(v, torch.Tensor) and v.is_cuda:
                                v.record_stream(st)
                ids.extend(b['id'])
                out.append(rows)
        finally:
            for st in streams[1:]:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)
        for r in out:
            r.record_stream(main)
        return ids, (torch.cat(out, 0) if out else torch.empty((0, N_COLS)))


def prefetch(iterable, depth=2):
    """The reference's feeder thread + queue (feeder.py:281-435: reader threads fill a bounded queue the session dequeues from) for
    one producer: `iterable` is consumed on a background thread, at most `depth` items ahead of the caller; items arrive in order,
    an exception in the producer is re-raised in the caller at the same position.  depth <= 0: no thread."""
    if depth <= 0:
        for x in iterable:
            yield x
        return
    import queue
    import threading
    q = queue.Queue(maxsize=depth)
    end, stop = object(), threading.Event()

    def put(x):
        while not stop.is_set():
            try:
                q.put(x, timeout=0.1)
                return True
            except queue.Full:
                pass
        return False

    def run():
        try:
            for x in iterable:
                if not put((None, x)):
                    return
            put((None, end))
        except BaseException as e:                              # noqa: B902 (handed to the consumer)
            put((e, None))

    t = threading.Thread(target=run, name='sag-feeder', daemon=True)
    t.start()
    try:
        while True:
            err, x = q.get()
            if err is not None:
                raise err
            if x is end:
                return
            yield x
    finally:
        stop.set()                                              # an abandoned generator releases its producer


def folder_batches(folders, params, batch_size=16, channel_masks=None, device=None, drop_remainder=True, gpu_jpeg=True, prefetch_depth=2):
    """The evaluation feeder (reference feeder.py:366-420 with for_eval=True, eval.py:43-60): every `folders[i]` (a per-video
    folder, see readers.py) is read in order with the eval schedule -- every 10th entry of audio_pow.lst, no shuffling, no
    rotations, silent chunks kept -- and the samples are grouped into batches of `batch_size` like `dequeue_many`.
    channel_masks: {video id: (4,) mask} from meta/audio_layouts.txt (feeder.py:312-314), default all ones.  The last, short
    batch is dropped by default, like the reference (its `dequeue_many` never returns it, feeder.py:412-419): the visual towers use
    batch statistics, so rows computed from a smaller batch correspond to nothing the reference computes.  drop_remainder=False
    yields it anyway (its dict carries 'short_batch': True).  gpu_jpeg: the frames' jpg files are decoded on the GPU, a batch at a
    time (readers.JpegDecoder: bit-identical to the PIL decode of gpu_jpeg=False; baseline files only -- others raise).
    prefetch_depth: batches of files read ahead by a feeder thread (wav / jpg file reads overlap the GPU; 0: read inline)."""
    from . import readers, myutils
    from .definitions import VIDEO, FLOW
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    decoders = {}

    def flush(items):
        # frames travel as decoded (uint8; flow with its per-frame limits) and are prepared by the ingest kernel on the device
        b = {'id': [c['id'] for c in items],
             'ambix': torch.as_tensor(np.stack([c['ambix'] for c in items]).astype(np.float32)).to(dev),
             'mask': torch.as_tensor(np.stack([c['mask'] for c in items]).astype(np.float32)).to(dev)}
        for k in (VIDEO, FLOW):
            if k in params.encoders and gpu_jpeg:
                files = [f for c in items for f in c[k]]
                if k not in decoders:
                    h, w = readers.jpeg_info(files[0])[:2]
                    decoders[k] = readers.JpegDecoder(batch_size * len(items[0][k]), h, w, device=dev)
                frames = decoders[k].decode(files)
                b[k] = frames.view(len(items), len(items[0][k]), frames.shape[1], frames.shape[2], 3)
            elif k in params.encoders:
                b[k] = torch.as_tensor(np.ascontiguousarray(np.stack([c[k] for c in items]))).to(dev)
        if FLOW in params.encoders:
            b['flow_limits'] = torch.as_tensor(np.stack([np.asarray(c['flow_limits'], np.float64).reshape(-1, 2)[0] for c in items])).to(dev)
        return b

    def host_batches():                                        # host side only: file reads, wav parsing (runs on the feeder thread)
        pending = []
        for folder in folders:
            r = readers.SampleReader(folder, ambi_order=params.ambi_order, audio_rate=params.audio_rate, video_rate=params.video_rate,
                                     context=params.context, duration=0.1, return_video=VIDEO in params.encoders,
                                     img_prep=None, return_flow=FLOW in params.encoders, skip_silence_thr=None,
                                     shuffle=False, random_rotations=False, skip_rate=10, raw_flow=True, jpeg_files=gpu_jpeg)
            mask = np.ones(4) if channel_masks is None else np.asarray(channel_masks.get(r.video_id, np.ones(4)))
            for c in r.loop_chunks():
                c['mask'] = mask
                pending.append(c)
                if len(pending) == batch_size:
                    yield pending
                    pending = []
        if pending and not drop_remainder:
            yield pending

    for items in prefetch(host_batches(), prefetch_depth):
        b = flush(items)
        if len(items) < batch_size:
            b['short_batch'] = True
        yield b


def evaluate_model_dir(model_dir, subset_fn=None, overwrite=False, db_dir=None, audio_layouts_fn='meta/audio_layouts.txt', batch_size=16,
                       drop_remainder=True, precision=None, device=None):
    """eval.py:29-215 `main`: the model of `model_dir` (train-params.txt + checkpoint, restored by name) over the per-video
    folders of the dataset directory (`db_dir`, default the one recorded in train-params.txt) restricted to `subset_fn`, with
    the eval schedule and the channel masks of `audio_layouts_fn`; writes `<model_dir>/eval-detailed.txt` (refusing to
    overwrite it unless asked, eval.py:31-32) and returns (sample ids, rows (N, 28) CUDA).  Like the reference's queue, the
    last, short batch is dropped (drop_remainder=False evaluates it too)."""
    import os
    from . import myutils, readers
    from .deploy import W2XYZ
    eval_fn = os.path.join(model_dir, 'eval-detailed.txt')
    assert os.path.exists(model_dir), 'Model dir does not exist.'
    assert overwrite or not os.path.exists(eval_fn), 'Evaluation file already exists.'
    params = myutils.load_params(model_dir)
    folders = readers.sample_folders(db_dir if db_dir is not None else params.db_dir, subset_fn)
    masks = readers.load_channel_masks(audio_layouts_fn) if audio_layouts_fn is not None and os.path.exists(audio_layouts_fn) else None
    model = W2XYZ(model_dir, params=params, precision=precision, device=device).model      # same construction / restore as eval.py:75-118
    batches = folder_batches(folders, params, batch_size=batch_size, channel_masks=masks, device=model.device, drop_remainder=drop_remainder)
    ids, rows = evaluate_batches(model, batches, audio_rate=params.audio_rate, rms_maps=True)
    write_eval_detailed(eval_fn, ids, rows)
    return ids, rows


def write_eval_detailed(path, sample_ids, rows):
    """eval.py:212-215: header `SampleID | names`, then `<id> | v1 ... v28` (space-pipe-space separator)."""
    rows = np.asarray(rows.detach().cpu() if isinstance(rows, torch.Tensor) else rows)
    with open(path, 'w') as f:
        f.write('SampleID | {}\n'.format(' '.join(ALL_METRICS)))
        for sid, r in zip(sample_ids, rows):
            f.write('{} | {}\n'.format(sid, ' '.join([str(float(v)) for v in r])))


def parse_eval_detailed_file(fn):
    """parse_eval_results.py:9-30: an eval-detailed.txt back into per-video arrays -- ({video: (n, 28) values}, {video: (n,) times},
    column names), each video's windows sorted by time."""
    lines = open(fn).read().splitlines()
    metrics = lines[0].split(' | ')[1].split()
    vals, times = OrderedDict(), OrderedDict()
    for line in lines[1:]:
        if not line.strip():
            continue
        sid, row = line.split(' | ')
        vid, t = sid.split()
        times.setdefault(vid, []).append(float(t))
        vals.setdefault(vid, []).append([float(v) for v in row.split()])
    for vid in sorted(vals):
        order = np.argsort(np.asarray(times[vid]), kind='stable')
        times[vid] = np.asarray(times[vid])[order]
        vals[vid] = np.asarray(vals[vid])[order]
    return OrderedDict((v, vals[v]) for v in sorted(vals)), OrderedDict((v, times[v]) for v in sorted(vals)), metrics


def parse_eval_results(fn, samples_per_sec=4800):
    """parse_eval_results.py:32-51, the table the paper reports: per video the mean over its windows of sqrt(mse * 4800),
    stft, sqrt(env_mse^2 * 4800), sqrt(emd^2 * 4800), then the mean over videos.  Returns OrderedDict MSE / STFT / ENV / EMD;
    `python -m spatialaudiogen_b200.evaluate <eval-detailed.txt>` prints it like the reference script."""
    vals, _, keys = parse_eval_detailed_file(fn)
    out = OrderedDict()
    for label, mt in (('MSE', 'mse/avg'), ('STFT', 'stft/avg'), ('ENV', 'env_mse/avg'), ('EMD', 'emd/dir')):
        i = keys.index(mt)
        if mt in ('emd/dir', 'env_mse/avg'):
            per_video = [np.sqrt(v[:, i] ** 2 * samples_per_sec).mean() for v in vals.values()]
        elif mt == 'mse/avg':
            per_video = [np.sqrt(v[:, i] * samples_per_sec).mean() for v in vals.values()]
        else:
            per_video = [v[:, i].mean() for v in vals.values()]
        out[label] = float(np.mean(per_video))
    return out


def summarize(rows):
    rows = np.asarray(rows.detach().cpu() if isinstance(rows, torch.Tensor) else rows)
    return OrderedDict((k, float(np.nanmean(rows[:, i])) if np.isfinite(rows[:, i]).any() else float('nan'))
                       for k, i in _COL.items())


def parse_arguments(argv=None):
    """eval.py:14-26 -- the same arguments."""
    import argparse
    import sys
    parser = argparse.ArgumentParser(description='Evaluate a model snapshot (reference eval.py), or print the table of an eval-detailed.txt.')
    parser.add_argument('model_dir', help='Directory to store model -- or an eval-detailed.txt to summarise (parse_eval_results.py).')
    parser.add_argument('--subset_fn', default='')
    parser.add_argument('--batch_size', default=16, type=int)
    parser.add_argument('--overwrite', action='store_true')
    parser.add_argument('--gpu', type=int, default=0, help='GPU id')
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    if len(args.subset_fn) == 0:
        args.subset_fn = None
    return args


def main(args):
    import os
    if os.path.isfile(args.model_dir):           # parse_eval_results.py
        for label, v in parse_eval_results(args.model_dir).items():
            print('{} = {:.3f}'.format(label.ljust(4), v))
        return
    with torch.cuda.device(args.gpu):            # eval.py:29-215
        evaluate_model_dir(args.model_dir, subset_fn=args.subset_fn, overwrite=args.overwrite, batch_size=args.batch_size)


if __name__ == '__main__':                       # python -m spatialaudiogen_b200.evaluate MODEL_DIR [--subset_fn ...]  |  <eval-detailed.txt>
    main(parse_arguments())
