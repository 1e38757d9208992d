#!/usr/bin/env python
"""bench.py -- ambisonic audio seconds per second of the spatialaudiogen inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--precision P] [--frames u8|f32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (sag_forward: STFT -> audio/video towers -> U-Net decoder -> masked iSTFT ->
32->3 mixing, reference model.py:356-434 behind deploy.py:141 / eval.py:145) over one batch of B=32 synthetic
0.1 s windows (52 799 mono samples @48 kHz + one 224x448x3 RGB frame each), followed by the per-window evaluation
metrics (reference model.py:110-154, eval.py:145 fetches both).  At N>1 every rank processes its own batches (whole
batches shard, weights replicate: SURVEY.md 8e) and ONE all-gather of the per-window metric rows closes the timed
region.  value = 0.1 s x windows processed by all ranks / max-over-ranks device time.

Lines printed (rank 0 only, one JSON object each run):
  default arm      : value (inputs resident in HBM), e2e (host pinned buffers -> H2D -> forward -> D2H of the waveform,
                     through SptAudioGen.inference_ops, the call a user makes), roofline of the dominant kernel family
                     (dense contractions), cpu_baseline (the CPU oracle on this box's host cores, N=1 only).
  --impl reference : the reference's CPU path.  TensorFlow 1.4 / python2 cannot be installed, so the
                     stand-in is oracle/sag_oracle.py (PyTorch-CPU restatement of the TF graph) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# NCCL's own log lines (version banner, NCCL_DEBUG=INFO) go to stderr: stdout carries the one JSON line only
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')

import numpy as np
import torch

METRIC = 'ambisonic_audio_seconds_per_second'
UNIT = 'audio_s/s'
SND_SIZE, SND_DUR, RATE = 52799, 4800, 48000
FRAME = (224, 448)
WINDOW_S = 0.1
# executed / reference conv-stack FLOPs per window: A = 2.386, each ResNet tower 7.254 GFLOP
CONV_GFLOP_REF = {'audio': 2.386365440, 'tower': 7.254245376}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help='BASELINE.json configs, numbered like SURVEY.md 8d: 1 = B=1 audio-only; 2 (default, the configuration the '
                         'metric is quoted on) = audio+video B=32; 3 = audio+video+flow B=32; 4 = YT-All-shaped clip stream, clip-sharded; '
                         '5 = eval pass over 10k windows with all metric columns + RMS maps')
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--encoders', default=None)
    ap.add_argument('--frames', default='u8', choices=['u8', 'f32'],
                    help='video / flow frames as the readers decode them (uint8, prepared on the device) or as prepared float32 tensors')
    ap.add_argument('--layer-table', default=None, help='write the per-layer CUDA-event table of the profiled forward to this JSON file')
    ap.add_argument('--precision', default=os.environ.get('SAG_BENCH_PRECISION', 'auto'))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-baseline-seconds', type=float, default=15.0)
    ap.add_argument('--rotate', type=int, default=4, help='distinct input batches cycled through (L2 hygiene)')
    ap.add_argument('--lanes', type=int, default=int(os.environ.get('SAG_LANES', '3')),
                    help='forwards in flight: consecutive batches alternate over this many (handle, workspace, stream) lanes, so the '
                         'short grids of one batch (FCs, decoder, BN passes) fill the SMs the other batch leaves idle')
    args = ap.parse_args()
    dflt = {1: (1, 'audio'), 2: (32, 'audio,video'), 3: (32, 'audio,video,flow'), 4: (32, 'audio,video'), 5: (32, 'audio,video')}[args.config]
    if args.batch is None:
        args.batch = dflt[0]
    if args.encoders is None:
        args.encoders = dflt[1]
    return args


# ---- synthetic inputs (SURVEY.md 8d "Synthetic value distributions") ----------------------------------------------
def synth_batch(B, encoders, seed):
    rng = np.random.RandomState(seed)
    t = np.arange(SND_SIZE)
    phase = rng.uniform(0, 2 * np.pi, size=(B, 1))

    def wave():
        return np.clip(0.1 * rng.randn(B, SND_SIZE) + 0.3 * np.sin(2 * np.pi * 440. * t / RATE + phase), -1, 1).astype(np.float32)

    out = {'audio': wave()[:, :, None]}
    # targets: independently generated Y,Z,X for the centre 0.1 s (eval.py:70)
    out['target'] = np.stack([wave()[:, RATE // 2:RATE // 2 + SND_DUR] for _ in range(3)], axis=2)
    if 'video' in encoders:
        # the frame as the reader decodes it (uint8) and as myutils.img_prep_fcn prepares it (x/255 - 0.5, myutils.py:88-89)
        out['video_u8'] = rng.randint(0, 256, size=(B, 1) + FRAME + (3,)).astype(np.uint8)
        out['video'] = (out['video_u8'] / 255. - 0.5).astype(np.float32)
    if 'flow' in encoders:
        # mag ~ U[0,20), theta ~ U[0,2pi), stored the way scraping/preprocess.py stores flow: 8-bit (angle, -, magnitude) + the
        # frame's (min, max); the float32 frame is FlowReader.get_by_index's de-quantisation of it (feeder.py:147-161)
        mag = rng.uniform(0, 20, size=(B, 1) + FRAME)
        th = rng.uniform(0, 2 * np.pi, size=(B, 1) + FRAME)
        lims = np.stack([mag.min(axis=(1, 2, 3)), mag.max(axis=(1, 2, 3))], 1).astype(np.float64)
        q = np.zeros((B, 1) + FRAME + (3,), np.uint8)
        q[..., 0] = np.round(th / (2 * np.pi) * 255.)
        q[..., 2] = np.round((mag - lims[:, 0].reshape(-1, 1, 1, 1)) / (lims[:, 1] - lims[:, 0]).reshape(-1, 1, 1, 1) * 255.)
        chunk = q[:, 0].astype(np.float32)
        chunk[:, :, :, 2] *= (lims[:, 1] - lims[:, 0]).reshape((-1, 1, 1)) / 255.
        chunk[:, :, :, 2] += lims[:, 0].reshape((-1, 1, 1))
        chunk[:, :, :, 0] *= (2 * np.pi) / 255.
        chunk[:, :, :, 1] = chunk[:, :, :, 2] * np.sin(chunk[:, :, :, 0])
        chunk[:, :, :, 0] = chunk[:, :, :, 2] * np.cos(chunk[:, :, :, 0])
        out['flow_u8'], out['flow_limits'], out['flow'] = q, lims, chunk[:, None]
    return out


def conv_gflop_per_window(encoders):
    return CONV_GFLOP_REF['audio'] + CONV_GFLOP_REF['tower'] * (('video' in encoders) + ('flow' in encoders))


# ---- clocks sampler (B200_PROFILING.md "clocks DURING the timed region") ------------------------------------------
class ClockSampler(object):
    Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix='sag_clocks_', suffix='.csv')
        self.proc = None
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0, t1):
        """Median SM clock / reasons over wall interval [t0, t1] (falls back to all samples if none land inside)."""
        res = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return res
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        try:
            import datetime
            for line in open(self.path):
                p = [s.strip() for s in line.split(',')]
                if len(p) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                    rows.append((ts, float(p[2]), float(p[3]), p[6], p[7], p[8], p[9]))
                except Exception:
                    continue
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        inside = [r for r in rows if t0 <= r[0] <= t1 + 0.1]
        use = inside if inside else rows
        if not use:
            return res
        res['samples'] = len(inside)
        res['sm_mhz'] = float(np.median([r[1] for r in use]))
        res['sm_max_mhz'] = float(max(r[2] for r in use))
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for k, n in enumerate(names):
            if any(r[3 + k].lower().startswith('active') for r in use):
                res['reasons'].append(n)
        return res


# ---- the CPU arm: oracle port of the reference's TF1 CPU deploy path ----------------------------------------------
def oracle_model(encoders, seed=1234):
    from oracle import sag_oracle as O                    # bench.py's cpu legs are allowed to execute the oracle
    from spatialaudiogen_b200 import weights as Wt
    W = Wt.init_weights(encoders, separation='unet_mask', seed=seed, resnet_npy=RESNET_NPY if os.path.exists(RESNET_NPY) else None)
    return O.SptAudioGen(W, encoders=encoders, separation='unet_mask')


def cpu_time_forward(model, batch, encoders, b):
    kw = {k: batch[k][:b] for k in ('video', 'flow') if k in encoders}
    t = time.perf_counter()
    model.inference_ops(batch['audio'][:b], **kw)
    return time.perf_counter() - t


def cpu_baseline(encoders, B, budget_s):
    """Oracle forward at the bench batch size on all host threads; median of the timed forwards."""
    torch.set_num_threads(os.cpu_count() or 1)
    m = oracle_model(encoders)
    batch = synth_batch(B, encoders, 99)
    cpu_time_forward(m, batch, encoders, B)                # warm-up (oneDNN primitive caches)
    ts = []
    t_start = time.perf_counter()
    while len(ts) < 3 or (time.perf_counter() - t_start < budget_s and len(ts) < 50):
        ts.append(cpu_time_forward(m, batch, encoders, B))
    med = float(np.median(ts))
    return {'value': WINDOW_S * B / med, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d forwards of one batch of %d windows (%s), median %.3f s each; oracle/sag_oracle.py '
                      '(PyTorch-CPU fp32 restatement; TF 1.4 is not installable)' % (len(ts), B, '+'.join(encoders), med)}


def run_reference(args, encoders):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    m = oracle_model(encoders)
    batch = synth_batch(B, encoders, 1234)
    t32 = cpu_time_forward(m, batch, encoders, B)          # untimed probe (also warms caches)
    t32 = min(t32, cpu_time_forward(m, batch, encoders, B))
    total = args.steps + args.warmup
    b = B
    if total * t32 > 150.0:                                # bound the whole run to a few minutes
        b = max(1, int(B * 150.0 / (total * t32)))
    for _ in range(args.warmup):
        cpu_time_forward(m, batch, encoders, b)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_time_forward(m, batch, encoders, b)
    el = time.perf_counter() - t
    val = WINDOW_S * b * args.steps / el
    sample = 'each step = oracle forward of %d of the %d windows of a batch (%s) on %d host threads' % (
        b, B, '+'.join(encoders), torch.get_num_threads())
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(encoders, B, default_precision_name(args), args, 1),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': 'reference = TF1.4 CPU graph; not installable here, so the PyTorch-CPU oracle port stands in'}
    print(json.dumps(line))
    return 0


def default_precision_name(args):
    if args.precision != 'auto':
        return args.precision
    return os.environ.get('SAG_PRECISION', 'mixed')


RESNET_NPY = os.path.join(ROOT, 'tests', 'golden', '_ref', 'resnet18.npy')   # staged by __graft_entry__.build() (git-ignored)

WORKLOADS = {
    1: 'configs[0]: single 0.1 s window, audio-only encoder, random weights',
    2: 'configs[1]: audio+video encoders, batch 32, 224x448 RGB, 1xB200, fp32-grade parity vs reference',
    3: 'configs[2]: audio+video+flow encoders, batch 32',
    4: 'configs[3]: YT-All-shaped synthetic stream (285 clips, lognormal lengths, one window per second), clip-sharded, one '
       'all-gather of metric rows',
    5: 'configs[4]: eval pass over 10k synthetic windows: forward + stft/lsd/mse/snr/env_mse/amplitude + 84-direction RMS maps, '
       'rows gathered and written in eval-detailed.txt format',
}


def workload_config(encoders, B, precision, args, world):
    return {'workload': '%s -- %s encoders, unet_mask separation, batch %d, 0.1 s @48 kHz mono%s per window'
                        % (WORKLOADS[args.config], '+'.join(encoders), B, ' + 224x448x3 frame per visual tower' if len(encoders) > 1 else ''),
            'baseline_config': args.config, 'batch_per_gpu': B, 'encoders': encoders, 'precision': precision,
            'frames': ('uint8 frames (as decoded from disk), prepared on the device' if args.frames == 'u8' else 'prepared float32 frames')
                      if len(encoders) > 1 else None,
            'weights': 'xavier random init (seed 1234); resnet towers: %s' % (
                'the reference\'s resnet18.npy (model.py:198)' if os.path.exists(RESNET_NPY) else 'xavier (resnet18.npy not staged)'),
            'lanes': max(1, args.lanes),
            'step': 'sag_forward + sag_metrics over one batch' + (' (replayed as one CUDA graph: host-launch-bound batch size)' if B <= 16 else ''),
            'l2': 'inputs rotate over %d distinct batches and each step streams >1 GB of activations through the '
                  'workspace (> 126 MB L2)' % args.rotate,
            'parallelism': 'clip-sharded, weights replicated, one all-gather of metric rows at the end of the pass'}


def layer_table(model, L):
    """Per-launch records of the last profiled forward: name, category, us, useful / issued GFLOP, tile, split."""
    import ctypes as C
    lib = L.lib()
    cats = ['conv', 'deconv', 'fc', 'stft', 'istft', 'pointwise', 'mix']
    out = []
    buf = C.create_string_buffer(64)
    for i in range(lib.sag_num_profile_records(model._h)):
        cat, tile, split = C.c_int(), C.c_int(), C.c_int()
        us, fl, iss, by = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        L.check(lib.sag_get_profile_record(model._h, i, buf, 64, C.byref(cat), C.byref(us), C.byref(fl), C.byref(iss), C.byref(by),
                                           C.byref(tile), C.byref(split)))
        out.append({'name': buf.value.decode(), 'cat': cats[cat.value], 'us': us.value, 'gflop': fl.value / 1e9,
                    'gflop_issued': iss.value / 1e9, 'mb': by.value / 1e6, 'tile': tile.value, 'split': split.value})
    return out


# ---- our arm -------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    encoders = [e.strip() for e in args.encoders.split(',') if e.strip()]
    if args.impl == 'reference':
        return run_reference(args, encoders)

    import torch.distributed as dist
    from spatialaudiogen_b200 import SptAudioGen, weights as Wt, _lib as L, evaluate as E, dist as D

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: libsag.so has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    precision = args.precision
    if precision == 'auto':
        # the per-layer plan measured in tests/test_gpu_bench_config.py::test_precision_plan_error_table_b32: bf16x3 (fp32-grade
        # split operands) everywhere except the U-Net decoder's deconv5..2 (one bf16 product); waveform error at this
        # configuration 6.6e-5 of max|y| with either -- the north-star tolerance is 1e-3.  SAG_PRECISION / --precision override.
        precision = os.environ.get('SAG_PRECISION', 'mixed')
    model = SptAudioGen(1, encoders=encoders, separation='unet_mask', precision=precision, device=dev)
    model.load_weights(Wt.init_weights(encoders, separation='unet_mask', seed=1234,
                                       resnet_npy=RESNET_NPY if os.path.exists(RESNET_NPY) else None))

    # R distinct synthetic batches, resident in HBM (value) and in pinned host memory (e2e)
    n_lanes = max(1, args.lanes)
    R = (max(1, args.rotate) + n_lanes - 1) // n_lanes * n_lanes   # (a multiple of the lanes: slot r always runs on lane r % lanes)
    args.rotate = R
    u8 = args.frames == 'u8'
    vkey, fkey = ('video_u8', 'flow_u8') if u8 else ('video', 'flow')
    host, devb = [], []
    for r in range(R):
        b = synth_batch(B, encoders, 1234 + 1000 * rank + r)
        hb = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()}
        host.append(hb)
        devb.append({k: v.to(dev) for k, v in hb.items() if k in ('audio', 'target', vkey, fkey, 'flow_limits')})
    out = torch.empty((B, SND_DUR, 3), dtype=torch.float32, device=dev)
    out_host = torch.empty((B, SND_DUR, 3), dtype=torch.float32).pin_memory()
    # lanes: lane 0 is (model, out, the current stream); every further lane has its own handle (same weights), workspace,
    # output buffer and stream
    for kv in filter(None, os.environ.get('SAG_BENCH_OPTS', '').split(',')):      # development: "overlap=0,cta_pair=1"
        model.set_option(kv.split('=')[0], int(kv.split('=')[1]))
    twins = model._lanes(n_lanes)                         # [model, twin, ...]: same configuration, options and weights
    prio = [int(x) for x in os.environ.get('SAG_LANE_PRIO', '0,0,0,0,0,0,0,0').split(',')]      # development: stream priority per lane
    lanes = [(model, out, None)] + [(m2, torch.empty_like(out), torch.cuda.Stream(device=dev, priority=prio[i + 1]))
                                    for i, m2 in enumerate(twins[1:])]

    def fwd(d, lane=0):
        lanes[lane][0].forward_into(d['audio'], d.get(vkey), d.get(fkey), lanes[lane][1], d.get('flow_limits') if u8 else None)

    # ---- the pass: which batches this rank processes ----
    if args.config == 4:                                  # YT-All-shaped stream: whole batches of one clip, clips round-robin
        sched = D.shard(D.eval_schedule(D.yt_all_clip_lengths(), B), rank, world)
        batch_ids = [(c, w0) for (c, w0, _) in sched]
    elif args.config == 5:                                # 10k windows = 312 whole batches of 32, dealt round-robin
        batch_ids = [(i, 0) for i in range(10000 // B) if i % world == rank]
    else:
        batch_ids = [(rank, i * B) for i in range(args.steps)]
    n_batches = len(batch_ids)
    n_max = n_batches
    if world > 1:
        t = torch.tensor([n_batches], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_max = int(t.item())
        t = torch.tensor([n_batches], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n_total = int(t.item())
    else:
        n_total = n_batches
    maps_on = args.config == 5
    rows = torch.zeros((max(n_batches, 1), B, E.N_COLS), dtype=torch.float32, device=dev)   # eval-detailed.txt columns per window
    ids = torch.tensor([[c, w0 + j] for (c, w0) in batch_ids for j in range(B)], dtype=torch.int64).reshape(-1, 2).to(dev)
    ss = RATE // 2

    def step_eager(d, store=None, lane=0):
        fwd(d, lane)
        # the metric set of SURVEY.md 8d config 5 (config 5 adds the 84-direction RMS maps of [W | pred] and [W | gt])
        r, _ = E.metric_rows(lanes[lane][1], d['target'], mono=d['audio'][:, ss:ss + SND_DUR] if maps_on else None, audio_rate=RATE,
                             rms_maps=maps_on, mel_lsd=False, emd=False)
        if store is not None:
            store.copy_(r)

    # Small batches are bound by the host's launch rate (~70 kernels per step), not by the GPU: replay the whole step --
    # forward + metrics, programmatic dependent launches included -- as one CUDA graph per rotating input slot.
    graph_env = os.environ.get('SAG_BENCH_GRAPH', '1')             # 0: never, 1: host-bound batch sizes (<= 16), 2: always
    use_graph = (B <= 16 or graph_env == '2') and graph_env != '0'
    graphs, rows_slot = [], []
    if use_graph:
        try:
            for r in range(R):
                rows_slot.append(torch.zeros((B, E.N_COLS), dtype=torch.float32, device=dev))
                step_eager(devb[r], rows_slot[r], r % n_lanes)   # warm-up: plans the batch, sets kernel attributes
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step_eager(devb[r], rows_slot[r], r % n_lanes)
                graphs.append(g)
        except RuntimeError as e:
            sys.stderr.write('CUDA graph capture failed (%s): eager launches\n' % e)
            use_graph, graphs = False, []

    def step(i, store=None):
        r = i % R
        lane = r % n_lanes
        with torch.cuda.stream(lanes[lane][2] or torch.cuda.current_stream()):
            if use_graph:
                graphs[r].replay()
                if store is not None:
                    store.copy_(rows_slot[r])
            else:
                step_eager(devb[r], store, lane)

    def fork_lanes():                                     # the lanes' streams start after everything queued on the current stream ...
        ev = torch.cuda.Event()
        ev.record()
        for _, _, s in lanes[1:]:
            s.wait_event(ev)

    def join_lanes():                                     # ... and the current stream continues after everything queued on them
        for _, _, s in lanes[1:]:
            ev = torch.cuda.Event()
            ev.record(s)
            torch.cuda.current_stream().wait_event(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fork_lanes()
    for i in range(max(3, args.warmup) * n_lanes):
        step(i)
    join_lanes()
    if world > 1:                                        # warm the collective too
        D.gather_rows(rows[:n_batches].reshape(-1, E.N_COLS), ids, max_rows=n_max * B)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    time.sleep(0.3 if sampler else 0.0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    fork_lanes()
    for i in range(n_batches):
        step(i, rows[i])
    join_lanes()
    all_rows, all_ids = rows[:n_batches].reshape(-1, E.N_COLS), ids
    if world > 1:                                        # the pass's single collective: all ranks' metric rows
        all_rows, all_ids = D.gather_rows(rows[:n_batches].reshape(-1, E.N_COLS), ids, max_rows=n_max * B)
    e1.record()
    barrier()
    w1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(w0, w1) if sampler else None
    launches_fwd = int(L.lib().sag_last_launch_count(model._h))
    # all ranks; + metrics_kernel (+ 2 sh_rms launches with the maps); torch's own fill / copy kernels are not counted
    launches = n_total * (launches_fwd + 1 + (2 if maps_on else 0))
    if rank == 0:
        assert all_rows.shape == (n_total * B, E.N_COLS), (all_rows.shape, n_total)
    value = WINDOW_S * B * n_total / (ms * 1e-3)
    extra = {}
    if args.config in (4, 5) and rank == 0:               # the rows in the reference's eval-detailed.txt format (eval.py:212-215)
        t0 = time.perf_counter()
        path = os.path.join(tempfile.gettempdir(), 'sag_bench_eval-detailed.txt')
        E.write_eval_detailed(path, ['clip%d %.1f' % (int(c), 0.5 + int(w)) for c, w in all_ids.cpu().tolist()], all_rows)
        extra = {'eval_detailed_rows': int(all_rows.shape[0]), 'eval_detailed_write_ms': 1e3 * (time.perf_counter() - t0),
                 'batches_total': n_total, 'batches_max_per_rank': n_max}

    # ---- e2e: host buffers in, host waveform out, through the public operator API ----
    # SptAudioGen.inference_stream is the driver loop around sess.run (deploy.py:112-148): every step's inputs are
    # copied from pinned host memory and every step's waveform is read back to the host inside the timed region; the
    # copies of neighbouring steps overlap the forward on a second stream.
    e2e_keys = {'audio': 'audio'}
    if 'video' in encoders:
        e2e_keys['video'] = vkey
    if 'flow' in encoders:
        e2e_keys['flow'] = fkey
        if u8:
            e2e_keys['flow_limits'] = 'flow_limits'
    # (at least 100 steps: the loop keeps several batches in flight, and a 20-step sample is mostly pipeline fill and drain)
    n_e2e = min(n_batches, 200) if args.config in (4, 5) else max(args.steps, 100)

    def host_batches(n):
        for i in range(n):
            yield {k: host[i % R][src] for k, src in e2e_keys.items()}

    def run_e2e(n):
        acc = 0.0
        for y in model.inference_stream(host_batches(n), depth=int(os.environ.get('SAG_STREAM_DEPTH', '3')), lanes=n_lanes,
                                        use_graph={'0': False, '2': True}.get(graph_env)):
            acc += float(y[0, 0, 0])                      # the caller consumes each waveform on the host
        return acc

    run_e2e(4 * n_lanes + 3)                             # warm-up: every slot of the loop (pinned / device buffers, graphs) is touched
    barrier()
    for _ in range(int(os.environ.get('SAG_E2E_REPEAT', '0'))):      # development: spread of the e2e sample within one process
        t0 = time.perf_counter()
        run_e2e(n_e2e)
        torch.cuda.synchronize()
        sys.stderr.write('e2e repeat: %.1f audio-s/s\n' % (WINDOW_S * B * n_e2e / (time.perf_counter() - t0)))
    e0.record()
    run_e2e(n_e2e)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([n_e2e], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms2 = float(ms2.item())
    h2d = sum(int(host[0][src].numel() * host[0][src].element_size()) for src in e2e_keys.values())
    e2e = {'value': WINDOW_S * B * int(cnt.item()) / (ms2 * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': int(out_host.numel() * 4), 'ms_per_step': ms2 / max(n_e2e, 1), 'steps': n_e2e,
           'api': 'SptAudioGen.inference_stream(pinned host batches, lanes=%d) -> host (B,4800,3) waveforms; copies overlap compute' % n_lanes}

    # ---- roofline of the dominant kernel family, timed live with CUDA events on the launching stream ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    model.set_option('profile', 1)
    cats = ['conv', 'deconv', 'fc', 'stft', 'istft', 'pointwise', 'mix']
    P = 3
    agg = {c: [0.0, 0.0, 0.0, 0, 0.0] for c in cats}       # ms, useful flops, bytes, launches, issued flops
    table = None
    for i in range(P):
        fwd(devb[i % R])
        torch.cuda.synchronize()
        recs = layer_table(model, L)
        for r in recs:
            a = agg[r['cat']]
            a[0] += r['us'] * 1e-3 / P
            a[1] += r['gflop'] * 1e9 / P
            a[2] += r['mb'] * 1e6 / P
            a[4] += r['gflop_issued'] * 1e9 / P
        for c in cats:
            agg[c][3] = sum(1 for r in recs if r['cat'] == c)
        if table is None:
            table = recs
        else:                                            # keep the fastest of the P passes per record
            for t, r in zip(table, recs):
                t['us'] = min(t['us'], r['us'])
    model.set_option('profile', 0)
    dense_ms = agg['conv'][0] + agg['deconv'][0] + agg['fc'][0]
    dense_fl = agg['conv'][1] + agg['deconv'][1] + agg['fc'][1]
    dense_issued = agg['conv'][4] + agg['deconv'][4] + agg['fc'][4]
    dense_n = agg['conv'][3] + agg['deconv'][3] + agg['fc'][3]
    conv_stack_fl = agg['conv'][1] + agg['deconv'][1]
    peak_tf = peaks.get('bf16_tflops_sustained')
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (measured)'
    if peak_tf is None:
        peak_tf, peak_src = 1400.0, 'fallback (B200_PROFILING.md sustained)'
    ach = dense_fl / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else 0.0
    traffic, traffic_src = None, None
    for name in ('r2_dominant_kernel_traffic.json', 'r1_dominant_kernel_traffic.json'):
        try:                                              # DRAM bytes per launch of the same kernel family, from ncu
            tj = json.load(open(os.path.join(ROOT, 'profiles', name)))
            if encoders == ['audio', 'video'] and B == 32 and precision in ('bf16x3', 'mixed'):
                traffic, traffic_src = tj['traffic_bytes_per_launch'], 'profiles/%s (%s)' % (name, tj.get('source', 'ncu'))
            break
        except Exception:
            continue
    products = {'bf16x3': 3, 'bf16': 1, 'mixed': 3}.get(precision)
    roofline = {'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                'traffic': traffic, 'traffic_source': traffic_src,
                'algorithmic_bytes_per_launch': (agg['conv'][2] + agg['deconv'][2] + agg['fc'][2]) / max(dense_n, 1),
                'kernel': 'gather_gemm_umma_kernel (tcgen05 conv / sub-pixel transposed conv / FC contractions; splitk_reduce included in the time)',
                'launches_per_step': dense_n, 'ms_per_step': dense_ms,
                # useful = each product of the reference graph that can reach the output, once; issued = what the
                # kernel multiplies (zero taps / border cells of the sub-pixel transposed convs, conv1's K padded 147 -> 256)
                'executed_gflop_per_step': dense_fl / 1e9, 'conv_stack_gflop_per_step': conv_stack_fl / 1e9,
                'issued_gflop_per_step': dense_issued / 1e9,
                'mma_products_per_useful_product': products,
                'tensor_pipe_frac_of_peak_counting_every_mma': (ach * products / peak_tf) if products else None,
                'reference_graph_gflop_per_step': conv_gflop_per_window(encoders) * B, 'peak_source': peak_src,
                'breakdown_ms_per_step': {c: round(agg[c][0], 4) for c in cats},
                'hbm_kernels_gbs': {c: (agg[c][2] / (agg[c][0] * 1e-3) / 1e9 if agg[c][0] > 0 else 0.0)
                                    for c in ('stft', 'istft', 'pointwise', 'mix')},
                'hbm_peak_gbs': peaks.get('hbm_gbs')}
    if rank == 0 and table:
        for t in table:
            t['tflops'] = (t['gflop'] / (t['us'] * 1e-6) / 1e3) if t['us'] > 0 and t['gflop'] > 0 else 0.0
        if args.layer_table:
            json.dump({'config': args.config, 'batch': B, 'encoders': encoders, 'precision': precision,
                       'note': 'CUDA-event time of each launch scope of one profiled forward (events between launches add ~2-8 us of '
                               'serialisation per scope; the timed step uses programmatic dependent launch without them)',
                       'layers': table}, open(args.layer_table, 'w'), indent=1)
        top = sorted(table, key=lambda t: -t['us'])[:6]
        roofline['slowest_launches'] = [{k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items()} for t in top]

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': n_max, 'warmup': max(3, args.warmup),
            'ms_per_step': ms / max(n_max, 1), 'higher_is_better': True, 'scaling': 'weak' if args.config in (1, 2, 3) else 'strong',
            'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'bf16': 'bf16', 'bf16x3': 'bf16x3(f32-grade)', 'mixed': 'bf16x3(f32-grade), decoder deconv5-2 bf16'}.get(precision, precision),
            'data': 'synthetic', 'config': workload_config(encoders, B, precision, args, world),
            'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'clocks': clocks}
    line.update(extra)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config in (1, 2, 3):
        line['cpu_baseline'] = cpu_baseline(encoders, B, args.cpu_baseline_seconds)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
