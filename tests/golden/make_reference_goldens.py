"""Golden vectors produced by the REFERENCE'S OWN CODE for the rows of the hot path whose reference implementation is
plain numpy / scipy: run in the build container, where /root/reference is mounted; the resulting
tests/golden/reference_goldens.npz travels with the repo, /root/reference does not.

    python tests/golden/make_reference_goldens.py            # rewrites tests/golden/reference_goldens.npz

The reference is python-2 source, so its modules are loaded through an import hook that applies purely syntactic
python-2 -> 3 shims to the text (print statements, `raise X, msg`, izip, implicit relative imports) and stubs the
third-party packages that are not installed (tensorflow, scikits.audiolab, resampy, ...).  No function body is
re-implemented: every number below comes out of the reference's functions.

What is pinned (and by which reference code):
  a1   derived constants                 model.py:24-60          SptAudioGen.__init__
  a12  Hilbert-envelope distance         myutils.py:109-116      compute_envelope_dist
  a13  mesh, SH matrix, RMS energy maps  distance.py:9-52, common.py:121-178, decoder.py:9-28, position.py:5-38
  a2   framed STFT                       myutils.py:119-147      stft            } the reference's graph-building code run
  a8   inverse STFT                      myutils.py:181-211      istft           } eagerly on a numpy stand-in for the dozen
  a11  STFT distance, LSD, MSE, SNR      model.py:62-154         evaluation_ops  } elementary TF ops it calls (fake_tf)
  f2   clip-edge padding / file offsets  feeder.py:50-105        AudioReader.get (wav decoding stubbed by arrays)
       flow de-quantisation              feeder.py:138-161       FlowReader.get_by_index
  --   train-params.txt parsing          myutils.py:40-85        load_params
"""
import importlib.abc
import importlib.util
import os
import re
import sys
import tempfile
import types
from unittest import mock

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_goldens.npz')
STUBS = ['scikits', 'scikits.audiolab', 'resampy', 'pyemd', 'librosa', 'skimage', 'skimage.io', 'skimage.transform',
         'matplotlib', 'matplotlib.pyplot']


def py2_to_py3(src):
    out = []
    for line in src.splitlines():
        m = re.match(r'^(\s*)print\s+(?!\()(.*)$', line)
        if m and not m.group(2).startswith('='):
            line = '%sprint(%s)' % (m.group(1), m.group(2).rstrip(','))
        line = re.sub(r'^(\s*)raise\s+(\w+)\s*,\s*(.+)$', r'\1raise \2(\3)', line)
        line = line.replace('from itertools import izip', 'izip = zip').replace('.iteritems()', '.items()')
        line = re.sub(r'^from common import', 'from pyutils.ambisonics.common import', line)
        line = re.sub(r'^from decoder import', 'from pyutils.ambisonics.decoder import', line)
        line = re.sub(r'^from scipy.misc import imresize', 'imresize = None', line)
        # python-2 integer division at the three sites of myutils.stft / istft where both operands are ints
        line = line.replace('range(0, wind_size, wind_size / n_overlap)', 'range(0, wind_size, wind_size // n_overlap)')
        line = line.replace('skip = n_freqs / n_overlap', 'skip = n_freqs // n_overlap')
        out.append(line)
    return '\n'.join(out) + '\n'


class RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `myutils`, `feeder`, `model`, `definitions`, `pyutils.*` from /root/reference through py2_to_py3."""
    ROOTS = ('myutils', 'feeder', 'model', 'definitions', 'pyutils')

    def find_spec(self, name, path, target=None):
        if name.split('.')[0] not in self.ROOTS:
            return None
        base = os.path.join(REF, *name.split('.'))
        if os.path.isdir(base):
            return importlib.util.spec_from_loader(name, self, is_package=True)
        if os.path.exists(base + '.py'):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        base = os.path.join(REF, *module.__name__.split('.'))
        if os.path.isdir(base):
            module.__path__ = [base]
            fn = os.path.join(base, '__init__.py')
            if not os.path.exists(fn):
                return
        else:
            fn = base + '.py'
        # pyutils/tflib is the TF layer library: nothing of it is executed here
        if module.__name__.startswith('pyutils.tflib'):
            module.__getattr__ = lambda k: mock.MagicMock()
            module.__path__ = []
            return
        exec(compile(py2_to_py3(open(fn).read()), fn, 'exec'), module.__dict__)


class _T(np.ndarray):
    """numpy array that answers the two TF tensor calls the reference's graph code makes on its inputs."""
    def get_shape(self):
        return _Shape(self.shape)

    def __getitem__(self, idx):                           # tensor[i] stays a (0-d) tensor
        r = np.ndarray.__getitem__(self, idx)
        return r if isinstance(r, np.ndarray) else np.asarray(r).view(_T)


class _Shape(tuple):
    def as_list(self):
        return [int(v) for v in self]


def _t(a):
    return np.asarray(a).view(_T)


def fake_tf():
    """A dozen elementary TF-1 ops evaluated eagerly with numpy, enough to run the reference's myutils.stft / istft /
    stft_for_loss and model.evaluation_ops graph-building code verbatim on arrays.  (tf.fft / tf.ifft transform the last
    axis; complex64 results are rounded from numpy's double-precision transform.)"""
    import contextlib
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.complex64 = np.float32, np.complex64
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), [int(v) for v in shape]))
    tf.stack = lambda xs, axis=0: _t(np.stack([np.asarray(x) for x in xs], axis))
    tf.concat = lambda xs, axis: _t(np.concatenate([np.asarray(x) for x in xs], axis))
    tf.unstack = lambda x, axis=0: [_t(v) for v in np.moveaxis(np.asarray(x), axis, 0)]
    tf.constant = lambda v, dtype=None: _t(np.asarray(v, dtype))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(np.asarray(x), axis))
    tf.cast = lambda x, dtype: _t(np.asarray(x).astype(dtype))
    tf.transpose = lambda x, perm: _t(np.transpose(np.asarray(x), perm))
    tf.fft = lambda x: _t(np.fft.fft(np.asarray(x), axis=-1).astype(np.complex64))
    tf.ifft = lambda x: _t(np.fft.ifft(np.asarray(x), axis=-1).astype(np.complex64))
    tf.real = lambda x: _t(np.real(np.asarray(x)))
    tf.abs = lambda x: _t(np.abs(np.asarray(x)))
    tf.log = lambda x: _t(np.log(np.asarray(x, dtype=np.float32) if np.isscalar(x) else np.asarray(x)))
    tf.sqrt = lambda x: _t(np.sqrt(np.asarray(x)))
    tf.maximum = lambda a, b: _t(np.maximum(np.asarray(a), b))
    tf.ones = lambda shape: _t(np.ones([int(v) for v in shape], np.float32))
    tf.add_n = lambda xs: _t(sum(np.asarray(x) for x in xs))
    tf.reduce_mean = lambda x, axis=None: _t(np.mean(np.asarray(x), axis=axis))
    tf.reduce_sum = lambda x, axis=None: _t(np.sum(np.asarray(x), axis=axis))
    tf.variable_scope = lambda *a, **k: contextlib.nullcontext()
    return tf


def install():
    sys.modules['tensorflow'] = fake_tf()
    for s in STUBS:
        if s not in sys.modules:
            sys.modules[s] = mock.MagicMock()
    sys.meta_path.insert(0, RefFinder())


def main():
    assert os.path.isdir(REF), 'run this where /root/reference is mounted'
    install()
    G = {}
    rng = np.random.RandomState(20261017)

    # ---- a1: derived constants (model.py:24-60) ----------------------------------------------------------------
    import model as Rm
    cfgs = [(1, 48000, 10, 1.0, 0.1, 0.025), (1, 44100, 10, 1.0, 0.1, 0.025), (1, 48000, 10, 0.5, 0.2, 0.05), (2, 16000, 5, 2.0, 0.2, 0.016)]
    rows = []
    for order, ar, vr, ctx, dur, win in cfgs:
        m = Rm.SptAudioGen(order, audio_rate=ar, video_rate=vr, context=ctx, sample_duration=dur, encoders=['audio'],
                           separation='none', params=Rm.SptAudioGenParams(sep_fft_window=win))
        rows.append([order, ar, vr, ctx, dur, win, m.num_ambi_channels, m.snd_contx, m.snd_dur, m.snd_size, m.wind_size])
    G['a1_configs_and_dims'] = np.asarray(rows, np.float64)

    # ---- a13: mesh, SH matrix, RMS maps (distance.py, common.py, decoder.py, position.py) ---------------------------
    from pyutils.ambisonics import distance as Rd, common as Rc, position as Rp
    for res in (30, 5):
        phi, nu = Rd.spherical_mesh(res)
        G['a13_phi_mesh_%d' % res], G['a13_nu_mesh_%d' % res] = phi, nu
    phi, nu = Rd.spherical_mesh(30)
    pos = [Rp.Position(p, n, 1., 'polar') for p, n in zip(phi.reshape(-1), nu.reshape(-1))]
    G['a13_sh_matrix_30'] = Rc.spherical_harmonics_matrix(pos, 1)
    ambi = (rng.randn(4800, 4) * np.array([0.2, 0.1, 0.05, 0.15])).astype(np.float32)
    G['a13_ambi'] = ambi
    for res in (30, 5):
        vis = Rd.SphericalAmbisonicsVisualizer(ambi.astype(np.float64), 48000, window=0.1, angular_res=float(res))
        G['a13_rms_map_%d' % res] = vis.get_next_frame()
    masked = ambi.astype(np.float64) * np.array([1., 1., 0., 1.])                  # a WXY clip (feeder.py:312-314, eval.py:147-148)
    G['a13_rms_map_30_wxy'] = Rd.SphericalAmbisonicsVisualizer(masked, 48000, window=0.1, angular_res=30.).get_next_frame()

    # ---- a12: Hilbert-envelope distance (myutils.py:109-116) -------------------------------------------------------
    import myutils as Ru
    gt = (rng.randn(4800, 3) * 0.1).astype(np.float32)
    pred = (gt + rng.randn(4800, 3) * 0.03).astype(np.float32)
    G['a12_gt'], G['a12_pred'] = gt, pred
    G['a12_env_dist'] = Ru.compute_envelope_dist(pred, gt)

    # ---- load_params (myutils.py:40-85) --------------------------------------------------------------------------------
    d = tempfile.mkdtemp()
    text = ("encoders: ['audio', 'video']\nseparation: UNET_MASK\nambi_order: 1\naudio_rate: 48000\nvideo_rate: 10\ncontext: 1.0\n"
            "sample_dur: 0.1\nlr: 0.0001\nn_iters: 100000\nbatch_size: 32\nlr_decay: 0.5\nlr_iters: 30000\nloc_units: [512, 512]\n")
    open(os.path.join(d, 'train-params.txt'), 'w').write(text)
    p = Ru.load_params(d)
    G['params_text'] = np.asarray(text)
    keys = sorted(k for k in vars(p) if isinstance(getattr(p, k), (int, float, str, list)))
    G['params_repr'] = np.asarray(repr([(k, getattr(p, k)) for k in keys]))

    # ---- f2: AudioReader.get / FlowReader.get_by_index (feeder.py:50-161) -----------------------------------------------
    import feeder as Rf
    rate, nfiles, nch = 200, 4, 4                                                # tiny clip: 4 files of 1 s at 200 Hz
    clip = np.round(rng.uniform(-0.5, 0.5, size=(nfiles * rate, nch)) * 1024) / 1024
    Rf.load_wav = lambda fn, r=None: (clip[int(os.path.basename(fn)[:6]) * rate:(int(os.path.basename(fn)[:6]) + 1) * rate], rate)
    ar = object.__new__(Rf.AudioReader)
    ar.audio_folder, ar.num_files, ar.rate, ar.num_channels, ar.duration, ar.num_frames = '/x', nfiles, rate, nch, nfiles, nfiles * rate
    G['f2_clip'] = clip
    cases = [(0.0, 219), (-0.25, 219), (1.37, 219), (3.0, 219), (3.6, 219), (0.5, 120), (2.995, 30)]
    G['f2_audio_cases'] = np.asarray(cases, np.float64)
    for i, (t0, size) in enumerate(cases):
        G['f2_audio_out_%d' % i] = ar.get(t0, size)
    G['f2_audio_rot'] = ar.get(0.5, 64, rotation=0.7)
    fr = object.__new__(Rf.FlowReader)
    raw = rng.randint(0, 256, size=(2, 6, 8, 3)).astype(np.uint8)

    class FakeReader(object):
        rate = 10.

        def get_by_index(self, start_time, size, rotation=None):
            return raw.copy()
    fr.reader, fr.rate = FakeReader(), 10.
    fr.lims = np.stack([np.linspace(0.5, 1.5, 40), np.linspace(10., 30., 40)], 1)
    G['f2_flow_raw'], G['f2_flow_lims'] = raw, fr.lims
    G['f2_flow_out'] = fr.get_by_index(1.3, 2)

    # ---- a2 / a8 / a11: the reference's own graph code for stft / istft / evaluation_ops, run eagerly (fake_tf) --------
    x = np.round(np.clip(0.1 * rng.randn(1, 1, 52799) + 0.3 * np.sin(2 * np.pi * 440 * np.arange(52799) / 48000.), -1, 1) * 4096) / 4096
    G['a2_audio_q12'] = np.round(x * 4096).astype(np.int16)                       # exactly representable input
    s = np.asarray(Ru.stft(_t(x.astype(np.float32)), 1024, 4))
    assert s.shape == (1, 1, 200, 1024)
    G['a2_stft_bins_stride37'] = s[0, 0, :, ::37]                                   # all 200 frames, every 37th bin
    G['a2_stft_abs_sum_per_frame'] = np.abs(s[0, 0]).sum(-1)
    xs = np.round(rng.randn(2, 3, 1000) * 256) / 256
    G['a2_small_in'] = xs.astype(np.float32)
    G['a2_small_stft'] = np.asarray(Ru.stft(_t(xs.astype(np.float32)), 64, 4))
    z = (rng.randn(2, 30, 64) + 1j * rng.randn(2, 30, 64)).astype(np.complex64)
    G['a8_small_in'] = z
    G['a8_small_istft'] = np.asarray(Ru.istft(_t(z), 4))
    G['a8_istft_of_stft_frames_89_117'] = np.asarray(Ru.istft(_t(s[0, :, 89:117]), 4))   # (1, 6400): 0.5 * x[23552:29952]
    gt2 = np.round(rng.randn(2, 4800, 3) * 0.1 * 4096) / 4096
    pr2 = np.round((gt2 + rng.randn(2, 4800, 3) * 0.03) * 4096) / 4096
    G['a11_gt_q12'], G['a11_pred_q12'] = np.round(gt2 * 4096).astype(np.int16), np.round(pr2 * 4096).astype(np.int16)
    mask = np.array([[1., 1., 1.], [1., 0., 1.]], np.float32)
    G['a11_mask'] = mask
    ref_model = Rm.SptAudioGen(1, encoders=['audio'], separation='none')
    metrics, stft_ps, lsd_ps, mse_ps, snr_ps = ref_model.evaluation_ops(_t(pr2.astype(np.float32)), _t(gt2.astype(np.float32)), None, _t(mask))
    G['a11_stft_ps'], G['a11_lsd_ps'], G['a11_mse_ps'], G['a11_snr_ps'] = [np.asarray(v) for v in (stft_ps, lsd_ps, mse_ps, snr_ps)]
    G['a11_metric_names'] = np.asarray(repr(list(metrics.keys())))
    G['a11_metric_values'] = np.asarray([float(np.asarray(v)) for v in metrics.values()])

    np.savez_compressed(OUT, **G)
    print('wrote %s: %d arrays, %d bytes' % (OUT, len(G), os.path.getsize(OUT)))


if __name__ == '__main__':
    main()
