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
       EMD columns (emd/dir, emd/dir2)   distance.py:100-143     emd / ambix_emd around a pyemd stand-in (its defining LP)
  a2   framed STFT                       myutils.py:119-147      stft            } the reference's graph-building code run
  a8   inverse STFT                      myutils.py:181-211      istft           } eagerly on a numpy stand-in for the dozen
  a11  STFT distance, LSD, MSE, SNR      model.py:62-154         evaluation_ops  } elementary TF ops it calls (fake_tf)
  f2   clip-edge padding / file offsets  feeder.py:50-105        AudioReader.get (wav decoding stubbed by arrays)
       flow de-quantisation              feeder.py:138-161       FlowReader.get_by_index
       frame indexing, rotation roll     feeder.py:106-132       VideoReader.get_by_index (jpg decoding stubbed by arrays)
       chunk schedule + reader requests  feeder.py:164-278       SampleReader (file readers recorded)
  --   train-params.txt parsing          myutils.py:40-85        load_params
  a3-a7, a9, a10  encoders, bottleneck,  model.py:161-434, pyutils/tflib/wrappers/core.py:10-220,
       localisation, mask decoder, mix   pyutils/tflib/models/image/resnet.py:110-249   (SptAudioGen.inference_ops)
       -- the reference's model-building code (scopes, variable names and shapes, layer order, strides, paddings, crops,
       concat / tile / reshape order, mask and mixing arithmetic) executed eagerly: tf.variable_scope / model_variable
       serve the weights BY THE NAME the reference's code asks for, and the TF kernels it calls (nn.convolution,
       conv2d_transpose, max_pool, contrib batch_norm, matmul ...) are evaluated by fake_tf_graph below in float64 from
       TensorFlow-1.4's documented semantics.  Those op semantics are restated here, not executed from TensorFlow;
       everything above them is the reference's own code.  restore_pretrained (resnet.py:238-249) is run against the
       reference's resnet18.npy, which pins the tower's variable names and shapes.
  f4   overlay maps, stereo down-mix     myutils.py:224-311      gen_360video (ffmpeg, video files, colour map and resize recorded)
  --   deploy loop                       deploy.py:90-152        W2XYZ.deploy (batches of 10, zero-padded tail, mono crop,
       row layout) around the model code above; the disk reader and tf.Session are the only stand-ins
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
        # model.py: python-2 list-returning zip and integer division; core.py: TF dtype attribute; resnet.py: numpy>=1.17
        if 'in reversed(zip(' in line and line.rstrip().endswith(')):'):
            line = line.replace('in reversed(zip(', 'in reversed(list(zip(').rstrip()[:-1] + '):'
        line = line.replace('self.snd_dur/sz[1]', 'self.snd_dur//sz[1]')
        line = re.sub(r'ss = self\.(model\.)?snd_contx / 2$', r'ss = self.\1snd_contx // 2', line)
        line = line.replace('x.dtype.base_dtype', 'x.dtype')
        line = line.replace("np.load(os.path.join(PWD, 'resnet18.npy')).all()",
                            "np.load(os.path.join(PWD, 'resnet18.npy'), allow_pickle=True, encoding='latin1').item()")
        out.append(line)
    return '\n'.join(out) + '\n'


class RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `myutils`, `feeder`, `model`, `definitions`, `pyutils.*` from /root/reference through py2_to_py3."""
    ROOTS = ('myutils', 'feeder', 'model', 'definitions', 'pyutils', 'deploy')

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
        # pyutils/tflib is the TF layer library: only wrappers/core.py and models/image/resnet.py are executed (the
        # package __init__ files pull in the trainer, recurrent layers, ...: skipped)
        name = module.__name__
        if name.startswith('pyutils.tflib') and name not in ('pyutils.tflib.wrappers.core', 'pyutils.tflib.models.image.resnet'):
            if name == 'pyutils.tflib.wrappers':                  # wrappers/__init__.py: `from core import fully_connected, ...`
                core = importlib.import_module('pyutils.tflib.wrappers.core')
                for k in ('fully_connected', 'conv_2d', 'deconv_2d', 'conv_1d'):
                    setattr(module, k, getattr(core, k))
            elif os.path.isdir(base):
                pass                                              # plain namespace: submodules import through the finder
            else:
                module.__getattr__ = lambda k: mock.MagicMock()   # e.g. models/image/preprocessing.py
            return
        module.__file__ = fn
        exec(compile(py2_to_py3(open(fn).read()), fn, 'exec'), module.__dict__)


class _T(np.ndarray):
    """numpy array that answers the two TF tensor calls the reference's graph code makes on its inputs."""
    def get_shape(self):
        return _Shape(self.shape)

    def __getitem__(self, idx):                           # tensor[i] stays a (0-d) tensor
        r = np.ndarray.__getitem__(self, idx)
        return r if isinstance(r, np.ndarray) else np.asarray(r).view(_T)


class _Shape(tuple):
    ndims = property(len)

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


def _same_pads(n, k, s):
    # TF 'SAME': output ceil(n / s); total padding max((out - 1) * s + k - n, 0), the odd unit goes after
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def _windows(x, kh, kw, sh, sw, padding, fill):
    if padding == 'SAME':
        ph, pw = _same_pads(x.shape[1], kh, sh), _same_pads(x.shape[2], kw, sw)
        x = np.pad(x, ((0, 0), ph, pw, (0, 0)), constant_values=fill)
    else:
        assert padding == 'VALID'
    oh, ow = (x.shape[1] - kh) // sh + 1, (x.shape[2] - kw) // sw + 1
    for a in range(kh):
        for b in range(kw):
            yield a, b, x[:, a:a + (oh - 1) * sh + 1:sh, b:b + (ow - 1) * sw + 1:sw, :]


def fake_tf_graph(tf, weights, created):
    """The variable / layer half of the stand-in: enough of TF-1.4 to run the reference's model.inference_ops,
    wrappers/core.py and resnet.py verbatim.  Variables are looked up in `weights` by the full scoped name the reference's
    code produces (KeyError / shape assertion if it asks for anything else) and logged in `created`; the NHWC kernels are
    tap-by-tap float64 numpy loops (deliberately not the oracle's torch calls)."""
    import contextlib
    scope = []

    class Scope(object):
        def __init__(self, name):
            self.name, self.original_name_scope = name, name + '/'

    @contextlib.contextmanager
    def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None):
        scope.append(name_or_scope if name_or_scope is not None else default_name)
        try:
            yield Scope('/'.join(scope))
        finally:
            scope.pop()

    class Var(object):                                                # what tf.get_collection hands restore_pretrained
        def __init__(self, name, shape):
            self.op, self.shape = types.SimpleNamespace(name=name), tuple(shape)

    def model_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True):
        full = '/'.join(scope + [name])
        arr = np.asarray(weights[full], np.float64)
        assert shape is None or tuple(int(v) for v in shape) == arr.shape, (full, shape, arr.shape)
        created.append(Var(full, arr.shape))
        return _t(arr)

    def convolution(input, filter, strides=None, dilation_rate=None, padding='VALID'):
        assert dilation_rate is None
        x, w = np.asarray(input, np.float64), np.asarray(filter, np.float64)
        kh, kw = w.shape[:2]
        out = 0.
        for a, b, win in _windows(x, kh, kw, int(strides[0]), int(strides[1]), padding, 0.):
            out = out + win @ w[a, b]
        return _t(out)

    def conv2d_transpose(value, filter, output_shape, strides, padding='SAME'):
        assert padding == 'VALID'                                      # filter: (kh, kw, out, in)
        x, w = np.asarray(value, np.float64), np.asarray(filter, np.float64)
        sh, sw = int(strides[1]), int(strides[2])
        out = np.zeros([int(v) for v in output_shape])
        n, h, wd = x.shape[:3]
        assert out.shape == (n, (h - 1) * sh + w.shape[0], (wd - 1) * sw + w.shape[1], w.shape[2])
        for a in range(w.shape[0]):
            for b in range(w.shape[1]):
                out[:, a:a + (h - 1) * sh + 1:sh, b:b + (wd - 1) * sw + 1:sw, :] += x @ w[a, b].T
        return _t(out)

    def max_pool(value, ksize, strides, padding):
        x, out = np.asarray(value, np.float64), None
        for a, b, win in _windows(x, ksize[1], ksize[2], strides[1], strides[2], padding, -np.inf):
            out = win.copy() if out is None else np.maximum(out, win)
        return _t(out)

    def batch_norm(x, decay=0.999, center=True, scale=False, epsilon=0.001, param_initializers=None, is_training=True,
                   trainable=True, reuse=None, scope=None):
        # tf.contrib.layers.batch_norm: variables beta, gamma, moving_mean, moving_variance under `scope`; with
        # is_training the batch moments (biased variance) over every axis but the last normalise the activations.
        with variable_scope(scope, 'BatchNorm'):
            c = [int(np.asarray(x).shape[-1])]
            beta = model_variable('beta', c) if center else 0.
            gamma = model_variable('gamma', c) if scale else 1.
            mean, var = model_variable('moving_mean', c), model_variable('moving_variance', c)
        x = np.asarray(x, np.float64)
        if is_training:
            ax = tuple(range(x.ndim - 1))
            mean, var = x.mean(ax), x.var(ax)
        return _t((x - np.asarray(mean)) / np.sqrt(np.asarray(var) + epsilon) * np.asarray(gamma) + np.asarray(beta))

    tf.variable_scope = variable_scope
    tf.nn = types.SimpleNamespace(relu=lambda x: _t(np.maximum(np.asarray(x), 0.)), bias_add=lambda x, b: _t(np.asarray(x) + np.asarray(b)),
                                  convolution=convolution, conv2d_transpose=conv2d_transpose, max_pool=max_pool)
    tf.matmul = lambda a, b: _t(np.asarray(a) @ np.asarray(b))
    tf.tile = lambda x, m: _t(np.tile(np.asarray(x), [int(v) for v in m]))
    tf.sigmoid = lambda x: _t(1. / (1. + np.exp(-np.asarray(x))))
    tf.identity = lambda x: x
    tf.truncated_normal_initializer = tf.constant_initializer = lambda *a, **k: None
    tf.contrib, tf.GraphKeys, tf.__path__ = mock.MagicMock(), mock.MagicMock(), []
    tf.get_collection = lambda key: list(created)
    assigned = []

    def assign(var, value):
        assert tuple(np.shape(value)) == var.shape, (var.op.name, np.shape(value), var.shape)
        assigned.append(var.op.name)
        return var.op.name
    tf.assign, tf.assigned = assign, assigned
    mods = {'tensorflow.contrib': tf.contrib, 'tensorflow.contrib.framework': None, 'tensorflow.contrib.framework.python': None,
            'tensorflow.contrib.framework.python.ops': types.SimpleNamespace(variables=types.SimpleNamespace(model_variable=model_variable)),
            'tensorflow.contrib.layers.python': None,
            'tensorflow.contrib.layers.python.layers': types.SimpleNamespace(utils=types.SimpleNamespace(
                collect_named_outputs=lambda coll, alias, x: x, last_dimension=lambda shape, min_rank=1: int(shape[-1]),
                n_positive_integers=lambda n, v: (int(v),) * n if np.isscalar(v) else tuple(int(i) for i in v))),
            'tensorflow.contrib.layers': types.SimpleNamespace(l2_regularizer=lambda s: None, batch_norm=batch_norm)}
    for k, v in mods.items():
        m = types.ModuleType(k)
        m.__dict__.update(vars(v) if isinstance(v, types.SimpleNamespace) else {})
        m.__path__ = []
        sys.modules[k] = m
    sys.modules['tensorflow.contrib'].layers = sys.modules['tensorflow.contrib.layers']
    return tf


WEIGHTS, CREATED = {}, []


def install():
    sys.modules['tensorflow'] = fake_tf_graph(fake_tf(), WEIGHTS, CREATED)
    for s in STUBS:
        if s not in sys.modules:
            sys.modules[s] = mock.MagicMock()
    sys.meta_path.insert(0, RefFinder())


MODEL_SEED = 7


def model_inputs(seed, batch):
    """Seeded inputs of the full-size model case (the test regenerates them from the same recipe)."""
    r = np.random.RandomState(seed)
    n = 52799
    tone = 0.2 * np.sin(2 * np.pi * 523.25 * np.arange(n) / 48000.)[None, :, None]
    audio = (np.round(np.clip(0.1 * r.randn(batch, n, 1) + tone, -1, 1) * 4096) / 4096).astype(np.float32)
    video = (r.randint(0, 256, size=(batch, 1, 224, 448, 3)) / 255.).astype(np.float32)
    flow = (r.randint(0, 256, size=(batch, 1, 224, 448, 3)) / 255. - 0.5).astype(np.float32)
    return audio, video, flow


def f4_clip():
    """3 s of first-order ambisonics whose dominant direction moves (the test regenerates it from the same recipe)."""
    clip = (np.random.RandomState(11).randn(3 * 48000, 4) * np.array([0.2, 0.1, 0.03, 0.15])).astype(np.float32)
    clip[48000:96000, 1] *= 3.
    return clip


DEPLOY_WINDOWS = 11


def deploy_inputs(seed, n):
    """n consecutive windows of a 4-channel clip + video frames (the test regenerates them from the same recipe)."""
    r = np.random.RandomState(seed)
    amb = np.round(np.clip(0.1 * r.randn(n, 52799, 4), -1, 1) * 4096) / 4096
    vid = (r.randint(0, 256, size=(n, 1, 224, 448, 3)) / 255.).astype(np.float32)
    return amb, vid


def torch_f64():
    import torch
    return torch.float64


def tf_assigned():
    out = list(sys.modules['tensorflow'].assigned)
    del sys.modules['tensorflow'].assigned[:]
    return out


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

    # a13 EMD columns (distance.py:100-143 `emd` / `ambix_emd`): the reference's wrapper (ground distance, the two mass
    # normalisations, frame loop) around a pyemd stand-in that solves pyemd.emd's defining EMD-hat LP with scipy / HiGHS
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import sag_oracle as O
    sys.modules['pyemd'] = types.SimpleNamespace(emd=lambda a, b, d: O.emd_hat_lp(a, b, d))
    ambi2 = (0.8 * ambi + np.random.RandomState(5).randn(4800, 4) * np.array([0.02, 0.06, 0.05, 0.03])).astype(np.float32)
    G['a13_ambi2'] = ambi2
    G['a13_emd_dir_dir2'] = np.asarray(Rd.ambix_emd(ambi2.astype(np.float64), ambi.astype(np.float64), 48000, ang_res=30))

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

    # ---- f2: VideoReader.get_by_index (feeder.py:106-132): frame indexing, img_prep, rotation roll; jpg decoding stood in ----
    vfolder = tempfile.mkdtemp()
    vframes = np.random.RandomState(4).randint(0, 256, size=(30, 6, 16, 3)).astype(np.uint8)
    for i in range(30):
        open(os.path.join(vfolder, '%06d.jpg' % i), 'w').close()
    Rf.sio = types.SimpleNamespace(imread=lambda fn: vframes[int(os.path.basename(fn)[:6])])
    vr = Rf.VideoReader(vfolder, 10, Ru.img_prep_fcn())
    G['f2_video_frames'] = vframes
    vcases = [(0.0, 1, None), (1.26, 3, None), (-0.3, 2, None), (2.0, 1, 0.7), (0.5, 2, -2.0), (1.0, 1, 3.1)]
    G['f2_video_cases'] = np.asarray(repr(vcases))
    G['f2_video_meta'] = np.asarray([vr.num_frames, vr.duration, vr.rate] + list(vr.frame_shape), np.float64)
    for i, (t0, size, rot) in enumerate(vcases):
        G['f2_video_out_%d' % i] = vr.get_by_index(t0, size, rot)

    # ---- f2: SampleReader's chunk schedule and the reader calls it makes (feeder.py:164-278), file readers recorded -----------
    calls = []

    class FakeReader(object):
        frame_shape = (224, 448, 3)

        def __init__(self, *a, **k):
            calls.append(('init', type(self).__name__) + tuple(os.path.basename(str(v)) if isinstance(v, str) else v for v in a[:3] if not callable(v)))

        def get(self, start, size, rotation=None):
            calls.append(('get', float(start), int(size), rotation))
            return np.zeros((size, 4))

        def get_by_index(self, start, size, rotation=None):
            calls.append((type(self).__name__, float(start), int(size), rotation))
            return np.zeros((size, 2, 2, 3))
    Rf.AudioReader, Rf.VideoReader, Rf.FlowReader = (type(n, (FakeReader,), {}) for n in ('AudioReader', 'VideoReader', 'FlowReader'))
    folder = os.path.join(tempfile.mkdtemp(), 'clip_xyz')
    os.makedirs(folder)
    sched = ''.join('%.2f %.6f\n' % (0.55 + 0.1 * i, p) for i, p in enumerate(np.random.RandomState(3).uniform(0, 0.02, 60)))
    open(os.path.join(folder, 'audio_pow.lst'), 'w').write(sched)
    G['f2_sched_audio_pow_lst'] = np.asarray(sched)
    cases = [dict(shuffle=False, random_rotations=False), dict(shuffle=False, random_rotations=False, skip_rate=10),
             dict(shuffle=False, random_rotations=False, skip_silence_thr=0.01, return_flow=True),
             dict(shuffle=False, random_rotations=False, start_time=2.0, sample_duration=1.5, skip_silence_thr=None, skip_rate=None),
             dict(shuffle=False, random_rotations=False, start_time=0.3, sample_duration=1.0),
             dict(shuffle=False, random_rotations=False, num_threads=4, thread_id=2, return_video=False),
             dict(shuffle=False, random_rotations=False, audio_rate=16000, video_rate=5, context=2.0, duration=0.2, skip_rate=7)]
    out = []
    for kw in cases:
        del calls[:]
        sr = Rf.SampleReader(folder, **kw)
        got = [sr.get() for _ in range(3)]
        out.append(dict(kwargs=kw, chunks_t=list(sr.chunks_t), audio_size=sr.audio_size, video_size=sr.video_size,
                        ids=[c['id'] if c else None for c in got], calls=list(calls)))
    G['f2_sched_cases'] = np.asarray(repr(out))

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

    # ---- f4: gen_360video's overlay / down-mix arithmetic (myutils.py:224-311); ffmpeg, the video files, the colour map and
    # the image resize are stood in by recorders, everything between them is the reference's code ------------------------
    import pyutils.iolib.audio as Ria
    import pyutils.iolib.video as Riv
    clip = f4_clip()
    rec = {'alpha': [], 'index': [], 'stereo': None}

    class FakeVideoReader(object):
        def __init__(self, fn, rate=None):
            self.fps, self.frame_shape, self.left = 10, (37, 72, 3), 23

        def get(self):
            self.left -= 1
            return np.zeros(self.frame_shape, np.uint8) if self.left >= 0 else None

    def fake_resize(img, shape):
        rec['alpha' if img.shape[-1] == 1 else 'index'].append(np.array(img))
        return np.zeros(tuple(shape) + (img.shape[-1],))
    Riv.VideoReader, Riv.VideoWriter = FakeVideoReader, lambda fn, fps: types.SimpleNamespace(write_frame=lambda f: None)
    Ria.load_wav = lambda fn, rate=None: (clip.astype(np.float64), 48000)
    Ria.save_wav = lambda fn, data, rate: rec.__setitem__('stereo', np.array(data))
    sys.modules['matplotlib'].pyplot.cm.YlOrRd = lambda v: np.stack([v, v, v, v], 1)      # colour = index / 255
    sys.modules['skimage.transform'].resize = fake_resize
    real_os = Ru.os
    Ru.os = types.SimpleNamespace(system=lambda cmd: 0, remove=lambda fn: None, chdir=lambda d: None, getcwd=real_os.getcwd, path=real_os.path)
    try:
        Ru.gen_360video('a.wav', 'v.mp4', 'out.mp4', inject_meta=True, overlay_map=True, binauralize=True)
    finally:
        Ru.os = real_os
    assert len(rec['alpha']) == len(rec['index']) == 23
    G['f4_rms_frames'] = np.stack(rec['alpha'], 0)[..., 0].astype(np.float32)             # the clipped 2*rms - 0.7 maps
    G['f4_colour_index'] = np.round(np.stack(rec['index'], 0)[..., 0] * 255).astype(np.uint8)
    G['f4_stereo_stride97'] = rec['stereo'][::97]

    # ---- a3-a7, a9, a10: the reference's model-building code, run eagerly (fake_tf_graph) -----------------------------------
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from spatialaudiogen_b200 import weights as PW
    from oracle import sag_oracle as O
    from pyutils.tflib.models.image import resnet as Rr
    pre = np.load(os.path.join(REF, 'pyutils/tflib/models/image/resnet18.npy'), allow_pickle=True, encoding='latin1').item()
    for tag, encoders in (('a', ['audio']), ('avf', ['audio', 'video', 'flow'])):
        W = PW.init_weights(encoders, 'unet_mask', seed=MODEL_SEED, stress=True)
        WEIGHTS.clear(), WEIGHTS.update(W)
        del CREATED[:]
        audio, video, flow = model_inputs(MODEL_SEED, 2)
        ref_model = Rm.SptAudioGen(1, encoders=list(encoders), separation='unet_mask', params=Rm.SptAudioGenParams())
        kw = dict(video=_t(video.astype(np.float64)), flow=_t(flow.astype(np.float64))) if 'video' in encoders else {}
        ambix = np.asarray(ref_model.inference_ops(_t(audio.astype(np.float64)), is_training=False, **kw))
        names = [(v.op.name, v.shape) for v in CREATED]
        assert sorted(n for n, _ in names) == sorted(W.keys()), set(W.keys()) ^ set(n for n, _ in names)
        assert all(W[n].shape == s for n, s in names)
        G['m_%s_variables' % tag] = np.asarray(repr(names))
        G['m_%s_weight_checksum' % tag] = np.asarray([float(np.asarray(v, np.float64).sum()) for v in W.values()])
        G['m_%s_ambix' % tag] = ambix.astype(np.float32)
        G['m_%s_sep_stride97' % tag] = np.asarray(ref_model.sep_channels)[:, 0, :, ::97].astype(np.float32)
        G['m_%s_loc_w_stride480' % tag] = np.asarray(ref_model.loc_channels[0])[:, ::480].astype(np.float32)
        G['m_%s_loc_b_stride480' % tag] = np.asarray(ref_model.loc_channels[1])[:, ::480].astype(np.float32)
        G['m_%s_inp_spect_sum' % tag] = np.asarray(ref_model.inp_spect).sum(-1).astype(np.float32)
        if 'video' in encoders:
            for k in ('video', 'flow'):
                G['m_%s_%s_conv1_stride' % (tag, k)] = np.asarray(ref_model.ends[k + '_encoder/conv'])[:, ::8, ::8, ::4].astype(np.float32)
                G['m_%s_%s_conv5_2_stride16' % (tag, k)] = np.asarray(ref_model.ends[k + '_encoder/conv5_2'])[..., ::16].astype(np.float32)
            # visual_encoding_ops ran restore_pretrained against the reference's resnet18.npy (resnet.py:238-249):
            tower = sorted(n[len('video_encoder/'):] for n, _ in names if n.startswith('video_encoder/'))
            assert sorted(tf_assigned()) == sorted(n for n, _ in names if n.split('/')[0] in ('video_encoder', 'flow_encoder'))
            assert set(tower) <= set(pre.keys()) and all(tuple(pre[k].shape) == W['video_encoder/' + k].shape for k in tower)
            G['m_resnet18_npy_keys_used'] = np.asarray(repr(tower))
        # generation-time cross-check of the oracle (float64) on the same inputs
        om = O.SptAudioGen(W, 1, encoders=list(encoders), separation='unet_mask', dtype=torch_f64())
        oa = om.inference_ops(audio, *( [video, flow] if 'video' in encoders else [])).numpy()
        err = np.abs(oa - ambix).max() / np.abs(ambix).max()
        print('model[%s]: %d variables, |ambix| max %.3g, oracle(float64) vs reference-code rel err %.2e' % (tag, len(names), np.abs(ambix).max(), err))
        assert err < 1e-4

    # ---- deploy.py:90-152: the reference's deploy loop (batches of 10, zero-padded tail, mono crop, row layout) around
    # its own model code; the disk reader and the session are the only stand-ins -----------------------------------------
    import deploy as Rdep
    enc = ['audio', 'video']
    W = PW.init_weights(enc, 'unet_mask', seed=MODEL_SEED + 1, stress=True)
    WEIGHTS.clear(), WEIGHTS.update(W)
    amb, vid = deploy_inputs(MODEL_SEED + 1, DEPLOY_WINDOWS)
    chunks = [{'id': 'clip', 'ambix': amb[i], 'video': vid[i]} for i in range(DEPLOY_WINDOWS)]

    class FakeSampleReader(object):
        def __init__(self, folder, **kw):
            assert kw['return_video'] and not kw['return_flow'] and not kw['shuffle'] and kw['duration'] == 0.1
            self.chunks_t, self.queue = [kw['start_time'] + 0.05 + 0.1 * i for i in range(len(chunks))], list(chunks)

        def get(self):
            return self.queue.pop(0) if self.queue else None
    Rdep.SampleReader = FakeSampleReader
    dep = object.__new__(Rdep.W2XYZ)                                   # __init__ builds placeholders / Saver / Session
    dep.params = types.SimpleNamespace(ambi_order=1, audio_rate=48000, video_rate=10, context=1.0, encoders=enc)
    dep.duration, dep.batch_size = 0.1, 10                              # deploy.py:49-50
    dep.model = Rm.SptAudioGen(ambi_order=1, audio_rate=48000, video_rate=10, context=1.0, sample_duration=0.1, encoders=list(enc),
                               separation='unet_mask', params=Rm.SptAudioGenParams())
    dep.tba = {'audio': 'ph_audio', 'video': 'ph_video'}
    batches = []

    def sess_run(fetch, feed_dict):
        batches.append(int(feed_dict['ph_audio'].shape[0]))
        return np.asarray(dep.model.inference_ops(is_training=False, audio=_t(np.asarray(feed_dict['ph_audio'], np.float64)),
                                                  video=_t(np.asarray(feed_dict['ph_video'], np.float64))))
    dep.sess, dep.ambi_pred_t = types.SimpleNamespace(run=sess_run), None
    rows = dep.deploy('/clip', 3.0, DEPLOY_WINDOWS * 0.1)
    assert batches == [10, 10] and rows.shape == (DEPLOY_WINDOWS * 4800, 4) and rows.dtype == np.float64
    assert np.array_equal(rows[:, 0], amb[:, 24000:28800, 0].reshape(-1))           # W column = the mono crop, exactly
    G['deploy_pred_stride3'] = rows[::3, 1:].astype(np.float32)                    # Y, Z, X of every third output sample
    od = O.deploy_assemble(O.SptAudioGen(W, 1, encoders=enc, separation='unet_mask', dtype=torch_f64()), amb, video_windows=vid)
    err = np.abs(od - rows).max() / np.abs(rows).max()
    print('deploy loop: %d windows, oracle.deploy_assemble(float64) vs reference deploy.py rel err %.2e' % (DEPLOY_WINDOWS, err))
    assert err < 1e-4

    np.savez_compressed(OUT, **G)
    print('wrote %s: %d arrays, %d bytes' % (OUT, len(G), os.path.getsize(OUT)))


if __name__ == '__main__':
    main()
