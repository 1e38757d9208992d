"""CPU tests of the host-side mirror of the reference interface that need no GPU: parameter parsing with the
reference's fall-backs (myutils.py:40-85), constants (definitions.py), checkpoint layout helpers."""
import os

import numpy as np
import pytest

from spatialaudiogen_b200 import definitions as Df
from spatialaudiogen_b200 import weights as Wt


def _load_params(model_dir):
    # myutils imports the ctypes binding lazily enough for this to work without a GPU
    from spatialaudiogen_b200 import myutils
    return myutils.load_params(model_dir)


def test_load_params_with_reference_fallbacks(tmp_path):
    (tmp_path / 'train-params.txt').write_text(
        "encoders: ['audio', 'video']\nseparation: UNET_MASK\nambi_order: 1\naudio_rate: 48000\nvideo_rate: 10\n"
        "n_iters: 1000000\nbatch_size: 32\ncontext: 1.0\nsample_dur: 0.1\nlr: 0.0001\nlr_decay: 0.5\nlr_iters: 250000.0\n")
    p = _load_params(str(tmp_path))
    assert p.encoders == ['audio', 'video'] and p.separation == 'unet_mask'
    assert (p.ambi_order, p.audio_rate, p.video_rate, p.batch_size) == (1, 48000, 10, 32)
    # fall-backs differ from definitions.py on purpose (old checkpoints rely on them)
    assert p.num_sep_tracks == 64 and p.loc_units == [256, 256] and p.freq_mask_units == [] and p.context_units == [64, 128, 128]
    assert p.fft_window == 0.025
    (tmp_path / 'train-params.txt').write_text(
        "encoders: ['audio']\nseparation: none\nambi_order: 1\naudio_rate: 48000\nvideo_rate: 10\nn_iters: 10\nbatch_size: 2\n"
        "context: 1.0\nsample_dur: 0.1\nlr: 0.1\nlr_decay: 0.5\nlr_iters: 5\nnum_sep_tracks: 32\nloc_units: [512, 512]\nfft_window: 0.025\n")
    p = _load_params(str(tmp_path))
    assert p.num_sep_tracks == 32 and p.loc_units == [512, 512] and p.encoders == ['audio']


def test_definitions_and_checkpoint_layout():
    assert (Df.AUDIO, Df.VIDEO, Df.FLOW) == ('audio', 'video', 'flow') and Df.FREQ_MASK == 'unet_mask'
    assert Df.NUM_SEP_TRACKS_DEF == 32 and Df.LOC_FCUNITS_DEF == [512, 512] and Df.SEP_FFT_WINDOW_DEF == 0.025
    sh = Wt.variable_shapes(['audio', 'video', 'flow'])
    assert len(sh) == 214 and sum(int(np.prod(s)) for s in sh.values()) == 49005763
    assert sh['separation/deconv1/weights'] == (7, 16, 32, 64) and sh['localization/fc3/weights'] == (512, 99)
    assert sh['bottleneck/video-fc/weights'] == (12544, 512) and sh['video_encoder/conv3_1/shortcut/weights'] == (1, 1, 64, 128)
    sh1 = Wt.variable_shapes(['audio'], separation='none', sep_num_tracks=1)
    assert sh1['localization/fc3/weights'] == (512, 6) and not any(k.startswith('separation/') for k in sh1)


def test_emd_hat_host_solver_matches_the_defining_lp():
    """sag_emd_hat (host code in libsag.so, no GPU involved) against the oracle's LP restatement of pyemd.emd."""
    import numpy as np
    from oracle import sag_oracle as O
    from spatialaudiogen_b200 import metrics as M
    rng = np.random.RandomState(1)
    phi, nu = M.spherical_mesh(30.)
    assert phi.shape == (7, 12)
    p = np.stack((np.cos(nu) * np.cos(phi), np.cos(nu) * np.sin(phi), np.sin(nu)), 0).reshape(3, -1)
    D = np.arccos(np.clip(p.T.dot(p), -1, 1))
    first, second = rng.rand(5, 84), rng.rand(5, 84) * np.array([[0.5], [1.0], [2.0], [1.0], [1.3]])
    second[1] *= first[1].sum() / second[1].sum()                       # one balanced pair
    got = M.emd_hat(first, second, D)
    for k in range(5):
        assert abs(got[k] - O.emd_hat_lp(first[k], second[k], D)) < 1e-9
    # identities: zero to itself; a unit of mass moved between two nodes costs their distance; excess mass costs max(D)
    assert M.emd_hat(first[0], first[0], D)[0] < 1e-6                   # (arccos of a rounded dot product leaves ~1e-8 on the diagonal)
    e3, e40 = np.eye(84)[3], np.eye(84)[40]
    assert abs(M.emd_hat(e3, e40, D)[0] - D[3, 40]) < 1e-12
    assert abs(M.emd_hat(2 * e3, e3, D)[0] - D.max()) < 1e-12
    assert abs(M.emd_hat(2 * e3, e3, D, extra_mass_penalty=0.25)[0] - 0.25) < 1e-12
    import pytest
    with pytest.raises(ValueError):
        M.emd_hat(-e3, e3, D)


def _make_video_folder(root, seconds=4, flow=False, seed=0):
    """A per-video folder in the layout scraping/preprocess.py writes (readers.py docstring); returns (folder, ambix)."""
    import numpy as np
    from PIL import Image
    from spatialaudiogen_b200 import readers as R
    folder = os.path.join(root, 'vidA')
    for sub in ('ambix', 'video') + (('flow',) if flow else ()):
        os.makedirs(os.path.join(folder, sub))
    rng = np.random.RandomState(seed)
    full = np.round(rng.uniform(-0.5, 0.5, size=(seconds * 48000, 4)) * 32768) / 32768      # exactly representable in PCM16
    for i in range(seconds):
        R.save_wav(os.path.join(folder, 'ambix', '%06d.wav' % i), full[i * 48000:(i + 1) * 48000], 48000)
    for i in range(seconds * 10):
        Image.fromarray(rng.randint(0, 256, size=(224, 448, 3)).astype(np.uint8)).save(os.path.join(folder, 'video', '%06d.jpg' % i), quality=95)
        if flow:
            Image.fromarray(rng.randint(0, 256, size=(224, 448, 3)).astype(np.uint8)).save(os.path.join(folder, 'flow', '%06d.jpg' % i), quality=95)
    if flow:
        np.save(os.path.join(folder, 'flow', 'flow_limits.npy'), np.stack([np.full(seconds * 10, 1.0), np.full(seconds * 10, 21.0)], 1))
    with open(os.path.join(folder, 'audio_pow.lst'), 'w') as f:
        for k in range(seconds):
            f.write('%.1f %.3f\n' % (0.5 + k, 0.05 if k == 1 else 0.3))
    return folder, full


def test_sample_reader_follows_the_reference_schedule_and_padding(tmp_path):
    """feeder.py:50-278 on a synthetic folder: chunk ids / times from audio_pow.lst, audio windows centred on the chunk
    with zero padding at both clip edges, frame index int(t * 10), flow de-quantisation, silence / duration filters."""
    import numpy as np
    from PIL import Image
    from spatialaudiogen_b200 import readers as R
    folder, full = _make_video_folder(str(tmp_path), seconds=4, flow=True)
    r = R.SampleReader(folder, shuffle=False, random_rotations=False, return_flow=True, img_prep=lambda x: x / 255. - 0.5)
    assert r.chunks_t == [0.5, 1.5, 2.5, 3.5] and r.audio_size == 52799 and r.video_size == 1
    c = r.get()
    assert c['id'] == 'vidA 0.5' and c['ambix'].shape == (52799, 4) and c['video'].shape == (1, 224, 448, 3)
    assert np.array_equal(c['ambix'], full[:52799])                      # window [t - 0.5, t + 0.6) s, bit-exact PCM16 decode
    img = np.asarray(Image.open(os.path.join(folder, 'video', '000005.jpg')).convert('RGB'))
    assert np.array_equal(c['video'][0], img / 255. - 0.5)               # frame int(0.5 * 10), myutils.img_prep_fcn
    raw = np.asarray(Image.open(os.path.join(folder, 'flow', '000005.jpg')).convert('RGB')).astype(np.float32)
    mag = raw[:, :, 2] * (21.0 - 1.0) / 255. + 1.0                       # feeder.py:153-160
    ang = raw[:, :, 0] * (2 * np.pi) / 255.
    assert np.allclose(c['flow'][0, :, :, 2], mag) and np.allclose(c['flow'][0, :, :, 0], mag * np.cos(ang), atol=1e-4)
    assert np.allclose(c['flow'][0, :, :, 1], mag * np.sin(ang), atol=1e-4)
    r.get(); r.get()
    c = r.get()                                                          # t = 3.5: 1 s of audio left, 4799 zeros after it
    assert np.array_equal(c['ambix'][:48000], full[3 * 48000:]) and not c['ambix'][48000:].any()
    assert r.get() is None
    a = R.AudioReader(os.path.join(folder, 'ambix'), 48000).get(-0.25, 52799)        # before the start: 12000 zeros first
    assert not a[:12000].any() and np.array_equal(a[12000:], full[:52799 - 12000])
    assert R.SampleReader(folder, shuffle=False, skip_silence_thr=0.2, return_video=False).chunks_t == [0.5, 2.5, 3.5]
    assert R.SampleReader(folder, shuffle=False, start_time=1.0, sample_duration=2.0, return_video=False).chunks_t == [1.5, 2.5]
    rot = R.AudioReader(os.path.join(folder, 'ambix'), 48000).get(0.0, 100, rotation=np.pi / 2)   # yaw by 90 deg: Y' = X, X' = -Y
    assert np.allclose(rot[:, 1], full[:100, 3]) and np.allclose(rot[:, 3], -full[:100, 1]) and np.allclose(rot[:, [0, 2]], full[:100, [0, 2]])


def test_sample_folders_and_channel_masks(tmp_path):
    """feeder.py:12-47 (FilenameProvider) and feeder.py:312-314 (audio layout masks)."""
    from spatialaudiogen_b200 import readers as R
    db = tmp_path / 'db'
    for vid in ('aaa', 'bbb', 'ccc'):
        os.makedirs(str(db / vid))
    subset = tmp_path / 'subset.lst'
    subset.write_text('ccc\naaa\nzzz\n')
    got = R.sample_folders(str(db), str(subset))
    assert got == [os.path.join(str(db), y) for y in os.listdir(str(db)) if y in ('aaa', 'ccc')] and len(got) == 2
    assert len(R.sample_folders(str(db))) == 3
    os.makedirs(str(tmp_path / 'empty'))
    with pytest.raises(ValueError):
        R.sample_folders(str(tmp_path / 'empty'))
    with pytest.raises(IOError):
        R.sample_folders(str(db), str(tmp_path / 'missing.lst'))
    lay = tmp_path / 'audio_layouts.txt'
    lay.write_text('aaa WXYZ\nbbb WXY\n\nccc WXYZ\n')
    m = R.load_channel_masks(str(lay))
    assert sorted(m) == ['aaa', 'bbb', 'ccc'] and list(m['bbb']) == [1., 1., 0., 1.] and list(m['aaa']) == [1., 1., 1., 1.]


def test_evaluate_model_dir_wiring(tmp_path, monkeypatch):
    """eval.py:29-215 `main`, host side only (the model and the GPU loops are stood in): folder list from db_dir + subset,
    channel masks, refusal to overwrite eval-detailed.txt, file layout."""
    import torch
    from spatialaudiogen_b200 import evaluate as E, deploy as D
    md, db = tmp_path / 'model', tmp_path / 'db'
    os.makedirs(str(md))
    for vid in ('v1', 'v2'):
        os.makedirs(str(db / vid))
    (md / 'train-params.txt').write_text("encoders: ['audio']\nseparation: unet_mask\nambi_order: 1\naudio_rate: 48000\nvideo_rate: 10\n"
                                         "context: 1.0\nsample_dur: 0.1\nlr: 0.0001\nn_iters: 10\nbatch_size: 32\nlr_decay: 0.5\nlr_iters: 30000\n"
                                         "db_dir: %s\n" % str(db))
    (tmp_path / 'subset.lst').write_text('v2\n')
    (tmp_path / 'layouts.txt').write_text('v2 WXY\n')
    seen = {}

    class FakeW2XYZ(object):
        def __init__(self, model_dir, params=None, precision=None, device=None):
            self.model = type('M', (), {'device': 'cpu'})()
            seen['model_dir'] = model_dir

    def fake_folder_batches(folders, params, batch_size=16, channel_masks=None, device=None, drop_remainder=True):
        seen.update(folders=folders, masks=channel_masks, batch_size=batch_size, drop=drop_remainder)
        return iter([])

    def fake_evaluate_batches(model, batches, audio_rate=48000, rms_maps=False):
        seen.update(audio_rate=audio_rate, rms_maps=rms_maps)
        return ['v2 0.5', 'v2 1.5'], torch.arange(56, dtype=torch.float32).reshape(2, 28)
    monkeypatch.setattr(D, 'W2XYZ', FakeW2XYZ)
    monkeypatch.setattr(E, 'folder_batches', fake_folder_batches)
    monkeypatch.setattr(E, 'evaluate_batches', fake_evaluate_batches)
    ids, rows = E.evaluate_model_dir(str(md), subset_fn=str(tmp_path / 'subset.lst'), audio_layouts_fn=str(tmp_path / 'layouts.txt'))
    assert seen['folders'] == [os.path.join(str(db), 'v2')] and list(seen['masks']['v2']) == [1., 1., 0., 1.]
    assert seen['rms_maps'] is True and seen['audio_rate'] == 48000 and seen['batch_size'] == 16
    lines = open(str(md / 'eval-detailed.txt')).read().splitlines()
    assert lines[0].startswith('SampleID | amplitude/predicted amplitude/gt mse/avg') and len(lines) == 3
    assert lines[1].startswith('v2 0.5 | 0.0 1.0 2.0') and len(lines[1].split(' | ')[1].split()) == 28
    with pytest.raises(AssertionError):
        E.evaluate_model_dir(str(md), subset_fn=str(tmp_path / 'subset.lst'))
    E.evaluate_model_dir(str(md), subset_fn=str(tmp_path / 'subset.lst'), overwrite=True, audio_layouts_fn=None)
    assert seen['masks'] is None


def test_stage_methods_glue_matches_oracle(monkeypatch):
    """stages.StageOps (the reference's audio_encoder_ops / visual_encoding_ops / bottleneck_ops / localization_ops /
    separation_ops): the crops, reshapes, tiles, concat order and the mask arithmetic between the dense primitives, with
    the four primitives (one C-ABI call each on the GPU) stood in by the oracle's CPU ops, against oracle.SptAudioGen."""
    import types
    import torch
    from oracle import sag_oracle as O
    from spatialaudiogen_b200 import stages as S
    enc = ['audio', 'video']
    W = Wt.init_weights(enc, separation='unet_mask', seed=5, stress=True)
    om = O.SptAudioGen(W, 1, encoders=enc, separation='unet_mask')

    class Fake(S.StageOps):
        pass
    m = Fake()
    m.ambi_order, m.separation, m.snd_contx, m.snd_dur, m.params = 1, 'unet_mask', om.snd_contx, om.snd_dur, om.params
    ss, tt = om.encoder_crop()
    mss, mtt, mskip = om.mask_crop()
    m.dims = types.SimpleNamespace(enc_ss=ss, enc_tt=tt, mask_ss=mss, mask_tt=mtt, mask_skip=mskip, final_crop=om.final_crop())
    m._conv = lambda scope, x, stride, same, relu: O.conv_2d(om.W, scope, x, stride, 'SAME' if same else 'VALID', relu)
    m._deconv = lambda scope, x, stride, relu: O.deconv_2d(om.W, scope, x, stride, relu)
    m._fc = lambda scope, x, relu=True: O.fully_connected(om.W, scope, x, relu)
    m._resnet = lambda scope, x: O.resnet18(om.W, scope, x)[0]
    monkeypatch.setattr(S.myutils, 'istft', O.istft)
    rng = np.random.RandomState(1)
    audio = torch.as_tensor((rng.randn(2, 52799, 1) * 0.1).astype(np.float32))
    video = torch.as_tensor(rng.rand(2, 1, 224, 448, 3).astype(np.float32) - 0.5)
    ref = om.inference_ops(audio, video)
    mono = audio.permute(0, 2, 1)
    s = O.stft(mono, om.wind_size, 4)
    a_enc = m.audio_encoder_ops(s)
    assert len(a_enc) == 6 and all(torch.equal(a, b) for a, b in zip(a_enc, om.ends['audio_encoder']))
    v = m.visual_encoding_ops(video, is_training=False, finetune=True, scope='video_encoder')
    feats = m.bottleneck_ops({'audio': a_enc, 'video': v}, True)
    assert torch.equal(feats, om.ends['bottleneck'])
    w, b = m.localization_ops(feats)
    assert tuple(w.shape) == (2, 4800, 3, 1, 32) and tuple(b.shape) == (2, 4800, 3, 1)
    assert torch.equal(w, om.loc_channels[0]) and torch.equal(b, om.loc_channels[1])
    x_sep = m.separation_ops(mono, s, a_enc, feats)
    assert tuple(x_sep.shape) == (2, 1, 32, 4800) and torch.equal(x_sep, om.sep_channels)
    y = (w * x_sep.permute(0, 3, 1, 2).unsqueeze(2)).sum(4).sum(3) + b[:, :, :, 0]          # model.py:424-432
    assert torch.equal(y, ref)
    m.separation = 'none'
    assert torch.equal(m.separation_ops(mono, s, None, None), mono[:, :, 24000:28800].unsqueeze(1))


def test_prefetch_feeder_thread_keeps_order_and_propagates_errors():
    """evaluate.prefetch (the reference's feeder thread + bounded queue, feeder.py:281-435, for one producer)."""
    import threading
    import time
    import pytest
    from spatialaudiogen_b200 import evaluate as E
    main = threading.get_ident()
    seen = []

    def produce(n, fail_at=None):
        for i in range(n):
            seen.append(threading.get_ident())
            if i == fail_at:
                raise KeyError('item %d' % i)
            yield i

    assert list(E.prefetch(produce(50), 3)) == list(range(50)) and all(t != main for t in seen)      # on another thread, in order
    del seen[:]
    assert list(E.prefetch(produce(5), 0)) == list(range(5)) and all(t == main for t in seen)        # depth 0: inline
    got = []
    with pytest.raises(KeyError):
        for x in E.prefetch(produce(10, fail_at=4), 2):
            got.append(x)
    assert got == [0, 1, 2, 3]                                                                       # the error arrives in position
    # bounded: the producer runs at most depth (+1 in hand) items ahead of a slow consumer; an abandoned generator stops it
    del seen[:]
    g = E.prefetch(produce(1000), 2)
    assert next(g) == 0
    time.sleep(0.3)
    assert len(seen) <= 5
    g.close()
    time.sleep(0.3)
    n = len(seen)
    time.sleep(0.2)
    assert len(seen) == n and n < 20


def test_parse_eval_results_matches_the_reference_script(tmp_path):
    """parse_eval_results.py (the paper's MSE / STFT / ENV / EMD table from eval-detailed.txt): golden = the reference's own script run
    on the same file (tests/golden/make_parse_eval_golden.py); also run live when /root/reference is present."""
    import json
    import subprocess
    import sys
    import numpy as np
    from spatialaudiogen_b200 import evaluate as E
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'parse_eval_results.json')))
    fn = str(tmp_path / 'eval-detailed.txt')
    E.write_eval_detailed(fn, g['ids'], np.asarray(g['rows']))
    vals, times, keys = E.parse_eval_detailed_file(fn)
    assert keys == E.ALL_METRICS and list(vals) == ['vidA', 'vidB', 'vidC'] and vals['vidA'].shape == (7, 28)
    assert all(np.all(np.diff(t) > 0) for t in times.values())                       # each video's windows sorted by time
    table = E.parse_eval_results(fn)
    mine = ''.join('{} = {:.3f}\n'.format(k.ljust(4), v) for k, v in table.items())
    assert mine == g['reference_stdout']
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = subprocess.run([sys.executable, '-m', 'spatialaudiogen_b200.evaluate', fn], capture_output=True, text=True, cwd=root)
    assert cli.returncode == 0 and cli.stdout == g['reference_stdout']
    ref = '/root/reference/parse_eval_results.py'
    if os.path.exists(ref):
        live = subprocess.run([sys.executable, ref, fn], capture_output=True, text=True)
        assert live.returncode == 0 and live.stdout == mine


def test_myutils_prep_functions():
    """myutils.py:88-93: img_prep_fcn and flow_prep_fcn (nearest-neighbour resize to 224 x 448)."""
    import numpy as np
    import pytest
    from spatialaudiogen_b200 import myutils as M
    x = np.random.RandomState(0).randint(0, 256, (112, 224, 3)).astype(np.uint8)
    y = M.flow_prep_fcn()(x)
    assert y.shape == (224, 448, 3) and y.dtype == np.uint8
    assert np.array_equal(y[::2, ::2], x) and np.array_equal(y[1::2, 1::2], x)      # every source pixel becomes a 2 x 2 block
    same = np.zeros((224, 448, 3), np.uint8)
    assert M.flow_prep_fcn()(same) is same
    with pytest.raises(TypeError):
        M.flow_prep_fcn()(x.astype(np.float32))
    assert np.allclose(M.img_prep_fcn()(np.array([0, 255, 51])), [-0.5, 0.5, -0.3])
    assert callable(M.compute_lsd_dist) and callable(M.compute_envelope_dist)


def test_command_lines_mirror_the_reference_arguments():
    """deploy.py:14-38 and eval.py:14-26: same arguments, same post-processing of them."""
    from spatialaudiogen_b200 import deploy as D, evaluate as E
    a = D.parse_arguments(['snap', 'data/frames/vid', 'vid.mp4', '--deploy_start', '3.5', '--deploy_duration', '0', '--output_fn', 'out/x', '--VR'])
    assert (a.model_dir, a.input_folder, a.video, a.deploy_start, a.deploy_duration, a.output_fn, a.VR, a.gpu) == \
        ('snap', 'data/frames/vid', 'vid.mp4', 3.5, None, 'out/x', True, 0)
    assert D.parse_arguments(['snap', 'folder']).deploy_duration == 10.
    e = E.parse_arguments(['snap', '--batch_size', '8', '--overwrite'])
    assert (e.model_dir, e.subset_fn, e.batch_size, e.overwrite, e.gpu) == ('snap', None, 8, True, 0)
    assert E.parse_arguments(['snap', '--subset_fn', 'meta/subsets/YT-All.test.1.lst']).subset_fn == 'meta/subsets/YT-All.test.1.lst'
