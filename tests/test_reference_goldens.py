"""The oracle (and the product's host-side code) against vectors computed by the REFERENCE'S OWN functions
(tests/golden/reference_goldens.npz, generated in the build container by tests/golden/make_reference_goldens.py from
/root/reference through a syntactic python-2 shim).  This pins the rows whose reference implementation is plain numpy /
scipy: a1 constants, a12 envelope distance, a13 mesh / SH matrix / energy maps, f2 reader logic, load_params -- and,
with the reference's TF graph-building code run eagerly on a numpy stand-in for the elementary ops it calls, a2 stft,
a8 istft and the a11 evaluation metrics."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import sag_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_goldens.npz'))


def test_a1_derived_constants_match_reference_init():
    for row in G['a1_configs_and_dims']:
        order, ar, vr, ctx, dur, win = int(row[0]), int(row[1]), int(row[2]), float(row[3]), float(row[4]), float(row[5])
        m = O.SptAudioGen({}, ambi_order=order, audio_rate=ar, video_rate=vr, context=ctx, sample_duration=dur, encoders=['audio'],
                          separation='none', params=O.SptAudioGenParams(sep_fft_window=win))
        assert [m.num_ambi_channels, m.snd_contx, m.snd_dur, m.snd_size, m.wind_size] == [int(v) for v in row[6:]]


def test_a13_mesh_sh_matrix_and_energy_maps_match_reference():
    for res in (30, 5):
        phi, nu = O.spherical_mesh(res)
        assert np.array_equal(phi, G['a13_phi_mesh_%d' % res]) and np.array_equal(nu, G['a13_nu_mesh_%d' % res])
    phi, nu = O.spherical_mesh(30)
    pts = [O._position_polar_roundtrip(p, n) for p, n in zip(phi.reshape(-1), nu.reshape(-1))]
    Y = O.spherical_harmonics_matrix([p[0] for p in pts], [p[1] for p in pts], 1)
    assert np.abs(Y - G['a13_sh_matrix_30']).max() < 1e-15
    ambi = G['a13_ambi'].astype(np.float64)
    for res in (30, 5):
        assert np.abs(O.ambix_rms_map(ambi, float(res)) - G['a13_rms_map_%d' % res]).max() < 1e-14
    assert np.abs(O.ambix_rms_map(ambi * np.array([1., 1., 0., 1.]), 30.) - G['a13_rms_map_30_wxy']).max() < 1e-14


def test_a13_emd_columns_match_reference_wrapper():
    """emd/dir and emd/dir2 of one window: the reference's distance.ambix_emd / emd (ground distance, normalisations, frame
    loop) with pyemd.emd stood in by its defining LP.  Checked: the oracle's restatement, and the product's host path
    (metrics.ambix_emd_from_maps -> sag_emd_hat, the min-cost-flow solver in libsag.so; no GPU involved) on the
    reference's own energy maps."""
    maps_of = lambda a: O.ambix_rms_map(a, 30.)                        # (itself pinned by the test above)
    ref = G['a13_emd_dir_dir2']
    a2, a1 = G['a13_ambi2'].astype(np.float64), G['a13_ambi'].astype(np.float64)
    got = O.ambix_emd(a2, a1, 30.)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)
    from spatialaudiogen_b200 import metrics as M
    d1, d2 = M.ambix_emd_from_maps(maps_of(a2)[None], maps_of(a1)[None], 30.)
    assert np.allclose([d1[0], d2[0]], ref, rtol=1e-9, atol=1e-12)


def test_a12_envelope_distance_matches_reference():
    got = O.compute_envelope_dist(G['a12_pred'], G['a12_gt'])
    # (the reference hands scipy.signal.hilbert float32 signals: its transform runs in single precision)
    assert np.allclose(np.asarray(got), G['a12_env_dist'], rtol=1e-6, atol=0)


def test_load_params_matches_reference(tmp_path):
    from spatialaudiogen_b200 import myutils
    open(str(tmp_path / 'train-params.txt'), 'w').write(str(G['params_text']))
    p = myutils.load_params(str(tmp_path))
    for k, v in ast.literal_eval(str(G['params_repr'])):
        assert getattr(p, k) == v, (k, getattr(p, k), v)


def test_f2_audio_reader_padding_and_flow_dequantisation_match_reference(tmp_path):
    """The product's host readers on the same tiny clip (4 one-second files at 200 Hz) the reference's AudioReader.get
    was run on, and on the same quantised flow frames."""
    from spatialaudiogen_b200 import readers as R
    from scipy.io import wavfile
    clip, rate = G['f2_clip'], 200
    folder = str(tmp_path / 'ambix')
    os.makedirs(folder)
    for i in range(4):                                                   # float wav files: exact round trip of the array
        wavfile.write(os.path.join(folder, '%06d.wav' % i), rate, clip[i * rate:(i + 1) * rate].astype(np.float64))
    ar = R.AudioReader(folder, rate, ambi_order=1)
    assert (ar.num_files, ar.num_channels, ar.num_frames) == (4, 4, 800)
    for i, (t0, size) in enumerate(G['f2_audio_cases']):
        got = ar.get(float(t0), int(size))
        assert got.shape == G['f2_audio_out_%d' % i].shape and np.array_equal(got, G['f2_audio_out_%d' % i]), (i, t0, size)
    assert np.abs(ar.get(0.5, 64, rotation=0.7) - G['f2_audio_rot']).max() < 1e-15
    fr = object.__new__(R.FlowReader)

    class FakeReader(object):
        rate = 10.

        def get_by_index(self, start_time, size, rotation=None):
            return G['f2_flow_raw'].copy()
    fr.reader, fr.rate, fr.lims = FakeReader(), 10., G['f2_flow_lims']
    assert np.array_equal(fr.get_by_index(1.3, 2), G['f2_flow_out'])


def _audio_a2():
    return (G['a2_audio_q12'].astype(np.float64) / 4096).astype(np.float32)               # (1, 1, 52799)


def test_a2_a8_stft_istft_match_reference_graph_code():
    s = O.stft(torch.as_tensor(_audio_a2()), 1024, 4)
    assert tuple(s.shape) == (1, 1, 200, 1024)
    s = s.numpy()
    ref = G['a2_stft_bins_stride37']
    assert np.abs(s[0, 0, :, ::37] - ref).max() < 3e-6 * np.abs(ref).max()
    assert np.allclose(np.abs(s[0, 0]).sum(-1), G['a2_stft_abs_sum_per_frame'], rtol=1e-5)
    small = O.stft(torch.as_tensor(G['a2_small_in']), 64, 4).numpy()
    assert small.shape == G['a2_small_stft'].shape and np.abs(small - G['a2_small_stft']).max() < 3e-6 * np.abs(G['a2_small_stft']).max()
    y = O.istft(torch.as_tensor(G['a8_small_in']), 4).numpy()
    assert y.shape == G['a8_small_istft'].shape and np.abs(y - G['a8_small_istft']).max() < 3e-6 * np.abs(G['a8_small_istft']).max()
    # istft(stft(x)[89:117]) = 0.5 * x[23552:29952], in the reference's own numbers
    ref = G['a8_istft_of_stft_frames_89_117']
    assert ref.shape == (1, 6400) and np.abs(ref[0] - 0.5 * _audio_a2()[0, 0, 23552:29952]).max() < 1e-6
    y = O.istft(torch.as_tensor(O.stft(torch.as_tensor(_audio_a2()), 1024, 4).numpy()[0, :, 89:117]), 4).numpy()
    assert np.abs(y - ref).max() < 1e-6


def _a11_inputs():
    return ((G['a11_pred_q12'].astype(np.float64) / 4096).astype(np.float32), (G['a11_gt_q12'].astype(np.float64) / 4096).astype(np.float32),
            G['a11_mask'])


def test_a11_evaluation_ops_match_reference_graph_code():
    pred, gt, mask = _a11_inputs()
    m = O.SptAudioGen({}, encoders=['audio'], separation='none')
    metrics, stft_ps, lsd_ps, mse_ps, snr_ps = m.evaluation_ops(pred, gt, None, mask)
    for got, key in ((stft_ps, 'a11_stft_ps'), (lsd_ps, 'a11_lsd_ps'), (mse_ps, 'a11_mse_ps'), (snr_ps, 'a11_snr_ps')):
        assert np.allclose(np.asarray(got), G[key], rtol=2e-5), key
    names = ast.literal_eval(str(G['a11_metric_names']))
    for k, v in zip(names, G['a11_metric_values']):
        assert abs(float(metrics[k]) - v) < 3e-5 * max(1.0, abs(v)), (k, float(metrics[k]), v)


@pytest.mark.gpu
def test_gpu_stft_istft_metrics_against_reference_graph_code():
    """The CUDA STFT / inverse STFT / metrics kernels directly against the numbers of the reference's own code."""
    from spatialaudiogen_b200 import myutils, metrics as M
    s = myutils.stft(torch.as_tensor(_audio_a2()).cuda(), 1024, 4)
    ref = G['a2_stft_bins_stride37']
    assert np.abs(s[0, 0, :, ::37].cpu().numpy() - ref).max() < 3e-6 * np.abs(ref).max()
    small = myutils.stft(torch.as_tensor(G['a2_small_in']).cuda(), 64, 4).cpu().numpy()
    assert np.abs(small - G['a2_small_stft']).max() < 3e-6 * np.abs(G['a2_small_stft']).max()
    y = myutils.istft(torch.as_tensor(G['a8_small_in']).cuda(), 4).cpu().numpy()
    assert np.abs(y - G['a8_small_istft']).max() < 3e-6 * np.abs(G['a8_small_istft']).max()
    y = myutils.istft(s[0, :, 89:117].contiguous(), 4).cpu().numpy()
    assert np.abs(y - G['a8_istft_of_stft_frames_89_117']).max() < 2e-6
    pred, gt, mask = _a11_inputs()
    r = M.window_metrics(torch.as_tensor(pred).cuda(), torch.as_tensor(gt).cuda())
    for key, name in (('stft', 'a11_stft_ps'), ('lsd', 'a11_lsd_ps'), ('mse', 'a11_mse_ps'), ('snr', 'a11_snr_ps')):
        assert np.allclose(r[key].cpu().numpy(), G[name], rtol=2e-4), key


@pytest.mark.gpu
def test_gpu_kernels_against_reference_goldens():
    """The CUDA kernels behind rows a12 / a13 directly against the reference's numbers (no oracle in between)."""
    from spatialaudiogen_b200 import metrics as M
    env = M.window_metrics(torch.as_tensor(G['a12_pred'])[None].cuda(), torch.as_tensor(G['a12_gt'])[None].cuda())['env'][0].cpu().numpy()
    assert np.allclose(env, G['a12_env_dist'], rtol=2e-4)
    ambi = torch.as_tensor(G['a13_ambi'])[None].cuda()
    for res in (30, 5):
        got = M.ambix_rms_map(ambi, float(res))[0].cpu().numpy()
        assert np.abs(got - G['a13_rms_map_%d' % res]).max() < 1e-5 * G['a13_rms_map_%d' % res].max()


# ---- a3-a7, a9, a10: the reference's model-building code (model.py, wrappers/core.py, resnet.py) run eagerly -------------

MODEL_SEED = 7


def _model_inputs(seed, batch):                                       # same recipe as make_reference_goldens.model_inputs
    r = np.random.RandomState(seed)
    n = 52799
    tone = 0.2 * np.sin(2 * np.pi * 523.25 * np.arange(n) / 48000.)[None, :, None]
    audio = (np.round(np.clip(0.1 * r.randn(batch, n, 1) + tone, -1, 1) * 4096) / 4096).astype(np.float32)
    video = (r.randint(0, 256, size=(batch, 1, 224, 448, 3)) / 255.).astype(np.float32)
    flow = (r.randint(0, 256, size=(batch, 1, 224, 448, 3)) / 255. - 0.5).astype(np.float32)
    return audio, video, flow


def _model_case(tag):
    from spatialaudiogen_b200 import weights as PW
    encoders = ['audio'] if tag == 'a' else ['audio', 'video', 'flow']
    W = PW.init_weights(encoders, 'unet_mask', seed=MODEL_SEED, stress=True)
    chk = np.asarray([float(np.asarray(v, np.float64).sum()) for v in W.values()])
    assert np.array_equal(chk, G['m_%s_weight_checksum' % tag]), 'weights.init_weights changed: regenerate the goldens'
    audio, video, flow = _model_inputs(MODEL_SEED, 2)
    return encoders, W, audio, (video, flow) if tag == 'avf' else ()


def _maxrel(got, ref):
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize('tag', ['a', 'avf'])
def test_checkpoint_layout_is_what_the_reference_code_creates(tag):
    """Every variable the reference's model code asked tf for -- full scoped name and shape -- is what the product's
    checkpoint layout (weights.variable_shapes) declares, and nothing else; the ResNet towers' names are keys of the
    reference's resnet18.npy (restore_pretrained)."""
    from spatialaudiogen_b200 import weights as PW
    created = dict(ast.literal_eval(str(G['m_%s_variables' % tag])))
    mine = PW.variable_shapes(['audio'] if tag == 'a' else ['audio', 'video', 'flow'], 'unet_mask')
    assert {k: tuple(v) for k, v in mine.items()} == created
    assert len(created) == (30 if tag == 'a' else 214)
    if tag == 'avf':
        tower = ast.literal_eval(str(G['m_resnet18_npy_keys_used']))
        assert tower == sorted(k[len('video_encoder/'):] for k in mine if k.startswith('video_encoder/'))


@pytest.mark.parametrize('tag', ['a', 'avf'])
def test_oracle_forward_matches_reference_model_code(tag):
    """oracle.SptAudioGen.inference_ops (float32, torch ops) against the ambisonics, separated tracks, localisation
    weights and tower activations that the reference's own inference_ops produced on the same weights and inputs."""
    encoders, W, audio, vis = _model_case(tag)
    m = O.SptAudioGen(W, 1, encoders=encoders, separation='unet_mask')
    y = m.inference_ops(audio, *vis).numpy()
    assert y.shape == G['m_%s_ambix' % tag].shape == (2, 4800, 3)
    assert _maxrel(y, G['m_%s_ambix' % tag]) < 2e-5
    assert _maxrel(m.sep_channels.numpy()[:, 0, :, ::97], G['m_%s_sep_stride97' % tag]) < 2e-5
    assert _maxrel(m.loc_channels[0].numpy()[:, ::480], G['m_%s_loc_w_stride480' % tag]) < 2e-5
    assert _maxrel(m.loc_channels[1].numpy()[:, ::480], G['m_%s_loc_b_stride480' % tag]) < 2e-5
    assert _maxrel(m.ends['stft'].abs().numpy().sum(-1), G['m_%s_inp_spect_sum' % tag]) < 1e-5
    if tag == 'avf':
        for k in ('video', 'flow'):
            assert _maxrel(m.ends[k + '_encoder/conv'].numpy()[:, ::8, ::8, ::4], G['m_avf_%s_conv1_stride' % k]) < 2e-5
            assert _maxrel(m.ends[k + '_encoder/conv5_2'].numpy()[..., ::16], G['m_avf_%s_conv5_2_stride16' % k]) < 5e-5


@pytest.mark.gpu
@pytest.mark.parametrize('tag,precision,tol', [('a', 'fp32', 1e-4), ('a', 'bf16x3', 1e-3), ('avf', 'fp32', 1e-4), ('avf', 'bf16x3', 1e-3)])
def test_gpu_forward_against_reference_model_code(tag, precision, tol):
    """The CUDA forward (two-kernel inference_ops and the fused deploy / eval hot loop forward_into) directly against the
    output of the reference's own model code -- no oracle in between.  Tolerance: north_star's 1e-3 relative on the
    waveform for the tensor-core path, 1e-4 for the fp32 FFMA path."""
    from spatialaudiogen_b200.model import SptAudioGen
    encoders, W, audio, vis = _model_case(tag)
    m = SptAudioGen(1, encoders=encoders, separation='unet_mask', precision=precision).load_weights(W)
    a = torch.as_tensor(audio).cuda()
    kw = dict(video=torch.as_tensor(vis[0]).cuda(), flow=torch.as_tensor(vis[1]).cuda()) if vis else {}
    y = m.inference_ops(a, **kw).cpu().numpy()
    assert _maxrel(y, G['m_%s_ambix' % tag]) < tol
    assert _maxrel(m.sep_channels.cpu().numpy()[:, 0, :, ::97], G['m_%s_sep_stride97' % tag]) < tol
    out = torch.empty(2, 4800, 3, device='cuda')
    m.forward_into(a, kw.get('video'), kw.get('flow'), out)
    torch.cuda.synchronize()
    assert _maxrel(out.cpu().numpy(), G['m_%s_ambix' % tag]) < tol


# ---- deploy.py:90-152: the reference's deploy loop around its own model code ---------------------------------------------

DEPLOY_WINDOWS = 11


def _deploy_inputs(seed, n):                                          # same recipe as make_reference_goldens.deploy_inputs
    r = np.random.RandomState(seed)
    amb = np.round(np.clip(0.1 * r.randn(n, 52799, 4), -1, 1) * 4096) / 4096
    vid = (r.randint(0, 256, size=(n, 1, 224, 448, 3)) / 255.).astype(np.float32)
    return amb, vid


def test_oracle_deploy_assemble_matches_reference_deploy_loop():
    """oracle.deploy_assemble (the CPU checker of the deploy row) against the reference's own deploy loop + model code."""
    from spatialaudiogen_b200 import weights as PW
    enc = ['audio', 'video']
    W = PW.init_weights(enc, 'unet_mask', seed=MODEL_SEED + 1, stress=True)
    amb, vid = _deploy_inputs(MODEL_SEED + 1, DEPLOY_WINDOWS)
    rows = O.deploy_assemble(O.SptAudioGen(W, 1, encoders=enc, separation='unet_mask'), amb, video_windows=vid)
    assert rows.dtype == np.float64 and rows.shape == (DEPLOY_WINDOWS * 4800, 4)
    assert np.array_equal(rows[:, 0], amb[:, 24000:28800, 0].reshape(-1))
    assert _maxrel(rows[::3, 1:], G['deploy_pred_stride3']) < 5e-5


@pytest.mark.gpu
def test_gpu_w2xyz_against_reference_deploy_loop():
    """W2XYZ.deploy_windows against the rows the reference's own W2XYZ.deploy produced for 11 windows of an audio+video
    model: a full batch of 10 and a tail of 1 zero-padded to 10 (with batch-statistics BN in the video tower the padding
    changes the tail's output, so this pins it), W = the exact mono crop, rows [W, Y, Z, X] in float64."""
    from spatialaudiogen_b200 import weights as PW
    from spatialaudiogen_b200.deploy import W2XYZ
    from types import SimpleNamespace
    enc = ['audio', 'video']
    W = PW.init_weights(enc, 'unet_mask', seed=MODEL_SEED + 1, stress=True)
    amb, vid = _deploy_inputs(MODEL_SEED + 1, DEPLOY_WINDOWS)
    params = SimpleNamespace(encoders=enc, separation='unet_mask', ambi_order=1, audio_rate=48000, video_rate=10, context=1.0,
                             num_sep_tracks=32, fft_window=0.025, context_units=[64, 128, 128], freq_mask_units=[256], loc_units=[512, 512])
    rows = W2XYZ(params=params, weights=W).deploy_windows(amb, video_windows=vid)
    assert rows.dtype == np.float64 and rows.shape == (DEPLOY_WINDOWS * 4800, 4)
    assert np.array_equal(rows[:, 0], amb[:, 24000:28800, 0].reshape(-1))
    ref = G['deploy_pred_stride3']
    assert _maxrel(rows[::3, 1:], ref) < 1e-3
    tail = slice(10 * 4800 // 3 + 1, None)                             # the zero-padded batch on its own
    assert _maxrel(rows[::3, 1:][tail], ref[tail]) < 1e-3


# ---- f4: gen_360video's overlay / down-mix arithmetic (myutils.py:224-311) --------------------------------------------

def _f4_clip():                                                        # same recipe as make_reference_goldens.f4_clip
    clip = (np.random.RandomState(11).randn(3 * 48000, 4) * np.array([0.2, 0.1, 0.03, 0.15])).astype(np.float32)
    clip[48000:96000, 1] *= 3.
    return clip


def test_f4_overlay_maps_and_stereo_downmix_match_reference_gen_360video():
    """The reference's gen_360video run with ffmpeg / video files / colour map / image resize stood in by recorders: the
    heat-map frames it blends over the video (decimation by 5, 5-degree mesh, min-max normalisation with the 0.005 guard,
    5-frame linear blend, 2*rms - 0.7 clipped), the colour-map index it derives from them, and the "binauralize" stereo
    down-mix -- against the oracle's and the product's host arithmetic."""
    from spatialaudiogen_b200 import myutils
    clip = _f4_clip()
    ref = G['f4_rms_frames']
    assert ref.shape == (23, 37, 72)                                    # the fake video ran out after 23 frames
    got = O.energy_map_frames(clip, 48000, 10.)
    assert got.shape == (25, 37, 72) and np.abs(got[:23] - ref).max() < 2e-6
    idx = np.minimum((got[:23] * 255).astype(int), 255)                 # myutils.py:273-274
    assert np.mean(idx != G['f4_colour_index']) < 1e-4                  # (float32-rounded golden: a handful of ties at most)
    st = myutils.ambix_to_stereo(clip)
    assert np.abs(st[::97] - G['f4_stereo_stride97']).max() < 1e-12 and abs(np.abs(st).max() - 0.95) < 1e-12


@pytest.mark.gpu
def test_gpu_overlay_maps_against_reference_gen_360video():
    """myutils.energy_map_frames (the K8 energy-map kernel + the host blend) directly against the reference's frames."""
    from spatialaudiogen_b200 import myutils
    got = myutils.energy_map_frames(_f4_clip(), 48000, 10.)
    assert got.shape == (25, 37, 72) and np.abs(got[:23] - G['f4_rms_frames']).max() < 1e-3


# ---- f2: SampleReader's chunk schedule and reader calls (feeder.py:164-278) -----------------------------------------------

def test_f2_sample_reader_schedule_and_calls_match_reference(tmp_path, monkeypatch):
    """readers.SampleReader on the audio_pow.lst the reference's SampleReader was run on, for the same argument sets
    (skip_rate, silence threshold, start / duration window, thread slices, other rates): identical chunk times, window
    sizes, sample ids and the same (start, size, rotation) requests to the audio / video / flow readers, in the same order.
    (One deliberate difference: with return_video=False the product does not open the video folder at all.)"""
    from spatialaudiogen_b200 import readers as R
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
    for n in ('AudioReader', 'VideoReader', 'FlowReader'):
        monkeypatch.setattr(R, n, type(n, (FakeReader,), {}))
    folder = str(tmp_path / 'clip_xyz')
    os.makedirs(folder)
    open(os.path.join(folder, 'audio_pow.lst'), 'w').write(str(G['f2_sched_audio_pow_lst']))
    cases = ast.literal_eval(str(G['f2_sched_cases']))
    assert len(cases) == 7
    for ref in cases:
        del calls[:]
        sr = R.SampleReader(folder, **ref['kwargs'])
        got = [sr.get() for _ in range(3)]
        assert sr.chunks_t == ref['chunks_t'], ref['kwargs']
        assert (sr.audio_size, sr.video_size) == (ref['audio_size'], ref['video_size'])
        assert [c['id'] if c else None for c in got] == ref['ids']
        want = [c for c in ref['calls'] if ref['kwargs'].get('return_video', True) or c[:2] != ('init', 'VideoReader')]
        assert calls == want, ref['kwargs']


def test_f2_video_reader_indexing_and_rotation_match_reference(tmp_path, monkeypatch):
    """readers.VideoReader.get_by_index on the frames the reference's VideoReader was run on (jpg decoding stood in on both
    sides): first frame = max(int(t * rate), 0), img_prep, single-frame axis, rotation as a roll of the width axis."""
    from spatialaudiogen_b200 import readers as R, myutils
    frames = G['f2_video_frames']
    for i in range(frames.shape[0]):
        open(str(tmp_path / ('%06d.jpg' % i)), 'w').close()
    monkeypatch.setattr(R, '_imread', lambda fn: frames[int(os.path.basename(fn)[:6])])
    vr = R.VideoReader(str(tmp_path), 10, myutils.img_prep_fcn())
    assert [vr.num_frames, vr.duration, vr.rate] + list(vr.frame_shape) == list(G['f2_video_meta'])
    for i, (t0, size, rot) in enumerate(ast.literal_eval(str(G['f2_video_cases']))):
        got = vr.get_by_index(t0, size, rot)
        assert got.shape == G['f2_video_out_%d' % i].shape and np.array_equal(got, G['f2_video_out_%d' % i]), (t0, size, rot)
