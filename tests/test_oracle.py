"""CPU tests that pin the oracle: analytic invariants, the ResNet-18 known-answer fixture,
fp64-vs-fp32 self agreement.  The reference has no tests of its own for this path (parity unpinned)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sag_oracle as O
from spatialaudiogen_b200 import weights as Wt

HERE = os.path.dirname(__file__)


def _audio(B, seed=0, n=52799):
    rng = np.random.RandomState(seed)
    t = np.arange(n)
    ph = rng.uniform(0, 2 * np.pi, size=(B, 1))
    x = 0.1 * rng.randn(B, n) + 0.3 * np.sin(2 * np.pi * 440 * t / 48000. + ph)
    return np.clip(x, -1, 1).astype(np.float32)[:, :, None]


def test_derived_constants():
    m = O.SptAudioGen({}, encoders=['audio'], separation='unet_mask')
    assert (m.snd_contx, m.snd_dur, m.snd_size, m.wind_size, m.num_ambi_channels) == (48000, 4800, 52799, 1024, 4)
    assert m.encoder_crop() == (46, 173)
    assert m.mask_crop() == (89, 117, 46)
    assert m.final_crop() == 448


def test_stft_frame_indexing_bit_exact():
    """frame t covers samples [256t, 256t+1024); 200 frames; last sample read is 51967."""
    x = torch.zeros(1, 1, 52799)
    s = O.stft(x, 1024, 4)
    assert tuple(s.shape) == (1, 1, 200, 1024)
    for n in [0, 255, 256, 1023, 1024, 30000, 51967, 51968, 52798]:
        x = torch.zeros(1, 1, 52799)
        x[0, 0, n] = 1.0
        e = O.stft(x, 1024, 4).abs().sum(-1)[0, 0]          # energy per frame
        hit = set(torch.nonzero(e > 0).flatten().tolist())
        w = O.hann(1024, torch.float32)
        exp = {t for t in range(200) if 256 * t <= n < 256 * t + 1024 and w[n - 256 * t] > 0}
        assert hit == exp, (n, hit, exp)


def test_stft_matches_direct_dft_and_sinusoid_peak():
    x = torch.as_tensor(_audio(1, 3)[:, :, 0])[:, None, :].double()
    s = O.stft(x, 1024, 4)
    w = torch.as_tensor(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024))
    for t in (0, 46, 117, 199):
        ref = torch.fft.fft(x[0, 0, 256 * t:256 * t + 1024] * w)
        assert torch.allclose(s[0, 0, t], ref, atol=1e-9)
    k, A = 37, 0.7
    n = torch.arange(52799).double()
    x = (A * torch.cos(2 * np.pi * k * n / 1024))[None, None]
    mag = O.stft(x, 1024, 4).abs()[0, 0, 10]
    assert abs(mag[k] - 256 * A) < 1e-6 and abs(mag[1024 - k] - 256 * A) < 1e-6
    assert int(mag.argmax()) in (k, 1024 - k)


def test_istft_of_stft_gain_half_and_crop():
    """iSTFT(STFT(x)) with unit mask = 0.5*x[23552:29952]; final crop [448:5248] = 0.5*x[24000:28800]."""
    x = torch.as_tensor(_audio(2, 1)[:, :, 0]).double()[:, None, :]
    s = O.stft(x, 1024, 4)[:, :, 89:117]
    y = O.istft(s.unsqueeze(2), 4)                         # (B,1,1,6400)
    assert tuple(y.shape) == (2, 1, 1, 6400)
    assert torch.allclose(y[:, 0, 0], 0.5 * x[:, 0, 23552:29952], atol=1e-12)
    assert torch.allclose(y[:, 0, 0, 448:5248], 0.5 * x[:, 0, 24000:28800], atol=1e-12)


def test_real_ifft_of_masked_equals_irfft_of_symmetrised_mask():
    g = torch.Generator().manual_seed(0)
    S = torch.fft.fft(torch.randn(4, 1024, generator=g, dtype=torch.float64))
    m = torch.rand(4, 1024, generator=g, dtype=torch.float64)
    ref = torch.fft.ifft(S * m).real
    idx = (-torch.arange(1024)) % 1024
    msym = 0.5 * (m + m[:, idx])
    out = torch.fft.irfft((S * msym)[:, :513], n=1024)
    assert torch.allclose(ref, out, atol=1e-12)


def test_zero_decoder_gives_quarter_mono_and_zero_fc3_gives_bias():
    W = Wt.init_weights(['audio'], stress=True, seed=3)
    for k in W:
        if k.startswith('separation/deconv'):
            W[k] = np.zeros_like(W[k])
    a = _audio(2, 5)
    m = O.SptAudioGen(W, encoders=['audio'], separation='unet_mask', dtype=torch.float64)
    m.inference_ops(a)
    mono = torch.as_tensor(a[:, 24000:28800, 0]).double()
    assert torch.allclose(m.sep_channels[:, 0], 0.25 * mono[:, None, :].expand(-1, 32, -1), atol=1e-12)
    W['localization/fc3/weights'] = np.zeros_like(W['localization/fc3/weights'])
    m = O.SptAudioGen(W, encoders=['audio'], separation='unet_mask', dtype=torch.float64)
    y = m.inference_ops(a)
    b3 = torch.as_tensor(W['localization/fc3/biases']).double().reshape(3, 33)
    exp = b3[:, :32].sum(1)[None, None, :] * 0 + (b3[:, :32][None, None] * (0.25 * mono)[:, :, None, None]).sum(-1) + b3[:, 32]
    assert torch.allclose(y, exp, atol=1e-10)


def test_tf_same_padding_table():
    """conv1 7x7/2 at 224x448 pads (2,3,2,3); 3x3/2 pads (0,1,0,1); 3x3/1 pads (1,1)."""
    assert O._same_pads(224, 7, 2) == (2, 3) and O._same_pads(448, 7, 2) == (2, 3)
    assert O._same_pads(56, 3, 2) == (0, 1) and O._same_pads(112, 3, 2) == (0, 1)
    assert O._same_pads(56, 3, 1) == (1, 1)
    assert O._same_pads(112, 3, 2) == (0, 1)       # max-pool 112->56


def test_conv_transpose_definition():
    """y[b,i*sh+p,j*sw+q,co] += x[b,i,j,ci]*w[p,q,co,ci] against a literal loop."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 4, 2, generator=g, dtype=torch.float64)
    w = torch.randn(3, 5, 3, 2, generator=g, dtype=torch.float64)
    sh, sw = 2, 2
    y = torch.zeros(1, (3 - 1) * sh + 3, (4 - 1) * sw + 5, 3, dtype=torch.float64)
    for i in range(3):
        for j in range(4):
            for p in range(3):
                for q in range(5):
                    y[0, i * sh + p, j * sw + q] += w[p, q] @ x[0, i, j]
    assert torch.allclose(O.tf_conv2d_transpose_valid(x, w, (sh, sw)), y, atol=1e-12)


def test_shapes_full_model_and_param_count():
    W = Wt.init_weights(['audio', 'video', 'flow'])
    assert len(W) == 214 and Wt.num_params(W) == 49005763
    W = Wt.init_weights(['audio', 'video'], stress=True)
    m = O.SptAudioGen(W, encoders=['audio', 'video'], separation='unet_mask')
    v = np.random.RandomState(1).rand(2, 1, 224, 448, 3).astype(np.float32) - 0.5
    y = m.inference_ops(_audio(2), video=v)
    assert tuple(y.shape) == (2, 4800, 3)
    assert [tuple(t.shape[1:]) for t in m.ends['audio_encoder']] == \
        [(127, 1024, 1), (31, 127, 32), (15, 31, 64), (7, 14, 128), (5, 10, 256), (3, 6, 512)]
    assert tuple(m.ends['video_encoder/conv5_2'].shape) == (2, 7, 14, 512)
    assert tuple(m.ends['bottleneck'].shape) == (2, 3, 1536)


def test_fp32_vs_fp64_agreement_bounds_tolerance():
    W = Wt.init_weights(['audio'], stress=True, seed=7)
    a = _audio(2, 9)
    y32 = O.SptAudioGen(W, encoders=['audio'], separation='unet_mask').inference_ops(a).double()
    y64 = O.SptAudioGen(W, encoders=['audio'], separation='unet_mask', dtype=torch.float64).inference_ops(a)
    rel = (y32 - y64).abs().max() / y64.abs().max()
    assert rel < 1e-4, rel


def test_sh_matrix_closed_form_and_mesh():
    phi, nu = O.spherical_mesh(30.)
    assert phi.shape == (7, 12)
    assert np.isclose(phi[0, 0], 150 / 180. * np.pi) and np.isclose(phi[0, -1], -np.pi)
    Y = O.spherical_harmonics_matrix(phi.reshape(-1), nu.reshape(-1), 1)
    ref = np.stack([np.ones(84), np.sin(phi) .reshape(-1) * np.cos(nu).reshape(-1), np.sin(nu).reshape(-1),
                    np.cos(phi).reshape(-1) * np.cos(nu).reshape(-1)], 1)
    assert np.allclose(Y, ref, atol=1e-12)


def test_metrics_shapes_and_identities():
    m = O.SptAudioGen({}, encoders=['audio'], separation='unet_mask', dtype=torch.float64)
    rng = np.random.RandomState(0)
    gt = rng.randn(4, 4800, 3) * 0.1
    metrics, stft_ps, lsd_ps, mse_ps, snr_ps = m.evaluation_ops(gt, gt, None, np.ones((4, 3)))
    assert tuple(stft_ps.shape) == tuple(lsd_ps.shape) == tuple(mse_ps.shape) == tuple(snr_ps.shape) == (4, 3)
    assert float(stft_ps.abs().max()) == 0 and float(lsd_ps.abs().max()) == 0 and float(mse_ps.abs().max()) == 0
    ps = torch.as_tensor((gt ** 2).sum(1))
    assert torch.allclose(snr_ps, 10 * torch.log10((ps + 0.1) / 0.1))
    pred = gt + 0.01
    _, _, _, mse_ps, _ = m.evaluation_ops(pred, gt, None, np.ones((4, 3)))
    assert torch.allclose(mse_ps, torch.full((4, 3), 1e-4, dtype=torch.float64))
    assert tuple(O.stft_for_loss(torch.zeros(2, 4800, 3), 1200, 2).shape) == (2, 3, 3, 2048)
    assert tuple(O.stft(torch.zeros(2, 3, 4800), 1200, 2).shape) == (2, 3, 6, 1200)


def test_envelope_distance_of_am_tone():
    n = np.arange(4800)
    env = 0.5 + 0.3 * np.cos(2 * np.pi * 5 * n / 4800)
    x = env * np.cos(2 * np.pi * 600 * n / 4800)
    d = O.compute_envelope_dist(np.stack([x] * 3, 1), np.zeros((4800, 3)))
    assert np.allclose(d, np.sqrt(np.mean(env ** 2)), rtol=1e-6)


def test_deploy_assembly_rows_and_zero_padded_tail():
    W = Wt.init_weights(['audio'], stress=True, seed=11)
    m = O.SptAudioGen(W, encoders=['audio'], separation='unet_mask')
    a = _audio(3, 2)
    out = O.deploy_assemble(m, a, batch_size=2)
    assert out.shape == (3 * 4800, 4) and out.dtype == np.float64
    assert np.array_equal(out[:, 0], a[:, 24000:28800, 0].reshape(-1).astype(np.float64))


def test_resnet_known_answer_fixture():
    """Semantic KAT of the oracle's ResNet-18 with the reference's own weights/images (container only)."""
    fx = json.load(open(os.path.join(HERE, 'golden', 'resnet18_kat.json')))
    expect = {'cat.jpg': 'cat', 'poodle.png': 'poodle', 'tiger.jpeg': 'tiger', 'puzzle.jpeg': 'puzzle',
              'laska.png': 'weasel', 'dog.png': 'dog'}
    for fn, word in expect.items():
        assert any(word in n for n in fx[fn]['top5_names']), (fn, fx[fn]['top5_names'])
    ref = '/root/reference/pyutils/tflib/models/image'
    if not os.path.exists(os.path.join(ref, 'resnet18.npy')):
        pytest.skip('reference assets not on this box; fixture content checked above')
    import importlib.util
    spec = importlib.util.spec_from_file_location('mk', os.path.join(HERE, 'golden', 'make_resnet_kat.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    from PIL import Image
    pre = np.load(os.path.join(ref, 'resnet18.npy'), allow_pickle=True, encoding='latin1').item()
    img = np.array(Image.open(os.path.join(ref, 'test_images', 'tiger.jpeg')).convert('RGB'))
    x = mk.resize_bilinear_tf(mk.central_crop(img, 0.875), 224, 224) / 255.
    x = (x - torch.tensor([0.485, 0.456, 0.406])) / torch.tensor([0.229, 0.224, 0.225])
    logits, _ = O.resnet18(O.Weights(pre), '', x[None], bn_train=False, truncate_at=None)
    assert torch.argsort(-logits[0])[:5].tolist() == fx['tiger.jpeg']['top5']


def test_mel_lsd_oracle_against_an_independent_librosa_compatible_implementation():
    """myutils.compute_lsd_dist calls librosa.feature.melspectrogram (librosa 0.6.0, absent here).  The oracle's restatement
    (Slaney mel filter bank, centred reflect-padded periodic-Hann power spectrogram) is cross-checked against the independent
    librosa-compatible implementation that ships with `transformers` (audio_utils.mel_filter_bank(norm='slaney',
    mel_scale='slaney') + audio_utils.spectrogram): filter bank to 1e-12, distances to 1e-6."""
    au = pytest.importorskip('transformers.audio_utils')
    sr, n_fft = 48000, 2048
    mine = O._mel_filter_bank(sr, n_fft, 128, 0.0, 12000.0)
    theirs = au.mel_filter_bank(1 + n_fft // 2, 128, 0.0, 12000.0, sr, norm='slaney', mel_scale='slaney')
    assert theirs.shape == (1025, 128) and np.abs(mine - theirs.T).max() < 1e-12 and mine.max() > 1e-3
    rng = np.random.RandomState(0)
    y = 0.1 * rng.randn(4800) + 0.3 * np.sin(2 * np.pi * 440 * np.arange(4800) / sr)
    pred = np.stack([y, 0.9 * y, y + 0.01 * rng.randn(4800)], 1)
    gt = np.stack([1.1 * y, y, y], 1)
    win = au.window_function(n_fft, 'hann', periodic=True)

    def mel(v):
        return au.spectrogram(v.astype(np.float64), win, n_fft, 512, fft_length=n_fft, power=2.0, center=True, pad_mode='reflect',
                              mel_filters=theirs, mel_floor=0.0, dtype=np.float64)

    def ps(x):
        return 10 * np.log(np.abs(x) + 1e-2) / np.log(10.)

    ref = np.array([np.sqrt(np.mean((ps(mel(gt[:, i])) - ps(mel(pred[:, i]))) ** 2)) for i in range(3)])
    assert np.abs(O.compute_lsd_dist(pred, gt, sr) - ref).max() < 1e-6
