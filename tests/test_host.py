"""CPU tests of the host-side mirror of the reference interface that need no GPU: parameter parsing with the
reference's fall-backs (myutils.py:40-85), constants (definitions.py), checkpoint layout helpers."""
import numpy as np

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
