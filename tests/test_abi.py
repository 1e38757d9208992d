"""CPU checks of the drop-in boundary: libsag.so builds/loads here (nvcc cross-compiles without a GPU) and exports every
symbol include/sag.h declares with the prototype the ctypes binding expects; no compute call is made."""
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'sag.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sag_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_bound_and_exported():
    from spatialaudiogen_b200 import _lib as L
    from spatialaudiogen_b200 import build as B
    B.build()
    declared = _declared()
    assert len(declared) >= 30
    assert sorted(L.PROTOTYPES) == declared, set(declared) ^ set(L.PROTOTYPES)
    out = subprocess.check_output(['nm', '-D', '--defined-only', L.LIB_PATH]).decode()
    exported = set(re.findall(r' T (sag_[a-z0-9_]+)', out))
    assert set(declared) <= exported, set(declared) - exported
    lib = L.lib()                       # dlopen + prototypes
    for name in declared:
        assert hasattr(lib, name)
    assert b'sm_100a' in lib.sag_version()


def test_config_struct_layout_matches_header():
    from spatialaudiogen_b200 import _lib as L
    import ctypes as C
    cfg = L.sag_config()
    assert L.lib().sag_config_default(C.byref(cfg)) == 0
    assert (cfg.ambi_order, cfg.audio_rate, cfg.video_rate, cfg.sep_num_tracks) == (1, 48000, 10, 32)
    assert (cfg.context, cfg.sample_duration, cfg.sep_fft_window) == (1.0, 0.1, 0.025)
    assert list(cfg.loc_fc_units)[:2] == [512, 512] and (cfg.frame_h, cfg.frame_w) == (224, 448)


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'spatialaudiogen_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'sag_oracle' not in text and 'import oracle' not in text and 'from oracle' not in text, fn
    import torch
    if not torch.cuda.is_available():
        from spatialaudiogen_b200 import SptAudioGen
        with pytest.raises(RuntimeError):
            SptAudioGen(1, encoders=['audio'], separation='unet_mask')
