"""ctypes binding of libsag.so (include/sag.h).  PyTorch tensors are only the buffer type: every call passes raw
device pointers + the current CUDA stream.  There is no CPU fallback: a missing library raises at import of the
first op, and every non-zero return code raises with the library's message."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libsag.so')

SAG_PREC_FP32, SAG_PREC_BF16, SAG_PREC_BF16X3, SAG_PREC_MIXED = 0, 2, 3, 4
PRECISIONS = {'fp32': SAG_PREC_FP32, 'bf16': SAG_PREC_BF16, 'bf16x3': SAG_PREC_BF16X3, 'mixed': SAG_PREC_MIXED}
SAG_FRAMES_F32, SAG_FRAMES_U8 = 0, 1
SAG_SEP_NONE, SAG_SEP_UNET_MASK = 0, 1


def default_precision():
    """Arithmetic of the dense contractions when the caller does not choose: 'bf16x3' = tcgen05 tensor cores with
    every fp32 operand split into bf16 hi+lo (3 MMAs per K step, fp32 accumulate in TMEM) -- meets the <= 1e-3 waveform
    parity of the fp32 reference; 'fp32' is the exact FFMA path, 'bf16' the single-MMA fast mode."""
    return os.environ.get('SAG_PRECISION', 'bf16x3')
SAG_EINVAL, SAG_ECUDA, SAG_ENOMEM, SAG_ESTATE, SAG_EUNSUPPORTED = -1, -2, -3, -4, -5


class sag_config(C.Structure):
    _fields_ = [('ambi_order', C.c_int32), ('audio_rate', C.c_int32), ('video_rate', C.c_int32),
                ('context', C.c_double), ('sample_duration', C.c_double),
                ('enc_audio', C.c_int32), ('enc_video', C.c_int32), ('enc_flow', C.c_int32),
                ('separation', C.c_int32), ('sep_num_tracks', C.c_int32), ('n_loc_fc', C.c_int32),
                ('loc_fc_units', C.c_int32 * 4), ('sep_fft_window', C.c_double), ('precision', C.c_int32),
                ('frame_h', C.c_int32), ('frame_w', C.c_int32)]


class sag_dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ('snd_contx', 'snd_dur', 'snd_size', 'wind_size', 'num_ambi_channels', 'n_stft_frames', 'enc_ss',
                 'enc_tt', 'mask_ss', 'mask_tt', 'mask_skip', 'final_crop', 'feat_dim')]


_P, _F, _I, _L, _S = C.c_void_p, C.c_float, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); mirrors include/sag.h one to one (tests check every symbol is exported)
PROTOTYPES = {
    'sag_last_error': (C.c_char_p, []),
    'sag_version': (C.c_char_p, []),
    'sag_config_default': (_I, [C.POINTER(sag_config)]),
    'sag_crc32c': (C.c_uint32, [_P, _S, C.c_uint32]),
    'sag_create': (_I, [C.POINTER(_P), C.POINTER(sag_config)]),
    'sag_destroy': (_I, [_P]),
    'sag_get_dims': (_I, [_P, C.POINTER(sag_dims)]),
    'sag_set_option': (_I, [_P, C.c_char_p, _I]),
    'sag_load_weight': (_I, [_P, C.c_char_p, _P, C.POINTER(_L), _I]),
    'sag_num_weights_expected': (_I, [_P]),
    'sag_weight_name': (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_I)]),
    'sag_finalize_weights': (_I, [_P, _P]),
    'sag_workspace_bytes': (_S, [_P, _I]),
    'sag_forward': (_I, [_P, _P, _P, _P, _P, _P, _S, _I, _P]),
    'sag_forward_frames': (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _S, _I, _P]),
    'sag_get_tensor': (_I, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_L), C.POINTER(_I), C.POINTER(_L)]),
    'sag_get_tensor_format': (_I, [_P, C.c_char_p, C.POINTER(_I), C.POINTER(_L)]),
    'sag_num_tensors': (_I, [_P]),
    'sag_tensor_name': (_I, [_P, _I, C.c_char_p, _I]),
    'sag_last_launch_count': (_I, [_P]),
    'sag_plan_contraction': (_I, [_I, _I, _L, C.POINTER(_I), C.POINTER(_I)]),
    'sag_plan_stream_k': (_I, [_I, _I, _L]),
    'sag_stream_k_schedule': (_I, [_L, _I, _I, _I, C.POINTER(_I), _I]),
    'sag_get_profile': (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I)]),
    'sag_num_profile_records': (_I, [_P]),
    'sag_get_profile_record': (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(_I)]),
    'sag_stft': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P]),
    'sag_istft': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'sag_conv2d': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    'sag_deconv2d': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    'sag_fc': (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _I, _P]),
    'sag_batchnorm_train': (_I, [_P, _L, _I, _P, _P, _P, _I, _P, _P, _P]),
    'sag_maxpool_3x3s2_same': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'sag_resnet18': (_I, [_P, C.c_char_p, _P, _I, _P, _P, _S, _P]),
    'sag_mix': (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    'sag_metrics_scratch_bytes': (_S, [_I, _I]),
    'sag_metrics': (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'sag_sh_rms_dims': (_I, [_F, C.POINTER(_I), C.POINTER(_I)]),
    'sag_sh_rms': (_I, [_P, _I, _I, _F, _P, _P]),
    'sag_mel_lsd': (_I, [_P, _P, _I, _I, _I, _P, _P]),
    'sag_emd_hat': (_I, [_P, _P, _I, _P, C.c_double, _I, _P]),
    'sag_jpeg_info': (_I, [_P, _S, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    'sag_jpeg_coefficients': (_I, [_P, _S, _P, _S, C.POINTER(_I), C.POINTER(_I), _P]),
    'sag_jpeg_create': (_I, [C.POINTER(_P), _I, _I, _I]),
    'sag_jpeg_destroy': (None, [_P]),
    'sag_jpeg_decode': (_I, [_P, C.POINTER(_P), C.POINTER(_S), _I, _P, _I, _P]),
    'sag_jpeg_set_option': (_I, [_P, C.c_char_p, _I]),
    'sag_jpeg_sync_rounds': (_I, [_P, C.POINTER(_I), _I]),
    'sag_jpeg_coefficients_parallel': (_I, [_P, _S, _I, _P, _S, C.POINTER(_I)]),
}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if libsag.so has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libsag.so is not built (%s): run `python -m spatialaudiogen_b200.build` -- '
                               'this package has no CPU or PyTorch fallback' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class SagError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, 'libsag error %d: %s' % (code, msg))
        self.code = code


def check(code):
    if code != 0:
        msg = lib().sag_last_error().decode('utf-8', 'replace')
        if code == SAG_EINVAL:
            raise ValueError('libsag: ' + msg)
        raise SagError(code, msg)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError('expected a CUDA tensor (libsag has no CPU path), got %r' % (type(t),))
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return C.c_void_p(t.data_ptr())


def f32(t, device=None):
    """float32 contiguous CUDA view/copy of a tensor or array-like."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if device is None:
        device = t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float32).contiguous()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
