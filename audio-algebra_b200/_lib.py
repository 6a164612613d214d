"""ctypes binding of libaa_b200.so (declarations mirror include/aa_b200.h one to one)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AA_B200_LIB") or os.path.join(_HERE, "libaa_b200.so")   # AA_B200_LIB: dev override (variant builds)


class AaError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python audio-algebra_b200/build.py` "
        "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this package.")

lib = C.CDLL(LIB_PATH)

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

_SIGS = {
    "aa_version": (_i, []),
    "aa_last_error": (C.c_char_p, []),
    "aa_check_device": (_i, []),
    "aa_launch_count": (_i64, []),
    "aa_stft_plan_create": (_i, [C.POINTER(_p), _i, _i, _i, _p, _i, _f, _f, _f, _p]),
    "aa_stft_plan_destroy": (_i, [_p]),
    "aa_stft_out_shape": (_i, [_p, _i64, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "aa_stft_complex_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_power_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_complex_tf_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_power_tf_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_mel_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_mel_tf_f32": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "aa_stft_mel_tf_supported": (_i, [_p]),
    "aa_stft_mel_tf_f32_host": (_i, [_p, _p, _i64, _i64, _i, _p, _i64]),
    "aa_magdphase_f32": (_i, [_p, _i64, _i64, _i64, _p, _p]),
    "aa_magdphase_ex_f32": (_i, [_p, _i64, _i64, _i64, _i, _i, _p, _p]),
    "aa_istft_workspace_floats": (_i64, [_i64, _i, _i64]),
    "aa_istft_f32": (_i, [_p, _i64, _i, _i, _i, _i64, _i64, _i64, _i64, _p, _p, _i64, _p, _p]),
    "aa_griffinlim_update_c64": (_i, [_p, _p, _p, _p, _i64, C.c_float, _i, _p]),
    "aa_inverse_mel_f32": (_i, [_p, _p, _i64, _i, _i, _i64, _i64, _i64, _i64, _p, _p]),
    "aa_magdphase_decode_f32": (_i, [_p, _i64, _i64, _i64, _i, _p, _p, _p]),
    "aa_stft_mel_f32_host": (_i, [_p, _p, _i64, _i64, _i, _p, _i64]),
}

EXPORTS = tuple(_SIGS)


def _declare(extra=None):
    for name, (res, args) in (extra or _SIGS).items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args


_declare()


def register(sigs):
    "later modules (losses, projector, encoder) add their entry points here"
    _SIGS.update(sigs)
    _declare(sigs)


def check(rc):
    if rc != 0:
        raise AaError(f"libaa_b200 error {rc}: {lib.aa_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise AaError(f"{what} must live on a CUDA device (got {t.device}); this package has no CPU path")


_device_ok = {}


def ensure_device(device=None):
    """Raises unless `device` (default: current) is a CUDA device this library can run on."""
    if not torch.cuda.is_available():
        raise AaError("no CUDA device: audio_algebra_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _device_ok:
        with torch.cuda.device(idx):
            check(lib.aa_check_device())
        _device_ok[idx] = True
    return idx
