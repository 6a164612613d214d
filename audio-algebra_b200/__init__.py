"""audio-algebra hot path on B200: same Python surface as the reference's `audio_algebra` package
(given_models encoders, aa_mixer / aa_effects projector + mixing + losses), backed by
hand-written sm_100a CUDA kernels in libaa_b200.so (C ABI: include/aa_b200.h).

There is no CPU or PyTorch fallback: importing the compute modules without the built library, or
calling them without a B200, raises."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (loads libaa_b200.so or raises)
from .given_models import (GivenModelClass, SpectrogramAE, MagSpectrogramAE, MagDPhaseSpectrogramAE,  # noqa: F401
                           MelSpectrogramAE)
