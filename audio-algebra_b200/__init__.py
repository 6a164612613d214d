"""audio-algebra hot path on B200: same Python surface as the reference's `audio_algebra` package
(given_models encoders, aa_mixer / aa_effects projector + mixing + losses), backed by
hand-written sm_100a CUDA kernels in libaa_b200.so (C ABI: include/aa_b200.h).

There is no CPU or PyTorch fallback: importing the compute modules without the built library, or
calling them without a B200, raises."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (loads libaa_b200.so or raises)
from .given_models import (GivenModelClass, SpectrogramAE, MagSpectrogramAE, MagDPhaseSpectrogramAE,  # noqa: F401
                           MelSpectrogramAE, DVAEWrapper, StackedDiffAEWrapper, encode_all)
from .aa_mixer import (EmbedBlock, AudioAlgebra, get_stems_faders, do_mixing, mseloss, vicreg_var_loss,  # noqa: F401
                       vicreg_var_loss_l2, vicreg_cov_loss, off_diagonal, latent_lincomb)
from .DiffusionDVAE import DiffusionDVAE, SoundStreamXLEncoder  # noqa: F401
from . import aa_mixer, aa_effects, latent_ops, pca, given_models, parallel, training, StackedAELatentDiffusionCond  # noqa: F401
