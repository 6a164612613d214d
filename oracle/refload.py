"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference modules.

Only usable in the build container, where /root/reference is mounted.  It is used by
oracle/make_golden.py to generate the committed fixtures under tests/golden/ and by
tests that pin the oracle restatement against the live reference (skipped when the
mount is absent, e.g. on the GPU box).  Nothing in the product package imports this.

The reference's modules import third-party packages that are not installed here
(pytorch_lightning, laion_clap, aeiou, audio-diffusion, ... -- SURVEY.md section 8c).
We pre-populate sys.modules with permissive stubs so that the in-tree arithmetic
(given_models.py STFT wrappers, aa_mixer.py / aa_effects.py projector + losses)
imports and runs unmodified.
"""
import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("AA_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """Module whose unknown attributes resolve to child stubs / dummy callables."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _StubObj(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child


class _StubObj:
    def __init__(self, name="stub", *a, **k):
        self._name = name

    def __call__(self, *a, **k):
        return _StubObj(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubObj(f"{self._name}.{name}")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


_STUB_MODULES = [
    "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.distributed",
    "pytorch_lightning.callbacks", "laion_clap", "laion_clap.training", "laion_clap.training.data",
    "aeiou", "aeiou.core", "aeiou.hpc", "aeiou.viz", "aeiou.datasets",
    "autoencoders", "autoencoders.models", "autoencoders.soundstream", "nwt_pytorch",
    "diffusion", "diffusion.pqmf", "diffusion.model", "encoders", "encoders.encoders",
    "decoders", "decoders.diffusion_decoder", "dvae", "dvae.residual_memcodes",
    "audio_encoders_pytorch", "ema_pytorch", "audio_diffusion_pytorch",
    "audio_diffusion_pytorch.modules", "k_diffusion", "prefigure", "prefigure.prefigure",
    "IPython", "IPython.display", "matplotlib", "matplotlib.pyplot", "accelerate", "rave", "gin",
    "webdataset", "fastcore", "fastcore.utils", "wandb", "tqdm", "tqdm.auto", "gdown",
    "a_unet", "a_unet.apex", "pedalboard", "pyloudnorm", "blocks", "blocks.utils",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "audio_algebra"))


def install_stubs():
    import torch.nn as nn

    for name in _STUB_MODULES:
        if name in sys.modules and not isinstance(sys.modules[name], _Stub):
            continue  # a real module is installed; keep it
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        sys.modules[name] = _Stub(name)
    for name in _STUB_MODULES:  # wire parents -> children
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _Stub):
                setattr(sys.modules[parent], child, sys.modules[name])
    pl = sys.modules["pytorch_lightning"]
    if isinstance(pl, _Stub):
        pl.LightningModule = nn.Module
        pl.Callback = object
    # audiomentations: datasets.py:48 uses these names as default args at class-definition time
    if "audiomentations" not in sys.modules:
        am = types.ModuleType("audiomentations")

        def _mk(n):
            return type(n, (), {"__init__": lambda self, p=1.0, **kw: None})

        names = ["Gain", "BandPassFilter", "BandStopFilter", "HighPassFilter", "LowPassFilter"]
        for n in names:
            setattr(am, n, _mk(n))
        am.__all__ = names
        sys.modules["audiomentations"] = am


_loaded = {}


def load_reference():
    """Returns dict(given_models=..., aa_mixer=..., aa_effects=...) of the reference's own modules."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    # GivenModelClass.__init__ creates ~/checkpoints (given_models.py:69-70): point HOME at a temp dir
    os.environ["HOME"] = tempfile.mkdtemp(prefix="aa_ref_home_")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _loaded["given_models"] = importlib.import_module("audio_algebra.given_models")
    _loaded["aa_mixer"] = importlib.import_module("audio_algebra.aa_mixer")
    try:
        _loaded["aa_effects"] = importlib.import_module("audio_algebra.aa_effects")
    except Exception as e:  # pragma: no cover - reported by make_golden
        _loaded["aa_effects_error"] = repr(e)
    return _loaded
