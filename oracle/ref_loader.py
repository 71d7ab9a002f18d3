"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference quantiser.

This module imports ``interactive_spectrogram_inpainting/vqvae/bottleneck.py``
straight from the read-only reference checkout (``/root/reference`` in the dev
container) so that (a) the restatement in ``oracle/quantizer_oracle.py`` can be
validated against the real thing and (b) ``oracle/make_golden.py`` can generate
the fixtures under ``tests/golden/``.

``/root/reference`` does NOT exist on the GPU box.  There the loader falls back to
``baseline/_ref`` -- byte-for-byte copies of the same files made in the dev container by
``oracle/stage_reference.py`` (git-ignored, shipped with the working tree, SHA-256 manifest)
-- so the GPU-side drop-in tests and ``bench.py --impl reference`` import the unmodified
reference classes too.  ``available()`` lets callers skip cleanly when neither exists.

The only thing we add is a one-class stub for the third-party module
``discretization`` (imported at bottleneck.py:27, used only by the
``QuantizedBottleneckWithRestarts`` class which is out of scope).
"""
import importlib.util
import os
import pathlib
import sys
import types

_STAGED_ROOT = pathlib.Path(__file__).resolve().parent.parent / "baseline" / "_ref"


def _find_root() -> pathlib.Path:
    env = os.environ.get("ISI_REFERENCE_ROOT")
    if env:
        return pathlib.Path(env)
    live = pathlib.Path("/root/reference")
    if (live / "interactive_spectrogram_inpainting" / "vqvae" / "bottleneck.py").is_file():
        return live
    return _STAGED_ROOT


REFERENCE_ROOT = _find_root()
_BOTTLENECK = (REFERENCE_ROOT / "interactive_spectrogram_inpainting" / "vqvae"
               / "bottleneck.py")

_cached = None


def available() -> bool:
    return _BOTTLENECK.is_file()


def load_reference_bottleneck():
    """Return the reference ``bottleneck`` module (bottleneck.py:1-166)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not present at {_BOTTLENECK}")
    import torch

    if "discretization" not in sys.modules:
        stub = types.ModuleType("discretization")
        stub.ProductVectorQuantizer = type(
            "ProductVectorQuantizer", (torch.nn.Module,), {})
        sys.modules["discretization"] = stub
    spec = importlib.util.spec_from_file_location("_isi_ref_bottleneck",
                                                  str(_BOTTLENECK))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def _stub_module(name):
    """A module whose every attribute is a dummy nn.Module subclass."""
    import torch

    class _Any(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            if f"{self.__name__}.{item}" in sys.modules:
                return sys.modules[f"{self.__name__}.{item}"]
            if item == "delegates":
                return lambda *a, **k: (lambda f: f)
            if item == "defaults":
                return types.SimpleNamespace(activation=torch.nn.ReLU)
            cls = type(item, (torch.nn.Module,), {})
            setattr(self, item, cls)
            return cls

    m = _Any(name)
    m.__path__ = []
    return m


def load_reference_vqvae_class():
    """Return the reference ``VQVAE`` class (vqvae.py:36), importing the package
    from the read-only checkout with stubs for the un-vendored third-party
    modules it names at import time (vqvae.py:11-13, encoder_decoder.py:8-15)."""
    if not available():
        raise FileNotFoundError("reference not present")
    load_reference_bottleneck()
    for name in ("GANsynth_pytorch", "GANsynth_pytorch.loader",
                 "GANsynth_pytorch.normalizer",
                 "GANsynth_pytorch.spectrograms_helper", "fastai",
                 "fastai.vision", "fastai.vision.models",
                 "fastai.vision.models.unet", "fastai.vision.models.xresnet",
                 "fastai.layers", "fastai.torch_core", "fastai.callback",
                 "fastai.callback.hook"):
        if name not in sys.modules:
            sys.modules[name] = _stub_module(name)
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.append(root)
    from interactive_spectrogram_inpainting.vqvae.vqvae import VQVAE
    return VQVAE
