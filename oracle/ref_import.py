"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference tree.

The reference (QianWangX/VidSeg_diffusion @ 3e96366) lives read-only at
/root/reference in the authoring container and does NOT exist on the GPU box;
``oracle/make_ref.py`` mirrors its ``sgm`` / ``scripts`` packages into the
git-ignored ``oracle/_ref/`` so that they travel there.  This module imports the
reference's own modules from whichever of the two exists.  Users: the
golden-vector generators under ``tests/golden/make_*.py``, ``oracle/reference_path.py``
(the reference arm / CPU baseline of bench.py and the full-size parity tests).

Packages the reference imports at module scope but which are absent from this
image are replaced by inert ``sys.modules`` stubs:
pytorch_lightning, omegaconf, kornia, open_clip, torchdata, webdataset,
matplotlib, xformers (absent -> the reference falls back to its SDPA path).
Nothing here is imported by the product package.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    env = os.environ.get("VIDSEG_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if os.path.isdir(os.path.join(cand, "sgm")) and os.path.isdir(os.path.join(cand, "scripts")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sgm"))


class _Anything:
    """Attribute sink: any attribute / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import a.b" works

    def _getattr(attr, _m=m):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything

    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


def install_stubs():
    import torch.nn as nn

    _stub("pytorch_lightning", LightningModule=nn.Module, LightningDataModule=object)
    _stub("omegaconf", ListConfig=list, DictConfig=dict, OmegaConf=_Anything)
    for name in (
        "kornia", "kornia.geometry", "open_clip", "torchdata", "torchdata.datapipes",
        "torchdata.datapipes.iter", "webdataset", "matplotlib", "matplotlib.pyplot",
        "fire", "imwatermark", "clip", "streamlit",
    ):
        _stub(name)


def import_reference(module: str):
    """Import ``module`` (e.g. ``sgm.modules.attention``) from the reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(module)
