"""Import shim that lets the UNMODIFIED reference (``/root/reference``) run on CPU in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``shapeformer_b200/`` may import this module; it is used by
``tests/golden/make_golden.py`` (fixture generation) and by the ``-m "not gpu"`` tests that pin ``oracle/sf_oracle.py``
against the reference's own modules, and by ``bench.py``'s CPU legs (``--impl reference`` / ``cpu_baseline``).
``/root/reference`` does not exist on the GPU box: there the shim resolves to ``oracle/_ref`` (a git-ignored copy of the
reference's ``*.py`` made by ``oracle/make_ref.py`` at build time); when neither exists ``available()`` returns False.

The reference imports a number of packages that are not installed here (``pytorch_lightning``, ``torch_scatter``,
``h5py``, ``igl``, ``mcubes``, ``fresnel``, ``matplotlib``, ``skimage`` ...), none of which is touched by the hot path
(SURVEY.md App. B).  They are replaced by inert mock modules; ``pytorch_lightning.LightningModule`` becomes a plain
``nn.Module`` and ``torch_scatter`` is restated with ``Tensor.scatter_reduce``.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

def _find_root():
    """$SFB200_REFERENCE_ROOT, else /root/reference (build container), else oracle/_ref (the git-ignored copy made by
    oracle/make_ref.py, which is what exists on the GPU box)."""
    env = os.environ.get("SFB200_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/shapeformer"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_root()

_MOCK_ROOTS = ("h5py", "igl", "mcubes", "fresnel", "matplotlib", "mpl_toolkits", "skimage", "trimesh", "plyfile",
               "open3d", "seaborn", "pathos", "bashlex", "imageio", "PIL", "cv2", "sklearn", "wandb", "omegaconf")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "shapeformer"))


class _Mock(types.ModuleType):
    """A module whose every attribute is another mock (callable, subclassable, iterable-empty)."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__all__ = []

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        obj = _MockObj(self.__name__ + "." + item)
        setattr(self, item, obj)
        return obj


class _MockObj:
    def __init__(self, name="mock", *a, **k):
        self._name = name if isinstance(name, str) else "mock"

    def __call__(self, *a, **k):
        return _MockObj(self._name + "()")

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _MockObj(self._name + "." + item)

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)

    def __getitem__(self, item):
        return _MockObj(self._name + "[]")


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MOCK_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Mock(spec.name)

    def exec_module(self, module):
        pass


def _install_pl_stub():
    import torch.nn as nn
    pl = types.ModuleType("pytorch_lightning")
    pl.__path__ = []

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            import torch
            return torch.device("cpu")

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class Callback:
        def __init__(self, *a, **k):
            pass

    class Trainer:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.Callback = Callback
    pl.Trainer = Trainer
    pl.seed_everything = lambda s=0, *a, **k: __import__("torch").manual_seed(s)
    sys.modules["pytorch_lightning"] = pl
    for sub in ("callbacks", "loggers", "utilities", "utilities.distributed", "plugins", "core", "core.lightning"):
        m = _Mock("pytorch_lightning." + sub)
        sys.modules["pytorch_lightning." + sub] = m
    sys.modules["pytorch_lightning.callbacks"].Callback = Callback
    sys.modules["pytorch_lightning.utilities.distributed"].rank_zero_only = lambda f: f
    sys.modules["pytorch_lightning.utilities"].rank_zero_only = lambda f: f


def _install_scatter_stub():
    """torch_scatter 2.0.7 call forms used at enc.py:72,103 (encoder only; off the hot path)."""
    import torch
    ts = types.ModuleType("torch_scatter")

    def _expand(index, src, dim):
        return index.expand_as(src) if index.dim() == src.dim() else index

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        index = _expand(index, src, dim)
        if out is None:
            shape = list(src.shape)
            shape[dim] = int(dim_size) if dim_size is not None else int(index.max()) + 1
            out = torch.zeros(shape, dtype=src.dtype, device=src.device)
            return out.scatter_reduce(dim, index, src, reduce="mean", include_self=False)
        res = torch.zeros_like(out).scatter_reduce(dim, index, src, reduce="mean", include_self=False)
        out.copy_(res)
        return out

    def scatter_max(src, index, dim=-1, out=None, dim_size=None):
        index = _expand(index, src, dim)
        shape = list(src.shape)
        shape[dim] = int(dim_size) if dim_size is not None else int(index.max()) + 1
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        res = res.scatter_reduce(dim, index, src, reduce="amax", include_self=False)
        return res, None

    ts.scatter_mean = scatter_mean
    ts.scatter_max = scatter_max
    sys.modules["torch_scatter"] = ts


_installed = False


def install():
    """Make ``import shapeformer...`` resolve to the reference tree.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.meta_path.insert(0, _MockFinder())
    _install_pl_stub()
    _install_scatter_stub()
    # xgutils.vis executes matplotlib code at import (xgutils/vis/visutil.py:17-45): replace the sub-package.
    vis = _Mock("xgutils.vis")
    vis.__all__ = ["visutil", "npfvis", "fresnelvis", "vis3d"]
    for n in vis.__all__:
        sub = _Mock("xgutils.vis." + n)
        sys.modules["xgutils.vis." + n] = sub
        setattr(vis, n, sub)
    sys.modules["xgutils.vis"] = vis
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_modules():
    """Return the reference's live hot-path modules (mingpt, dec, quantizer, shapeformer, representers, common, vqdif)."""
    install()
    import importlib
    mods = {}
    mods["mingpt"] = importlib.import_module("shapeformer.models.shapeformer.transformer.mingpt")
    mods["dec"] = importlib.import_module("shapeformer.models.vqdif.dec")
    mods["quantizer"] = importlib.import_module("shapeformer.models.vqdif.quantizer")
    mods["common"] = importlib.import_module("shapeformer.models.shapeformer.common")
    mods["representers"] = importlib.import_module("shapeformer.models.shapeformer.representers")
    mods["shapeformer"] = importlib.import_module("shapeformer.models.shapeformer.shapeformer")
    mods["vqdif"] = importlib.import_module("shapeformer.models.vqdif.vqdif")
    return mods
