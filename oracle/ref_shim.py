"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import/compat shim that lets the UNMODIFIED reference modules under
/root/reference be imported in this container (Python 3.12, torch 2.11, CPU
only) so that (a) the CPU restatement in ``oracle/oracle.py`` can be validated
against the real thing and (b) golden vectors for ``tests/golden`` can be
generated (``oracle/make_golden.py``).  /root/reference does not exist on the
GPU box, so nothing under ``tests -m gpu``, ``smoke()`` or ``bench.py`` may
import this file.

What the shim does (SURVEY.md section 8c lists the blockers):
  * stub modules for packages that are absent here and only used for
    visualisation / IO (open3d, geomdl, matplotlib, trimesh, transforms3d,
    lap, pykdtree, h5py, configobj, turtle, positional_encodings);
    ``lapsolver.solve_dense`` is mapped to scipy's Hungarian solver;
  * ``torch.matrix_rank`` (removed from torch) -> ``torch.linalg.matrix_rank``;
  * on a CPU-only box: ``Tensor.cuda`` / ``Module.cuda`` become no-ops,
    ``torch.device('cuda')`` inside src/PointNet.py resolves to CPU, and
    ``torch.eye(..., device=-1)`` (``get_device()`` of a CPU tensor) maps to CPU.
No reference source is copied; the modules are imported from where they lie.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SEDNET_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = (
    "open3d", "geomdl", "matplotlib", "trimesh", "transforms3d", "lap",
    "pykdtree", "h5py", "configobj", "turtle", "positional_encodings", "lapsolver",
    "ipdb",
)


class _Anything:
    """Callable/attribute sink used for every symbol of a stubbed package."""

    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        return _Anything(self._name + "()")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything(self._name + "." + item)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    __all__ = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything(self.__name__ + "." + item)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        name = module.__name__
        if name == "lapsolver":
            from scipy.optimize import linear_sum_assignment

            module.solve_dense = lambda cost: linear_sum_assignment(cost)
        elif name == "positional_encodings.torch_encodings":
            import torch

            class PositionalEncoding1D(torch.nn.Module):
                # only the registered buffer matters (state_dict key pos_enc.inv_freq, shape (128,))
                def __init__(self, channels):
                    super().__init__()
                    ch = int(-(-channels // 2) * 2)
                    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
                    self.register_buffer("inv_freq", inv_freq)

            module.PositionalEncoding1D = PositionalEncoding1D
            module.PositionalEncoding2D = PositionalEncoding1D
            module.PositionalEncoding3D = PositionalEncoding1D
            module.Summer = PositionalEncoding1D
        elif name == "open3d":
            module.utility = _Anything("open3d.utility")
            module.visualization = _Anything("open3d.visualization")
            module.geometry = _Anything("open3d.geometry")
            module.io = _Anything("open3d.io")
            module.__all__ = ["utility", "visualization", "geometry", "io"]  # `from open3d import *`


_installed = False


def install():
    """Install the stubs and torch compat patches (idempotent)."""
    global _installed
    if _installed:
        return
    _installed = True
    import torch

    sys.meta_path.insert(0, _StubFinder())
    torch.matrix_rank = torch.linalg.matrix_rank  # the old name still exists but raises
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _eye = torch.eye

        def eye(*a, **k):
            dev = k.get("device", None)
            if isinstance(dev, int) and dev < 0:
                k["device"] = "cpu"
            return _eye(*a, **k)

        torch.eye = eye
        _get_device = torch.get_device
        torch.get_device = lambda t: (None if _get_device(t) < 0 else _get_device(t))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
        sys.path.insert(1, os.path.join(REFERENCE_ROOT, "src"))


class _TorchCPUProxy:
    """Stands in for the ``torch`` global of src/PointNet.py on a CPU box so that
    ``torch.device('cuda')`` (src/PointNet.py:148,185) yields the CPU device."""

    def __init__(self, torch):
        self._t = torch

    def __getattr__(self, item):
        return getattr(self._t, item)

    def device(self, *a, **k):
        return self._t.device("cpu")


def load():
    """Return a namespace with the reference's hot-path modules imported unmodified."""
    install()
    import torch

    ns = types.SimpleNamespace()
    ns.PointNet = importlib.import_module("PointNet")  # generate_predictions_aug.py:40 adds src/ to sys.path
    if not torch.cuda.is_available():
        ns.PointNet.torch = _TorchCPUProxy(torch)
    ns.SEDNet = importlib.import_module("src.SEDNet")
    ns.mean_shift = importlib.import_module("src.mean_shift")
    ns.guard = importlib.import_module("src.guard")
    ns.fitting_utils = importlib.import_module("src.fitting_utils")
    ns.primitive_forward = importlib.import_module("src.primitive_forward")
    ns.primitives = importlib.import_module("src.primitives")
    ns.segment_utils = importlib.import_module("src.segment_utils")
    return ns


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
