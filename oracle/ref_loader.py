"""Import the UNMODIFIED reference (``/root/reference``) for oracle pinning.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` only
exists in the build container, never on the GPU box, so nothing that runs under
``-m gpu``, ``smoke()`` or ``bench.py`` may call :func:`load_reference`.

The reference's ``run_t2mapping.py`` imports SimpleITK and, through
``utils/t2map_utils.py``, pydicom / matplotlib / skimage (none installed here).
``fit_voxel`` (run_t2mapping.py:120-312), ``set_fit_params`` (:29-111) and
``compute_residuals`` (utils/t2map_utils.py:62-89) need only numpy + scipy, so
the missing I/O and plotting packages are replaced by empty stub modules
before the import (recipe: SURVEY.md Appendix B).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("T2FIT_REFERENCE_ROOT", "/root/reference")

_STUBS = ("SimpleITK", "pydicom", "matplotlib", "matplotlib.pyplot",
          "matplotlib.cm", "skimage", "skimage.restoration")

_ref = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "run_t2mapping.py"))


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__getattr__ = lambda attr: types.SimpleNamespace()
    sys.modules[name] = m
    return m


def load_reference():
    """Return the reference's ``run_t2mapping`` module, imported unmodified."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for n in _STUBS:
        if n not in sys.modules:
            _stub(n)
    sys.modules["skimage"].restoration = sys.modules["skimage.restoration"]
    sys.dont_write_bytecode = True          # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import run_t2mapping as ref             # noqa: E402
    _ref = ref
    return ref
