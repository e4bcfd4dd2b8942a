"""TEST INFRASTRUCTURE ONLY -- locate the UNMODIFIED reference and import it on the shim.

The reference modules (`/root/reference/flowfusion/{diffusion,flow,symplectic}.py`) are pure
Python and import fine once *some* module called ``torchdiffeq`` exists.  This
loader puts ``oracle/`` (which holds the restated ``torchdiffeq`` package) and the reference
checkout on ``sys.path`` and returns the three reference modules.  Nothing is copied.

The reference checkout only exists in the build container (``/root/reference``); on a GPU
box ``reference_available()`` is False and callers fall back to the self-contained port in
``oracle/port.py`` plus the committed golden vectors in ``tests/golden``.
"""
from __future__ import annotations

import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ("/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))


def shim_on_path() -> None:
    """Make ``import torchdiffeq`` resolve to the restatement in ``oracle/torchdiffeq``."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    mod = sys.modules.get("torchdiffeq")
    if mod is not None and not getattr(mod, "__file__", "").startswith(_HERE):
        raise RuntimeError("a different torchdiffeq is already imported: %r" % (mod,))


def reference_root():
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "flowfusion", "diffusion.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def load_reference():
    """Return ``(diffusion, flow, symplectic)`` modules of the unmodified reference."""
    root = reference_root()
    if root is None:
        raise ImportError("reference checkout not found (looked in %s)" % (_CANDIDATES,))
    shim_on_path()
    if root not in sys.path:
        sys.path.append(root)
    mods = [importlib.import_module("flowfusion." + m) for m in ("diffusion", "flow", "symplectic")]
    return tuple(mods)
