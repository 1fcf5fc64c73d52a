"""Import the UNMODIFIED reference modules for comparator / drop-in runs  --  TEST / BENCH INFRASTRUCTURE ONLY.

Resolves the reference tree (`baseline/_ref` staged by oracle/stage_reference.py, else /root/reference) and imports
`ModeT/models.py`, `ModeT/losses.py` or `ModeT-cu/models.py` under private module names, so that they never shadow (or
get shadowed by) the drop-in `models` module.  Returns None when no reference tree is available (callers skip)."""
from __future__ import annotations

import importlib.util
import os
import sys

from .stage_reference import staged_dir


def _load(path: str, name: str, extra_path=()):
    saved = list(sys.path)
    sys.path[:0] = list(extra_path)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path[:] = saved


def reference_models():
    """The reference's ModeT/models.py as a module, or None."""
    base = staged_dir()
    return None if base is None else _load(os.path.join(base, "ModeT", "models.py"), "_smile_ref_models")


def reference_losses():
    base = staged_dir()
    return None if base is None else _load(os.path.join(base, "ModeT", "losses.py"), "_smile_ref_losses")


def reference_models_cu():
    """ModeT-cu/models.py with the reference's own `modet` CUDA extension (built for sm_100 by stage_reference.py);
    None if the tree or the built extension is missing.  Needs a CUDA device to run."""
    base = staged_dir()
    if base is None:
        return None
    cu = os.path.join(base, "ModeT-cu")
    ext = os.path.join(cu, "modet")
    if not any(f.startswith("modet") and f.endswith(".so") for f in os.listdir(ext)):
        return None
    # `functional.py` does `from modet import ...` and models.py does `from functional import ...`: both by bare name
    saved = {k: sys.modules.pop(k, None) for k in ("functional", "modet")}
    try:
        return _load(os.path.join(cu, "models.py"), "_smile_ref_models_cu", extra_path=[cu, ext])
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
