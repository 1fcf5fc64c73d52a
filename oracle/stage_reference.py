#!/usr/bin/env python
"""Stage the UNMODIFIED reference for comparator runs  --  TEST / BENCH INFRASTRUCTURE ONLY.

/root/reference exists in the build container but not on the GPU box.  This script copies the two
directories of the reference that sit on the hot path (`ModeT/`, `ModeT-cu/`) into the git-ignored
`baseline/_ref/` (it travels to the box with the gpurun snapshot, like our own built .so) and builds
the reference's own CUDA extension `ModeT-cu/modet` for sm_100 there with the reference's own
`setup.py` (`TORCH_CUDA_ARCH_LIST=10.0`; nvcc cross-compiles without a GPU).  Nothing here is product
code: `bench.py --impl reference`, the informational GPU comparators in `bench.py` and
`tests/test_dropin_scripts.py` are the only consumers, and all of them skip when the directory is
absent.  No reference source enters the git history (`baseline/_ref/` is in .gitignore).

    python oracle/stage_reference.py [--no-ext]
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SMILE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def staged_dir() -> str | None:
    """Directory holding `ModeT/models.py` of the reference: the staged copy, else the original, else None."""
    for base in (DST, SRC):
        if os.path.isfile(os.path.join(base, "ModeT", "models.py")):
            return base
    return None


def ext_built() -> bool:
    return bool(glob.glob(os.path.join(DST, "ModeT-cu", "modet", "modet*.so")))


def stage(build_ext: bool = True, verbose: bool = True) -> str | None:
    if not os.path.isfile(os.path.join(SRC, "ModeT", "models.py")):
        return staged_dir()          # GPU box: use what travelled
    os.makedirs(DST, exist_ok=True)
    for sub in ("ModeT", "ModeT-cu"):
        dst = os.path.join(DST, sub)
        if not os.path.isdir(dst):
            shutil.copytree(os.path.join(SRC, sub), dst)
    if build_ext and not ext_built():
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="4")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=os.path.join(DST, "ModeT-cu", "modet"),
                           env=env, capture_output=True, text=True)
        with open(os.path.join(DST, "modet_ext_build.log"), "w") as f:
            f.write(r.stdout + "\n" + r.stderr)
        if verbose:
            print(f"reference modet extension build rc={r.returncode} (log: baseline/_ref/modet_ext_build.log)")
    return DST


if __name__ == "__main__":
    print(stage(build_ext="--no-ext" not in sys.argv))
