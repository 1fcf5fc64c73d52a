"""The fast kernels replace the IEEE division of the sampling-coordinate replay (ModeT/models.py:56) by a
correctly rounded 3-instruction sequence; this compiles the C checker under oracle/ and runs it (CPU only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_markstein_division_is_bit_identical_to_ieee_division(tmp_path):
    exe = str(tmp_path / "markstein_check")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "oracle", "markstein_check.c"),
                    "-lm"], check=True)
    out = subprocess.run([exe, "20000"], capture_output=True, text=True, check=True).stdout
    assert "bad=0" in out, out
    assert int(out.split("tot=")[1].split()[0]) > 2.5e7
