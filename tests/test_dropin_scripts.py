"""The drop-in claim, executed (north_star: "drops into ModeT/train.py and infer.py unchanged").

tests/dropin_runner.py runs the reference's own, unmodified scripts from the staged reference tree (baseline/_ref,
written by oracle/stage_reference.py; skipped when absent) with `dropin/` first on sys.path.  The CPU half (import,
construct, load a state_dict produced by the REFERENCE module with strict=True) runs in the normal suite; the two
scripts themselves need the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    sys.path.insert(0, ROOT)
    from oracle.stage_reference import staged_dir
    return staged_dir() is not None


def _run(mode, tmp_path, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_runner.py"), mode, str(tmp_path)],
                       capture_output=True, text=True, timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DROPIN ")]
    assert r.returncode == 0 and lines, f"rc={r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return json.loads(lines[-1][7:]), r.stdout


needs_ref = pytest.mark.skipif(not _have_reference(), reason="reference tree not staged (oracle/stage_reference.py)")


@needs_ref
def test_reference_state_dict_loads_strict_into_dropin(tmp_path):
    out, _ = _run("load", tmp_path, 300)
    assert out["models_file"].startswith(os.path.join(ROOT, "dropin"))
    assert out["missing"] == [] and out["unexpected"] == [] and out["same_keys"] and out["keys"] >= 70


@needs_ref
@pytest.mark.gpu
def test_reference_infer_py_runs_on_dropin(tmp_path):
    out, stdout = _run("infer", tmp_path, 900)
    assert out["models_file"].startswith(os.path.join(ROOT, "dropin"))
    assert out["our_kernel_launches"] > 100                     # the forward ran on our kernels
    assert stdout.count("Trans dsc:") == 2 and "Deformed DSC:" in stdout and "deformed det:" in stdout   # infer.py:92-99


@needs_ref
@pytest.mark.gpu
def test_reference_train_py_runs_on_dropin(tmp_path):
    out, stdout = _run("train", tmp_path, 1200)
    assert out["models_file"].startswith(os.path.join(ROOT, "dropin"))
    assert out["optimizer_steps"] == 5 and out["our_kernel_launches"] > 1000
    assert len(out["checkpoints"]) >= 1                          # save_checkpoint at the end of an epoch (train.py:155-160)
    assert any("loss" in l for l in out["log_tail"])


@needs_ref
def test_modet_cu_checkpoints_load_strict_both_ways():
    """ModeT-cu names the tap table `mdtN.v` (ModeT-cu/models.py:296-299): a checkpoint written by either ModeT_cu loads into
    the other with strict=True."""
    from oracle import reference_loader as rl
    from smilecode_b200 import models
    ref_cu = rl.reference_models_cu()
    if ref_cu is None:
        pytest.skip("reference modet extension not built")
    shape = (32, 32, 32)
    theirs, ours = ref_cu.ModeT_cu(shape), models.ModeT_cu(shape)
    assert sorted(theirs.state_dict().keys()) == sorted(ours.state_dict().keys())
    r1 = ours.load_state_dict(theirs.state_dict(), strict=True)
    r2 = theirs.load_state_dict(ours.state_dict(), strict=True)
    assert not r1.missing_keys and not r1.unexpected_keys and not r2.missing_keys and not r2.unexpected_keys
