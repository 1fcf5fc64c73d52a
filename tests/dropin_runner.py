#!/usr/bin/env python
"""Runs the reference's OWN scripts (ModeT/infer.py, ModeT/train.py -- executed unmodified from the staged tree) with the
drop-in `models` / `losses` modules first on sys.path.  

    python tests/dropin_runner.py {infer|train|load} WORKDIR

What the harness supplies, none of it touching the reference's code:
  * sys.path = [dropin/, <ref>/ModeT]: `from models import ModeT` and `import losses` resolve to smilecode_b200,
    `utils` and `data` stay the reference's own files;
  * stubs for the three third-party modules that are absent from this image (natsort, pystrum, matplotlib) and
    `collections.Sequence` (removed in Python 3.10; data/trans.py:19 still uses it);
  * `glob.glob('/LPBA_path/...')` redirected to a synthetic two-subject .pkl set in WORKDIR (the scripts hard-code the
    dataset path, train.py:45-46 / infer.py:50);
  * a checkpoint written from the REFERENCE module's state_dict() for infer.py to load with strict=True (infer.py:62-64);
  * train.py's 30-epoch loop is stopped by the harness after STOP_AFTER optimizer steps (two full epochs incl. validation
    and save_checkpoint) by raising from a wrapped optimizer.step.
Prints one `DROPIN {json}` line with what ran.
"""
import collections
import collections.abc
import glob as _glob
import json
import os
import pickle
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPE = (160, 192, 160)
STOP_AFTER = 5


class _Stop(Exception):
    pass


def install_stubs():
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence
    import re
    ns = types.ModuleType("natsort")
    key = lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]
    ns.natsorted = lambda seq: sorted(seq, key=key)
    sys.modules["natsort"] = ns
    import numpy as np
    nd = types.ModuleType("pystrum.pynd.ndutils")

    def volsize2ndgrid(volsize):                      # pystrum.pynd.ndutils.volsize2ndgrid: ndgrid of aranges, 'ij'
        return np.meshgrid(*[np.arange(e) for e in volsize], indexing="ij")
    nd.volsize2ndgrid = volsize2ndgrid
    ps, pynd = types.ModuleType("pystrum"), types.ModuleType("pystrum.pynd")
    ps.pynd, pynd.ndutils = pynd, nd
    sys.modules.update({"pystrum": ps, "pystrum.pynd": pynd, "pystrum.pynd.ndutils": nd})
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    tk, m3 = types.ModuleType("mpl_toolkits"), types.ModuleType("mpl_toolkits.mplot3d")
    m3.axes3d = object()
    tk.mplot3d = m3
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "mpl_toolkits": tk, "mpl_toolkits.mplot3d": m3})


def make_dataset(work):
    """Two synthetic subjects per split in the reference's pickle format (datasets.py:8-10): (img fp32, seg uint16)."""
    import numpy as np
    from smilecode_b200.data import SEG_TABLE
    from smilecode_b200.synth import make_pair
    for split in ("Train", "Val"):
        d = os.path.join(work, "LPBA_path", split)
        os.makedirs(d, exist_ok=True)
        moving, fixed = make_pair(SHAPE, batch=1, seed=24 if split == "Train" else 25)
        for i, vol in enumerate((moving, fixed)):
            img = vol[0, 0].numpy().astype(np.float32)
            # label volume: quantised intensity -> LPBA label codes (0 = background)
            seg = SEG_TABLE[np.minimum((img * 54).astype(np.int64) + (img > 0), 54)].astype(np.uint16)
            with open(os.path.join(d, f"S{i + 1:02d}.pkl"), "wb") as f:
                pickle.dump((img, seg), f)


def redirect_glob(work):
    real = _glob.glob

    def patched(pat, *a, **k):
        if isinstance(pat, str) and pat.startswith("/LPBA_path/"):
            pat = os.path.join(work, pat.lstrip("/"))
        return sorted(real(pat, *a, **k))
    _glob.glob = patched


def main():
    mode, work = sys.argv[1], os.path.abspath(sys.argv[2])
    from oracle.stage_reference import staged_dir      # checker-side helper: where the unmodified reference lives
    ref = staged_dir()
    if ref is None:
        print("DROPIN " + json.dumps({"skipped": "no reference tree (baseline/_ref or /root/reference)"}))
        return
    refdir = os.path.join(ref, "ModeT")
    os.makedirs(work, exist_ok=True)
    install_stubs()
    import torch
    out = {"mode": mode, "reference": refdir}

    if mode == "load":          # CPU half: import / construct / strict load, no kernels
        from oracle import reference_loader as rl
        ref_sd = rl.reference_models().ModeT(SHAPE, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1).state_dict()
        sys.path[:0] = [os.path.join(ROOT, "dropin"), refdir]
        import models
        assert models.__file__.startswith(os.path.join(ROOT, "dropin")), models.__file__
        m = models.ModeT(SHAPE, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1)
        res = m.load_state_dict(ref_sd)           # strict=True
        out.update(models_file=models.__file__, missing=list(res.missing_keys), unexpected=list(res.unexpected_keys),
                   keys=len(ref_sd), same_keys=sorted(m.state_dict().keys()) == sorted(ref_sd.keys()))
        print("DROPIN " + json.dumps(out))
        return

    make_dataset(work)
    redirect_glob(work)
    os.chdir(work)
    sys.path[:0] = [os.path.join(ROOT, "dropin"), refdir]
    from smilecode_b200 import _lib
    l0 = _lib.LAUNCHES
    if mode == "infer":
        from oracle import reference_loader as rl
        from smilecode_b200.synth import randomize_weights
        rm = rl.reference_models().ModeT(SHAPE, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1)
        randomize_weights(rm, seed=1234)
        folder = "experiments/modet-heads(84211)-rpe_headim_6_ncc_1_reg_1_lr_0.0001_54r/"     # infer.py:57-58
        os.makedirs(folder, exist_ok=True)
        torch.save({"state_dict": rm.state_dict()}, folder + "dsc0.700.pth.tar")
        del rm
        runpy.run_path(os.path.join(refdir, "infer.py"), run_name="__main__")
    elif mode == "train":
        steps = {"n": 0}
        real_step = torch.optim.Adam.step

        def counted(self, *a, **k):
            r = real_step(self, *a, **k)
            steps["n"] += 1
            if steps["n"] >= STOP_AFTER:
                raise _Stop()
            return r
        torch.optim.Adam.step = counted
        try:
            runpy.run_path(os.path.join(refdir, "train.py"), run_name="__main__")
        except _Stop:
            pass
        finally:
            torch.optim.Adam.step = real_step
            sys.stdout = sys.__stdout__            # train.py:58 replaces sys.stdout with its Logger
        out["optimizer_steps"] = steps["n"]
        out["checkpoints"] = sorted(os.listdir(os.path.join(work, "experiments", os.listdir(os.path.join(work, "experiments"))[0])))
        with open(_glob.glob(os.path.join(work, "logs", "*", "logfile.log"))[0]) as f:
            out["log_tail"] = f.read().strip().splitlines()[-6:]
    import models
    out["models_file"] = models.__file__
    out["our_kernel_launches"] = _lib.LAUNCHES - l0
    print("DROPIN " + json.dumps(out))


if __name__ == "__main__":
    main()
