"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly the symbols
include/smilecode_b200.h declares, the ctypes table mirrors the header, and the host-side modules
keep the reference's names / state_dict contract.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smilecode_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(smile_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(2)] = args
    return out


@pytest.fixture(scope="module")
def built():
    from smilecode_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    fns = header_functions()
    assert len(fns) >= 11
    for name in fns:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    handle.smile_version.restype = ctypes.c_int
    assert handle.smile_version() >= 100


def test_ctypes_table_mirrors_header(built):
    from smilecode_b200 import _lib
    fns = header_functions()
    compute = {k: v for k, v in fns.items() if k not in ("smile_version", "smile_last_error")}
    assert set(compute) == set(_lib.SIGNATURES)
    kind = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "double": ctypes.c_double}
    for name, args in compute.items():
        sig = _lib.SIGNATURES[name]
        assert len(sig) == len(args), name
        for a, ct in zip(args, sig):
            if "*" in a or a.startswith("smile_stream_t"):
                assert ct is ctypes.c_void_p, (name, a)
            else:
                base = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                assert ct is kind[base], (name, a)


def test_invalid_arguments_are_rejected_without_a_gpu(built):
    from smilecode_b200 import _lib
    with pytest.raises(_lib.SmileError, match="NULL"):
        _lib.call("smile_warp3d_fwd", None, None, None, 1, 1, 4, 4, 4, None)


def test_missing_library_fails_loudly(monkeypatch):
    from smilecode_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsmilecode_b200.so")
    with pytest.raises(_lib.SmileError, match="no CPU"):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    from smilecode_b200 import ops
    from smilecode_b200._lib import SmileError
    with pytest.raises(SmileError, match="CUDA"):
        ops.warp3d(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 3, 4, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smilecode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/ ", ""), f"{f} mentions the oracle"


def test_state_dict_contract():
    from smilecode_b200 import models
    from oracle import modet_oracle as orc
    m = models.ModeT((32, 32, 32), head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1)
    sd = m.state_dict()
    learnable = {k: tuple(v.shape) for k, v in m.named_parameters()}
    assert learnable == orc.param_shapes()
    assert sum(v.numel() for v in m.parameters()) == 1029670
    buffers = sorted(k for k in sd if k not in learnable)
    assert buffers == sorted([f"mdt{i}.grid" for i in range(1, 6)] + [f"transformer.{i}.grid" for i in range(4)])
    assert tuple(sd["transformer.1.grid"].shape) == (1, 3, 16, 16, 16)
    assert tuple(sd["mdt3.grid"].shape) == (3, 3, 3, 3)
    # ModeT-cu checkpoints name the tap table `v` [27,3]
    cu = {k.replace(".grid", ".v") if k.startswith("mdt") else k: (v.reshape(27, 3) if k.startswith("mdt") and k.endswith("grid") else v)
          for k, v in sd.items()}
    m2 = models.ModeT_cu((32, 32, 32))
    assert m2.mdt1.scale == 1
    m2.load_state_dict(cu, strict=True)
    # ... and ModeT_cu writes it that way too (the -cu contract used to hold in one direction only); either
    # class loads either spelling
    sd_cu = m2.state_dict()
    assert sorted(k for k in sd_cu if k.startswith("mdt") and not k.endswith("rpb")) == [f"mdt{i}.v" for i in range(1, 6)]
    assert tuple(sd_cu["mdt2.v"].shape) == (27, 3)
    m2.load_state_dict(sd, strict=True)          # ModeT checkpoint (grid) into ModeT_cu
    m.load_state_dict(sd_cu, strict=True)        # ModeT_cu checkpoint (v) into ModeT
    for name in ("Encoder", "ProjectionLayer", "ConvBlock", "ConvInsBlock", "VecInt", "ResizeTransform", "UpConvBlock",
                 "DeconvBlock", "CWM", "ModeTransformer", "SpatialTransformer"):
        assert hasattr(models, name)
    assert models.ModeTransformer(6, 1).scale == 6 ** -0.5
    with pytest.raises(ValueError):
        models.ModeT((32, 32, 32), num_heads=[6, 6, 3, 2, 1])


def test_dropin_shim_resolves():
    import importlib.util
    spec = importlib.util.spec_from_file_location("models", os.path.join(ROOT, "dropin", "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ModeT.__module__ == "smilecode_b200.models" and hasattr(mod, "ModeT_cu")
