"""CPU tier: the C-ABI library loads (no CUDA call) and exports every symbol that include/immesh_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "immesh_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(immesh_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from immesh_b200 import build
    lib_path = build.build_cuda()       # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 24
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_fails_loudly():
    import ctypes as C
    from immesh_b200 import api
    import torch
    if torch.cuda.is_available():
        return
    lib = api.load_library()
    cfg = api.MeshConfig()
    try:
        api.Mesh(cfg, lib=lib)
    except RuntimeError as e:
        assert "no CUDA device" in str(e) or "code -4" in str(e)
    else:
        raise AssertionError("the product library must not run without a GPU")
