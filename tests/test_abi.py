"""The C-ABI library loads and exports every symbol include/t3d.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "t3d.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(t3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from textureless_3d_reconstruction_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libt3d.so does not export {s}"
    # the ctypes table binds exactly the header's functions
    assert sorted(_lib.declared_symbols()) == syms


def test_version_and_error_string():
    from textureless_3d_reconstruction_b200 import _lib
    lib = _lib.load()
    assert lib.t3d_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_struct_layouts_match_header():
    from textureless_3d_reconstruction_b200 import _lib
    assert ctypes.sizeof(_lib.BackprojectParams) == 8 * 4 + 7 * 8 + 12 * 8
    assert ctypes.sizeof(_lib.TsdfParams) == 32
    assert ctypes.sizeof(_lib.FrameView) == 16 + 16 + 48 + 8        # depth, bgr | K | T_cw | conf_mask
    assert ctypes.sizeof(_lib.IcpResult) == 16 * 8 + 16 + 8 + 8


def test_no_cpu_fallback_without_gpu():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from textureless_3d_reconstruction_b200.runtime import get_context
    with pytest.raises(RuntimeError):
        get_context()
    from textureless_3d_reconstruction_b200 import _lib
    assert not _lib.load().t3d_create(0)          # fails loudly, message set
    assert "no CPU fallback" in _lib.last_error() or "CUDA" in _lib.last_error()


def test_product_never_imports_oracle():
    pkg = ROOT / "textureless_3d_reconstruction_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "t3d_oracle" not in text.replace(
            "oracle/t3d_oracle.c", ""), f"{f} references the oracle"
