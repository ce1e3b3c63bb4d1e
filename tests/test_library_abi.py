"""CPU-side checks of the C-ABI boundary: the library loads and exports every
symbol include/mmb200.h declares.  No compute calls (no GPU here)."""
import os
import re

from magellanmapper_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="mmb200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmb_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mmb200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert lib.mmb_version() == 1
    tools = _declared_symbols("mmb200_tools.h")
    assert set(tools) == set(_lib.TOOLS_SIGNATURES)
    for name in tools:
        assert hasattr(lib, name), f"{name} declared in mmb200_tools.h but not exported"


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.MmbCand) == 20
    assert C.sizeof(_lib.MmbPreprocParams) == 7 * 8


def test_no_cpu_fallback_without_cuda():
    """The product path must fail loudly when there is no device."""
    import pytest
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from magellanmapper_b200 import gpu
    with pytest.raises(RuntimeError):
        gpu.require_cuda()
