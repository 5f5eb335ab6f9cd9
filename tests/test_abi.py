"""CPU-side: the C-ABI library loads and exports every symbol declared in include/calclens_b200.h (no compute)."""
import ctypes
import os

from calclens_b200 import _lib


def test_library_loads_and_exports_header_symbols():
    L = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/calclens_b200.h is not exported" % n
    assert L.clb_abi_version() == 2


def test_binding_covers_header():
    L = _lib.load()
    assert sorted(L._clb_signatures) == _lib.exported_symbols()


def test_library_is_sm100a_only():
    """the shared object must carry sm_100a SASS (no multi-arch fat binary, no PTX-only JIT fallback)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    archs = {ln.split(".")[-2] for ln in out.stdout.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except ImportError as e:
        assert "no fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
    importlib.reload(_lib)
