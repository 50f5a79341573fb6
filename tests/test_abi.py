"""CPU: the C-ABI library builds/loads and exports every symbol include/v2e_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

from v2e_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "v2e_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2e_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert "v2e_emu_step" in names and "v2e_emu_create" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/v2e_b200.h but not exported: %s" % missing


def test_binding_table_covers_header():
    names = _declared()
    unbound = [n for n in names if n not in _lib._SIGS]
    assert not unbound, "no ctypes signature for %s" % unbound


def test_version_call_without_gpu():
    lib = _lib.load()
    assert lib.v2e_version() >= 100
    assert lib.v2e_last_error() is not None


def test_struct_sizes_match_header_layout():
    # V2eEmuCfg: 4 int32, 8 double, 2 int32, uint64, 2 int32, 2 double, 2 int32
    assert ctypes.sizeof(_lib.V2eEmuCfg) == 16 + 64 + 8 + 8 + 8 + 16 + 8
    assert ctypes.sizeof(_lib.V2eFrameInfo) == 40


def test_emulator_refuses_cpu_device():
    import pytest
    from v2e_b200 import EventEmulator
    with pytest.raises(RuntimeError):
        EventEmulator(device="cpu")
