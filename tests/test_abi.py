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
    # V2eEmuCfg: 4 int32, 8 double, 2 int32, uint64, 2 int32, 2 double, 2 int32, 2 uint32, 4 int32
    assert ctypes.sizeof(_lib.V2eEmuCfg) == 16 + 64 + 8 + 8 + 8 + 16 + 8 + 8 + 16
    assert ctypes.sizeof(_lib.V2eFrameInfo) == 40


def test_library_reports_the_struct_layouts_the_binding_mirrors():
    """_lib.load() refuses a stale library: the sizes compiled into the .so must equal ctypes'."""
    lib = _lib.load()
    v, a, b, c = (ctypes.c_int(0) for _ in range(4))
    assert lib.v2e_abi_info(ctypes.byref(v), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
    assert v.value == _lib.ABI_VERSION == lib.v2e_version()
    assert (a.value, b.value, c.value) == (ctypes.sizeof(_lib.V2eEmuCfg), ctypes.sizeof(_lib.V2eFrameInfo),
                                           ctypes.sizeof(_lib.V2eUNetWeights))


def test_sink_keywords_are_delegated_or_ignored_not_refused(tmp_path, monkeypatch):
    """v2e.py:545-563 always passes the sink keywords. Constructing with them must not raise:
    they go to the reference's writers when v2ecore imports, otherwise they are ignored with a warning."""
    import torch
    from v2e_b200 import emulator as em_mod
    monkeypatch.setattr(em_mod._lib, "load", lambda *a, **k: object())
    e = em_mod.EventEmulator(device="cuda", output_folder=str(tmp_path), dvs_text="ev", dvs_aedat2="ev",
                             dvs_h5=None, show_dvs_model_state=None, output_width=346, output_height=260)
    try:
        import v2ecore.output.ae_text_output  # noqa: F401
        have_ref = True
    except Exception:
        have_ref = False
    assert (e.dvs_text is not None) == have_ref
    e._finalizer.detach()
    if e._sinks is not None:
        e._sinks.close()


def test_emulator_refuses_cpu_device():
    import pytest
    from v2e_b200 import EventEmulator
    with pytest.raises(RuntimeError):
        EventEmulator(device="cpu")
