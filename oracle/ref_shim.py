"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (SensorsINI/v2e).

Only usable where /root/reference exists (the build container, not the GPU box).
It injects empty stand-ins for GUI / file-format modules that the reference
imports at module top level but that the hot path never calls
(reference: v2ecore/emulator.py:14,17; v2ecore/output/aedat4_output.py:10;
v2ecore/v2e_utils.py:9-12), then imports the reference packages untouched.
Used by oracle/make_golden.py to generate tests/golden/*.npz and by tests that
cross-check the oracle port when the reference tree is present.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("V2E_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "v2ecore", "emulator.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns (emulator_module, emulator_utils_module, model_module, slomo_module)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _stub("h5py")
    _stub("screeninfo", get_monitors=lambda: [])
    _stub("dv_processing")
    _stub("easygui")
    tk = _stub("tkinter")
    fd = _stub("tkinter.filedialog")
    tk.filedialog = fd
    tk.Tk = object
    fd.askdirectory = lambda *a, **k: None
    _stub("engineering_notation", EngNumber=lambda x: x)
    import cv2
    cv2.destroyAllWindows = lambda: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import v2ecore.emulator as emu
    import v2ecore.emulator_utils as emu_utils
    import v2ecore.model as model
    import v2ecore.slomo as slomo
    return emu, emu_utils, model, slomo
