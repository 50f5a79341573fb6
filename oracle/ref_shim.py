"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (SensorsINI/v2e).

Loads the reference from /root/reference (the build container) or, where that does not exist (the
GPU box), from the verbatim copy oracle/make_ref.py vendored into oracle/_ref/ (git-ignored, travels with
the snapshot). It injects empty stand-ins for GUI / file-format modules that the reference
imports at module top level but that the hot path never calls
(reference: v2ecore/emulator.py:14,17; v2ecore/output/aedat4_output.py:10;
v2ecore/v2e_utils.py:9-12), then imports the reference packages untouched.
Used by oracle/make_golden.py to generate tests/golden/*.npz and by tests that
cross-check the oracle port when the reference tree is present.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pick_root():
    for r in (os.environ.get("V2E_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if r and os.path.isfile(os.path.join(r, "v2ecore", "emulator.py")):
            return r
    return os.environ.get("V2E_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "v2ecore", "emulator.py"))


def reference_kind() -> str:
    """'_ref' when the vendored verbatim copy is in use, 'reference' for /root/reference itself."""
    return "_ref" if os.path.abspath(REFERENCE_ROOT) == os.path.join(_HERE, "_ref") else "reference"


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
