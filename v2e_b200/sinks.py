"""Event-sink row conversions on the device.

The reference converts the emitted rows on the host, per frame, inside generate_events
(emulator.py:953-965 -> v2ecore/output/aedat2_output.py:133-188). Here the packed device rows are
converted by one streaming kernel and the caller copies file-ready words: 8 bytes per event for
AEDAT-2.0 instead of 16 for the float rows. File headers and the writers themselves stay the
reference's (out of scope)."""
import ctypes

import torch

from . import _lib

# aedat2_output.py:38-60: (width, height) -> yShiftBits, xShiftBits, polShiftBits; flipx = flipy = True
AEDAT2_LAYOUTS = {(346, 260): (22, 12, 11), (240, 180): (22, 12, 11), (640, 480): (11, 1, 0)}


def _check(events):
    if not (isinstance(events, torch.Tensor) and events.is_cuda and events.dtype == torch.float32
            and events.dim() == 2 and events.shape[1] == 4 and events.is_contiguous()):
        raise ValueError("events must be a contiguous CUDA float32 tensor [N, 4]")


def events_to_h5_rows(events):
    """[N,4] float32 rows [t,x,y,p] -> [N,4] uint32 rows [t_us, x, y, p01] (emulator.py:953-959)."""
    _check(events)
    out = torch.empty((events.shape[0], 4), dtype=torch.int32, device=events.device)
    lib = _lib.load()
    with torch.cuda.device(events.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(events.device).cuda_stream)
        _lib.check(lib.v2e_events_to_h5_rows(ctypes.c_void_p(events.data_ptr()), events.shape[0],
                                             ctypes.c_void_p(out.data_ptr()), st))
    return out.view(torch.int32)    # bit pattern of the uint32 rows (torch has no uint32 arithmetic)


def events_to_aedat2(events, output_width=346, output_height=260):
    """[N,4] float32 rows -> ([2N] int32 big-endian words ready for file.write(), number of ON events),
    as AEDat2Output.appendEvents builds them (aedat2_output.py:133-165). Raises ValueError for a size the
    reference's writer does not support (aedat2_output.py:61-63)."""
    _check(events)
    key = (int(output_width), int(output_height))
    if key not in AEDAT2_LAYOUTS:
        raise ValueError("AEDAT-2.0 output width=%d height=%d not supported" % key)
    ys, xs, ps = AEDAT2_LAYOUTS[key]
    n = events.shape[0]
    out = torch.empty((2 * n,), dtype=torch.int32, device=events.device)
    n_on = torch.zeros((1,), dtype=torch.int64, device=events.device)
    lib = _lib.load()
    with torch.cuda.device(events.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(events.device).cuda_stream)
        _lib.check(lib.v2e_events_to_aedat2(ctypes.c_void_p(events.data_ptr()), n, key[0], key[1], xs, ys, ps, 1, 1,
                                            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(n_on.data_ptr()), st))
    return out, n_on
