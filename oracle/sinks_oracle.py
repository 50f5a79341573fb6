"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's event-sink row conversions.

  h5_rows(events)        emulator.py:953-959 (uint32 rows of the HDF5 "events" dataset)
  aedat2_words(events)   v2ecore/output/aedat2_output.py:133-165 (big-endian address / timestamp words)

Pinned by tests/test_sinks.py against tests/golden/sinks_aedat2.npz, whose payload the reference's own
AEDat2Output wrote (oracle/make_golden_sinks.py). Never imported by v2e_b200/."""
import numpy as np

# aedat2_output.py:38-60: (width, height) -> (yShiftBits, xShiftBits, polShiftBits); flipx = flipy = True
LAYOUTS = {(346, 260): (22, 12, 11), (240, 180): (22, 12, 11), (640, 480): (11, 1, 0)}


def h5_rows(events):
    t = np.array(events, dtype=np.float32)
    t[:, 0] = t[:, 0] * 1e6
    t[t[:, 3] == -1, 3] = 0
    return t.astype(np.uint32)


def aedat2_words(events, width=346, height=260):
    ys, xs, ps = LAYOUTS[(width, height)]
    t = (1e6 * events[:, 0]).astype(np.int32)
    x = (width - 1) - events[:, 1].astype(np.int32)
    y = (height - 1) - events[:, 2].astype(np.int32)
    p = ((events[:, 3] + 1) / 2).astype(np.int32)
    a = (x << xs | y << ys | p << ps)
    out = np.empty(2 * events.shape[0], dtype=np.int32)
    out[0::2] = a
    out[1::2] = t
    return out.byteswap(), int(np.count_nonzero(p))
