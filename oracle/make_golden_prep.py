"""TEST INFRASTRUCTURE ONLY -- tests/golden/prep_cv2.npz: what OpenCV itself (cv2.resize INTER_AREA + cvtColor
BGR2GRAY, the two calls of v2e.py:719-729) returns for small seeded frames, one case per code path of the resizer
(fractional shrink, integer 2x2 / 3x3 / 4x2 boxes, grey and colour input, with and without v2e's --crop).

    python oracle/make_golden_prep.py          # needs cv2 (this container: opencv-python-headless 4.13.0)
"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prep_cv2.npz")
# (source H, W, channels), (out W, out H), crop (left, right, top, bottom) or None
CASES = [((101, 77, 3), (50, 33), None), ((72, 128, 3), (35, 26), None), ((64, 96, 1), (48, 32), None),
         ((90, 120, 3), (40, 30), None), ((80, 64, 3), (32, 20), None), ((77, 131, 1), (46, 35), None),
         ((110, 150, 3), (35, 26), (10, 5, 7, 13)), ((60, 60, 3), (60, 60), None), ((54, 96, 3), (35, 26), (0, 0, 2, 0))]


def reference_prep(frame, out_wh, crop):
    """v2e.py:696-729 with cv2 itself."""
    f = frame
    if crop is not None:
        c_l = crop[0] if crop[0] > 0 else 0
        c_r = -crop[1] if crop[1] > 0 else None
        c_t = crop[2] if crop[2] > 0 else 0
        c_b = -crop[3] if crop[3] > 0 else None
        f = f[c_t:c_b, c_l:c_r]
    if (f.shape[1], f.shape[0]) != tuple(out_wh):
        f = cv2.resize(src=np.ascontiguousarray(f), dsize=tuple(out_wh), fx=out_wh[0] / frame.shape[1],
                       fy=out_wh[1] / frame.shape[0], interpolation=cv2.INTER_AREA)
    if f.ndim == 3:
        f = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    return f


def main():
    rng = np.random.default_rng(2024)
    out = {"cv2_version": np.array(cv2.__version__), "n_cases": np.array(len(CASES))}
    for i, (shape, wh, crop) in enumerate(CASES):
        fr = rng.integers(0, 256, (2,) + (shape if shape[2] == 3 else shape[:2]), dtype=np.uint8)
        out["in_%d" % i] = fr
        out["wh_%d" % i] = np.array(wh)
        out["crop_%d" % i] = np.array(crop if crop is not None else (-1, -1, -1, -1))
        out["out_%d" % i] = np.stack([reference_prep(f, wh, crop) for f in fr])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
