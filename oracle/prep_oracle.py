"""TEST INFRASTRUCTURE ONLY -- numpy restatement of v2e's stage-1 input preparation (v2e.py:687-737):

    frame[c_t:c_b, c_l:c_r]  ->  cv2.resize(dsize=(W, H), interpolation=cv2.INTER_AREA)  ->  cv2.cvtColor(BGR2GRAY)

for 8-bit frames. The arithmetic lives in OpenCV (opencv-python, unpinned by the reference; 4.13.0 in this container):
  * INTER_AREA, integer scale factors (modules/imgproc/src/resize.cpp, resizeAreaFast_Invoker): integer box sum times
    the float32 1/area, rounded half-to-even (saturate_cast<uchar> = cvRound); the 2x2 case takes the SIMD path
    (s0 + s1 + s2 + s3 + 2) >> 2;
  * INTER_AREA, fractional shrink (computeResizeAreaTab + ResizeArea_Invoker): per axis a table of (source index,
    destination index, float32 weight); rows are accumulated in float32 -- horizontally buf[dx] += S[sx] * alpha in
    table order, vertically sum[dx] += beta * buf[dx] in row order, separate multiply and add -- and rounded
    half-to-even at the end;
  * BGR2GRAY (color_rgb.simd.hpp, RGB2Gray<uchar>): (B * 3735 + G * 19235 + R * 9798 + (1 << 14)) >> 15.
Pinned by tests/test_prep.py against tests/golden/prep_*.npz (oracle/make_golden_prep.py ran cv2 itself).
Enlarging (a scale factor below 1, where OpenCV switches INTER_AREA to its bilinear code) is not restated.
"""
import math

import numpy as np

BY15, GY15, RY15 = 3735, 19235, 9798          # 0.114, 0.587, 0.299 in 15-bit fixed point


def area_tab(ssize, dsize, scale):
    """computeResizeAreaTab (resize.cpp): list of (si, di, float32 alpha)."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((sx1 - 1, dx, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((sx, dx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((sx2, dx, np.float32(min(min(fsx2 - sx2, 1.), cell) / cell)))
    return tab


def _round_u8(x):
    """saturate_cast<uchar>(float): round half to even, clamp."""
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def resize_area_u8(img, dsize_wh):
    """cv2.resize(img, dsize_wh, interpolation=cv2.INTER_AREA) for uint8 [H, W] or [H, W, C], shrinking only."""
    a = img if img.ndim == 3 else img[:, :, None]
    sh, sw, cn = a.shape
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])
    if (dw, dh) == (sw, sh):
        return img.copy()
    scale_x, scale_y = sw / dw, sh / dh
    if scale_x < 1 or scale_y < 1:
        raise NotImplementedError("INTER_AREA enlargement (OpenCV's bilinear path) is not restated")
    isx, isy = int(round(scale_x)), int(round(scale_y))
    eps = np.finfo(np.float64).eps
    if abs(scale_x - isx) < eps and abs(scale_y - isy) < eps:
        s = a[:dh * isy, :dw * isx].astype(np.int64).reshape(dh, isy, dw, isx, cn).sum((1, 3))
        if isx == 2 and isy == 2:
            out = ((s + 2) >> 2).astype(np.uint8)
        else:
            out = _round_u8(s.astype(np.float32) * np.float32(1.0 / (isx * isy)))
    else:
        xtab, ytab = area_tab(sw, dw, scale_x), area_tab(sh, dh, scale_y)
        af = a.astype(np.float32)
        rows = {}
        out_f = np.zeros((dh, dw, cn), np.float32)
        started = np.zeros(dh, bool)
        for sy, dy, beta in ytab:
            if sy not in rows:
                buf = np.zeros((dw, cn), np.float32)
                for sx, dx, alpha in xtab:
                    buf[dx] = buf[dx] + af[sy, sx] * alpha          # float32 multiply, then float32 add
                rows[sy] = buf
            t = beta * rows[sy]
            out_f[dy] = t if not started[dy] else out_f[dy] + t
            started[dy] = True
        out = _round_u8(out_f)
    return out if img.ndim == 3 else out[:, :, 0]


def bgr2gray_u8(img):
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    return ((b * BY15 + g * GY15 + r * RY15 + (1 << 14)) >> 15).astype(np.uint8)


def prep_frame(frame, out_wh=None, crop=None):
    """v2e.py:696-729 for one frame: crop (left, right, top, bottom), INTER_AREA resize, BGR -> luma."""
    f = frame
    if crop is not None:
        c_l = crop[0] if crop[0] > 0 else 0
        c_r = -crop[1] if crop[1] > 0 else None
        c_t = crop[2] if crop[2] > 0 else 0
        c_b = -crop[3] if crop[3] > 0 else None
        f = f[c_t:c_b, c_l:c_r]
    if out_wh is not None and (f.shape[1], f.shape[0]) != tuple(out_wh):
        f = resize_area_u8(np.ascontiguousarray(f), out_wh)
    if f.ndim == 3:
        f = bgr2gray_u8(f)
    return f
