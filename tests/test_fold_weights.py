"""CPU: v2e_conv_up2_fold_weights (host code of the C ABI) -- the x2 bilinear up-sampling folded into a 3x3 filter
must reproduce conv2d(interpolate(x, 2, 'bilinear')) on the interior of the image (the 2-pixel frame is the
frame kernel's job on the GPU). Layout: fp16 [C_pad/64][2 px][3 b][6 q][Cout_pad][64], q -> (py = q & 1, a = 1 - q // 2)."""
import ctypes

import numpy as np
import torch

from v2e_b200 import _lib


def test_folded_filter_equals_upsample_then_conv_on_the_interior():
    lib = _lib.load()
    torch.manual_seed(3)
    C, Co, Cp, h, w = 64, 20, 32, 7, 9
    x = torch.randn(1, C, h, w, dtype=torch.float64)
    W = torch.randn(Co, C, 3, 3, dtype=torch.float32) / 24
    fold = np.zeros((1 * 2 * 3 * 6 * Cp * 64,), np.float16)
    wh = np.ascontiguousarray(W.numpy())
    rc = lib.v2e_conv_up2_fold_weights(wh.ctypes.data_as(ctypes.c_void_p), Co, C, Cp, C,
                                       fold.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    Wf = torch.from_numpy(fold.astype(np.float64)).reshape(2, 3, 6, Cp, 64)        # [px][b][q][co][c]
    assert (Wf[:, :, :, Co:] == 0).all()                                         # padded output channels
    ref = torch.nn.functional.conv2d(
        torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), W.double(), padding=1)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            acc = torch.zeros(1, Co, h, w, dtype=torch.float64)
            for a in (-1, 0, 1):
                q = 2 * (1 - a) + py
                for b in range(3):
                    acc += torch.einsum("oc,bchw->bohw", Wf[px, b, q, :Co], xp[:, :, a + 1:a + 1 + h, b:b + w])
            out[:, :, py::2, px::2] = acc
    err = (out - ref).abs()[:, :, 2:-2, 2:-2].max().item()
    assert err < 3e-3, err            # fp16 rounding of the folded filter (K = 576 terms of ~0.04 * 1)
    assert (out - ref).abs().max().item() > 0.05      # ... and the frame really is different (clamping / zero padding)
