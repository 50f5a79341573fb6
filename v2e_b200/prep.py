"""Stage 1 of v2e.py on the device (v2e.py:687-737): crop, cv2.resize(INTER_AREA) and BGR -> luma
of 8-bit source frames, bit-exact with OpenCV (csrc/prep.cu; restated and pinned in oracle/prep_oracle.py). The
frames stay in HBM for SuperSloMo.interpolate_frames / the pixel model instead of going through .npy files."""
import ctypes

import numpy as np
import torch

from . import _lib


class InputPrep:
    def __init__(self, src_size, out_size, channels=3, crop=None, device="cuda:0"):
        """src_size / out_size: (width, height) as cv2 counts them; channels: 1 (grey) or 3 (BGR, what cv2.VideoCapture
        delivers); crop: v2e's --crop (left, right, top, bottom) or None."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.src_w, self.src_h = int(src_size[0]), int(src_size[1])
        self.out_w, self.out_h = int(out_size[0]), int(out_size[1])
        self.channels = int(channels)
        c = (0, 0, 0, 0) if crop is None else tuple(int(v) for v in crop)
        if len(c) != 4:
            raise ValueError("--crop must have 4 elements")        # v2e.py:640-644
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.v2e_prep_create(self.src_w, self.src_h, self.channels, c[0], c[1], c[2], c[3],
                                                self.out_w, self.out_h, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.v2e_prep_destroy(self._h)
            self._h = None

    def __call__(self, frames):
        """frames: uint8 [N, H, W, C] (or [N, H, W] for grey; ndarray or tensor). Returns [N, out_h, out_w] uint8 on the device."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        want = (self.src_h, self.src_w) + ((self.channels,) if self.channels == 3 else ())
        if frames.dtype != torch.uint8 or tuple(frames.shape[1:]) != want:
            raise ValueError("frames must be uint8 [N, %s]" % ", ".join(str(v) for v in want))
        fr = frames.to(self.device, non_blocking=True).contiguous()
        out = torch.empty((fr.shape[0], self.out_h, self.out_w), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.v2e_prep_run(self._h, ctypes.c_void_p(fr.data_ptr()), fr.shape[0],
                                             ctypes.c_void_p(out.data_ptr()), st))
        return out
