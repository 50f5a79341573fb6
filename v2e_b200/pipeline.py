"""Stage 2 + stage 3 of v2e.py (v2e.py:741-846) without the temp folders: source frames ->
SuperSloMo up-sampling -> DVS events, everything device-resident.

The reference hands frames from SloMo to the emulator as 8-bit PNG files in a temp dir
(slomo.py:440-444 -> v2e.py:832 read_image); here the uint8 frames stay in HBM. Times follow
v2e.py:794-797: interpTimes (units of source-frame intervals) scaled to the clip's duration.
"""
import numpy as np
import torch

from .emulator import EventEmulator
from .slomo import SuperSloMo


class V2EPipeline:
    def __init__(self, slomo: SuperSloMo, emulator: EventEmulator):
        self.slomo = slomo
        self.emulator = emulator

    def run(self, frames_u8, src_duration_s, t_offset=0.0, return_device=False, copy=False):
        """frames_u8: [N,H,W] uint8 source frames covering `src_duration_s` seconds.
        Returns (events [M,4] float32, frame offsets, interp_times_s, n_interp_frames). Host rows are a
        view of the emulator's pinned staging buffer unless copy=True (valid until the next call)."""
        interp, times, avg_u = self.slomo.interpolate_frames(frames_u8)
        f = src_duration_s / (np.max(times) - np.min(times))          # v2e.py:794-797
        t = t_offset + f * times
        ev, offs = self.emulator.generate_events_batch(interp, t, return_device=return_device, copy=copy)
        return ev, offs, t, interp.shape[0]
