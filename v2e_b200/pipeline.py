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

    def run_clip_sharded(self, frames_u8, src_duration_s, t_offset=0.0, group=None):
        """ONE clip over the ranks of `group` (BASELINE config 5 layout). Every rank passes the
        same source frames; the emulator must have been built with shard=(rank, world, group).
          1. SloMo over this rank's frame pairs (parallel.pair_range) -- no halo, weights replicated;
          2. all-to-all of uint8 row bands (parallel.exchange_frame_bands);
          3. pixel model on this rank's rows of every frame (one all-reduce(MAX) of an int32 per frame).
        Returns (rows [M_r, 4] float32 host array of THIS rank's pixel rows (global y), interp_times_s,
        n_interp_frames). Union over ranks = the events of the clip; parallel.gather_event_streams /
        merge_by_time assemble them where one stream is wanted."""
        import torch.distributed as dist
        from . import parallel
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if self.emulator.shard is None:
            raise RuntimeError("run_clip_sharded needs EventEmulator(shard=(rank, world, group))")
        if isinstance(frames_u8, np.ndarray):
            frames_u8 = torch.from_numpy(np.ascontiguousarray(frames_u8))
        n, H, W = frames_u8.shape
        if n - 1 < world:
            raise ValueError("fewer frame pairs than ranks")
        if self.slomo.auto_upsample:
            raise NotImplementedError("auto_upsample picks U per batch (slomo.py:366-372); a sharded clip "
                                      "needs one U for the time axis: pass upsampling_factor")
        p0, p1 = parallel.pair_range(n - 1, rank, world)
        local, times_l, _ = self.slomo.interpolate_frames(frames_u8[p0:p1 + 1])
        U = int(self.slomo.upsampling_factor)
        times = np.arange((n - 1) * U) * (1.0 / U)                       # slomo.py:391-395 for the whole clip
        assert np.allclose(times_l + p0, times[p0 * U:p1 * U])
        bands = parallel.exchange_frame_bands(local, H, group=group, halo=self.emulator.cs_halo_rows(H))
        f = src_duration_s / (np.max(times) - np.min(times))            # v2e.py:794-797
        t = t_offset + f * times
        if self.emulator.rng_mode == "device" or not (self.emulator.leak_rate_hz > 0 or self.emulator.shot_noise_rate_hz > 0):
            # chunks of frames through the multi-frame kernels: one all-reduce(MAX) of the frame maxima per chunk
            # (frame by frame -- one all-reduce each -- for a chunk the refractory filter touches, and for the
            # centre-surround model, whose Euler iteration exchanges halo rows)
            rows, _ = self.emulator.generate_events_band_batch(bands, t, H)
            return rows, t, bands.shape[0]
        out = []
        for k in range(bands.shape[0]):
            ev = self.emulator.generate_events_band(bands[k], t[k], H)
            if ev is not None:
                out.append(ev)
        rows = np.concatenate(out, 0) if out else np.zeros((0, 4), np.float32)
        return rows, t, bands.shape[0]
