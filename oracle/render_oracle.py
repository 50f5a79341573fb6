"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's EventRenderer.render_events_to_frames
(v2ecore/renderer.py:161-390, accumulate_event_frame :392-430, hist2d_numba_seq v2ecore/v2e_utils.py:474-486), with
its quirks kept, because a drop-in must return the same frames:
  * the frame being filled is dropped at the start of every call (`self.currentFrame = None`, :270);
  * the last event of a packet is never rendered (`end = numEvents - 1`, :300-303; slices are end-exclusive);
  * DURATION: boundaries by searchsorted(ts, frame start, 'left') / (ts, next start, 'right') over the whole packet;
  * a finished frame is clip(hist_on - hist_off, +-full_scale_count), returned as (frame + fs) / (2 fs) in float64.
Pinned by tests/test_render.py against tests/golden/render_ref.npz (oracle/make_golden_render.py ran the unmodified class).
AREA_COUNT (a sequential, data-dependent scan, renderer.py:246-261) is restated too.
"""
import numpy as np

DURATION, COUNT, AREA_COUNT, SOURCE = 1, 2, 3, 4


class RenderOracle:
    def __init__(self, full_scale_count=3, exposure_mode=DURATION, exposure_value=1 / 300.0, area_dimension=None):
        self.mode, self.value, self.fs = exposure_mode, exposure_value, full_scale_count
        self.area_dimension = area_dimension
        self.interval = 1 / (1 / exposure_value) if exposure_mode == DURATION else None      # renderer.py:92-93
        self.cur_start = None
        self.area_counts = None

    def _frame(self, ev, H, W):
        on = np.zeros((H, W)); off = np.zeros((H, W))
        for t, x, y, p in ev:
            i, j = float(y), float(x)
            if 0 <= i < H and 0 <= j < W:
                (on if p == 1 else off)[int(i), int(j)] += 1
        return np.clip(on - off, -self.fs, self.fs)

    def render(self, ev, H, W):
        if ev is None or ev.shape[0] == 0:
            return None
        ts = ev[:, 0]
        n = len(ts)
        if self.mode == DURATION:
            if self.cur_start is None:
                self.cur_start = ts[0]
            nxt = self.cur_start + self.interval
        if self.mode == AREA_COUNT and self.area_counts is None:
            self.area_counts = np.zeros((1 + W // self.area_dimension, 1 + H // self.area_dimension), dtype=int)
        out, idx, done = [], 0, False
        while not done:
            if self.mode == DURATION:
                start = int(np.searchsorted(ts[idx:], self.cur_start, side="left"))
                end = int(np.searchsorted(ts[idx:], nxt, side="right"))
            elif self.mode == COUNT:
                start, end = idx, idx + int(self.value)
            elif self.mode == AREA_COUNT:
                start = idx
                e = start
                for e in range(start, n):
                    x, y = int(ev[e, 1] // self.area_dimension), int(ev[e, 2] // self.area_dimension)
                    c = 1 + self.area_counts[x, y]
                    self.area_counts[x, y] = c
                    if c >= int(self.value):
                        self.area_counts = np.zeros_like(self.area_counts)
                        break
                end = e
            else:
                start, end = 0, n
            if end >= n - 1:
                done, end = True, n - 1
            frame = self._frame(ev[start:end], H, W)
            if not done or self.mode == SOURCE:
                if self.mode == DURATION:
                    self.cur_start += self.interval
                    nxt = self.cur_start + self.interval
                elif self.mode in (COUNT, AREA_COUNT):
                    idx = end
                out.append((frame + self.fs) / float(self.fs * 2))
        return np.stack(out) if out else None
