"""Pixel-model kernel times (CUDA events through v2e_emu_profile): update / filter / emit us per frame."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from v2e_b200 import EventEmulator, _lib
for (H, W) in ((720, 1280), (260, 346)):
    T = 64
    fr = bench.source_clip(H, W, 2 * T + 1, px_per_frame=1)[:T]
    em = EventEmulator(device="cuda:0", rng_mode="device", seed=3, max_frames_per_step=T, **bench.CLI_DEFAULTS)
    em.event_rows_hint = 40 * 1024 * 1024
    frd = torch.from_numpy(fr).cuda()
    for rep in range(4):
        if rep == 3:
            _lib.check(em._lib.v2e_emu_profile(em._h, 1))
        rows, offs = em.generate_events_batch(frd, np.arange(T) / 300.0 + rep * T / 300.0, return_device=True)
    ms3, n3 = (ctypes.c_float * 4)(), (ctypes.c_int * 4)()
    _lib.check(em._lib.v2e_emu_profile_read4(em._h, ms3, n3, em._stream()))
    us = [ms3[i] / max(n3[i], 1) * 1e3 for i in range(4)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(em._lib.v2e_emu_profile(em._h, 0))
    torch.cuda.synchronize()
    e0.record()
    for rep in range(4, 8):
        rows, offs = em.generate_events_batch(frd, np.arange(T) / 300.0 + rep * T / 300.0, return_device=True)
    e1.record(); torch.cuda.synchronize()
    print("%dx%d: update %.2f us  filter %.2f us  emit %.2f us  (null bracket %.2f us) | whole frame step %.2f us | ev/frame %.0f | update %.0f GB/s (47 B/px)" % (
        W, H, us[0], us[1], us[2], us[3], e0.elapsed_time(e1) * 1e3 / (4 * T), rows.shape[0] / T, H * W * 47 / us[0] / 1e3))
    em.cleanup()
