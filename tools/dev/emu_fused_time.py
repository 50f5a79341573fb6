"""Development: times of the multi-frame pixel-model path (v2e_emu_time_fused) next to the frame-by-frame kernels."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from v2e_b200 import EventEmulator, _lib
for (H, W, T) in ((720, 1280, 80), (260, 346, 300)):
    fr = bench.source_clip(H, W, 2 * T + 1, px_per_frame=1)[:T + 1]
    frd = torch.from_numpy(fr).cuda()
    for fused in (True, False):
        em = EventEmulator(device="cuda:0", rng_mode="device", seed=3, max_frames_per_step=T, fused=fused, **bench.CLI_DEFAULTS)
        em.event_rows_hint = 40 * 1024 * 1024
        for rep in range(3):
            rows, offs = em.generate_events_batch(frd[1:] if rep else frd, (np.arange(T + (0 if rep else 1)) + rep * (T + 1)) / 300.0, return_device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(3, 7):
            rows, offs = em.generate_events_batch(frd[1:], (np.arange(T) + rep * (T + 1)) / 300.0, return_device=True)
        e1.record(); torch.cuda.synchronize()
        us_frame = e0.elapsed_time(e1) * 1e3 / (4 * T)
        msg = "%dx%d T=%d fused=%s: %.2f us/frame through generate_events_batch, %.0f ev/frame" % (W, H, T, fused, us_frame, rows.shape[0] / T)
        if fused:
            a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
            em._lib.v2e_emu_fused_stats(em._h, ctypes.byref(a), ctypes.byref(b))
            uc, uu = ctypes.c_float(0), ctypes.c_float(0)
            ts = (ctypes.c_double * T)(*[float(em.t_previous) + (k + 1) / 300.0 for k in range(T)])
            _lib.check(em._lib.v2e_emu_time_fused(em._h, ctypes.c_void_p(frd[1:].data_ptr()), 0, T, ts, float(em.t_previous),
                                                  ctypes.c_void_p(em._ev_dev.data_ptr()), em._ev_dev.shape[0], 10,
                                                  ctypes.byref(uc), ctypes.byref(uu), em._stream()))
            msg += " | chunks %d rejected %d | time_fused: chunk %.1f us (%.2f us/frame), update kernel %.1f us (%.2f us/frame) -> %.0f GB/s at 53 B/px/frame" % (
                a.value, b.value, uc.value, uc.value / T, uu.value, uu.value / T, H * W * 53 / (uc.value / T) / 1e3)
        print(msg, flush=True)
        em.cleanup()
