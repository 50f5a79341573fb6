"""Development: times of the multi-frame pixel-model path (v2e_emu_time_fused) next to the frame-by-frame kernels."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from v2e_b200 import EventEmulator, _lib
only_fused = '--fused-only' in sys.argv
for (H, W, T) in ((720, 1280, 80), (260, 346, 300)):
    fr = bench.source_clip(H, W, T + 1, px_per_frame=1)          # loops: frame T equals frame 0
    frd = torch.from_numpy(fr).cuda()
    for fused in ((True,) if only_fused else (True, False)):
        em = EventEmulator(device="cuda:0", rng_mode="device", seed=3, max_frames_per_step=T, fused=fused, **bench.CLI_DEFAULTS)
        em.event_rows_hint = 40 * 1024 * 1024
        k = 0
        def run():
            global k
            n = T + (0 if k else 1)
            r = em.generate_events_batch(frd[1:] if k else frd, (np.arange(n) + k) / 300.0, return_device=True)
            k += n
            return r
        for rep in range(3):
            rows, offs = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(4):
            rows, offs = run()
        e1.record(); torch.cuda.synchronize()
        us_frame = e0.elapsed_time(e1) * 1e3 / (4 * T)
        _lib.check(em._lib.v2e_emu_profile(em._h, 1))
        run()
        ms3, n3 = (ctypes.c_float * 4)(), (ctypes.c_int * 4)()
        _lib.check(em._lib.v2e_emu_profile_read4(em._h, ms3, n3, em._stream()))
        _lib.check(em._lib.v2e_emu_profile(em._h, 0))
        msg = "%dx%d T=%d fused=%s: %.2f us/frame through generate_events_batch, %.0f ev/frame | brackets (us per chunk or frame): update %.1f count/filter %.1f emit %.1f" % (
            W, H, T, fused, us_frame, rows.shape[0] / T, ms3[0] * 1e3 / (1 if fused else T), ms3[1] * 1e3 / (1 if fused else T), ms3[2] * 1e3 / (1 if fused else T))
        if fused:
            a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
            em._lib.v2e_emu_fused_stats(em._h, ctypes.byref(a), ctypes.byref(b))
            uc, uu = ctypes.c_float(0), ctypes.c_float(0)
            ts = (ctypes.c_double * T)(*[(k + j) / 300.0 for j in range(T)])
            _lib.check(em._lib.v2e_emu_time_fused(em._h, ctypes.c_void_p(frd[1:].data_ptr()), 0, T, ts, float(em.t_previous),
                                                  ctypes.c_void_p(em._ev_dev.data_ptr()), em._ev_dev.shape[0], 10,
                                                  ctypes.byref(uc), ctypes.byref(uu), em._stream()))
            msg += " | chunks %d rejected %d | time_fused: chunk %.1f us (%.2f us/frame), update kernel %.1f us (%.2f us/frame) -> %.0f GB/s at 53 B/px/frame" % (
                a.value, b.value, uc.value, uc.value / T, uu.value, uu.value / T, H * W * 53 / (uc.value / T) / 1e3)
        print(msg, flush=True)
        em.cleanup()
