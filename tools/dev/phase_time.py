"""Where a bench step goes: SloMo phase vs pixel-model phase (CUDA events), conv share via the profile hooks."""
import ctypes, sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import bench
from v2e_b200 import EventEmulator, SuperSloMo, V2EPipeline, _lib
H, W, NS, U = 720, 1280, 9, 10
src = torch.from_numpy(bench.source_clip(H, W, NS)).cuda()
sl = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=8, device="cuda:0", state_dicts=bench.slomo_weights())
em = EventEmulator(device="cuda:0", rng_mode="device", seed=1, max_frames_per_step=80, **bench.CLI_DEFAULTS)
em.event_rows_hint = 48 * 1024 * 1024
clip_s = 8 / 30.0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for rep in range(6):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    ev[0].record()
    interp, times, _ = sl.interpolate_frames(src)
    w1 = time.perf_counter()
    ev[1].record()
    t = rep * clip_s + clip_s / (times.max() - times.min()) * times
    rows, offs = em.generate_events_batch(interp, t, return_device=True)
    w2 = time.perf_counter()
    ev[2].record()
    torch.cuda.synchronize()
    print("rep %d: slomo %.2f ms (host enqueue %.2f ms)  pixel model %.2f ms (host %.2f ms)  events %d" % (
        rep, ev[0].elapsed_time(ev[1]), (w1 - w0) * 1e3, ev[1].elapsed_time(ev[2]), (w2 - w1) * 1e3, rows.shape[0]))
