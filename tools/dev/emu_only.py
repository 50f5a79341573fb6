"""Emulator-only loop for profiling: 1280x720, CLI defaults, device RNG."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bench import source_clip, CLI_DEFAULTS
from v2e_b200 import EventEmulator
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
import bench
fr = bench.source_clip(720, 1280, 2 * T + 1, px_per_frame=1)[:T]
em = EventEmulator(device="cuda:0", rng_mode="device", seed=3, max_frames_per_step=64, **CLI_DEFAULTS)
em.event_rows_hint = 40 * 1024 * 1024
frd = torch.from_numpy(fr).cuda()
for rep in range(3):
    rows, offs = em.generate_events_batch(frd, np.arange(T) / 300.0 + rep * T / 300.0, return_device=True)
torch.cuda.synchronize()
print('events/frame', rows.shape[0] / T)
