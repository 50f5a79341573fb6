"""Development: small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).
    compute-sanitizer --tool memcheck python tools/dev/sanitize_smoke.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
import bench
from v2e_b200 import EventEmulator
from v2e_b200.prep import InputPrep
from v2e_b200.renderer import EventRenderer, ExposureMode

g.smoke()                                   # replay-mode pixel model (update / filter / shot / emit) + SloMo (strip + per-tap convs)
# multi-frame pixel-model path, with noise, ragged size, a rejected chunk, capacity growth
for kw, speed in ((bench.CLI_DEFAULTS, 1), (dict(cutoff_hz=200, refractory_period_s=0.004, pos_thres=0.05, neg_thres=0.05, leak_rate_hz=0.1, shot_noise_rate_hz=2), 10)):
    fr = bench.source_clip(37, 53, 41, px_per_frame=speed, up=8)[:20]
    em = EventEmulator(device="cuda:0", rng_mode="device", seed=3, max_frames_per_step=7, **kw)
    em.event_rows_hint = 64
    rows, offs = em.generate_events_batch(fr, np.arange(20) / 300.0)
    print("fused path:", rows.shape[0], "events")
    em.cleanup()
# centre-surround model (single GPU ring of 2)
em = EventEmulator(device="cuda:0", cs_lambda_pixels=4, cs_tau_p_ms=2.0, cutoff_hz=200, leak_rate_hz=0, shot_noise_rate_hz=0)
fr = bench.source_clip(40, 56, 9, px_per_frame=2, up=8)[:4]
for k in range(4):
    em.generate_events(fr[k], k * 5e-4)
print("csdvs steps", em.cs_steps_taken)
em.cleanup()
# stage-1 prep and renderer
rng = np.random.default_rng(0)
p = InputPrep((131, 77), (46, 35), channels=3, crop=(3, 2, 1, 0))
print("prep", p(rng.integers(0, 256, (2, 77, 131, 3), dtype=np.uint8)).float().mean().item())
p.close()
r = EventRenderer(exposure_mode=ExposureMode.DURATION, exposure_value=0.002)
ev = np.stack([np.sort(rng.uniform(0, 0.01, 5000)), rng.integers(0, 32, 5000), rng.integers(0, 24, 5000), np.where(rng.random(5000) < 0.5, 1, -1)], 1).astype(np.float32)
print("render", r.render_events_to_frames(ev, 24, 32, return_frames=True).shape)
torch.cuda.synchronize()
print("sanitize_smoke done")
