import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
from test_emulator_gpu import texture_frames, _emulator
from emu_oracle import OracleEmulator
H, W, T = 720, 1280, 6
kw = dict(sigma_thres=0.0, cutoff_hz=300, leak_rate_hz=0, shot_noise_rate_hz=0)
fr = texture_frames(H, W, T, seed=8, speed=2.0)
ts = [k * 2e-3 for k in range(T)]
for hint in (None, 40_000_000):
    em = _emulator(rng_mode="device", **kw)
    em.event_rows_hint = hint
    rows, offs = em.generate_events_batch(fr, ts)
    d = (em.lp_log_frame - em.base_log_frame).abs()
    bad = torch.nonzero(d >= 0.2)
    print('hint',hint,'max',d.max().item(),'nbad',len(bad), bad[:5].tolist(), 'events', len(rows), 'buf', em._ev_dev.shape[0])
    if len(bad):
        y,x = bad[0].tolist()
        y0=max(0,min(H-8,y-4)); x0=max(0,min(W-8,x-4))
        orc=OracleEmulator(**kw)
        for i in range(T): orc.generate_events(fr[i,y0:y0+8,x0:x0+8], ts[i])
        print(' gpu lp',em.lp_log_frame[y,x].item(),'base',em.base_log_frame[y,x].item())
        print(' orc lp',orc.lp[y-y0,x-x0],'base',orc.base[y-y0,x-x0], 'pix', fr[:,y,x])
