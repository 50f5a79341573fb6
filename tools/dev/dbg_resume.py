import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
from test_emulator_gpu import texture_frames, _emulator
from emu_oracle import OracleEmulator
from helpers import canonical
kw = dict(sigma_thres=0.0, cutoff_hz=0, leak_rate_hz=0, shot_noise_rate_hz=0)
H, W, T = 48, 64, 5
fr = texture_frames(H, W, T, seed=1, speed=2.0)
ts = [k * 1e-3 for k in range(T)]
orc = OracleEmulator(**kw)
want = [orc.generate_events(f, t) for f, t in zip(fr, ts)]
for hint in (1<<20, 64):
    em = _emulator(rng_mode="device", **kw)
    em.event_rows_hint = hint
    rows, offs = em.generate_events_batch(fr, ts)
    print('hint',hint,'offs',offs, 'buf', em._ev_dev.shape)
    for i in range(1,T):
        a=canonical(rows[offs[i]:offs[i+1]]); b=canonical(want[i])
        print(' frame',i,len(a),len(b),'equal',np.array_equal(a,b))
        if len(a)==len(b) and not np.array_equal(a,b):
            bad=np.nonzero((a!=b).any(1))[0]
            print('  first bad rows', bad[:5]); print(a[bad[:3]]); print(b[bad[:3]])
            sa=set(map(tuple,a)); sb=set(map(tuple,b)); print('  only in got',len(sa-sb),'only in want',len(sb-sa), list(sa-sb)[:3], list(sb-sa)[:3])
            print('  dup in got', len(a)-len(sa))
