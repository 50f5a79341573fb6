"""GPU harness: halo-resident row conv kernel vs torch; probes the descriptor base-offset policy."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools/dev')
from test_conv import L, pad16, cout_pad, to_nhwc16, dev

def pack_w_row(w, C1, C2, KC):
    Cout, Cin, KH, KW = w.shape
    C1p, C2p = pad16(C1), (pad16(C2) if C2 else 0)
    Cp = cout_pad(Cout)
    full = torch.zeros((Cp, KH * KW, C1p + C2p), dtype=torch.float16, device=w.device)
    wt = w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).half()
    full[:Cout, :, :C1] = wt[:, :, :C1]
    if C2: full[:Cout, :, C1p:C1p + C2] = wt[:, :, C1:]
    slabs = (C1p + C2p) // KC
    # [Cp][taps][slabs][KC] -> [slabs][taps][Cp][KC]
    return full.reshape(Cp, KH * KW, slabs, KC).permute(2, 1, 0, 3).contiguous(), Cp

def run_case(N, H, W, C1, C2, Cout, K, bo, out_mode=0, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    x1 = torch.randn((N, C1, H, W), generator=g).to(dev)
    x2 = torch.randn((N, C2, H, W), generator=g).to(dev) if C2 else None
    w = (torch.randn((Cout, C1 + C2, K, K), generator=g) / np.sqrt((C1 + C2) * K * K)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    a1 = to_nhwc16(x1); a2 = to_nhwc16(x2) if C2 else None
    Cp = cout_pad(Cout)
    KC = L.v2e_conv_strip_pick_kc(a1.shape[-1], a2.shape[-1] if C2 else 0, Cp, K, K, W)
    if KC == 0:
        print('skip (no KC)', N, H, W, C1, C2, Cout, K); return True
    wp, Cp = pack_w_row(w, C1, C2, KC)
    bp = torch.zeros(Cp, device=dev); bp[:Cout] = b
    out = torch.full((N, H, W, Cp if out_mode == 0 else 8), float('nan'), dtype=torch.float16 if out_mode == 0 else torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    rc = L.v2e_conv2d_lrelu_sm100_strip(p(a1), a1.shape[-1], p(a2), a2.shape[-1] if C2 else 0, p(wp), p(bp), Cp, K, K,
                                        N, H, W, p(out), Cp, out_mode, min(Cout, 8), ctypes.c_float(0.1), st)
    if rc != 0:
        print('ERR', rc, L.v2e_last_error()); return False
    torch.cuda.synchronize()
    xin = torch.cat([x1, x2], 1) if C2 else x1
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(xin.half().float(), w.half().float(), b, padding=K // 2), 0.1).permute(0, 2, 3, 1)
    got = out[..., :min(Cout, out.shape[-1])].float(); refc = ref[..., :got.shape[-1]]
    err = (got - refc).abs(); tol = 2e-3 * refc.abs() + 2e-3
    ok = bool(torch.isfinite(got).all() and (err <= tol).all())
    print('bo%d N%d %dx%d C%d+%d->%d k%d KC%d mode%d: max_err %.3e frac_bad %.4f %s' % (bo, N, H, W, C1, C2, Cout, K, KC, out_mode,
          err[torch.isfinite(err)].max().item() if torch.isfinite(err).any() else float('nan'), (~(err <= tol)).float().mean().item(), 'OK' if ok else 'FAIL'))
    return ok

if __name__ == '__main__':
    cases = [(1, 4, 512, 64, 0, 32, 3), (1, 40, 512, 32, 0, 32, 3), (1, 33, 530, 16, 0, 32, 3), (1, 61, 600, 64, 0, 64, 3),
             (1, 70, 512, 32, 0, 32, 7), (2, 64, 640, 12, 0, 32, 7), (2, 75, 576, 32, 32, 32, 3), (1, 48, 640, 32, 0, 64, 5),
             (1, 256, 520, 32, 0, 32, 7), (1, 60, 512, 32, 0, 5, 3, None, 1), (3, 5, 513, 64, 0, 32, 3)]
    res = {0: True, 1: True}
    for bo in (0,):
        for c in cases:
            c = list(c)
            mode = 0
            if len(c) > 7: mode = c[8]; c = c[:7]
            try:
                res[bo] &= run_case(*c, bo, mode)
            except Exception as e:
                print('EXC', c, e); res[bo] = False
    print('RESULT bo1', res[1], 'bo0', res[0])
