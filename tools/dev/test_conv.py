"""GPU harness: tcgen05 implicit-GEMM conv vs torch conv2d on fp16-rounded operands."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from v2e_b200 import _lib
L = _lib.load()
dev = torch.device('cuda')

def pad16(c): return (c + 15) // 16 * 16
def cout_pad(c):
    p = pad16(c)
    if p <= 16: return 16
    if p <= 32: return 32
    if p <= 64: return 64
    return (p + 127) // 128 * 128

def to_nhwc16(x):                      # [N,C,H,W] fp32 -> [N,H,W,Cp] fp16
    N, C, H, W = x.shape
    out = torch.zeros((N, H, W, pad16(C)), dtype=torch.float16, device=x.device)
    out[..., :C] = x.permute(0, 2, 3, 1).half()
    return out.contiguous()

def pack_w(w, C1, C2):                 # [Cout, C1r+C2r, KH, KW] -> [Cout_pad][KH*KW*(C1p+C2p)]
    Cout, Cin, KH, KW = w.shape
    C1r, C2r = C1, C2
    C1p, C2p = pad16(C1r), (pad16(C2r) if C2r else 0)
    Cp = cout_pad(Cout)
    out = torch.zeros((Cp, KH * KW, C1p + C2p), dtype=torch.float16, device=w.device)
    wt = w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).half()
    out[:Cout, :, :C1r] = wt[:, :, :C1r]
    if C2r: out[:Cout, :, C1p:C1p + C2r] = wt[:, :, C1r:]
    return out.reshape(Cp, -1).contiguous(), Cp

def run_case(N, H, W, C1, C2, Cout, K, out_mode=0, seed=0, verbose=True):
    g = torch.Generator(device='cpu').manual_seed(seed)
    x1 = torch.randn((N, C1, H, W), generator=g).to(dev)
    x2 = torch.randn((N, C2, H, W), generator=g).to(dev) if C2 else None
    w = (torch.randn((Cout, C1 + C2, K, K), generator=g) / np.sqrt((C1 + C2) * K * K)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    a1 = to_nhwc16(x1); a2 = to_nhwc16(x2) if C2 else None
    wp, Cp = pack_w(w, C1, C2)
    bp = torch.zeros(Cp, device=dev); bp[:Cout] = b
    if out_mode == 0:
        out = torch.full((N, H, W, Cp), float('nan'), dtype=torch.float16, device=dev)
    else:
        out = torch.full((N, H, W, 8), float('nan'), dtype=torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    rc = L.v2e_conv2d_lrelu_sm100(p(a1), a1.shape[-1], p(a2), a2.shape[-1] if C2 else 0, p(wp), p(bp), Cp, K, K,
                                  N, H, W, p(out), Cp, out_mode, min(Cout, 8), ctypes.c_float(0.1), st)
    if rc != 0:
        print('ERR', rc, L.v2e_last_error()); return False
    torch.cuda.synchronize()
    xin = torch.cat([x1, x2], 1) if C2 else x1
    ref = torch.nn.functional.conv2d(xin.half().float(), w.half().float(), b, padding=K // 2)
    ref = torch.nn.functional.leaky_relu(ref, 0.1).permute(0, 2, 3, 1)
    got = out[..., :min(Cout, out.shape[-1])].float()
    refc = ref[..., :got.shape[-1]]
    err = (got - refc).abs()
    tol = 2e-3 * refc.abs() + 2e-3
    ok = bool(torch.isfinite(got).all() and (err <= tol).all())
    if verbose:
        print('N%d %dx%d C%d+%d->%d k%d mode%d: max_err %.3e (ref max %.2f) nan %d %s' % (
            N, H, W, C1, C2, Cout, K, out_mode, err.max().item(), refc.abs().max().item(),
            int((~torch.isfinite(got)).sum()), 'OK' if ok else 'FAIL'))
        if not ok:
            bad = torch.nonzero(~(err <= tol))
            print('  first bad idx', bad[:4].tolist(), 'got', got[tuple(bad[0])].item(), 'ref', refc[tuple(bad[0])].item())
    return ok

if __name__ == '__main__':
    cases = [
        (1, 8, 16, 64, 0, 64, 3), (1, 8, 16, 64, 0, 64, 1), (1, 16, 32, 64, 0, 64, 3), (1, 24, 40, 128, 0, 128, 3),
        (2, 17, 23, 64, 0, 32, 3), (1, 32, 32, 32, 0, 32, 7), (1, 32, 48, 32, 0, 64, 5), (1, 16, 32, 16, 0, 32, 7),
        (1, 16, 32, 2, 0, 32, 7), (1, 16, 32, 12, 0, 32, 7), (1, 8, 10, 512, 0, 512, 3), (1, 16, 20, 512, 512, 512, 3),
        (1, 64, 80, 32, 32, 32, 3), (1, 32, 40, 64, 64, 64, 3), (1, 32, 32, 32, 0, 5, 3, 1), (1, 32, 32, 32, 0, 4, 3, 1),
        (2, 64, 96, 256, 0, 128, 3),
    ]
    allok = True
    for c in cases:
        try:
            allok &= run_case(*c)
        except Exception as e:
            print('EXC', c, e); allok = False
    print('ALL OK' if allok else 'SOME FAILED')
    if allok:
        # timing at a big layer: 1280x704, 64->64 5x5 (half res) and 32->32 7x7
        for (N, H, W, C1, Cout, K) in [(1, 704, 1280, 32, 32, 7), (1, 352, 640, 64, 64, 5), (1, 176, 320, 128, 128, 3), (1, 44, 80, 512, 512, 3)]:
            x1 = torch.randn((N, C1, H, W), device=dev); w = torch.randn((Cout, C1, K, K), device=dev) * 0.01
            a1 = to_nhwc16(x1); wp, Cp = pack_w(w, C1, 0); bp = torch.zeros(Cp, device=dev)
            out = torch.empty((N, H, W, Cp), dtype=torch.float16, device=dev)
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            f = lambda: L.v2e_conv2d_lrelu_sm100(p(a1), C1, None, 0, p(wp), p(bp), Cp, K, K, N, H, W, p(out), Cp, 0, 8, ctypes.c_float(0.1), st)
            for _ in range(3): f()
            torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
            fl = 2.0 * N * H * W * Cout * C1 * K * K
            print('%dx%d %d->%d k%d: %.3f ms  %.1f TFLOP/s' % (H, W, C1, Cout, K, ms, fl / ms / 1e9))
