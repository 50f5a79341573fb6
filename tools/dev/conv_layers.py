"""Per-layer timing of the UNet convolutions at 1280x704, batch B (standalone conv ABI, random data)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools/dev')
from test_conv import L, pad16, cout_pad, dev
USE_ROW = (len(sys.argv) > 2 and sys.argv[2] == 'row')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 704, 1280
layers = [("conv1", 12, 0, 32, 7, 0), ("conv2", 32, 0, 32, 7, 0),
          ("down1.c1", 32, 0, 64, 5, 1), ("down1.c2", 64, 0, 64, 5, 1),
          ("down2.c1", 64, 0, 128, 3, 2), ("down2.c2", 128, 0, 128, 3, 2),
          ("down3.c1", 128, 0, 256, 3, 3), ("down3.c2", 256, 0, 256, 3, 3),
          ("down4.c1", 256, 0, 512, 3, 4), ("down4.c2", 512, 0, 512, 3, 4),
          ("down5.c1", 512, 0, 512, 3, 5), ("down5.c2", 512, 0, 512, 3, 5),
          ("up1.c1", 512, 0, 512, 3, 4), ("up1.c2", 512, 512, 512, 3, 4),
          ("up2.c1", 512, 0, 256, 3, 3), ("up2.c2", 256, 256, 256, 3, 3),
          ("up3.c1", 256, 0, 128, 3, 2), ("up3.c2", 128, 128, 128, 3, 2),
          ("up4.c1", 128, 0, 64, 3, 1), ("up4.c2", 64, 64, 64, 3, 1),
          ("up5.c1", 64, 0, 32, 3, 0), ("up5.c2", 32, 32, 32, 3, 0), ("conv3", 32, 0, 5, 3, 0)]
tot_ms = 0; tot_fl = 0
print("B=%d" % B)
for name, c1, c2, co, k, lvl in layers:
    h, w = H >> lvl, W >> lvl
    c1p, c2p, cp = pad16(c1), (pad16(c2) if c2 else 0), cout_pad(co)
    a1 = torch.randn((B, h, w, c1p), device=dev).half()
    a2 = torch.randn((B, h, w, c2p), device=dev).half() if c2 else None
    wt = (torch.randn((cp, k * k * (c1p + c2p)), device=dev) * 0.02).half()
    bias = torch.zeros(cp, device=dev)
    mode = 1 if name == "conv3" else 0
    out = torch.empty((B, h, w, 8 if mode else cp), dtype=torch.float32 if mode else torch.float16, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    kc = L.v2e_conv_strip_pick_kc(c1p, c2p, cp, k, k, w) if USE_ROW else 0
    if kc:
        f = lambda: L.v2e_conv2d_lrelu_sm100_strip(p(a1), c1p, p(a2), c2p, p(wt), p(bias), cp, k, k, B, h, w, p(out), cp, mode, min(co, 8), ctypes.c_float(0.1), st)
        name = name + '*'
    else:
        f = lambda: L.v2e_conv2d_lrelu_sm100(p(a1), c1p, p(a2), c2p, p(wt), p(bias), cp, k, k, B, h, w, p(out), cp, mode, min(co, 8), ctypes.c_float(0.1), st)
    for _ in range(2): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * h * w * co * (c1 + c2) * k * k
    tot_ms += ms; tot_fl += fl
    print("%-10s %4dx%-4d C%4d+%-3d->%3d k%d: %7.3f ms %7.1f TFLOP/s (%4.1f%% of time so far)" % (name, h, w, c1, c2, co, k, ms, fl / ms / 1e9, 0))
print("total %.2f ms  %.1f GFLOP  %.1f TFLOP/s" % (tot_ms, tot_fl / 1e9, tot_fl / tot_ms / 1e9))
