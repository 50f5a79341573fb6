"""STRIP2_DEBUG build only: issuer wait breakdown per strip2 layer."""
import ctypes, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools/dev')
from test_conv import L, pad16, cout_pad, dev
B, H, W = 8, 704, 1280
layers = [("conv1", 12, 0, 32, 7, 0), ("conv2", 32, 0, 32, 7, 0), ("down1.c1", 32, 0, 64, 5, 1), ("down1.c2", 64, 0, 64, 5, 1),
          ("up4.c1", 128, 0, 64, 3, 1), ("up5.c1", 64, 0, 32, 3, 0), ("up5.c2", 32, 32, 32, 3, 0), ("conv3", 32, 0, 5, 3, 0)]
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
    f = lambda: L.v2e_conv2d_lrelu_sm100_strip(p(a1), c1p, p(a2), c2p, p(wt), p(bias), cp, k, k, B, h, w, p(out), cp, mode, min(co, 8), ctypes.c_float(0.1), st)
    for _ in range(2): f()
    torch.cuda.synchronize()
    L.v2e_strip2_debug_dump()
    e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
    f()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.3f ms" % e0.elapsed_time(e1)); sys.stdout.flush()
    L.v2e_strip2_debug_dump()
