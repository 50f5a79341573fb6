"""Timing of the fused up-sampling convolution (up5.conv1 shape: 64 -> 32 at 1280x704, batch 8)."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools/dev')
from test_conv import L, pad16, cout_pad, dev
B, H, W, C, Co = 8, 704, 1280, 64, 32
a = torch.randn((B, H // 2, W // 2, C), device=dev).half()
w = (torch.randn((Co, C, 3, 3)) / 24).float()
wp = torch.zeros((32, 9 * C), dtype=torch.float16)
wp[:Co] = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).half()
wp = wp.to(dev)
fold = np.zeros((1 * 2 * 3 * 6 * 32 * 64,), np.float16)
wh = np.ascontiguousarray(w.numpy())
assert L.v2e_conv_up2_fold_weights(wh.ctypes.data_as(ctypes.c_void_p), Co, C, 32, C, fold.ctypes.data_as(ctypes.c_void_p)) == 0
fold_d = torch.from_numpy(fold).to(dev)
bias = torch.zeros(32, device=dev)
out = torch.empty((B, H, W, 32), dtype=torch.float16, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
f = lambda: L.v2e_conv2d_up2_lrelu_sm100(p(a), C, p(fold_d), p(wp), p(bias), 32, B, H, W, p(out), 32, ctypes.c_float(0.1), st)
for _ in range(3): assert f() == 0
torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
fl = 2.0 * B * H * W * Co * C * 9
print("fused up2+conv: %.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9))
