"""TEST INFRASTRUCTURE ONLY -- float32 torch restatement of the SuperSloMo path.

Not imported by the product (v2e_b200/). Floating-point path, so the checker is a plain PyTorch
fp32 reference of the same ops, written functionally against a state_dict:

  unet_forward         v2ecore/model.py:198-226 (+ down :72-76, up :137-154)
  backwarp             v2ecore/model.py:268-300
  interpolate_frames   v2ecore/slomo.py:330-444 (CUDA-branch transforms, slomo.py:157-162), with the
                       data loader's PIL LANCZOS resize (dataloader.py:136-147) and the PIL BILINEAR
                       output resize (slomo.py:438), frames kept in memory instead of .npy/.png files

Pinned by tests/test_slomo_oracle.py against tests/golden/slomo_*.npz, which oracle/make_golden_slomo.py
produced by running the unmodified reference classes (SuperSloMo.interpolate through temp folders,
with its CUDA-branch transforms forced on the CPU device).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

LAYER_NAMES = (["conv1", "conv2"] +
               ["down%d.conv%d" % (d, c) for d in range(1, 6) for c in (1, 2)] +
               ["up%d.conv%d" % (u, c) for u in range(1, 6) for c in (1, 2)] +
               ["conv3"])


def layer_shapes(in_ch, out_ch):
    ch = [32, 64, 128, 256, 512, 512]
    dk = [5, 3, 3, 3, 3]
    s = [(32, in_ch, 7), (32, 32, 7)]
    for d in range(5):
        s += [(ch[d + 1], ch[d], dk[d]), (ch[d + 1], ch[d + 1], dk[d])]
    uo, ui = [512, 256, 128, 64, 32], [512, 512, 256, 128, 64]
    for k in range(5):
        s += [(uo[k], ui[k], 3), (uo[k], 2 * uo[k], 3)]
    s += [(out_ch, 32, 3)]
    return s


def make_test_weights(seed, in_ch, out_ch, head_gain=1.0):
    """Variance-preserving random weights (He-normal for LeakyReLU(0.1)) so that every layer matters
    to the output; the real checkpoint (SuperSloMo39.ckpt, README.md:95-96) is not available offline."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    gain = math.sqrt(2.0 / (1 + 0.1 ** 2))
    shapes = layer_shapes(in_ch, out_ch)
    for i, (name, (co, ci, k)) in enumerate(zip(LAYER_NAMES, shapes)):
        std = gain / math.sqrt(ci * k * k)
        if i == len(shapes) - 1:
            std *= head_gain
        sd[name + ".weight"] = torch.randn((co, ci, k, k), generator=g) * std
        sd[name + ".bias"] = torch.randn((co,), generator=g) * 0.05
    return sd


def _cl(sd, name, x, pad):
    return F.leaky_relu(F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad), 0.1)


def unet_forward(sd, x):
    x = _cl(sd, "conv1", x, 3)
    s1 = _cl(sd, "conv2", x, 3)
    skips = [s1]
    x = s1
    for d, k in zip(range(1, 6), (5, 3, 3, 3, 3)):
        x = F.avg_pool2d(x, 2)
        x = _cl(sd, "down%d.conv1" % d, x, k // 2)
        x = _cl(sd, "down%d.conv2" % d, x, k // 2)
        skips.append(x)
    x = skips.pop()                       # output of down5
    for u in range(1, 6):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = _cl(sd, "up%d.conv1" % u, x, 1)
        x = _cl(sd, "up%d.conv2" % u, torch.cat((x, skips.pop()), 1), 1)
    return _cl(sd, "conv3", x, 1)


def backwarp(img, flow):
    B, _, H, W = img.shape
    gx, gy = np.meshgrid(np.arange(W), np.arange(H))
    gx = torch.tensor(gx).unsqueeze(0).expand(B, -1, -1).float()
    gy = torch.tensor(gy).unsqueeze(0).expand(B, -1, -1).float()
    x = gx + flow[:, 0]
    y = gy + flow[:, 1]
    x = 2 * (x / W - 0.5)
    y = 2 * (y / H - 0.5)
    return F.grid_sample(img, torch.stack((x, y), dim=3), align_corners=False)


def net_dims(W, H):
    return int(W / 32) * 32, int(H / 32) * 32


def pil_resize(arr_u8, size_wh, filt):
    return np.asarray(Image.fromarray(arr_u8).resize(size_wh, filt))


def load_pair_tensors(frames_u8, dim_wh):
    """dataloader.py:136-147 + slomo.py:157-160: LANCZOS resize, /255, -0.428 -> [N,1,Hd,Wd]"""
    out = []
    for f in frames_u8:
        r = pil_resize(np.ascontiguousarray(f), dim_wh, Image.LANCZOS)
        out.append(torch.from_numpy(r.astype(np.float32) / 255.0 - 0.428))
    return torch.stack(out).unsqueeze(1), np.stack([pil_resize(np.ascontiguousarray(f), dim_wh, Image.LANCZOS)
                                                    for f in frames_u8])


def to_u8(ft):
    """revNormalize + ToPILImage: (x + 0.428).mul(255).byte() on the CPU (truncate, wrap mod 256)."""
    return ((ft + 0.428) * 255.0).to(torch.int32).bitwise_and(255).to(torch.uint8)


@torch.no_grad()
def interp_batch(sd_fc, sd_at, I0, I1, U):
    """slomo.py:343-433 for one batch; returns flowOut and the list over k of (intrpOut, Ft_p)."""
    flow = unet_forward(sd_fc, torch.cat((I0, I1), 1))
    F01, F10 = flow[:, :2], flow[:, 2:]
    outs = []
    for k in range(U):
        t = (k + 0.5) / U
        temp = -t * (1 - t)
        Ft0 = temp * F01 + (t * t) * F10
        Ft1 = ((1 - t) * (1 - t)) * F01 + temp * F10
        g0 = backwarp(I0, Ft0)
        g1 = backwarp(I1, Ft1)
        intrp = unet_forward(sd_at, torch.cat((I0, I1, F01, F10, Ft1, Ft0, g1, g0), 1))
        Ft0f = intrp[:, :2] + Ft0
        Ft1f = intrp[:, 2:4] + Ft1
        V0 = torch.sigmoid(intrp[:, 4:5])
        V1 = 1 - V0
        g0f = backwarp(I0, Ft0f)
        g1f = backwarp(I1, Ft1f)
        Ft = ((1 - t) * V0 * g0f + t * V1 * g1f) / ((1 - t) * V0 + t * V1)
        outs.append((intrp, Ft))
    return flow, outs


@torch.no_grad()
def interpolate_frames(frames_u8, sd_fc, sd_at, U, batch_size=1, auto_upsample=False):
    """In-memory restatement of SuperSloMo.interpolate. frames_u8: [N,H,W]. Returns
    (out_u8 [M,H,W], interpTimes, avgUpsampling)."""
    N, H, W = frames_u8.shape
    dim = net_dims(W, H)
    I, _ = load_pair_tensors(frames_u8, dim)
    n_pairs = N - 1
    out = {}
    times = []
    in_ctr = out_ctr = 0
    ups = []
    while in_ctr < n_pairs:
        b = min(batch_size, n_pairs - in_ctr)
        I0, I1 = I[in_ctr:in_ctr + b], I[in_ctr + 1:in_ctr + b + 1]
        flow = unet_forward(sd_fc, torch.cat((I0, I1), 1))
        u = U
        if auto_upsample:
            sp = torch.cat((torch.sqrt(flow[:, 0] ** 2 + flow[:, 1] ** 2).flatten(1),
                            torch.sqrt(flow[:, 2] ** 2 + flow[:, 3] ** 2).flatten(1)), 1)
            u = int(np.ceil(sp.max().item()))
            if U is not None and U > u:
                u = U
        if u < 2:
            u = 2
        ups.append(u)
        _, outs = interp_batch(sd_fc, sd_at, I0, I1, u)
        for k, (_, Ft) in enumerate(outs):
            q = to_u8(Ft)
            for bi in range(b):
                out[out_ctr + u * bi + k] = pil_resize(q[bi, 0].numpy(), (W, H), Image.BILINEAR)
        times.append(in_ctr + np.array(range(u * b)) * (1 / u))
        in_ctr += b
        out_ctr += u * b
    frames = np.stack([out[i] for i in range(out_ctr)])
    return frames, np.concatenate(times), sum(ups) / len(ups)
