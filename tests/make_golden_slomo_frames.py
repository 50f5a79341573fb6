"""Synthetic clip shared by the SloMo tests (same generator as oracle/make_golden_slomo.py)."""
import numpy as np
import torch


def smooth_frames(N, H, W, seed, dx=3, dy=1, up=8):
    rng = np.random.default_rng(seed)
    ph, pw = H + dy * N + 2 * up, W + dx * N + 2 * up
    base = torch.from_numpy(rng.uniform(20, 235, (1, 1, ph // up + 3, pw // up + 3)).astype(np.float32))
    big = torch.nn.functional.interpolate(base, scale_factor=up, mode="bicubic", align_corners=False)[0, 0]
    big = big.clamp(0, 255).round().to(torch.uint8).numpy()
    return np.stack([big[k * dy:k * dy + H, k * dx:k * dx + W] for k in range(N)])
