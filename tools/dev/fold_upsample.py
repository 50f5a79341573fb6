"""Folding a x2 bilinear upsample (align_corners=False) into the 3x3 convolution that follows it:
conv3x3(up2(L)) == four phase-specific 3x3 convolutions over the LOW-resolution tensor L (interior pixels).
Checks the coefficient derivation against torch before it goes into the CUDA weight packer."""
import numpy as np, torch
torch.manual_seed(0)
B, C, Co, h, w = 1, 5, 4, 9, 11
L = torch.randn(B, C, h, w, dtype=torch.float64)
W = torch.randn(Co, C, 3, 3, dtype=torch.float64)
up = torch.nn.functional.interpolate(L, scale_factor=2, mode="bilinear", align_corners=False)
ref = torch.nn.functional.conv2d(up, W, padding=1)            # [B,Co,2h,2w]

def coef(i, k):
    """weight of low-res index k in upsampled index i (interior, no clamping): i=2m -> 0.25*L[m-1]+0.75*L[m];
    i=2m+1 -> 0.75*L[m]+0.25*L[m+1]"""
    m, odd = divmod(i, 2)
    if not odd:
        return {m - 1: 0.25, m: 0.75}.get(k, 0.0)
    return {m: 0.75, m + 1: 0.25}.get(k, 0.0)

# folded weights: out[2m+py][2j+px] = sum_{a,b in -1..1} Wf[py][px][a][b] . L[m+a][j+b]
Wf = torch.zeros(2, 2, 3, 3, Co, C, dtype=torch.float64)
for py in range(2):
    for px in range(2):
        for r in range(3):
            for s in range(3):
                iy, ix = py + r - 1, px + s - 1        # upsampled offsets relative to (2m, 2j)
                for a in (-1, 0, 1):
                    for b in (-1, 0, 1):
                        cy = coef(iy + 100, 50 + a)     # shift by 100 (even) to stay positive
                        cx = coef(ix + 100, 50 + b)
                        Wf[py, px, a + 1, b + 1] += W[:, :, r, s] * cy * cx
out = torch.zeros_like(ref)
Lp = torch.nn.functional.pad(L, (1, 1, 1, 1))
for py in range(2):
    for px in range(2):
        acc = torch.zeros(B, Co, h, w, dtype=torch.float64)
        for a in range(3):
            for b in range(3):
                acc += torch.einsum("oc,bchw->bohw", Wf[py, px, a, b], Lp[:, :, a:a + h, b:b + w])
        out[:, :, py::2, px::2] = acc
d = (out - ref).abs()
print("interior max err", d[:, :, 2:-2, 2:-2].max().item(), " border (2 px frame) max err", d.max().item())
# which vertical low-row offsets each output-row offset touches: input row k -> output rows 2k-2 .. 2k+3
for py in range(2):
    print("py", py, "nonzero a:", [a - 1 for a in range(3) if Wf[py, :, a].abs().sum() > 0])
