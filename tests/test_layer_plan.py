"""CPU: which convolution kernel every UNet layer gets (host logic of the C ABI, no GPU needed). Pins the layer
plan DESIGN.md 4.2 describes, so that a change of thresholds shows up as a test diff rather than as a silent
slow-down: strip kernels for the wide, narrow-channel layers; per-tap kernel elsewhere; the up-sampling fold
for up5.conv1 at 1280 px only."""
import pytest

from v2e_b200 import _lib


def pad16(c):
    return (c + 15) // 16 * 16


def cout_pad(c):
    p = pad16(c)
    return 16 if p <= 16 else 32 if p <= 32 else 64 if p <= 64 else (p + 127) // 128 * 128


# (name, cin1, cin2, cout, k, level) of UNet(12, 5) in forward order (model.py:184-196)
LAYERS = [("conv1", 12, 0, 32, 7, 0), ("conv2", 32, 0, 32, 7, 0),
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


def strip_layers(W):
    lib = _lib.load()
    out = {}
    for name, c1, c2, co, k, lvl in LAYERS:
        kc = lib.v2e_conv_strip_pick_kc(pad16(c1), pad16(c2) if c2 else 0, cout_pad(co), k, k, W >> lvl)
        if kc:
            out[name] = kc
    return out


def test_layer_plan_at_1280():
    assert strip_layers(1280) == {"conv1": 16, "conv2": 32, "down1.c1": 32, "down1.c2": 64, "up4.c1": 64,
                                  "up4.c2": 64, "up5.c1": 64, "up5.c2": 32, "conv3": 32}


def test_layer_plan_at_320():
    # 346x260 runs the networks at 320x256: only the full-resolution layers are wide enough (>= 256 px)
    assert strip_layers(320) == {"conv1": 16, "conv2": 32, "up5.c1": 64, "up5.c2": 32, "conv3": 32}


def test_small_frames_use_the_per_tap_kernel_only():
    assert strip_layers(128) == {}


@pytest.mark.parametrize("w_out,want", [(1280, 1), (512, 1), (320, 0)])
def test_upsampling_fold_applies_to_up5_conv1_on_wide_frames(w_out, want):
    lib = _lib.load()
    assert lib.v2e_conv_up2_supported_c(64, 32, w_out) == want        # up5.conv1: 64 -> 32
    assert lib.v2e_conv_up2_supported_c(128, 64, w_out // 2) == 0     # up4.conv1: folded weights do not fit
