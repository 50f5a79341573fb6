"""CPU: the synthetic inputs of bench.py are what SURVEY.md 8(d) / BASELINE.json name. The config-2 clip must be the
reference's own scripts/gradients.py pattern (checked against the script itself where /root/reference exists, and
against properties of im_function everywhere); the config-3 / config-5 block texture must be deterministic."""
import os
import sys

import numpy as np
import pytest

import bench


def test_gradient_clip_properties():
    fr = bench.gradient_clip(260, 346, 31)
    assert fr.shape == (31, 260, 346) and fr.dtype == np.uint8
    low, high = np.uint8((127 * 2) / 3), np.uint8(2 * (127 * 2) / 3)      # contrast 2 around the background 127
    assert fr.min() == low and fr.max() == high
    assert (fr == fr[:, :1, :]).all()                                  # constant along y
    # the bump's peak moves 300 px/s = 10 px per 30 fps frame
    peaks = [int(np.argmax(f[0, :int(0.5 * 346) + 10 * k + 2])) for k, f in enumerate(fr[:10])]
    assert np.all(np.diff(peaks) == 10), peaks


def test_gradient_clip_equals_reference_script():
    root = "/root/reference"
    if not os.path.isfile(os.path.join(root, "scripts", "gradients.py")):
        pytest.skip("reference tree not present")
    import ref_shim
    ref_shim.load_reference()
    sys.path.insert(0, os.path.join(root, "scripts"))
    try:
        import gradients as g
    finally:
        sys.path.pop(0)
    m = g.gradients.__new__(g.gradients)          # im_function only needs these attributes (gradients.py:117-140)
    m.bg, m.contrast, m.bump_width, m.w, m.h, m.speed_pps = 127, 2.0, 0.5, 346, 260, 300.0
    mine = bench.gradient_clip(260, 346, 8)
    for k in range(8):
        ref = m.im_function(np.arange(260)[:, None], np.arange(346)[None, :], k / 30.0)
        assert np.array_equal(mine[k], ref), k


def test_block_texture_clip_is_deterministic_and_translates():
    a = bench.block_texture_clip(64, 96, 5, seed=0)
    b = bench.block_texture_clip(64, 96, 5, seed=0)
    assert np.array_equal(a, b) and a.dtype == np.uint8
    assert np.array_equal(a[1][:-4, :-8], a[0][4:, 8:])                # (+8, +4) px per source frame
    assert len(np.unique(a[0])) > 50


def test_unet_activation_bytes_matches_a_hand_count():
    # one layer by hand: conv2 of UNet(12, 5) at 1280x704, batch 8: 32 channels in + 32 out, fp16
    tot = bench.unet_activation_bytes(12, 5, 704, 1280, 8)
    conv2 = 8 * 704 * 1280 * (32 + 32) * 2
    assert tot > 5 * conv2 and tot < 12 * conv2
