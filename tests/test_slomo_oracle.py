"""CPU: pins the float32 torch restatement (oracle/slomo_ref.py) against fixtures produced by the
unmodified reference classes (oracle/make_golden_slomo.py)."""
import hashlib

import numpy as np
import pytest
import torch

import slomo_ref
from helpers import GOLDEN_DIR
import os

CASES = ["slomo_64x96_u2_b1", "slomo_70x100_u3_b2", "slomo_96x130_auto"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def weights_for(g):
    seed = int(g["seed"])
    sd_fc = slomo_ref.make_test_weights(100 + seed, 2, 4, head_gain=25.0)
    sd_at = slomo_ref.make_test_weights(200 + seed, 12, 5, head_gain=0.3)
    h = hashlib.sha1()
    for sd in (sd_fc, sd_at):
        hh = hashlib.sha1()
        for k in sorted(sd):
            hh.update(sd[k].numpy().tobytes())
        h.update(hh.hexdigest().encode())
    dig = "".join(hashlib.sha1(b"".join(sd[k].numpy().tobytes() for k in sorted(sd))).hexdigest()
                  for sd in (sd_fc, sd_at))
    if dig != str(g["weights_sha1"]):
        pytest.skip("torch.randn on this host does not reproduce the fixture's weights")
    return sd_fc, sd_at


@pytest.mark.parametrize("name", CASES[:2])
def test_restatement_matches_reference_golden(name):
    g = load(name)
    sd_fc, sd_at = weights_for(g)
    torch.set_num_threads(min(8, torch.get_num_threads()))
    out, times, avg = slomo_ref.interpolate_frames(g["frames"], sd_fc, sd_at, int(g["U"]),
                                                   batch_size=int(g["batch_size"]), auto_upsample=bool(g["auto"]))
    assert out.shape == g["out"].shape
    assert np.array_equal(times, g["times"]) and avg == float(g["avg"])
    d = np.abs(out.astype(np.int32) - g["out"].astype(np.int32))
    # same float32 ops; conv reduction order may differ with the host's thread count: <= 1 DN, rarely
    assert d.max() <= 1 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())


def test_network_is_sensitive():
    """The fixture weights must make the UNets matter (flows of ~pixels, visibility not constant),
    otherwise output parity would only test the blend."""
    g = load(CASES[0])
    sd_fc, sd_at = weights_for(g)
    I, _ = slomo_ref.load_pair_tensors(g["frames"][:2], slomo_ref.net_dims(g["frames"].shape[2], g["frames"].shape[1]))
    flow, outs = slomo_ref.interp_batch(sd_fc, sd_at, I[:1], I[1:2], 2)
    assert flow.abs().mean() > 0.3
    intrp, ft = outs[0]
    assert intrp[:, 4].std() > 0.05
