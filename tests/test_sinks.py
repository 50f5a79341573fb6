"""Event-sink row conversions: oracle pinned against bytes written by the reference's
own AEDat2Output (tests/golden/sinks_aedat2.npz, oracle/make_golden_sinks.py); the CUDA kernels against the
oracle and the same bytes. Integer work: bit-exact."""
import os

import numpy as np
import pytest

import sinks_oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sinks_aedat2.npz"))
SIZES = [(346, 260), (240, 180), (640, 480)]


@pytest.mark.parametrize("size", SIZES)
def test_oracle_matches_reference_writer_bytes(size):
    w, h = size
    ev = G["events_%dx%d" % size]
    words, n_on = sinks_oracle.aedat2_words(ev, w, h)
    assert words.tobytes() == G["body_%dx%d" % size].tobytes()
    assert n_on == int(G["on_off_%dx%d" % size][0])


def test_h5_rows_oracle_values():
    ev = np.array([[0.0125, 3, 7, 1], [1.9999995, 345, 259, -1], [4000.5, 0, 0, 1]], np.float32)
    r = sinks_oracle.h5_rows(ev)
    assert r.dtype == np.uint32 and r[:, 1:].tolist() == [[3, 7, 1], [345, 259, 0], [0, 0, 1]]
    assert r[0, 0] == 12500 and r[2, 0] == np.uint32(np.float32(4000.5) * np.float32(1e6))


@pytest.mark.gpu
@pytest.mark.parametrize("size", SIZES)
def test_cuda_aedat2_matches_reference_bytes(size):
    import torch
    from v2e_b200 import sinks
    w, h = size
    ev = torch.from_numpy(G["events_%dx%d" % size]).cuda()
    words, n_on = sinks.events_to_aedat2(ev, w, h)
    assert words.cpu().numpy().tobytes() == G["body_%dx%d" % size].tobytes()
    assert int(n_on.item()) == int(G["on_off_%dx%d" % size][0])


@pytest.mark.gpu
def test_cuda_sinks_match_oracle_on_emulator_rows():
    """Rows as the pixel model emits them (ragged counts, float32 timestamps), 1280x720 coordinates for the
    HDF5 rows; an empty stream is a no-op."""
    import torch
    from v2e_b200 import sinks
    rng = np.random.default_rng(5)
    n = 200003
    ev = np.stack([np.sort(rng.uniform(0, 2000.0, n)).astype(np.float32), rng.integers(0, 1280, n).astype(np.float32),
                   rng.integers(0, 720, n).astype(np.float32), rng.choice([-1.0, 1.0], n).astype(np.float32)], 1)
    d = torch.from_numpy(ev).cuda()
    got = sinks.events_to_h5_rows(d).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, sinks_oracle.h5_rows(ev))
    ev2 = ev.copy()
    ev2[:, 1] = np.mod(ev2[:, 1], 346)
    ev2[:, 2] = np.mod(ev2[:, 2], 260)
    ev2[:, 0] = ev2[:, 0] / 1000.0
    w, n_on = sinks.events_to_aedat2(torch.from_numpy(ev2).cuda(), 346, 260)
    want, want_on = sinks_oracle.aedat2_words(ev2, 346, 260)
    assert np.array_equal(w.cpu().numpy(), want) and int(n_on.item()) == want_on
    e0 = torch.zeros((0, 4), dtype=torch.float32, device="cuda")
    assert sinks.events_to_h5_rows(e0).shape == (0, 4) and sinks.events_to_aedat2(e0)[0].shape == (0,)
    with pytest.raises(ValueError):
        sinks.events_to_aedat2(d, 1280, 720)          # aedat2_output.py:61-63: unsupported camera size
