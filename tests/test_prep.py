"""Stage-1 input preparation (v2e.py:687-737: crop, cv2.resize INTER_AREA, BGR -> luma).
CPU: the numpy oracle against OpenCV's own output (fixtures written by oracle/make_golden_prep.py, and cv2 live when
it imports). GPU: the CUDA kernel against the same fixtures and, at video sizes, against the oracle -- bit-exact."""
import os

import numpy as np
import pytest

import prep_oracle
from helpers import GOLDEN_DIR


def _cases():
    z = np.load(os.path.join(GOLDEN_DIR, "prep_cv2.npz"))
    for i in range(int(z["n_cases"])):
        crop = tuple(int(c) for c in z["crop_%d" % i])
        yield i, z["in_%d" % i], tuple(int(v) for v in z["wh_%d" % i]), (None if crop[0] < 0 else crop), z["out_%d" % i]


def test_oracle_matches_opencv_fixtures():
    n = 0
    for i, fr, wh, crop, want in _cases():
        for k in range(fr.shape[0]):
            got = prep_oracle.prep_frame(fr[k], wh, crop)
            assert got.dtype == np.uint8 and np.array_equal(got, want[k]), "case %d" % i
        n += 1
    assert n >= 8


def test_oracle_matches_live_opencv_at_video_sizes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for (sh, sw), (dw, dh) in [((720, 1280), (346, 260)), ((480, 640), (346, 260)), ((360, 640), (320, 180))]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
        assert np.array_equal(prep_oracle.prep_frame(img, (dw, dh)), ref)


def test_enlarging_is_refused():
    with pytest.raises(NotImplementedError):
        prep_oracle.resize_area_u8(np.zeros((10, 10), np.uint8), (20, 20))


@pytest.mark.gpu
def test_cuda_prep_matches_opencv_fixtures():
    import torch
    from v2e_b200.prep import InputPrep
    for i, fr, wh, crop, want in _cases():
        cn = 3 if fr.ndim == 4 else 1
        p = InputPrep((fr.shape[2], fr.shape[1]), wh, channels=cn, crop=crop)
        got = p(torch.from_numpy(fr).cuda()).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), "case %d" % i
        p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [((720, 1280), (346, 260)), ((1080, 1920), (1280, 720)), ((720, 1280), (640, 360)),
                                   ((780, 1038), (346, 260)), ((720, 1280), (320, 180))])
def test_cuda_prep_matches_oracle_at_video_sizes(sizes):
    import torch
    from v2e_b200.prep import InputPrep
    (sh, sw), (dw, dh) = sizes
    rng = np.random.default_rng(3)
    fr = rng.integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
    p = InputPrep((sw, sh), (dw, dh), channels=3)
    got = p(torch.from_numpy(fr).cuda()).cpu().numpy()
    for k in range(2):
        assert np.array_equal(got[k], prep_oracle.prep_frame(fr[k], (dw, dh))), sizes
    p.close()


@pytest.mark.gpu
def test_cuda_prep_refuses_enlarging():
    from v2e_b200 import _lib
    from v2e_b200.prep import InputPrep
    with pytest.raises(_lib.V2eError):
        InputPrep((100, 100), (200, 200), channels=1)
