"""CPU: pins the oracle (oracle/emu_oracle.c + emu_oracle.py) against fixtures that
the unmodified reference produced (oracle/make_golden.py). Bit-exact rows, order and state."""
import hashlib

import numpy as np
import pytest
import torch

from emu_oracle import OracleEmulator, lib
from helpers import (EMU_GOLDENS, EMU_GOLDENS_OPT, TapeRNG, assert_events_equal, canonical, load_golden,
                     split_events)


@pytest.mark.parametrize("name", EMU_GOLDENS + EMU_GOLDENS_OPT)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    rng = TapeRNG(g["tape"])
    # photoreceptor noise: the amplitude comes from an unseeded numpy generator in the reference
    # (emulator_utils.py:234-235); the fixture carries the values it used
    extra = {"pr_vrms_tape": list(g["pr_vrms"])} if "pr_vrms" in g else {}
    em = OracleEmulator(rng=rng, **extra, **g["kwargs"])
    want = split_events(g["events"], g["event_counts"])
    for i, (f, t) in enumerate(zip(g["frames"], g["times"])):
        ev = em.generate_events(f, float(t))
        assert_events_equal(ev, want[i], exact_order=True, ctx="%s frame %d" % (name, i))
    assert rng.exhausted()
    assert em.num_events_on == int(g["num_on"]) and em.num_events_off == int(g["num_off"])
    if "cs_steps_taken" in g:
        assert list(g["cs_steps_taken"]) == em.cs_steps_taken
    for key, arr in (("state_base_log_frame", em.base), ("state_lp_log_frame", em.lp),
                     ("state_timestamp_mem", em.tmem), ("state_cs_surround_frame", em.surround)):
        if key in g and arr is not None:
            assert g[key].dtype == arr.dtype, key
            assert np.array_equal(g[key], arr), key
    if em.noise_arr is not None:      # (the reference keeps a zero tensor when the option is off)
        assert np.array_equal(g["state_photoreceptor_noise_arr"], em.noise_arr)
    if em.hp is not None:
        # sinh: torch's CPU kernel (Sleef) and libm may differ in the last place; everything else is exact
        assert g["state_scidvs_highpass"].dtype == em.hp.dtype
        tol = 4e-15 if em.hp.dtype == np.float64 else 2e-6
        assert np.max(np.abs(g["state_scidvs_highpass"] - em.hp)) <= tol
        assert np.array_equal(g["state_scidvs_tau_arr"], em.tau_arr)


def test_oracle_moving_dot_config1_seeded():
    """BASELINE config 1 (scripts/moving_dot.py 64x64, class defaults, seed 42): 27 917 events.
    Uses torch's own CPU generator, so it is only meaningful where torch draws the same
    numbers as in the build container (same torch build and CPU dispatch level)."""
    g = load_golden("emu_moving_dot_c1")
    if str(g["cpu_capability"]) != torch.backends.cpu.get_cpu_capability() or \
            str(g["torch_version"]) != torch.__version__:
        pytest.skip("torch CPU RNG kernels differ from the fixture's host")
    em = OracleEmulator(seed=int(g["seed"]), **g["kwargs"])
    h = hashlib.sha1()
    counts = []
    for f, t in zip(g["frames"], g["times"]):
        ev = em.generate_events(f, float(t))
        counts.append(0 if ev is None else len(ev))
        if ev is not None:
            h.update(canonical(ev).tobytes())
    assert np.array_equal(np.array(counts), g["event_counts"])
    assert em.num_events_total == 27917 and em.num_events_on == 14124 and em.num_events_off == 13793
    assert h.hexdigest() == str(g["events_sha1_canonical"])


def test_linspace_restatement_matches_torch():
    """oracle_linspace_f32 vs torch.linspace(float32) (emulator.py:793-796)."""
    L = lib()
    rng = np.random.default_rng(0)
    for _ in range(300):
        tp = rng.uniform(0, 50)
        dt = rng.uniform(1e-5, 1e-2)
        n = int(rng.integers(1, 70))
        start, end = tp + dt / n, tp + dt
        ref = torch.linspace(start=start, end=end, steps=n, dtype=torch.float32).numpy()
        got = np.array([L.oracle_linspace_f32(start, end, n, i) for i in range(n)], np.float32)
        assert np.array_equal(ref, got)


def test_time_going_backwards_raises():
    em = OracleEmulator()
    em.generate_events(np.zeros((4, 4), np.uint8), 0.0)
    em.generate_events(np.zeros((4, 4), np.uint8), 1.0)
    with pytest.raises(ValueError):
        em.generate_events(np.zeros((4, 4), np.uint8), 0.5)
