"""Shared test helpers: golden-fixture loading, the RNG tape, event comparators."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EMU_GOLDENS = ["emu_class_default", "emu_cli_noisy", "emu_clean", "emu_scalar_thres_f64",
               "emu_refractory_multi", "emu_float_frames", "emu_static_leak_shot",
               "emu_ragged_13x37", "emu_csdvs", "emu_csdvs_120x176"]
# optional pixel models: SCIDVS (emulator.py:58-80, 719-725), photoreceptor noise (emulator.py:694-703)
EMU_GOLDENS_OPT = ["emu_scidvs", "emu_scidvs_f32", "emu_prnoise", "emu_prnoise_scidvs_csdvs"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files if not k.startswith("tape_0")}
    g["kwargs"] = json.loads(str(z["kwargs_json"]))
    if "tape_kinds" in z.files:
        g["tape"] = [(str(k), z["tape_%05d" % i]) for i, k in enumerate(z["tape_kinds"])]
    return g


class TapeRNG:
    """Replays the random draws the reference made when the golden was recorded
    (oracle/make_golden.py::Recorder), checking that the consumer asks for the same
    kind and size of draw in the same order."""

    def __init__(self, tape):
        self.tape = list(tape)
        self.pos = 0

    def _next(self, kind, shape=None):
        assert self.pos < len(self.tape), "RNG tape exhausted (asked for %s)" % kind
        k, arr = self.tape[self.pos]
        self.pos += 1
        assert k == kind, "draw %d: reference drew %s, consumer asked %s" % (self.pos - 1, k, kind)
        if shape is not None:
            assert tuple(arr.shape) == tuple(shape), (kind, arr.shape, shape)
        return torch.from_numpy(np.array(arr))

    def normal(self, mean, std, shape):
        return self._next("normal", shape)

    def randn(self, shape):
        return self._next("randn", shape)

    def rand(self, shape):
        return self._next("rand", shape)

    def randperm(self, n):
        t = self._next("randperm", (n,))
        return t.long()

    def exhausted(self):
        return self.pos == len(self.tape)


def split_events(events, counts):
    off = np.concatenate([[0], np.cumsum(counts)])
    return [events[off[i]:off[i + 1]] for i in range(len(counts))]


def canonical(ev):
    """Sort rows by (t, y, x, p) -- the order-insensitive form (SURVEY 8d parity criteria)."""
    if ev is None or len(ev) == 0:
        return np.zeros((0, 4), np.float32)
    k = np.lexsort((ev[:, 3], ev[:, 1], ev[:, 2], ev[:, 0]))
    return np.ascontiguousarray(ev[k])


def assert_events_equal(got, want, exact_order=True, t_tol=0.0, ctx=""):
    got = np.zeros((0, 4), np.float32) if got is None else got
    want = np.zeros((0, 4), np.float32) if want is None else want
    assert got.shape == want.shape, "%s: %s rows vs reference %s" % (ctx, got.shape, want.shape)
    if not exact_order:
        got, want = canonical(got), canonical(want)
    assert np.array_equal(got[:, 1:], want[:, 1:]), "%s: x/y/polarity differ" % ctx
    if t_tol == 0.0:
        assert np.array_equal(got[:, 0], want[:, 0]), "%s: timestamps differ" % ctx
    else:
        assert np.max(np.abs(got[:, 0] - want[:, 0]), initial=0.0) <= t_tol, "%s: timestamps" % ctx
