"""Multi-GPU tests over NCCL (one process per GPU). They need >= 2 CUDA devices and are skipped on a single-GPU
box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`. The same host logic is covered on
one GPU by the gloo tests in test_emulator_gpu.py / test_slomo_gpu.py (gloo stages CUDA tensors through the host, so
its ranks can share a device; NCCL cannot)."""
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d CUDA devices" % n)


def _port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _weights(seed):
    import slomo_ref
    return (slomo_ref.make_test_weights(seed, 2, 4, head_gain=25.0), slomo_ref.make_test_weights(seed + 1, 12, 5, head_gain=0.3))


def _clip_worker(rank, world, port, frames, kw, U, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from v2e_b200 import EventEmulator, SuperSloMo, V2EPipeline
        fc, at = _weights(5)
        dev = "cuda:%d" % rank
        sl = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=2, device=dev,
                        state_dicts={'state_dictFC': fc, 'state_dictAT': at})
        em = EventEmulator(device=dev, seed=9, rng_mode="device", shard=(rank, world, None), max_frames_per_step=5, **kw)
        rows, t, nf = V2EPipeline(sl, em).run_clip_sharded(frames, 0.2)
        q.put((rank, np.asarray(rows), nf, list(em.cs_steps_taken)))
        sl.cleanup()
        em.cleanup()
    finally:
        dist.destroy_process_group()


def _run(world, frames, kw, U):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _port()
    procs = [ctx.Process(target=_clip_worker, args=(r, world, port, frames, kw, U, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def _frames(n=6, H=96, W=128):
    rng = np.random.default_rng(4)
    big = np.kron(rng.integers(30, 220, (H // 8 + 2, W // 8 + 16)).astype(np.uint8), np.ones((8, 8), np.uint8))
    return np.stack([big[3:3 + H, 4 * k:4 * k + W] for k in range(n)])


def _single(frames, kw, U):
    from v2e_b200 import EventEmulator, SuperSloMo, V2EPipeline
    fc, at = _weights(5)
    sl = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=2,
                    state_dicts={'state_dictFC': fc, 'state_dictAT': at})
    em = EventEmulator(device="cuda:0", seed=9, rng_mode="device", max_frames_per_step=5, **kw)
    ev, offs, t, nf = V2EPipeline(sl, em).run(frames, 0.2)
    sl.cleanup()
    steps = list(em.cs_steps_taken)
    em.cleanup()
    return np.asarray(ev), nf, steps


def _key(e):
    return e[np.lexsort((e[:, 3], e[:, 1], e[:, 2], e[:, 0]))]


@pytest.mark.parametrize("kw", [
    # leak + shot noise from the device RNG (whole-frame Philox counters), refractory period that never engages
    dict(cutoff_hz=200, leak_rate_hz=0.1, shot_noise_rate_hz=2.0, refractory_period_s=0.0001, sigma_thres=0.02),
    # refractory filter active in some frames: chunks rejected, replayed frame by frame with one all-reduce each
    dict(cutoff_hz=200, leak_rate_hz=0, shot_noise_rate_hz=0, refractory_period_s=0.004, sigma_thres=0.02,
         pos_thres=0.05, neg_thres=0.05),
])
def test_nccl_one_clip_over_two_gpus_equals_single_gpu(kw):
    """BASELINE config 5 layout on real NCCL: SloMo sharded over frame pairs, ONE all_to_all of uint8 row bands
    (parallel.exchange_frame_bands' NCCL branch), pixel model over pixel rows with the frame maxima all-reduced per
    chunk. The union of the ranks' events must equal the single-GPU pipeline's events."""
    _need(2)
    frames = _frames()
    want, nf, _ = _single(frames, kw, 3)
    res = _run(2, frames, kw, 3)
    assert all(r[2] == nf for r in res)
    got = np.concatenate([r[1] for r in res], 0)
    assert got.shape == want.shape and want.shape[0] > 0
    assert np.array_equal(_key(got), _key(want))


def test_nccl_centre_surround_over_two_gpus_equals_single_gpu():
    """The centre-surround model over 2 GPUs (halo rows all-gathered per Euler chunk, maxima all-reduced): events and
    Euler steps per frame equal the single-GPU run."""
    _need(2)
    kw = dict(cs_lambda_pixels=4, cs_tau_p_ms=2.0, cutoff_hz=200, leak_rate_hz=0, shot_noise_rate_hz=0, sigma_thres=0.02,
              refractory_period_s=0.001)
    frames = _frames()
    want, nf, steps = _single(frames, kw, 3)
    res = _run(2, frames, kw, 3)
    got = np.concatenate([r[1] for r in res], 0)
    assert got.shape == want.shape and want.shape[0] > 0
    assert np.array_equal(_key(got), _key(want))
    for r in res:
        assert r[3] == steps
