"""bench.py --workload c5: BASELINE config 5 -- ONE 1280x720 clip over the N ranks, x50 SloMo, centre-surround
(CSDVS) pixel model (scripts/csdvs.sh parameters on a C3-style block texture).

  N = 1 : SuperSloMo + EventEmulator(cs_lambda_pixels=10, cs_tau_p_ms=0.5) on one GPU (frame-by-frame kernels: the
          surround's Euler iteration is a chain of stencil launches, emulator.py:1102-1124).
  N > 1 : V2EPipeline.run_clip_sharded -- SloMo over this rank's frame pairs (no halo: every (pair, t) is independent),
          ONE all-to-all of uint8 row bands (+ K halo rows), pixel model on the rank's rows: per chunk of K Euler
          steps an all-gather of 2 x K edge rows and an all-reduce(MAX) of K maxima, per frame an all-reduce(MAX) of
          the event maximum. The literal "spatial tiles + halo for the SloMo receptive field" of BASELINE.json is
          replaced by pair sharding (the UNet's receptive field spans several hundred pixels).
Rank 0 then replays the gathered interpolated frames through the single-GPU pixel model and compares event count and
an order-independent checksum of the rows with the sharded run ("verified_vs_single_gpu").
"""
import ctypes
import os

import numpy as np


def rows_checksum(rows):
    """Order-independent 63-bit checksum of packed float32 rows [t, x, y, p] (sum of per-row mixes of the bit patterns)."""
    import torch
    if rows.shape[0] == 0:
        return 0
    b = rows.contiguous().view(torch.int32).to(torch.int64)
    mix = (b[:, 0] * 1000003 + b[:, 1] * 10007 + b[:, 2] * 101 + b[:, 3]) & 0x7FFFFFFFFFFF
    return int(mix.sum().item() & 0x7FFFFFFFFFFFFFFF)


def run_config5(args, rank, world, local_rank, pk):
    import torch
    import torch.distributed as dist
    import bench
    from v2e_b200 import EventEmulator, SuperSloMo, V2EPipeline
    dev = torch.device("cuda", local_rank)
    devname = "cuda:%d" % local_rank
    H, W = args.height, args.width
    NS, U = int(os.environ.get("V2E_C5_SRC_FRAMES", "17")), int(os.environ.get("V2E_C5_U", "50"))
    n_pairs, n_frames = NS - 1, (NS - 1) * U
    clip_s = n_pairs / bench.SRC_FPS
    wts = bench.slomo_weights()
    src = torch.from_numpy(bench.block_texture_clip(H, W, NS, seed=0)).to(dev)
    params = dict(bench.C5_PARAMS)
    steps = max(1, min(args.steps, 2))
    warm = min(args.warmup, 1)
    batch = min(args.batch, max(1, n_pairs // world))

    def make(shard):
        sl = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=batch, device=devname,
                        state_dicts=wts)
        em = EventEmulator(device=devname, rng_mode="device", seed=4242, shard=shard, max_frames_per_step=64, **params)
        em.event_rows_hint = 8 * 1024 * 1024
        return V2EPipeline(sl, em)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total, chk, k = 0, 0, 0
    pipe = make((rank, world, None) if world > 1 else None)

    def one(timed):
        nonlocal total, chk, k
        t0 = k * clip_s * n_frames / (n_frames - 1)      # the next pass starts one frame interval after the last frame
        k += 1
        if world > 1:
            rows, t, nf = pipe.run_clip_sharded(src, clip_s, t_offset=t0)
            rows = torch.from_numpy(rows) if isinstance(rows, np.ndarray) else rows
        else:
            rows, offs, t, nf = pipe.run(src, clip_s, t_offset=t0, return_device=True)
        if timed:
            total += rows.shape[0]
        return rows

    for _ in range(warm):
        one(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(steps):
        last = one(True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    cnt = torch.tensor([float(total)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, total_all = tms.item(), cnt.item()
    cs_steps = list(pipe.emulator.cs_steps_taken)

    # ---- verification against the single-GPU pixel model on the same interpolated frames (first clip pass) ----
    verified = None
    if world > 1:
        pipe2 = make((rank, world, None))
        rows0, t0v, _ = pipe2.run_clip_sharded(src, clip_s, t_offset=0.0)
        r0 = torch.from_numpy(rows0).to(dev)
        my = torch.tensor([r0.shape[0], rows_checksum(r0)], device=dev, dtype=torch.int64)
        allv = [torch.zeros_like(my) for _ in range(world)]
        dist.all_gather(allv, my)
        n_sh = int(sum(int(v[0]) for v in allv))
        c_sh = int(sum(int(v[1]) for v in allv) & 0x7FFFFFFFFFFFFFFF)
        steps_sh = list(pipe2.emulator.cs_steps_taken)
        # all interpolated frames on rank 0: every rank synthesised a contiguous run of the clip
        from v2e_b200 import parallel
        p0, p1 = parallel.pair_range(n_pairs, rank, world)
        local, _, _ = pipe2.slomo.interpolate_frames(src[p0:p1 + 1])
        pieces = [torch.empty(((parallel.pair_range(n_pairs, r, world)[1] - parallel.pair_range(n_pairs, r, world)[0]) * U, H, W),
                              dtype=torch.uint8, device=dev) for r in range(world)] if rank == 0 else None
        if rank == 0:
            pieces[0].copy_(local)
            for r in range(1, world):
                dist.recv(pieces[r], src=r)
        else:
            dist.send(local.contiguous(), dst=0)
        if rank == 0:
            frames = torch.cat(pieces, 0)
            em1 = EventEmulator(device=devname, rng_mode="device", seed=4242, max_frames_per_step=64, **params)
            em1.event_rows_hint = 8 * 1024 * 1024
            rows1, _ = em1.generate_events_batch(frames, t0v, return_device=True)
            verified = {"events_sharded": n_sh, "events_single_gpu": int(rows1.shape[0]),
                        "checksum_equal": bool(rows_checksum(rows1) == c_sh),
                        "cs_steps_equal": bool(list(em1.cs_steps_taken) == steps_sh),
                        "equal": bool(rows1.shape[0] == n_sh and rows_checksum(rows1) == c_sh and
                                      list(em1.cs_steps_taken) == steps_sh)}
            em1.cleanup()
        pipe2.slomo.cleanup()
        pipe2.emulator.cleanup()
    pipe.slomo.cleanup()
    pipe.emulator.cleanup()
    if rank != 0:
        return None
    Wd, Hd = int(W / 32) * 32, int(H / 32) * 32
    fl = 2.0 * Hd * Wd * (330016 + 314048 / U)
    return {
        "metric": "Mevents/s", "value": total_all / (ms * 1e-3) / 1e6, "unit": "Mevents/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "fp16 tensor-core convs (fp32 accumulate) + f64 pixel state", "data": "synthetic",
        "config": {"workload": "config5_%dx%d_block_texture_%dsrc_frames_slomo_x%d_csdvs_one_clip_over_%d_gpus" % (
                       W, H, NS, U, world),
                   "params": params, "frames": n_frames,
                   "euler_steps_per_frame": {"min": int(min(cs_steps)) if cs_steps else None,
                                             "max": int(max(cs_steps)) if cs_steps else None,
                                             "mean": float(np.mean(cs_steps)) if cs_steps else None},
                   "halo_rows_per_exchange": pipe.emulator.cs_halo_rows(H) if world > 1 else 0,
                   "sharding": "SloMo over frame pairs, all-to-all of uint8 row bands (+halo), pixel model over pixel rows"},
        "interp_frames_per_s": steps * n_frames / (ms * 1e-3),
        "slomo_flops_per_interp_frame": fl,
        "events_per_px_per_frame": total_all / steps / (n_frames * H * W),
        "verified_vs_single_gpu": verified,
    }
