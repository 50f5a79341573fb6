#!/usr/bin/env python
"""bench.py -- headline benchmark of the v2e hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one synthetic clip (frames resident in HBM when the timed
region starts for `value`; in pinned host memory, copied inside the timed region, for `e2e`).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLI_DEFAULTS = dict(pos_thres=0.2, neg_thres=0.2, sigma_thres=0.03, cutoff_hz=300.0, leak_rate_hz=0.01,
                    shot_noise_rate_hz=0.001, refractory_period_s=0.0005)   # v2e_args.py:150-204


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def texture_clip(H, W, T, seed=0, dx=1, dy=0, up=16, lo=40.0, hi=215.0):
    """Smooth random texture (uniform noise, bicubic x16) translating 1 px per frame, uint8: with the
    CLI-default pixel parameters this gives ~0.1 events/pixel/frame (SURVEY 8d asks for >= 0.05)."""
    import torch
    rng = np.random.default_rng(seed)
    ph, pw = H + dy * T + 2 * up, W + dx * T + 2 * up
    base = torch.from_numpy(rng.uniform(lo, hi, (1, 1, ph // up + 3, pw // up + 3)).astype(np.float32))
    big = torch.nn.functional.interpolate(base, scale_factor=up, mode="bicubic", align_corners=False)[0, 0]
    big = big.clamp(0, 255).round().to(torch.uint8).numpy()
    out = np.empty((T, H, W), np.uint8)
    for k in range(T):
        j = k if k < T // 2 else T - 1 - k      # forward then backward: the clip loops without a jump
        out[k] = big[j * dy:j * dy + H, j * dx:j * dx + W]
    return out


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md "clocks line")."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_oracle(frames, times, kw, max_seconds=20.0):
    """The CPU oracle (scalar C restatement, 1 thread) timed on a bounded prefix of the same clip."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from emu_oracle import OracleEmulator
    em = OracleEmulator(seed=1, **kw)
    t0 = time.perf_counter()
    n = 0
    em.generate_events(frames[0], float(times[0]))
    for i in range(1, len(frames)):
        em.generate_events(frames[i], float(times[i]))
        n += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return dict(events=em.num_events_total, frames=n, seconds=dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    H, W, T = args.height, args.width, args.frames
    fps_src, U = 30.0, 10
    dt = 1.0 / (fps_src * U)                       # 10x slow-motion timestamps
    times = np.arange(T) * dt
    kw = dict(CLI_DEFAULTS)
    workload = "emulator_%dx%d_smooth_texture_1px_per_frame_T%d_cli_defaults_dt%.4gms" % (W, H, T, dt * 1e3)
    pk = peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        frames = texture_clip(H, W, min(T, 64), seed=0)
        vals = []
        for _ in range(args.warmup + args.steps):
            r = cpu_baseline_oracle(frames, times, kw, max_seconds=8.0)
            vals.append(r)
        vals = vals[args.warmup:]
        ev = sum(v["events"] for v in vals)
        sec = sum(v["seconds"] for v in vals)
        fr = sum(v["frames"] for v in vals)
        v = ev / sec / 1e6
        line = {"impl": "reference", "metric": "Mevents/s", "value": v, "unit": "Mevents/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / len(vals) * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload},
                "cpu_baseline": {"value": v, "unit": "Mevents/s", "cores": 1, "kind": "port",
                                 "sample": "%d frames/step of the same clip, scalar C oracle" % (fr // len(vals)),
                                 "frames_per_s": fr / sec},
                "e2e": {"value": v, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from v2e_b200 import EventEmulator, _lib
    import ctypes
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    frames_host = torch.from_numpy(texture_clip(H, W, T, seed=rank)).pin_memory()
    frames_dev = frames_host.to(dev)

    def fresh():
        em = EventEmulator(device="cuda:%d" % local_rank, rng_mode="device", seed=1234 + rank,
                           max_frames_per_step=64, **kw)
        em.event_rows_hint = 40 * 1024 * 1024
        return em

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clip_dt = T * dt

    def timed(e2e, steps, warmup):
        """One emulator, the (looping) clip fed `warmup + steps` times with advancing timestamps."""
        em = fresh()
        k = 0

        def one():
            nonlocal k
            t = times + k * clip_dt
            k += 1
            if e2e:
                fr = frames_host.to(dev, non_blocking=True)
                rows, offs = em.generate_events_batch(fr, t, return_device=False)
            else:
                rows, offs = em.generate_events_batch(frames_dev, t, return_device=True)
            return rows.shape[0]
        for _ in range(warmup):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 0
        for _ in range(steps):
            n += one()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        em.cleanup()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        cnt = torch.tensor([float(n)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return t.item(), cnt.item()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ev_dev = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, ev_e2e = timed(True, args.steps, max(1, args.warmup))

    # roofline of the dominant kernel (update): CUDA events inside the library around every launch
    em = fresh()
    em.generate_events_batch(frames_dev[:2], times[:2])
    _lib.check(em._lib.v2e_emu_profile(em._h, 1))
    ms3 = (ctypes.c_float * 3)()
    n3 = (ctypes.c_int * 3)()
    tot_ms = np.zeros(3)
    tot_n = np.zeros(3)
    f = 2
    while f < T:
        e = min(T, f + 64)
        em._run_step(frames_dev[f:e], _lib.U8, times[f:e])
        _lib.check(em._lib.v2e_emu_profile_read(em._h, ms3, n3, em._stream()))
        tot_ms += np.array(list(ms3)); tot_n += np.array(list(n3))
        em.t_previous = float(times[e - 1])
        f = e
    prof_events = em.num_events_total
    em.cleanup()
    upd_ms = tot_ms[0] / max(tot_n[0], 1)
    # algorithmic bytes of the update kernel per launch: read frame 1 + lp 8 + base 8 + thresholds 8
    # + noise_rate 4; write lp 8 + base 8 + record 2  = 47 B/px (float64 state, CLI defaults)
    upd_bytes = H * W * 47.0
    achieved = upd_bytes / (upd_ms * 1e-3) / 1e9
    step_bytes = H * W * 53.0 + 16.0 * prof_events / max(tot_n[0], 1)   # SURVEY 8(d), T=1 form
    frame_ms = tot_ms.sum() / max(tot_n[0], 1)

    if rank == 0:
        frames_sample = texture_clip(H, W, 24, seed=0)
        cb = cpu_baseline_oracle(frames_sample, times, kw, max_seconds=15.0)
        cpu_val = cb["events"] / cb["seconds"] / 1e6
        steps = args.steps
        value = ev_dev / (ms_dev * 1e-3) / 1e6
        e2e = ev_e2e / (ms_e2e * 1e-3) / 1e6
        kernels_per_frame = 3
        line = {
            "metric": "Mevents/s", "value": value, "unit": "Mevents/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "frames_per_step": T, "clips": world,
                       "events_per_px_per_frame": ev_dev / steps / world / (T * H * W),
                       "l2_policy": "inputs (%.0f MB of frames per step) larger than L2" % (T * H * W / 1e6),
                       "rng": "device philox", "sharding": "one independent clip per GPU, no data-path collective"},
            "frames_per_s": world * T * steps / (ms_dev * 1e-3),
            "e2e": {"value": e2e, "unit": "Mevents/s", "h2d_bytes_per_step": T * H * W,
                    "d2h_bytes_per_step": int(16 * ev_e2e / steps / world), "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(steps * T * kernels_per_frame + steps * ((T + 63) // 64)),
            "roofline": {"kernel": "emu_update_kernel<double,u8>", "bound": "hbm", "achieved": achieved,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": None, "peak_source": pk["source"], "bytes_per_launch": upd_bytes,
                         "us_per_launch": upd_ms * 1e3,
                         "frame_all_kernels": {"us": frame_ms * 1e3, "bytes": step_bytes,
                                               "frac": step_bytes / (frame_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                         "kernel_us": {"update": tot_ms[0] / max(tot_n[0], 1) * 1e3,
                                       "filter": tot_ms[1] / max(tot_n[1], 1) * 1e3,
                                       "emit": tot_ms[2] / max(tot_n[2], 1) * 1e3}},
            "cpu_baseline": {"value": cpu_val, "unit": "Mevents/s", "cores": 1, "kind": "port",
                             "sample": "%d frames of the same clip, scalar C oracle" % cb["frames"],
                             "frames_per_s": cb["frames"] / cb["seconds"]},
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
