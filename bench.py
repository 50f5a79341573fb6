#!/usr/bin/env python
"""bench.py -- headline benchmark of the v2e hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json: "Mevents/s + interpolated-frames/s ... 1280x720 at 10x slowdown"):
one clip of 9 source frames (1280x720 uint8, smooth random texture translating 10 px per source
frame) -> SuperSloMo x10 (batch 8) -> 80 interpolated frames -> DVS pixel model with v2e's CLI-default
parameters -> events. A "step" is one pass of that whole path over one clip.
  value : events/s with the source frames already resident in HBM, events left in HBM
  e2e   : same, source frames in pinned host memory copied in and the packed event rows copied out
          inside the timed region, through the package's public API (V2EPipeline.run)
With N>1 every rank processes its own clip (weak scaling, no data-path collective) and the event
streams are gathered with NCCL at the end of each step.
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CLI_DEFAULTS = dict(pos_thres=0.2, neg_thres=0.2, sigma_thres=0.03, cutoff_hz=300.0, leak_rate_hz=0.01,
                    shot_noise_rate_hz=0.001, refractory_period_s=0.0005)   # v2e_args.py:150-204
SRC_FPS = 30.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def source_clip(H, W, n_src, seed=0, px_per_frame=10, up=16, lo=40.0, hi=215.0):
    """n_src source frames: smooth random texture (uniform noise, bicubic x16) translating
    `px_per_frame` px per source frame, forward then backward so that the clip loops seamlessly."""
    import torch
    rng = np.random.default_rng(seed)
    half = n_src // 2
    pw = W + px_per_frame * half + 2 * up
    base = torch.from_numpy(rng.uniform(lo, hi, (1, 1, H // up + 5, pw // up + 5)).astype(np.float32))
    big = torch.nn.functional.interpolate(base, scale_factor=up, mode="bicubic", align_corners=False)[0, 0]
    big = big.clamp(0, 255).round().to(torch.uint8).numpy()
    out = np.empty((n_src, H, W), np.uint8)
    for k in range(n_src):
        j = k if k <= half else n_src - 1 - k
        out[k] = big[up:up + H, j * px_per_frame:j * px_per_frame + W]
    return out


def slomo_weights():
    """Seeded variance-preserving weights in the reference's checkpoint layout ('state_dictFC' /
    'state_dictAT'); the real SuperSloMo39.ckpt is not available offline (README.md:95-96)."""
    import slomo_ref
    return {"state_dictFC": slomo_ref.make_test_weights(1234, 2, 4, head_gain=25.0),
            "state_dictAT": slomo_ref.make_test_weights(4321, 12, 5, head_gain=0.3)}


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
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_sample(H, W, U_sample=2, seed=0):
    """CPU stand-in for the reference's own path (torch fp32 SloMo restatement + scalar C pixel model,
    oracle/): one frame pair, U_sample interpolated frames, at the benchmark's resolution."""
    import torch
    import slomo_ref
    from emu_oracle import OracleEmulator
    # torchrun pins OMP_NUM_THREADS=1; the CPU leg uses the physical cores (hyper-threads slow ATen's convs down)
    torch.set_num_threads(max(1, min(64, (os.cpu_count() or 2) // 2)))
    wts = slomo_weights()
    frames = source_clip(H, W, 9, seed=seed)[:2]
    t0 = time.perf_counter()
    out, times, _ = slomo_ref.interpolate_frames(frames, wts["state_dictFC"], wts["state_dictAT"], U_sample,
                                                 batch_size=1)
    t_slomo = time.perf_counter() - t0
    em = OracleEmulator(seed=1, **CLI_DEFAULTS)
    dt = 1.0 / (SRC_FPS * 10)
    t1 = time.perf_counter()
    em.generate_events(frames[0], 0.0)           # state init
    for i in range(out.shape[0]):
        em.generate_events(out[i], (i + 1) * dt)
    t_emu = time.perf_counter() - t1
    return dict(events=em.num_events_total, interp_frames=int(out.shape[0]), seconds=t_slomo + t_emu,
                slomo_s=t_slomo, emu_s=t_emu, threads=torch.get_num_threads())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--src-frames", type=int, default=9)
    ap.add_argument("--upsampling", type=int, default=10)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary-only", action="store_true", help="development: print only the 346x260 line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    H, W, NS, U = args.height, args.width, args.src_frames, args.upsampling
    n_interp = (NS - 1) * U
    clip_s = (NS - 1) / SRC_FPS
    workload = "%dx%d_smooth_texture_%dsrc_frames_slomo_x%d_b%d_emulator_cli_defaults" % (W, H, NS, U, args.batch)
    pk = peaks()
    Wd, Hd = int(W / 32) * 32, int(H / 32) * 32
    flops_per_interp = 2.0 * Hd * Wd * (330016 + 314048 / U)

    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for _ in range((1 if args.warmup > 0 else 0) + args.steps):
            vals.append(cpu_port_sample(H, W))
        vals = vals[1:] if len(vals) > args.steps else vals
        ev = sum(v["events"] for v in vals)
        sec = sum(v["seconds"] for v in vals)
        fr = sum(v["interp_frames"] for v in vals)
        v = ev / sec / 1e6
        line = {"impl": "reference", "metric": "Mevents/s", "value": v, "unit": "Mevents/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / len(vals) * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 convs / f64 pixel state", "data": "synthetic", "config": {"workload": workload},
                "interp_frames_per_s": fr / sec,
                "cpu_baseline": {"value": v, "unit": "Mevents/s", "cores": vals[0]["threads"], "kind": "port",
                                 "sample": "1 frame pair -> 2 interpolated frames + pixel model per step "
                                           "(torch fp32 SloMo restatement on all threads + scalar C pixel model)",
                                 "interp_frames_per_s": fr / sec},
                "e2e": {"value": v, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from v2e_b200 import EventEmulator, SuperSloMo, V2EPipeline, _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at the VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    src_host = torch.from_numpy(source_clip(H, W, NS, seed=rank)).pin_memory()
    src_dev = src_host.to(dev)
    wts = slomo_weights()

    def make_pipe():
        sl = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=args.batch,
                        device="cuda:%d" % local_rank, state_dicts=wts)
        em = EventEmulator(device="cuda:%d" % local_rank, rng_mode="device", seed=1234 + rank,
                           max_frames_per_step=n_interp, **CLI_DEFAULTS)
        em.event_rows_hint = 48 * 1024 * 1024
        return V2EPipeline(sl, em)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_events(rows):
        """NCCL gather of the packed event streams (the only collective of the job)."""
        n = torch.tensor([rows.shape[0]], device=dev, dtype=torch.int64)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = int(max(x.item() for x in ns))
        pad = torch.zeros((mx, 4), dtype=torch.float32, device=dev)
        pad[:rows.shape[0]] = rows
        out = torch.empty((world * mx, 4), dtype=torch.float32, device=dev) if rank == 0 else None
        dist.gather(pad, list(out.split(mx)) if rank == 0 else None, dst=0)

    def timed(e2e, steps, warmup):
        pipe = make_pipe()
        k = 0

        def one():
            nonlocal k
            t0 = k * clip_s
            k += 1
            if e2e:
                fr = src_host.to(dev, non_blocking=True)
                ev, offs, t, nf = pipe.run(fr, clip_s, t_offset=t0, return_device=(world > 1))
                if world > 1:
                    gather_events(ev)
                    ev = ev.cpu()
            else:
                ev, offs, t, nf = pipe.run(src_dev, clip_s, t_offset=t0, return_device=True)
                if world > 1:
                    gather_events(ev)
            return ev.shape[0]
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
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        cnt = torch.tensor([float(n)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return t.item(), cnt.item(), pipe

    if args.secondary_only:
        args.steps, args.warmup, args.no_profile, args.no_e2e = 1, 1, True, True
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ev_dev, pipe = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (conv_tc_kernel, tensor pipe) and of the pixel-model update
    prof = {}
    if rank == 0 and not args.no_profile:
        eng = pipe.slomo._engine
        em = pipe.emulator
        _lib.check(eng.lib.v2e_slomo_profile(eng._h, 1))
        _lib.check(em._lib.v2e_emu_profile(em._h, 1))
        k0 = args.steps + args.warmup
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        pipe.run(src_dev, clip_s, t_offset=k0 * clip_s, return_device=True)
        torch.cuda.synchronize()
        step_ms_prof = (time.perf_counter() - w0) * 1e3
        conv_ms, conv_n, conv_fl = ctypes.c_float(0), ctypes.c_int(0), ctypes.c_double(0)
        ms23, n23, fl23 = (ctypes.c_float * 23)(), (ctypes.c_int * 23)(), (ctypes.c_double * 23)()
        _lib.check(eng.lib.v2e_slomo_profile_read_layers(eng._h, ms23, n23, fl23, ctypes.byref(conv_ms),
                                                         ctypes.byref(conv_n), ctypes.byref(conv_fl), eng._stream()))
        names = ["conv1", "conv2"] + ["down%d.conv%d" % (d, c) for d in range(1, 6) for c in (1, 2)] + \
                ["up%d.conv%d" % (d, c) for d in range(1, 6) for c in (1, 2)] + ["conv3"]
        layers = []
        for i in range(23):
            if n23[i]:
                tf = fl23[i] / (ms23[i] * 1e-3) / 1e12
                layers.append({"layer": names[i], "launches": n23[i], "ms": ms23[i], "tflops": tf,
                               "frac": tf / pk["bf16_tflops_sustained"]})
        big = max(range(23), key=lambda i: ms23[i])
        ms3, n3 = (ctypes.c_float * 4)(), (ctypes.c_int * 4)()
        _lib.check(em._lib.v2e_emu_profile_read4(em._h, ms3, n3, em._stream()))
        _lib.check(eng.lib.v2e_slomo_profile(eng._h, 0))
        _lib.check(em._lib.v2e_emu_profile(em._h, 0))
        # the update kernel alone: 50 back-to-back launches on the clip's last source frame and the state the
        # step left, between ONE event pair (stores go to scratch arrays, so every launch does the real work)
        us_b2b = ctypes.c_float(0)
        dt_i = clip_s / n_interp
        t_end = float(em.t_previous)
        b2b_ok = True
        try:
            _lib.check(em._lib.v2e_emu_time_update(em._h, ctypes.c_void_p(src_dev[NS - 1].data_ptr()), 0, t_end + dt_i,
                                                   t_end, 50, ctypes.byref(us_b2b), em._stream()))
        except Exception as exc:          # keep the bench line: fall back to the per-kernel bracket
            sys.stderr.write("v2e_emu_time_update failed: %s\n" % exc)
            b2b_ok = False
        achieved = conv_fl.value / (conv_ms.value * 1e-3) / 1e12
        upd_us_bracket = ms3[0] / max(n3[0], 1) * 1e3
        upd_us = us_b2b.value if (b2b_ok and us_b2b.value > 0) else upd_us_bracket
        upd_bytes = H * W * 47.0
        prof = {
            "roofline": {"kernel": "conv_tc_kernel (all UNet convolutions of one step, summed)", "bound": "tensor",
                         "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_tflops_sustained"], "traffic": None,
                         "peak_source": pk["source"] + " (sustained 16-bit dense; burst %.1f)" % pk["bf16_tflops"],
                         "flops_per_step": conv_fl.value, "conv_ms_per_step": conv_ms.value,
                         "launches_per_step": conv_n.value, "share_of_step": conv_ms.value / step_ms_prof,
                         # the single largest layer, with the DRAM traffic ncu measured for one of its launches
                         # (profiles/r1_conv_strip2_ncu.md; algorithmic bytes = input + output activations)
                         "largest_layer": {"layer": names[big], "kernel": "conv_strip2_kernel<7,32>" if big == 1 else None,
                                           "ms_per_launch": ms23[big] / n23[big],
                                           "tflops": fl23[big] / (ms23[big] * 1e-3) / 1e12,
                                           "frac": fl23[big] / (ms23[big] * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                                           "traffic": 886.9e6 if big == 1 else None,
                                           "algorithmic_bytes": 2.0 * args.batch * Hd * Wd * 32 * 2 if big == 1 else None},
                         "layers": layers},
            "roofline_emulator": {"kernel": "emu_update_kernel<double,u8,philox>", "bound": "hbm",
                                  "achieved": upd_bytes / (upd_us * 1e-6) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": upd_bytes / (upd_us * 1e-6) / 1e9 / pk["hbm_gbs"], "traffic": None,
                                  "bytes_per_launch": upd_bytes, "us_per_launch": upd_us,
                                  "timing": ("50 back-to-back launches between one CUDA-event pair on the launching "
                                             "stream (v2e_emu_time_update: the clip's last source frame on the state the "
                                             "step left, stores out of place so that every launch does the real work)")
                                            if (b2b_ok and us_b2b.value > 0) else "per-kernel CUDA-event bracket inside the step",
                                  # per-kernel CUDA-event brackets inside the step (each bracket costs the floor below)
                                  "kernel_us": {"update": upd_us_bracket, "filter": ms3[1] / max(n3[1], 1) * 1e3,
                                                "emit": ms3[2] / max(n3[2], 1) * 1e3},
                                  # what the same CUDA-event bracket reports around an EMPTY kernel on this
                                  # stream: the floor of the method, included in every figure above
                                  "event_bracket_floor_us": ms3[3] / max(n3[3], 1) * 1e3,
                                  "traffic_note": "ncu (cold caches): 26.8 MB DRAM read, <1 MB written per launch -- "
                                                  "the 43 MB of state and frame are L2-resident between frames "
                                                  "(profiles/r1_emu_update_ncu.md)"},
        }
    pipe.slomo.cleanup()
    pipe.emulator.cleanup()
    del pipe
    torch.cuda.empty_cache()

    if args.no_e2e:
        ms_e2e, ev_e2e = ms_dev, ev_dev
    else:
        ms_e2e, ev_e2e, pipe2 = timed(True, args.steps, max(1, args.warmup))
        pipe2.slomo.cleanup()
        pipe2.emulator.cleanup()

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        # BASELINE configs[1] size (346x260, x10): same path, quoted beside the headline (SloMo runs at 320x256)
        H2, W2, NS2 = 260, 346, 31
        src2 = torch.from_numpy(source_clip(H2, W2, NS2, seed=7, px_per_frame=5, up=8)).to(dev)
        # small frames: all 30 pairs in one batch (batch_size is SuperSloMo's own knob, slomo.py:44-54), otherwise
        # the deep UNet levels (10x8 pixels) leave most SMs idle
        batch2 = NS2 - 1
        sl2 = SuperSloMo(model=None, auto_upsample=False, upsampling_factor=U, batch_size=batch2,
                         device="cuda:%d" % local_rank, state_dicts=wts)
        em2 = EventEmulator(device="cuda:%d" % local_rank, rng_mode="device", seed=99,
                            max_frames_per_step=(NS2 - 1) * U, **CLI_DEFAULTS)
        em2.event_rows_hint = 16 * 1024 * 1024
        p2 = V2EPipeline(sl2, em2)
        clip2 = (NS2 - 1) / SRC_FPS
        n2, reps2 = 0, 3
        for k in range(2):
            p2.run(src2, clip2, t_offset=k * clip2, return_device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(2, 2 + reps2):
            ev2, _, _, nf2 = p2.run(src2, clip2, t_offset=k * clip2, return_device=True)
            n2 += ev2.shape[0]
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1)
        secondary = {"workload": "%dx%d_smooth_texture_%dsrc_frames_slomo_x%d_b%d_emulator_cli_defaults" % (
                         W2, H2, NS2, U, batch2),
                     "value": n2 / (ms2 * 1e-3) / 1e6, "unit": "Mevents/s", "steps": reps2,
                     "ms_per_step": ms2 / reps2, "interp_frames_per_s": reps2 * (NS2 - 1) * U / (ms2 * 1e-3),
                     "events_per_px_per_frame": n2 / reps2 / ((NS2 - 1) * U * H2 * W2)}
        sl2.cleanup()
        em2.cleanup()
        del p2
        torch.cuda.empty_cache()

    if args.secondary_only:
        if rank == 0:
            print(json.dumps(secondary))
        return
    if rank == 0:
        cb = cpu_port_sample(H, W)
        cpu_val = cb["events"] / cb["seconds"] / 1e6
        steps = args.steps
        value = ev_dev / (ms_dev * 1e-3) / 1e6
        e2e = ev_e2e / (ms_e2e * 1e-3) / 1e6
        n_batches = -(-(NS - 1) // args.batch)
        launches_step = n_batches * (1 + 33 + U * (1 + 33 + 1 + 2) + 2) + 3 * n_interp + 2
        line = {
            "metric": "Mevents/s", "value": value, "unit": "Mevents/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 tensor-core convs (fp32 accumulate) + f64 pixel state",
            "data": "synthetic",
            "config": {"workload": workload, "interp_frames_per_step": n_interp, "clips": world,
                       "events_per_px_per_frame": ev_dev / steps / world / (n_interp * H * W),
                       "l2_policy": "activations of one UNet pass (>2 GB at batch 8) exceed L2",
                       "rng": "device philox", "weights": "seeded random, reference checkpoint layout",
                       "sharding": "one independent clip per GPU; NCCL gather of the event streams per step"},
            "interp_frames_per_s": world * n_interp * steps / (ms_dev * 1e-3),
            "slomo_flops_per_interp_frame": flops_per_interp,
            "e2e": {"value": e2e, "unit": "Mevents/s", "h2d_bytes_per_step": NS * H * W,
                    "d2h_bytes_per_step": int(16 * ev_e2e / steps / world), "ms_per_step": ms_e2e / steps,
                    "interp_frames_per_s": world * n_interp * steps / (ms_e2e * 1e-3)},
            "gpu_launches": int(steps * launches_step),
            "cpu_baseline": {"value": cpu_val, "unit": "Mevents/s", "cores": cb["threads"], "kind": "port",
                             "sample": "1 frame pair -> 2 interpolated frames + pixel model, same resolution "
                                       "(torch fp32 SloMo restatement %.1fs + scalar C pixel model %.1fs)" % (
                                           cb["slomo_s"], cb["emu_s"]),
                             "interp_frames_per_s": cb["interp_frames"] / cb["seconds"]},
            "clocks": clocks,
        }
        line.update(prof)
        if secondary is not None:
            line["secondary_346x260"] = secondary
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
