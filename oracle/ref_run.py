"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference (v2ecore, loaded by ref_shim from
/root/reference or the verbatim copy in oracle/_ref/) through its own public API on a bounded sample:

    stage 2  SuperSloMo.interpolate(source .npy folder, output folder, (W, H))   v2ecore/slomo.py:231
    stage 3  read_image(png) -> EventEmulator.generate_events(frame, t)           v2e.py:826-834

exactly the way v2e.py strings them together (v2e.py:741-846: .npy source frames in a temp folder, PNG
interpolated frames in another, interpTimes scaled to seconds, v2e.py:794-797). Used by bench.py's
`--impl reference` / cpu_baseline legs and by the end-to-end parity tests; never by v2e_b200/.
"""
import os
import shutil
import tempfile
import time

import numpy as np


def write_checkpoint(path, state_dicts):
    """The reference loads `torch.load(ckpt)['state_dictFC' / 'state_dictAT']` (slomo.py:225-227)."""
    import torch
    torch.save({"state_dictFC": state_dicts["state_dictFC"], "state_dictAT": state_dicts["state_dictAT"]}, path)


def run_reference(frames_u8, src_fps, U, batch_size, emu_kwargs, state_dicts, seed=1, device="cpu",
                  t_offset=0.0, keep_frames=False, threads=None):
    """frames_u8: [N, H, W] uint8 source frames at `src_fps`. Returns a dict with the events the reference
    produced (list of per-frame arrays if keep_frames), its interpolated frames (if keep_frames), and the
    wall-clock seconds of each stage. device: what EventEmulator is given ('cpu' for the CPU arm);
    SuperSloMo picks its own device from torch.cuda.is_available() (slomo.py:84-89)."""
    import logging
    import torch
    import ref_shim
    logging.disable(logging.WARNING)
    if threads:
        torch.set_num_threads(int(threads))
    emu_mod, _, _, slomo_mod = ref_shim.load_reference()
    from v2ecore.v2e_utils import all_images, read_image
    n, H, W = frames_u8.shape
    work = tempfile.mkdtemp(prefix="v2e_ref_")
    try:
        src_dir, out_dir = os.path.join(work, "src"), os.path.join(work, "interp")
        os.makedirs(src_dir)
        os.makedirs(out_dir)
        for i in range(n):                       # v2e.py:733-737 writes the source frames as <idx>.npy
            np.save(os.path.join(src_dir, "%06d.npy" % i), frames_u8[i])
        ckpt = os.path.join(work, "weights.ckpt")
        write_checkpoint(ckpt, state_dicts)
        t0 = time.perf_counter()
        sl = slomo_mod.SuperSloMo(model=ckpt, auto_upsample=False, upsampling_factor=int(U),
                                  batch_size=int(batch_size), video_path=None, preview=False)
        interp_times, avg_u = sl.interpolate(src_dir, out_dir, (W, H))
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        t_slomo = time.perf_counter() - t0
        files = all_images(out_dir)
        n_interp = len(files)
        # v2e.py:794-797: times in units of source-frame intervals -> seconds of the clip
        f = ((n - 1) / src_fps) / (np.max(interp_times) - np.min(interp_times))
        times = t_offset + f * interp_times
        t1 = time.perf_counter()
        em = emu_mod.EventEmulator(seed=seed, device=device, output_folder=None, **emu_kwargs)
        rows, frames = [], []
        n_ev = 0
        with torch.no_grad():
            for i in range(n_interp):
                fr = read_image(files[i])
                ev = em.generate_events(fr, float(times[i]))
                if ev is not None:
                    n_ev += ev.shape[0]
                if keep_frames:
                    rows.append(ev)
                    frames.append(fr.astype(np.uint8))
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        t_emu = time.perf_counter() - t1
        em.cleanup()
        sl.cleanup()
        return dict(events=int(n_ev), interp_frames=int(n_interp), seconds=t_slomo + t_emu, slomo_s=t_slomo,
                    emu_s=t_emu, threads=torch.get_num_threads(), times=times, rows=rows,
                    frames=np.stack(frames) if frames else None, avg_upsampling=avg_u,
                    slomo_device=sl.device, kind=ref_shim.reference_kind())
    finally:
        shutil.rmtree(work, ignore_errors=True)
        logging.disable(logging.NOTSET)
