"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/slomo_*.npz by running the UNMODIFIED
reference (v2ecore/slomo.py::SuperSloMo.interpolate through temp folders, v2ecore/model.py) in the
build container.

    python oracle/make_golden_slomo.py        # needs /root/reference

The reference's CPU branch omits the 0.428 mean normalisation (slomo.py:154-156) that its CUDA
branch applies; the fixture forces the CUDA-branch transforms on the CPU device,
everything else is the reference's own code path (DataLoader, PIL resizes, PNG round trip).
Weights: oracle/slomo_ref.make_test_weights(seed) saved in the reference's checkpoint format.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import slomo_ref  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def smooth_frames(N, H, W, seed, dx=3, dy=1, up=8):
    rng = np.random.default_rng(seed)
    ph, pw = H + dy * N + 2 * up, W + dx * N + 2 * up
    base = torch.from_numpy(rng.uniform(20, 235, (1, 1, ph // up + 3, pw // up + 3)).astype(np.float32))
    big = torch.nn.functional.interpolate(base, scale_factor=up, mode="bicubic", align_corners=False)[0, 0]
    big = big.clamp(0, 255).round().to(torch.uint8).numpy()
    return np.stack([big[k * dy:k * dy + H, k * dx:k * dx + W] for k in range(N)])


def run_reference(slomo_mod, frames, sd_fc, sd_at, U, batch_size, auto):
    import cv2
    from torchvision import transforms
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "src"), os.path.join(td, "dst")
        os.makedirs(src); os.makedirs(dst)
        for i, f in enumerate(frames):
            np.save(os.path.join(src, "%08d.npy" % i), f)
        ck = os.path.join(td, "ckpt.pt")
        torch.save({"state_dictFC": sd_fc, "state_dictAT": sd_at}, ck)
        s = slomo_mod.SuperSloMo(model=ck, auto_upsample=auto, upsampling_factor=U, batch_size=batch_size)
        # force the CUDA-branch transforms (slomo.py:157-162) on the CPU device
        normalize = transforms.Normalize(mean=[0.428], std=[1])
        rev = transforms.Normalize(mean=[-0.428], std=[1])
        s.to_tensor = transforms.Compose([transforms.ToTensor(), normalize])
        s.to_image = transforms.Compose([rev, transforms.ToPILImage()])
        H, W = frames.shape[1:]
        times, avg = s.interpolate(src, dst, (W, H))
        n = len(os.listdir(dst))
        out = np.stack([cv2.imread(os.path.join(dst, "%d.png" % i), cv2.IMREAD_GRAYSCALE) for i in range(n)])
    return out, np.asarray(times, np.float64), avg


def weight_digest(sd):
    import hashlib
    h = hashlib.sha1()
    for k in sorted(sd):
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def save_case(name, slomo_mod, N, H, W, U, batch_size, auto=False, seed=0):
    import logging
    logging.disable(logging.WARNING)
    frames = smooth_frames(N, H, W, seed)
    sd_fc = slomo_ref.make_test_weights(100 + seed, 2, 4, head_gain=25.0)
    sd_at = slomo_ref.make_test_weights(200 + seed, 12, 5, head_gain=0.3)
    out, times, avg = run_reference(slomo_mod, frames, sd_fc, sd_at, U, batch_size, auto)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, frames=frames, out=out, times=times, avg=np.array(avg), U=np.array(U),
                        batch_size=np.array(batch_size), auto=np.array(auto), seed=np.array(seed),
                        weights_sha1=np.array(weight_digest(sd_fc) + weight_digest(sd_at)),
                        torch_version=np.array(torch.__version__))
    print("%-24s in %s -> out %s  avgU %.2f  %.1f KB" % (name, frames.shape, out.shape, avg,
                                                        os.path.getsize(path) / 1024))


def main():
    _, _, _, slomo_mod = ref_shim.load_reference()
    torch.set_num_threads(8)
    save_case("slomo_70x100_u3_b2", slomo_mod, N=5, H=70, W=100, U=3, batch_size=2, seed=0)
    save_case("slomo_64x96_u2_b1", slomo_mod, N=3, H=64, W=96, U=2, batch_size=1, seed=1)
    save_case("slomo_96x130_auto", slomo_mod, N=5, H=96, W=130, U=2, batch_size=2, auto=True, seed=2)


if __name__ == "__main__":
    main()
