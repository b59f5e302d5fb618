"""Fixture for check_scene / ssim_matlab (models/utils/tools.py:27-30, models/pytorch_msssim/__init__.py:83-136):
frame pairs of several sizes and similarity levels with the SSIM the REFERENCE computes for them (CPU, fp32):
    python tests/golden/make_golden_scene.py
Per size the file holds three arrays (base, other, noise); `scene_pairs` below derives the pairs from them and is
imported by the tests, so generator and tests build identical inputs."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

SIZES = (("a", (96, 160)), ("b", (64, 64)), ("c", (135, 240)), ("d", (32, 32)), ("e", (100, 37)))


def scene_pairs(base, other, noise):
    """[(tag, x1, x2)] from the stored arrays (torch fp32 tensors [1,3,h,w])."""
    return [("same", base, base.clone()),
            ("near", base, (base + 0.02 * noise).clamp(0, 1)),
            ("mix30", base, 0.7 * base + 0.3 * other), ("mix60", base, 0.4 * base + 0.6 * other),
            ("mix75", base, 0.25 * base + 0.75 * other),
            ("cut", base, other), ("bytes", base * 255.0, other * 255.0),
            ("signed", base * 2 - 1, (0.8 * base + 0.2 * other) * 2 - 1)]


def main():
    sys.path.insert(0, "/root/reference")
    import models.pytorch_msssim as ms
    ms.device = torch.device("cpu")
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(21)

    def smooth(h, w):
        lo = torch.rand((1, 3, h // 8 + 2, w // 8 + 2), generator=g)
        return F.interpolate(lo, scale_factor=8, mode="bicubic", align_corners=False)[:, :, 4:4 + h, 4:4 + w].clamp(0, 1).contiguous()

    out = {}
    for name, (h, w) in SIZES:
        base, other, noise = smooth(h, w), smooth(h, w), torch.randn((1, 3, h, w), generator=g)
        out[name + "_base"], out[name + "_other"], out[name + "_noise"] = base.numpy(), other.numpy(), noise.numpy()
        for tag, x1, x2 in scene_pairs(base, other, noise):
            t1 = F.interpolate(x1, (32, 32), mode="bilinear", align_corners=False)      # tools.py:28-29
            t2 = F.interpolate(x2, (32, 32), mode="bilinear", align_corners=False)
            s = float(ms.ssim_matlab(t1, t2))
            out[f"{name}_{tag}_ssim"] = np.float32(s)
            print(name, tag, s)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scene_golden.npz"), **out)


if __name__ == "__main__":
    main()
