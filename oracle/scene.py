"""CPU restatement of scene detection (test infrastructure, see oracle/__init__.py).

  models/utils/tools.py:27-30                  check_scene: 32x32 bilinear thumbnails, ssim_matlab < threshold
  models/pytorch_msssim/__init__.py:9-11       gaussian(11, 1.5)
  models/pytorch_msssim/__init__.py:22-27      create_window_3d: outer products of the 1-D window
  models/pytorch_msssim/__init__.py:83-136     ssim_matlab: replicate-padded conv3d over the (C, H, W) volume

Plain torch fp32 on the CPU, the full 11^3 window as the reference builds it (not the separable form the kernel
uses).  Pinned by tests/golden/scene_golden.npz (tests/golden/make_golden_scene.py imports the reference)."""
from math import exp

import torch
import torch.nn.functional as F


def _window_3d(n):
    g = torch.Tensor([exp(-(x - n // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(n)])      # :9-11
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t())                                                                          # :23-24
    w3 = w2.unsqueeze(2) @ g.t()                                                              # :25
    return w3.expand(1, 1, n, n, n).contiguous()


def ssim_matlab(img1, img2):
    """pytorch_msssim/__init__.py:83-136 with the defaults check_scene uses."""
    max_val = 255 if torch.max(img1) > 128 else 1                                             # :86-89
    min_val = -1 if torch.min(img1) < -0.5 else 0                                             # :91-94
    L = max_val - min_val
    win = _window_3d(min(11, img1.shape[2], img1.shape[3]))
    a, b = img1.unsqueeze(1), img2.unsqueeze(1)

    def blur(x):
        return F.conv3d(F.pad(x, (5, 5, 5, 5, 5, 5), mode="replicate"), win)                  # :110-111

    mu1, mu2 = blur(a), blur(b)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = blur(a * a) - mu1_sq                                                                 # :117-119
    s2 = blur(b * b) - mu2_sq
    s12 = blur(a * b) - mu12
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    v1, v2 = 2.0 * s12 + C2, s1 + s2 + C2
    return (((2 * mu12 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2)).mean()                    # :127-130


def check_scene(x1, x2, scdet_threshold=0.3):
    """tools.py:27-30; returns (flag, ssim)."""
    t1 = F.interpolate(x1.float(), (32, 32), mode="bilinear", align_corners=False)
    t2 = F.interpolate(x2.float(), (32, 32), mode="bilinear", align_corners=False)
    s = float(ssim_matlab(t1, t2))
    return s < scdet_threshold, s
