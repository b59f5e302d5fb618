"""Host-side helpers around the hot path -- mirror of models/utils/tools.py (sizing, frame
conversion) and models/pytorch_msssim (scene detection).  Frame resizes run through
drba_resize_bilinear_f32; scene detection is 32x32 work (out of the hot path) and
stays a few torch ops."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .ops import resize_bilinear


def get_valid_net_inp_size(img, scale, div=64):
    """models/utils/tools.py:41-56: frames are RESIZED (not padded) to a multiple of div/scale."""
    h, w, _ = img.shape
    src_h, src_w, _ = img.shape
    if h * scale % div != 0:
        h = int((h * scale // div + 1) * div / scale)
    if w * scale % div != 0:
        w = int((w * scale // div + 1) * div / scale)
    return {'src_size': (src_h, src_w), 'dst_size': (h, w)}


def to_tensor(img, device):
    """tools.py:33-34: uint8 HWC (BGR, as decoded) -> float NCHW / 255 on the device."""
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).unsqueeze(0).float().to(device) / 255.


def to_cv2(img):
    """tools.py:37-38 (astype(uint8) wraps, as in the reference)."""
    return (img[0].cpu().float().numpy().transpose(1, 2, 0) * 255.).astype(np.uint8)


def resize(tensor, size):
    """tools.py:71-72."""
    if tuple(tensor.shape[2:]) == tuple(size):
        return tensor
    return resize_bilinear(tensor, size=size, align_corners=False)


def to_inp(npInp, dst_size, device):
    return resize(to_tensor(npInp, device), dst_size)


def to_out(tenInp, src_size):
    return to_cv2(resize(tenInp, src_size))


def _gauss1d(n, sigma=1.5):
    g = torch.tensor([math.exp(-(x - n // 2) ** 2 / float(2 * sigma ** 2)) for x in range(n)])
    return g / g.sum()


def ssim_matlab(img1, img2, window_size=11):
    """models/pytorch_msssim/__init__.py:83-136: SSIM with a 3-D Gaussian window over (C, H, W),
    replicate padding; written with the separable form of the same window."""
    mx, mn = float(img1.max()), float(img1.min())
    L = (255 if mx > 128 else 1) - (-1 if mn < -0.5 else 0)
    g = _gauss1d(min(window_size, img1.shape[2], img1.shape[3])).to(img1.device, img1.dtype)
    n = g.numel()
    pad = 5

    def blur(x):
        x = F.pad(x.unsqueeze(1), (pad,) * 6, mode='replicate')
        x = F.conv3d(x, g.view(1, 1, n, 1, 1))
        x = F.conv3d(x, g.view(1, 1, 1, n, 1))
        return F.conv3d(x, g.view(1, 1, 1, 1, n))

    mu1, mu2 = blur(img1), blur(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = blur(img1 * img1) - mu1_sq
    sigma2_sq = blur(img2 * img2) - mu2_sq
    sigma12 = blur(img1 * img2) - mu1_mu2
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    v1 = 2.0 * sigma12 + C2
    v2 = sigma1_sq + sigma2_sq + C2
    return (((2 * mu1_mu2 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2)).mean()


def check_scene(x1, x2, scdet_threshold=0.3):
    """tools.py:27-30."""
    x1 = F.interpolate(x1.float(), (32, 32), mode='bilinear', align_corners=False)
    x2 = F.interpolate(x2.float(), (32, 32), mode='bilinear', align_corners=False)
    return bool(ssim_matlab(x1, x2) < scdet_threshold)
