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


# ---- fused frame ingest / egress ----------------------------------------------------
def frame_ingest_u8(u8_hwc, dst_size, out=None):
    """to_inp (tools.py:59-62) in one kernel: device uint8 [h,w,3] -> float [1,3,H,W] = resize(x / 255)."""
    from . import _lib
    from ._torch_util import ptr, require_cuda, stream_ptr
    require_cuda(u8_hwc)
    assert u8_hwc.dtype == torch.uint8 and u8_hwc.dim() == 3 and u8_hwc.shape[2] == 3 and u8_hwc.is_contiguous()
    h, w = int(u8_hwc.shape[0]), int(u8_hwc.shape[1])
    H, W = int(dst_size[0]), int(dst_size[1])
    if out is None:
        out = torch.empty((1, 3, H, W), dtype=torch.float32, device=u8_hwc.device)
    with torch.cuda.device(u8_hwc.device):
        with _lib.launch("frame_ingest_u8", 1, nbytes=float(h * w * 3 + H * W * 12)):
            rc = _lib.lib().drba_frame_ingest_u8(ptr(u8_hwc), ptr(out), h, w, H, W, stream_ptr(u8_hwc.device))
    _lib.check(rc, "drba_frame_ingest_u8")
    return out


def frame_egress_u8(frame, src_size, out=None):
    """to_out (tools.py:65-68) in one kernel: float [1,3,H,W] -> device uint8 [h,w,3] = uint8(resize(x) * 255)."""
    from . import _lib
    from ._torch_util import ptr, require_cuda, stream_ptr
    require_cuda(frame)
    x = frame.float().contiguous()
    H, W = int(x.shape[2]), int(x.shape[3])
    h, w = int(src_size[0]), int(src_size[1])
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        with _lib.launch("frame_egress_u8", 1, nbytes=float(h * w * 3 + H * W * 12)):
            rc = _lib.lib().drba_frame_egress_u8(ptr(x), ptr(out), H, W, h, w, stream_ptr(x.device))
    _lib.check(rc, "drba_frame_egress_u8")
    return out


class FrameIO:
    """Double-buffered host <-> device frame traffic around the interpolation stream.

    upload(np/pinned uint8 frame) -> float net-input tensor: H2D copy + ingest kernel on a copy stream;
    download(float frame) -> pinned uint8 host buffer: egress kernel + D2H copy on a second copy stream.
    Both are ordered against the caller's current stream with events only (no host synchronisation);
    `drain()` waits for all downloads issued so far.  Replaces the reference's synchronous
    to_inp / to_out (tools.py:59-68), whose fp32 D2H moves 4x the bytes."""

    def __init__(self, src_size, dst_size, device, depth=4, out_depth=8):
        self.src, self.dst, self.device = tuple(src_size), tuple(dst_size), torch.device(device)
        self.h2d, self.d2h = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        h, w = self.src
        H, W = self.dst
        self._in_u8 = [torch.empty((h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self._in_f = [torch.empty((1, 3, H, W), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self._out_u8 = [torch.empty((h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(out_depth)]
        self._out_host = [torch.empty((h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(out_depth)]
        self._out_done = [None] * out_depth
        self._in_free = [None] * depth      # event: the compute stream has finished reading slot k
        self._i = self._o = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def upload(self, host_u8):
        """host_u8: pinned uint8 tensor [h,w,3] (or numpy array).  Returns the float frame; the current
        stream is made to wait for it."""
        if not torch.is_tensor(host_u8):
            host_u8 = torch.from_numpy(np.ascontiguousarray(host_u8))
        k = self._i % len(self._in_u8)
        self._i += 1
        cur = torch.cuda.current_stream(self.device)
        if self._in_free[k] is not None:
            self.h2d.wait_event(self._in_free[k])
        with torch.cuda.stream(self.h2d):
            self._in_u8[k].copy_(host_u8, non_blocking=True)
            frame_ingest_u8(self._in_u8[k], self.dst, out=self._in_f[k])
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        cur.wait_event(ev)
        self.h2d_bytes += host_u8.numel()
        self._last_slot = k
        return self._in_f[k]

    def release_inputs(self, frames=None):
        """Call after the work that reads uploaded frames has been enqueued on the current stream: marks their input
        slots as reusable once that work completes.  frames: the uploaded tensors that work reads (None = every slot).
        Naming them lets the next frame be uploaded while the window that does not touch its slot is still running
        (upload one window ahead: the H2D copy then overlaps the compute instead of stalling it)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        if frames is None:
            self._in_free = [ev] * len(self._in_free)
            return
        ptrs = {f.data_ptr() for f in frames}
        for k, buf in enumerate(self._in_f):
            if buf.data_ptr() in ptrs:
                self._in_free[k] = ev

    def download(self, frame):
        """Enqueue egress + D2H of a float frame produced on the current stream; returns the pinned host buffer
        (valid after drain() or after its own event)."""
        k = self._o % len(self._out_u8)
        self._o += 1
        if self._out_done[k] is not None:
            self._out_done[k].synchronize()         # host buffer k is about to be overwritten
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.d2h.wait_event(ev)
        with torch.cuda.stream(self.d2h):
            frame.record_stream(self.d2h)
            frame_egress_u8(frame, self.src, out=self._out_u8[k])
            self._out_host[k].copy_(self._out_u8[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h)
        self._out_done[k] = done
        self.d2h_bytes += self._out_host[k].numel()
        return self._out_host[k], done

    def drain(self):
        self.d2h.synchronize()
