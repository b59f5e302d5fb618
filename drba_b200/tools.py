"""Host-side helpers around the hot path -- mirror of models/utils/tools.py (sizing, frame
conversion, scene detection).  Frame resizes run through drba_resize_bilinear_f32, scene detection
(check_scene + models/pytorch_msssim ssim_matlab) through drba_check_scene_f32 (one kernel, no host sync)."""

import numpy as np
import torch

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


def check_scene_async(x1, x2, scdet_threshold=0.3, out=None):
    """tools.py:27-30 as ONE kernel (csrc/scene.cu: 32x32 bilinear thumbnails + ssim_matlab + compare), nothing
    synchronises.  x1, x2: [1,3,H,W] fp32 CUDA frames.  Returns (ssim, flag): device tensors of one element each
    (flag 1 = scene cut), or writes into `out = (ssim_tensor, flag_tensor)` (may be mapped pinned host memory)."""
    from . import _lib
    from ._torch_util import ptr, require_cuda, stream_ptr
    require_cuda(x1, x2)
    a, b = x1.float().contiguous(), x2.float().contiguous()
    if a.dim() != 4 or a.shape[0] != 1 or a.shape[1] != 3 or a.shape != b.shape:
        raise ValueError("check_scene expects two [1,3,H,W] frames of the same size")
    H, W = int(a.shape[2]), int(a.shape[3])
    if out is None:
        out = (torch.empty(1, dtype=torch.float32, device=a.device), torch.empty(1, dtype=torch.int32, device=a.device))
    with torch.cuda.device(a.device):
        with _lib.launch("check_scene", 1, nbytes=float(2 * 3 * 32 * 32 * 16)):
            rc = _lib.lib().drba_check_scene_f32(ptr(a), ptr(b), 0, 0, 1, H, W, float(scdet_threshold), ptr(out[0]), ptr(out[1]),
                                                 stream_ptr(a.device))
    _lib.check(rc, "drba_check_scene_f32")
    return out


def ssim_matlab(img1, img2, window_size=11):
    """models/pytorch_msssim/__init__.py:83-136 on two [1,3,32,32]-or-larger CUDA frames: the SSIM value as a
    one-element device tensor (the kernel works on 32x32 thumbnails, which is the only way the reference calls it)."""
    if tuple(img1.shape[2:]) != (32, 32) or window_size != 11:
        raise ValueError("drba_b200.tools.ssim_matlab serves check_scene's call: 32x32 frames, window 11")
    return check_scene_async(img1, img2, 0.0)[0]


def check_scene(x1, x2, scdet_threshold=0.3):
    """tools.py:27-30, same signature and result (a Python bool: this form waits for the kernel; the driver loops
    use SceneDetector, which does not)."""
    return bool(check_scene_async(x1, x2, scdet_threshold)[1].item())


class SceneDetector:
    """Scene flags without stalling the stream.  submit(a, b) enqueues the check of a frame pair on
    the current stream and returns a ticket; result(ticket) returns the bool.  The kernel writes its flag into pinned
    host memory that the device sees directly, so result() only waits on the ticket's event -- which has long
    completed when the driver loop asks one window later (the reference stalls three times per pair)."""

    def __init__(self, device, threshold=0.3, depth=64):
        self.device = torch.device(device)
        self.threshold = float(threshold)
        self.depth = depth
        self._ssim = torch.zeros(depth, dtype=torch.float32).pin_memory()
        self._flag = torch.zeros(depth, dtype=torch.int32).pin_memory()
        self._events = [None] * depth
        self._n = 0

    def submit(self, a, b):
        k = self._n % self.depth
        self._n += 1
        check_scene_async(a, b, self.threshold, out=(self._ssim[k:k + 1], self._flag[k:k + 1]))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events[k] = ev
        return k

    def result(self, ticket):
        self._events[ticket].synchronize()
        return bool(self._flag[ticket].item())

    def ssim(self, ticket):
        self._events[ticket].synchronize()
        return float(self._ssim[ticket].item())

    def __call__(self, a, b):
        return self.result(self.submit(a, b))


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

    upload(np/pinned uint8 frame) -> float net-input tensor: H2D copy on a copy stream, ingest kernel on the caller's;
    download(float frame) -> pinned uint8 host buffer: egress kernel on the caller's stream, D2H copy on a second copy stream.
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
        self._u8_free = [None] * depth      # event: the ingest kernel has consumed the uint8 staging buffer of slot k
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
        if self._u8_free[k] is not None:
            self.h2d.wait_event(self._u8_free[k])
        # Only the COPY rides the copy stream; the ingest kernel runs on the caller's stream.  A kernel on a side stream
        # competes for SMs with the persistent conv programs, whose grid barrier needs all their CTAs resident: measured,
        # ONE 13 us side-stream kernel per window costs 73 us of window time, the same kernel in stream 14 us.
        with torch.cuda.stream(self.h2d):
            self._in_u8[k].copy_(host_u8, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        cur.wait_event(ev)
        frame_ingest_u8(self._in_u8[k], self.dst, out=self._in_f[k])
        used = torch.cuda.Event()
        used.record(cur)
        self._u8_free[k] = used             # the staging buffer may be overwritten once the ingest kernel has read it
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
        cur = torch.cuda.current_stream(self.device)
        if self._out_done[k] is not None:
            cur.wait_event(self._out_done[k])       # the previous D2H out of device buffer k has finished (host already waited)
        frame_egress_u8(frame, self.src, out=self._out_u8[k])      # on the caller's stream (see upload)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.d2h.wait_event(ev)
        with torch.cuda.stream(self.d2h):
            self._out_host[k].copy_(self._out_u8[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h)
        self._out_done[k] = done
        self.d2h_bytes += self._out_host[k].numel()
        return self._out_host[k], done

    def drain(self):
        self.d2h.synchronize()
