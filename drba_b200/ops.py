"""Stand-alone sampling operators (backward warp, bilinear resize, RIFE flow inversion)."""
import torch

from . import _lib
from ._torch_util import Workspace, f32c, ptr, require_cuda, stream_ptr


def backwarp(tenInput, tenFlow, padding_mode="border"):
    """warplayer.warp (models/rife_426_heavy/warplayer.py:8-22, border) and MetricNet.backwarp
    (models/model_gmfss/MetricNet.py:10-20, zeros): bilinear, align_corners=True."""
    require_cuda(tenInput, tenFlow)
    dtype = tenInput.dtype
    x, f = f32c(tenInput), f32c(tenFlow)
    n, c, h, w = x.shape
    assert f.shape == (n, 2, h, w)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        with _lib.launch("backwarp", 1, nbytes=4.0 * n * h * w * (2 * c + 2)):
            rc = _lib.lib().drba_backwarp_f32(ptr(x), ptr(f), ptr(out), n, c, h, w,
                                              {"border": 0, "zeros": 1}[padding_mode], stream_ptr(x.device))
    _lib.check(rc, "drba_backwarp_f32")
    return out.to(dtype)


def resize_bilinear(x, size=None, scale_factor=None, align_corners=False):
    """F.interpolate(x, size= | scale_factor=, mode='bilinear', align_corners=...)."""
    require_cuda(x)
    dtype = x.dtype
    xf = f32c(x)
    n, c, h, w = xf.shape
    if size is not None:
        oh, ow = int(size[0]), int(size[1])
        rh, rw = h / oh, w / ow
    else:
        oh, ow = int(h * scale_factor), int(w * scale_factor)
        rh = rw = 1.0 / scale_factor
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        with _lib.launch("resize_bilinear", 1, nbytes=4.0 * n * c * (h * w + oh * ow)):
            rc = _lib.lib().drba_resize_bilinear_f32(ptr(xf), ptr(out), n, c, h, w, oh, ow,
                                                     int(bool(align_corners)), rh, rw, stream_ptr(xf.device))
    _lib.check(rc, "drba_resize_bilinear_f32")
    return out.to(dtype)


def rife_invert_flow(flow_t0):
    """models/rife.py:59-73: 2 * fill(-splat_avg(flow, flow), holes <- max(H, W))."""
    require_cuda(flow_t0)
    f = f32c(flow_t0)
    n, _, h, w = f.shape
    out = torch.empty_like(f)
    L = _lib.lib()
    with torch.cuda.device(f.device):
        need = L.drba_rife_invert_flow_workspace_bytes(n, h, w)
        ws = Workspace.get(need, f.device)
        with _lib.launch("rife_invert_flow", 2, nbytes=16.0 * n * h * w):
            rc = L.drba_rife_invert_flow_f32(ptr(f), ptr(out), n, h, w, ws.data_ptr(), need, stream_ptr(f.device))
    _lib.check(rc, "drba_rife_invert_flow_f32")
    return out
