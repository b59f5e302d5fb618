"""Distance Ratio Map operators -- host-side mirror of models/drm.py.

``get_drm_t`` (drm.py:10-62), ``calc_drm_rife`` (:65-107), ``calc_drm_gmfss`` (:110-155) and
``calc_drm_rife_auxiliary`` (:158-195) with the reference's signatures and result dicts.  Each
is one call into libdrba_b200.so (csrc/drm.cu): a fused scatter kernel + a resolve kernel
instead of ~35 elementwise launches, 4 softsplats and 2 boolean-mask host syncs.

The distance is computed in fp32 and the maps are returned in the flow's dtype, as the
reference does (tools.py:77-80).  Unlike the reference, intermediate ratios are kept in fp32
even when the flows arrive as fp16/bf16 (the reference rounds them to the flow dtype between
steps); with fp32 flows the arithmetic is operation-for-operation the reference's.
``only=`` (extension): name of the single map the caller needs; the other one is not computed.
"""
import torch

from . import _lib
from ._torch_util import Workspace, f32c, ptr, require_cuda, stream_ptr


def distance_calculator(_x):
    """models/utils/tools.py:77-80 (kept for API completeness; the fused kernels do not call it)."""
    dtype = _x.dtype
    u, v = _x[:, 0:1].float(), _x[:, 1:].float()
    return torch.sqrt(u ** 2 + v ** 2).to(dtype)


def get_drm_t(drm, t, precision=1e-3):
    require_cuda(drm)
    dtype = drm.dtype
    x = f32c(drm)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        with _lib.launch("get_drm_t", 1, nbytes=8.0 * x.numel()):
            rc = _lib.lib().drba_get_drm_t_f32(ptr(x), float(t), float(precision), ptr(out), x.numel(),
                                               stream_ptr(x.device))
    _lib.check(rc, "drba_get_drm_t_f32")
    return out.to(dtype)


def _drm_rife(t, flow10, flow12, metric10, metric12, linear, only):
    require_cuda(flow10, flow12, metric10, metric12)
    dtype = flow10.dtype
    f10, f12 = f32c(flow10), f32c(flow12)
    soft = metric10 is not None and metric12 is not None
    m10, m12 = (f32c(metric10), f32c(metric12)) if soft else (None, None)
    n, _, h, w = f10.shape
    assert f12.shape == f10.shape and f10.shape[1] == 2
    names = ("drm_t1_t01", "drm_t1_t12")
    assert only in (None,) + names
    outs = {k: (torch.empty((n, 1, h, w), dtype=torch.float32, device=f10.device)
                if only in (None, k) else None) for k in names}
    L = _lib.lib()
    with torch.cuda.device(f10.device):
        need = L.drba_drm_workspace_bytes(n, h, w)
        ws = Workspace.get(need, f10.device)
        nmaps = sum(v is not None for v in outs.values())
        with _lib.launch("drm_rife", 2, nbytes=float(n * h * w * (16 + (8 if soft else 0) + 4 * nmaps))):
            rc = L.drba_drm_rife_f32(float(t), ptr(f10), ptr(f12), ptr(m10), ptr(m12), int(bool(linear)),
                                     ptr(outs[names[0]]), ptr(outs[names[1]]), n, h, w,
                                     ws.data_ptr(), need, stream_ptr(f10.device))
    _lib.check(rc, "drba_drm_rife_f32")
    return {k: v.to(dtype) for k, v in outs.items() if v is not None}


def calc_drm_rife(t, flow10, flow12, linear=False, only=None):
    return _drm_rife(t, flow10, flow12, None, None, linear, only)


def calc_drm_rife_auxiliary(t, flow10, flow12, metric10, metric12, linear=False, only=None):
    return _drm_rife(t, flow10, flow12, metric10, metric12, linear, only)


def calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear=False):
    require_cuda(flow10, flow12, metric10, metric12)
    dtype = flow10.dtype
    f10, f12 = f32c(flow10), f32c(flow12)
    soft = metric10 is not None and metric12 is not None
    m10, m12 = (f32c(metric10), f32c(metric12)) if soft else (None, None)
    n, _, h, w = f10.shape
    assert f12.shape == f10.shape and f10.shape[1] == 2
    names = ("drm0t_t01", "drm1t_t01", "drm1t_t12", "drm2t_t12")
    outs = [torch.empty((n, 1, h, w), dtype=torch.float32, device=f10.device) for _ in names]
    L = _lib.lib()
    with torch.cuda.device(f10.device):
        need = L.drba_drm_workspace_bytes(n, h, w)
        ws = Workspace.get(need, f10.device)
        with _lib.launch("drm_gmfss", 3, nbytes=float(n * h * w * (16 + (8 if soft else 0) + 16))):
            rc = L.drba_drm_gmfss_f32(float(t), ptr(f10), ptr(f12), ptr(m10), ptr(m12), int(bool(linear)),
                                      *[ptr(o) for o in outs], n, h, w, ws.data_ptr(), need, stream_ptr(f10.device))
    _lib.check(rc, "drba_drm_gmfss_f32")
    return {k: v.to(dtype) for k, v in zip(names, outs)}
